"""Timeline of the pipelined end-to-end leg (two engines, two host threads): where does the time between the
device-resident rate and the end-to-end rate go?  torch.profiler (CUPTI) sees the library's kernels and copies."""
import os, sys, time, threading as th, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qca_b200
from qca_b200 import _lib
from torch.profiler import profile, ProfilerActivity

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
per_engine = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rules = qca_b200.Rules(n, range(2, 4), 2)
plist = qca_b200.states.plist("triple_blinker", rules)
namps = 1 << n
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
engines = [_lib.ExactEngine(rules, device=0, stream=s.cuda_stream) for s in streams]
hosts = [torch.empty(2 * namps, dtype=torch.float64).pin_memory() for _ in range(2)]
for e, h in zip(engines, hosts):
    e.set_product_state(plist); e.get_state_ptr(h.data_ptr(), namps)
gate = th.Lock()
marks = []
def run(k, count):
    e, h, s = engines[k], hosts[k], streams[k]
    for i in range(count):
        t0 = time.perf_counter(); e.set_state_ptr(h.data_ptr(), namps); t1 = time.perf_counter()
        gate.acquire(); t2 = time.perf_counter()
        try:
            e.measure(); e.step(1.0, 1); s.synchronize()
        finally:
            gate.release()
        t3 = time.perf_counter(); e.get_state_ptr(h.data_ptr(), namps); t4 = time.perf_counter()
        marks.append((k, i, t1 - t0, t2 - t1, t3 - t2, t4 - t3))
for k in range(2):
    run(k, 1)
torch.cuda.synchronize(); marks.clear()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    ts = [th.Thread(target=run, args=(k, per_engine)) for k in range(2)]
    [t.start() for t in ts]; [t.join() for t in ts]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
print(f"pipelined: {2 * per_engine} states in {wall:.3f} s = {wall / (2 * per_engine):.3f} s per state")
for m in sorted(marks, key=lambda m: (m[1], m[0])):
    print("engine %d state %d: upload %.3f s, wait for the SMs %.3f s, measure+step %.3f s, download %.3f s" % m)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
import collections
by = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    key = e.name.split("(")[0][:60]
    by[key][0] += 1; by[key][1] += e.device_time
for k, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{t / 1e3:10.1f} ms {c:6d}x {t / c:9.1f} us  {k}")
# pass-kernel duration while a copy is in flight vs not
copies = sorted((e.time_range.start, e.time_range.end) for e in ev if "Memcpy" in e.name and e.device_time > 500)
def overlaps(a, b):
    import bisect
    i = bisect.bisect_left(copies, (a, a))
    for j in (i - 1, i):
        if 0 <= j < len(copies) and copies[j][0] < b and copies[j][1] > a:
            return True
    return False
stat = {True: [0, 0.0], False: [0, 0.0]}
for e in ev:
    if "pass_kernel" in e.name:
        o = overlaps(e.time_range.start, e.time_range.end)
        stat[o][0] += 1; stat[o][1] += e.device_time
for o in (False, True):
    c, t = stat[o]
    if c:
        print(f"pass kernels {'WITH' if o else 'without'} a PCIe copy in flight: {c} launches, mean {t / c / 1e3:.3f} ms")
