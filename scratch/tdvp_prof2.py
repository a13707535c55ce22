"""Where does a chi=256 2TDVP step spend its time?  GPU-busy time by kernel (torch.profiler) against
wall time, plus phase timers.  usage: python scratch/tdvp_prof2.py [ncells] [chi] [1tdvp|2tdvp]"""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qca_b200
from torch.profiler import profile, ProfilerActivity

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 256
algorithm = sys.argv[3] if len(sys.argv) > 3 else "2tdvp"


def random_mps(n, chi, seed=0):
    rng = np.random.default_rng(seed)
    dims = [min(2 ** i, 2 ** (n - i), chi) for i in range(n + 1)]
    return qca_b200.MPS([(rng.standard_normal((2, dims[i], dims[i + 1])) + 1j * rng.standard_normal((2, dims[i], dims[i + 1]))) / np.sqrt(2 * dims[i]) for i in range(n)])


rules = qca_b200.Rules(n, range(1, 2), 1)
args = qca_b200.Args(rules=rules, step_size=0.005, algorithm=algorithm, max_bond_dim=chi, svd_epsilon=1e-14)
t_init = time.time()
algo = qca_b200.TDVP(random_mps(n, chi), qca_b200.MPO.hamiltonian_from_rules(rules), args)
torch.cuda.synchronize()
print(f"{algorithm}: constructor (canonicalisation + right environments) {time.time() - t_init:.2f} s")
for _ in range(2):
    algo.do_time_step()
torch.cuda.synchronize()
t0 = time.time()
algo.do_time_step()
torch.cuda.synchronize()
wall = time.time() - t0
print(f"{algorithm} N={n} chi={chi}: wall {wall:.3f} s/step, heff applications {algo.heff_applications // 3}")

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    algo.do_time_step()
    torch.cuda.synchronize()
ev = prof.key_averages()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in ev if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"GPU busy (sum of kernels) {tot:.1f} ms over {sum(r[2] for r in rows)} launches")
for k, ms, c in rows[:40]:
    print(f"{ms:9.2f} ms {c:7d}x  {ms / c * 1e3:8.1f} us  {k[:110]}")

# phase timers (synchronising: inflates, but shows the split)
import qca_b200.algorithms.tdvp as T
timers = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.time(); r = fn(*a, **k); torch.cuda.synchronize()
        timers[name] = timers.get(name, 0.0) + time.time() - t
        return r
    return w
T.gram_svd = timed("gram_svd", T.gram_svd)
algo._expm_apply = timed("expm_apply", algo._expm_apply)
algo._grow_left = timed("grow_env", algo._grow_left)
algo._grow_right = timed("grow_env", algo._grow_right)
torch.cuda.synchronize(); t0 = time.time()
algo.do_time_step(); torch.cuda.synchronize()
print(f"synchronised step {time.time() - t0:.3f} s; phases {dict((k, round(v, 3)) for k, v in timers.items())}")
