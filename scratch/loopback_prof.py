"""Per-pass timing of the SHARDED tile-pass kernel on one GPU (remote operands looped back to this
rank's own planes), next to the unsharded kernel on a register of the same local size.
usage: python scratch/loopback_prof.py [ncells] [world] [rank]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qca_b200
from qca_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rbits = world.bit_length() - 1


def run(tag, eng, plist):
    eng.set_product_state(plist)
    if tag != "local":
        eng.resolve_planes(*eng.plane_flags())
    eng.step(1.0, 1)
    torch.cuda.synchronize()
    eng.reset_stats()
    if os.environ.get("QCA_NCU") == tag.split()[0]:
        torch.cuda.profiler.start()      # ncu --profile-from-start off
    eng.step(1.0, 1)
    torch.cuda.synchronize()
    if os.environ.get("QCA_NCU") == tag.split()[0]:
        torch.cuda.profiler.stop()
    st = eng.stats()
    applies = st["pass_launches"] // st["passes_per_apply"]
    by_pass = [m / applies for m in st["profiled_ms_by_pass"][:st["passes_per_apply"]]]
    gb = st["pass_bytes"] / st["pass_launches"] / 1e9
    print(f"{tag}: local bits {st['local_bits']}, {st['pass_launches']} launches, avg ms by pass {[round(x, 3) for x in by_pass]}, "
          f"local GB/launch {gb:.2f}, local GB/s {gb / (st['profiled_pass_ms'] / st['profiled_pass_launches'] * 1e-3):.0f}, "
          f"remote GB/apply {st['remote_bytes'] / applies / 1e9:.2f}", flush=True)
    eng.close()


flags = _lib.QCA_FLAG_PROFILE
warm = _lib.ExactEngine(qca_b200.Rules(14, range(2, 4), 2), device=0)
warm.set_product_state(qca_b200.states.plist("blinker", qca_b200.Rules(14, range(2, 4), 2)))
warm.step(1.0, 1)
warm.close()
rules = qca_b200.Rules(n, range(2, 4), 2)
plist = qca_b200.states.plist("triple_blinker", rules)
eng = _lib.ExactEngine(rules, device=0, world_size=world, rank=rank, flags=flags)
eng.loopback_peers()
run(f"sharded N={n} world={world} rank={rank} (loopback)", eng, plist)
rules1 = qca_b200.Rules(n - rbits, range(2, 4), 2)
run("local", _lib.ExactEngine(rules1, device=0, flags=flags), qca_b200.states.plist("triple_blinker", rules1))
