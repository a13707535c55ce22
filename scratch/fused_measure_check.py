"""One-shot check of the experimental fused measurement (QCA_FLAG_FUSED_MEASURE) against the per-cell path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qca_b200
from qca_b200 import _lib

worst = 0.0
for n in (22, 24, 28):
    rules = qca_b200.Rules(n, range(2, 4), 2)
    plist = qca_b200.states.plist("triple_blinker", rules)   # basis state: one real plane in the rotated frame
    res = []
    for flags in (_lib.QCA_FLAG_LOOSE_BOUND | _lib.QCA_FLAG_PERCELL_MEASURE, _lib.QCA_FLAG_LOOSE_BOUND | _lib.QCA_FLAG_FUSED_MEASURE):
        eng = _lib.ExactEngine(rules, device=0, flags=flags)
        eng.set_product_state(plist)
        if eng.stats()["planes"] != 1:
            print("n", n, "two planes: fused path not used"); 
        eng.step(1.0, 2)                       # spreads over the connected component of the basis state
        eng.measure()
        t0 = time.perf_counter()
        out = eng.measure()
        dt = time.perf_counter() - t0
        res.append((out, dt, eng.stats()["planes"]))
        eng.close()
    d = max(np.abs(a - b).max() for a, b in zip(res[0][0], res[1][0]))
    worst = max(worst, d)
    print(f"N={n}: planes {res[0][2]}, max |fused - per-cell| = {d:.3e}, per-cell {res[0][1] * 1e3:.2f} ms, fused {res[1][1] * 1e3:.2f} ms", flush=True)
print("FUSED_OK" if worst < 1e-12 else "FUSED_MISMATCH", worst)
