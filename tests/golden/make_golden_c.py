#!/usr/bin/env python
"""Golden vectors at register sizes the reference cannot reach (its dense U stops near N = 13),
produced by the C oracle (``oracle/qca_oracle_c.c``: matrix-free Hermitian H, forward Chebyshev
series in complex arithmetic -- itself pinned to the reference's fixtures for N <= 14 by
``tests/test_oracle_golden.py``).  They tie the kernels that are actually benchmarked -- three tile
passes, sharded registers -- to an independent computation:

    python tests/golden/make_golden_c.py            # minutes per case on 8 cores

Stored: the populations and entropies before every step (the reference's measure-then-step loop,
quantum_game.py:82-119), nothing of size 2^N.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))

CASES = [
    # name, ncells, distance, lo, hi, state, effective step, measured rows
    ("cexact_triple22_d2", 22, 2, 2, 4, "triple_blinker", 1.0, 4),   # two tile passes; 2/4/8-rank worker size
    ("cexact_blinker20", 20, 1, 1, 2, "blinker", 1.0, 6),            # BASELINE configs[1] rule and size
    ("cexact_triple26_d2", 26, 2, 2, 4, "triple_blinker", 1.0, 3),   # three tile passes (the N=30 bench geometry)
]


def main() -> None:
    import qca_oracle as oracle
    import qca_oracle_c as oc
    only = os.environ.get("GOLDEN_ONLY")
    for (name, n, d, lo, hi, state, tau, rows) in CASES:
        if only and only not in name:
            continue
        t0 = time.time()
        s = oc.Stepper(n, d, lo, hi)
        s.set_product_state(oracle.initial_plist(state, n, d))
        pop, ent = np.zeros((rows, n)), np.zeros((rows, n))
        for k in range(rows):
            pop[k], ent[k] = s.measure()
            if k + 1 < rows:
                s.step(tau)
        spec = dict(kind="cexact", name=name, ncells=n, distance=d, lo=lo, hi=hi, state=state)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), spec=json.dumps(spec), population=pop,
                            single_site_entropy=ent, effective_step_size=tau, chebyshev_terms=s.terms,
                            norm2=float(np.vdot(s.psi, s.psi).real))
        print(f"golden: {name} rows={rows} terms={s.terms} {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
