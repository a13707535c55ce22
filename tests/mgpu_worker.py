"""Multi-GPU parity worker: run under torchrun (one rank per GPU).  Every rank drives its shard
through qca_b200.Exact / ShardedExactEngine; results are compared with a single-GPU engine (itself
pinned to the oracle by test_exact_gpu.py) and, at small sizes, with the CPU oracle directly."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch
import torch.distributed as dist

import qca_b200
import qca_oracle as oracle
from qca_b200 import _lib, sharding


def main() -> int:
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = []

    def check(name, cond, detail=""):
        if not cond:
            failures.append(f"[rank {rank}] {name} {detail}")

    rbits = world.bit_length() - 1
    # (ncells, distance, lo, hi, state): generic kernel, 2-pass and 3-pass fast kernel regimes
    cases = [(10 + rbits, 1, 1, 2, "single"), (9 + rbits, 2, 2, 4, "equal_superposition"),
             (15 + rbits, 1, 1, 2, "blinker"), (17 + rbits, 2, 2, 4, "triple_blinker"),
             (17 + rbits, 2, 1, 3, "gradient"), (19 + rbits, 1, 1, 2, "gradient"),   # scattered shard cells from here on
             (23 + rbits, 2, 2, 4, "triple_blinker")]
    for (n, d, lo, hi, state) in cases:
        rules = qca_b200.Rules(n, range(lo, hi), d)
        plist = qca_b200.states.plist(state, rules)
        eng = sharding.ShardedExactEngine(rules, device=local)
        assert eng.world == world and eng.local_amps == (1 << n) // world
        eng.set_product_state(plist)
        ref = _lib.ExactEngine(rules, device=local) if rank == 0 else None
        if ref is not None:
            ref.set_product_state(plist)
            check(f"planes n={n} {state}", ref.stats()["planes"] == eng.stats()["planes"])
        for k in range(2):
            got = eng.measure()
            if ref is not None:
                want = ref.measure()
                for a, b, nm in zip(got, want, ("pop", "dpop", "ent", "bond")):
                    if nm == "dpop":
                        continue
                    check(f"measure {nm} n={n} {state} step {k}", np.abs(a - b).max() < 1e-10, f"{np.abs(a - b).max():.3e}")
            eng.step(1.0, 1)
            if ref is not None:
                ref.step(1.0, 1)
        check(f"norm n={n} {state}", abs(eng.norm2() - 1.0) < 1e-11)
        if n <= 22:
            full = eng.get_state()
            if ref is not None:
                check(f"state n={n} {state}", np.abs(full - ref.get_state()).max() < 1e-11,
                      f"{np.abs(full - ref.get_state()).max():.3e}")
                if n <= 11:
                    _, _, _, _, psi_o = oracle.run_exact(state, n, d, lo, hi, 1.0, 2)
                    check(f"state vs oracle n={n}", np.abs(full - psi_o).max() < 1e-11)
            # H psi through the sharded test hook on a seeded complex vector
            rng = np.random.default_rng(n)
            v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
            hv = eng.apply_h(v)
            if ref is not None:
                check(f"apply_h n={n}", np.abs(hv - ref.apply_h(v)).max() < 1e-11)
        # general complex state: both planes, explicit upload of the full vector
        if n <= 22:
            rng = np.random.default_rng(7 * n)
            psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
            psi /= np.linalg.norm(psi)
            eng.set_state(psi)
            eng.step(0.5, 1)
            full = eng.get_state()
            if ref is not None:
                ref.set_state(psi)
                ref.step(0.5, 1)
                check(f"complex state n={n}", np.abs(full - ref.get_state()).max() < 1e-11)
        eng.close()
        if ref is not None:
            ref.close()
        dist.barrier()
    # rows written by the C oracle (tests/golden/make_golden_c.py: matrix-free Hermitian H, Chebyshev series on the
    # host) at 20..26 qubits: the sharded path against an independent computation, not against this library
    from conftest import golden_names, load_golden
    for name in golden_names("cexact"):
        spec, g = load_golden(name)
        n = spec["ncells"]
        if n - rbits < 13:
            continue
        rules = qca_b200.Rules(n, range(spec["lo"], spec["hi"]), spec["distance"])
        eng = sharding.ShardedExactEngine(rules, device=local)
        eng.set_product_state(qca_b200.states.plist(spec["state"], rules))
        rows = g["population"].shape[0]
        for k in range(rows):
            pop, _, ent, _ = eng.measure()
            dp, de = np.abs(pop - g["population"][k]).max(), np.abs(ent - g["single_site_entropy"][k]).max()
            check(f"C-oracle rows {name} row {k}", dp < 1e-10 and de < 1e-10, f"pop {dp:.3e} ent {de:.3e}")
            if k + 1 < rows:
                eng.step(float(g["effective_step_size"]), 1)
        if rank == 0:
            print(f"mgpu_worker world={world}: {name} {rows} rows vs C oracle checked "
                  f"({eng.stats()['passes_per_apply']} tile passes, {eng.stats()['local_bits']} local qubits)", flush=True)
        eng.close()
        dist.barrier()
    # the Exact plug-in picks the sharded engine up from torch.distributed
    rules = qca_b200.Rules(14 + rbits, range(1, 2), 1)
    algo = qca_b200.Exact(qca_b200.states.make("single", rules), None, qca_b200.Args(rules=rules, step_size=1.0), device=local)
    check("plug-in is sharded", isinstance(algo.engine, sharding.ShardedExactEngine))
    n = rules.ncells
    pop, dpop, ent, bond = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n + 1)
    algo.do_time_step()
    algo.measure(pop, dpop, ent, bond)
    check("plug-in mirror symmetry", np.abs(pop - pop[::-1]).max() < 1e-11 if n % 2 else True)
    check("plug-in norm", abs(algo.engine.norm2() - 1) < 1e-11)
    all_fail = sharding.gather_objects(failures)
    dist.destroy_process_group()
    if rank == 0:
        flat = [f for fs in all_fail for f in fs]
        print(f"mgpu_worker world={world}: {len(flat)} failures")
        for f in flat:
            print("  FAIL", f)
        return 1 if flat else 0
    return 0


if __name__ == "__main__":
    sys.exit(main())
