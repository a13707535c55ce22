#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py [/root/reference]

It imports the reference's own modules (``parameters``, ``tensor_networks``,
``states``, ``algorithms``) and records, for a set of small cases, the outputs of
the reference's public path: ``MPO.hamiltonian_from_rules``, ``MPO.as_matrix``,
``Exact`` / ``TDVP`` ``do_time_step`` and ``Algorithm.measure``.  ``quantum_game.py``
itself cannot be imported here (matplotlib/plotly are not installed), so the
measure-then-step loop of quantum_game.py:82-119 is replayed verbatim below
without the csv/plot side effects.

One python process per case: the reference keeps its arguments in a process-wide
singleton (parameters/parser.py:187-193) read at import time.
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

EXACT_CASES = [
    # name, ncells, distance, lo, hi, state, num_steps, step_size, plot_freq
    ("exact_single9", 9, 1, 1, 2, "single", 2000, 0.005, 1.0),           # BASELINE configs[0]
    ("exact_blinker10", 10, 1, 1, 2, "blinker", 1600, 0.005, 1.0),
    ("exact_triple11_d2", 11, 2, 2, 4, "triple_blinker", 1200, 0.005, 1.0),  # configs[3] rule, small N
    ("exact_eqsup8", 8, 1, 1, 3, "equal_superposition", 1600, 0.005, 1.0),
    ("exact_gradient9_half", 9, 1, 1, 2, "gradient", 1200, 0.005, 2.0),  # step 0.5
    ("exact_fullblinker10_d2", 10, 2, 1, 3, "full_blinker", 1000, 0.005, 1.0),
    ("exact_outer7_d3", 7, 3, 2, 5, "all_ket_1_but_outer", 1000, 0.005, 1.0),
    ("exact_single2", 2, 1, 1, 2, "single", 800, 0.005, 1.0),            # smallest chain
    # registers of >= 13 qubits: the sizes at which the B200 path runs its fast tile-pass kernel
    # (pass_kernel_v2), the Clenshaw stepper and the fused measurement -- dense 8192^2 / 16384^2 eigh
    # in the reference, minutes on the build box
    ("exact_triple13_d2", 13, 2, 2, 4, "triple_blinker", 1200, 0.005, 1.0),  # configs[3] rule, one tile pass
    ("exact_blinker14", 14, 1, 1, 2, "blinker", 800, 0.005, 1.0),            # configs[1] rule, two tile passes
]

TDVP_CASES = [
    # name, algorithm, ncells, distance, lo, hi, state, num_steps, step_size, chi, eps, plot_freq
    ("tdvp2_single8", "2tdvp", 8, 1, 1, 2, "single", 60, 0.005, 8, 5e-5, 10.0),
    ("tdvp2_blinker10_chi16", "2tdvp", 10, 1, 1, 2, "blinker", 60, 0.005, 16, 5e-5, 10.0),
    ("tdvp2_triple9_d2", "2tdvp", 9, 2, 2, 4, "triple_blinker", 40, 0.005, 8, 1e-6, 10.0),
    ("tdvp1_single8", "1tdvp", 8, 1, 1, 2, "single", 60, 0.005, 8, 5e-5, 10.0),
    ("tdvp2_eqsup7", "2tdvp", 7, 1, 1, 3, "equal_superposition", 40, 0.005, 8, 5e-5, 10.0),
    # well-conditioned inputs (no exactly-zero Schmidt values; the reference's own output is stable to
    # round-off there, unlike for 0/1 product states -- see DESIGN.md "TDVP parity")
    ("tdvp2_gradient8", "2tdvp", 8, 1, 1, 2, "gradient", 40, 0.005, 8, 5e-5, 10.0),
    ("tdvp1_eqsup7", "1tdvp", 7, 1, 1, 2, "equal_superposition", 40, 0.005, 8, 5e-5, 10.0),
    ("tdvp2_eqsup8_d2", "2tdvp", 8, 2, 2, 4, "equal_superposition", 30, 0.005, 6, 1e-6, 10.0),
    ("tdvp2_gradient9_chi4", "2tdvp", 9, 1, 1, 3, "gradient", 40, 0.01, 4, 1e-4, 5.0),
    # round 2: 12 measured rows each (plotting frequency 50 -> every 4th step); BASELINE configs[2]'s shape
    # (2tdvp, single, 15 cells) at a bond cap and step count the reference's dense H_eff finishes in seconds
    ("tdvp2_single15_chi8", "2tdvp", 15, 1, 1, 2, "single", 48, 0.005, 8, 5e-5, 50.0),
    ("tdvp2_gradient12_rows12", "2tdvp", 12, 1, 1, 2, "gradient", 48, 0.005, 8, 5e-5, 50.0),
    # (1tdvp pads every bond of this product state to the cap with Householder completions of zero columns: the
    #  reference's own numbers are reproducible to ~1e-7 only -- "padded" fixtures are compared at 1e-6)
    ("tdvp1_padded_gradient10_rows12", "1tdvp", 10, 1, 1, 2, "gradient", 48, 0.005, 8, 5e-5, 50.0),
]

HPSI_CASES = [
    # name, ncells, distance, lo, hi, seed
    ("hpsi_n10_d1_12", 10, 1, 1, 2, 1),
    ("hpsi_n10_d1_13", 10, 1, 1, 3, 2),
    ("hpsi_n11_d2_24", 11, 2, 2, 4, 3),
    ("hpsi_n10_d2_15", 10, 2, 1, 5, 4),
    ("hpsi_n9_d3_25", 9, 3, 2, 5, 5),
    ("hpsi_n5_d1_12", 5, 1, 1, 2, 6),
    ("hpsi_n8_d4_36", 8, 4, 3, 6, 7),
]


def child(spec: dict, ref: str) -> None:
    import numpy as np
    sys.path.insert(0, ref)
    argv = ["main.py", "--num-cells", str(spec["ncells"]), "--distance", str(spec["distance"]),
            "--activation-interval", str(spec["lo"]), str(spec["hi"])]
    if spec["kind"] != "hpsi":
        argv += ["--num-steps", str(spec["num_steps"]), "--step-size", str(spec["step_size"]),
                 "--plotting-frequency", str(spec["plot_freq"]), "--initial-states", spec["state"],
                 "--algorithm", spec.get("algorithm", "exact")]
    if spec["kind"] == "tdvp":
        argv += ["--max-bond-dim", str(spec["chi"]), "--svd-epsilon", str(spec["eps"])]
    sys.argv = argv
    from parameters import Parser
    args = Parser.instance()
    from tensor_networks import MPO, MPS
    H = MPO.hamiltonian_from_rules(args.rules)
    out = {"spec": json.dumps(spec)}
    n = args.rules.ncells

    if spec["kind"] == "hpsi":
        rng = np.random.default_rng(spec["seed"])
        v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        hm = H.as_matrix()
        out["hv"] = hm @ v
        out["w_bulk"] = H.W[1] if n > 2 else H.W[0]
        out["w_first"], out["w_last"] = H.W[0], H.W[-1]
        out["eig_max"] = np.linalg.eigvalsh(hm).max()
        np.savez_compressed(os.path.join(HERE, spec["name"] + ".npz"), **out)
        return

    import states
    from algorithms import Exact, TDVP
    def make_state():
        if spec["state"] == "single":  # default argument is frozen at import: states.py:34
            return states.single(position=int(n / 2))
        return getattr(states, spec["state"])()

    psi0 = make_state()
    out["psi0"] = psi0.as_vector()
    # quantum_game.py:70-73
    if args.algorithm == "exact":
        args.step_size = args.step_size * args.plot_step_interval
    out["effective_step_size"] = args.step_size
    out["plot_step_interval"] = args.plot_step_interval
    def run_reference():
        algo = (Exact if args.algorithm == "exact" else TDVP)(psi_0=make_state(), H=H, args=args)
        pop = np.zeros([args.plot_steps, n]); dpop = np.zeros_like(pop); sse = np.zeros_like(pop)
        bond = np.zeros([args.plot_steps, n + 1])
        # quantum_game.py:82-119 minus csv/npz/plot side effects
        for step in range(args.num_steps):
            if step % args.plot_step_interval == 0:
                k = step // args.plot_step_interval
                algo.measure(population=pop[k, :], d_population=dpop[k, :],
                             single_site_entropy=sse[k, :], bond_dims=bond[k, :])
                if args.algorithm == "exact":
                    algo.do_time_step()
            if args.algorithm != "exact":
                algo.do_time_step()
        return pop, dpop, sse, bond, (algo.psi.as_vector() if args.algorithm != "exact" else algo._psi)

    pop, dpop, sse, bond, psi_final = run_reference()
    out.update(population=pop, d_population=dpop, single_site_entropy=sse, bond_dims=bond)
    out["psi_final"] = psi_final
    if args.algorithm == "2tdvp":
        # The reference's 2TDVP depends on the arbitrary phases of the singular vectors np.linalg.svd
        # returns (stale environments after re-canonicalisation).  Record how far the UNMODIFIED
        # reference moves when the SVD is replaced by an equivalent one with random phases.
        stock = np.linalg.svd
        spread_pop, spread_sse = 0.0, 0.0
        for seed in (1, 2, 3, 4):
            rng = np.random.default_rng(seed)

            def svd(a, full_matrices=True, **kw):
                u, s, vh = stock(a, full_matrices=full_matrices, **kw)
                ph = np.exp(2j * np.pi * rng.random(len(s)))
                u, vh = u.copy(), vh.copy()
                u[:, :len(s)] *= ph
                vh[:len(s), :] *= ph.conj()[:, None]
                return u, s, vh
            np.linalg.svd = svd
            try:
                p2, _, s2, _, _ = run_reference()
            finally:
                np.linalg.svd = stock
            spread_pop = max(spread_pop, np.abs(p2 - pop).max())
            spread_sse = max(spread_sse, np.abs(s2 - sse).max())
        out["gauge_spread_population"] = spread_pop
        out["gauge_spread_entropy"] = spread_sse
    if spec["kind"] == "exact":
        from algorithms import Algorithm
        out["classical"] = Algorithm.classical_evolution(dpop[0, :], args.rules, args.plot_steps)
    np.savez_compressed(os.path.join(HERE, spec["name"] + ".npz"), **out)


def main() -> None:
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    specs = []
    for (name, n, d, lo, hi, state, steps, dt, pf) in EXACT_CASES:
        specs.append(dict(kind="exact", name=name, ncells=n, distance=d, lo=lo, hi=hi, state=state,
                          num_steps=steps, step_size=dt, plot_freq=pf))
    for (name, alg, n, d, lo, hi, state, steps, dt, chi, eps, pf) in TDVP_CASES:
        specs.append(dict(kind="tdvp", name=name, algorithm=alg, ncells=n, distance=d, lo=lo, hi=hi,
                          state=state, num_steps=steps, step_size=dt, chi=chi, eps=eps, plot_freq=pf))
    for (name, n, d, lo, hi, seed) in HPSI_CASES:
        specs.append(dict(kind="hpsi", name=name, ncells=n, distance=d, lo=lo, hi=hi, seed=seed))
    only = os.environ.get("GOLDEN_ONLY")
    for spec in specs:
        if only and only not in spec["name"]:
            continue
        print("golden:", spec["name"], flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", json.dumps(spec), ref],
                       check=True, cwd="/tmp")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(json.loads(sys.argv[2]), sys.argv[3])
    else:
        main()
