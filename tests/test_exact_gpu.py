"""GPU parity tests of the exact path, through the C ABI (ctypes) and the Exact plug-in class,
against (a) the fixtures produced by the unmodified reference, (b) the CPU oracle on seeded
inputs, (c) size-independent properties at sizes the oracle cannot reach.

Tolerance: 1e-10 absolute on populations and entropies (BASELINE.json north_star); state
vectors are compared at 1e-11.
"""
import numpy as np
import pytest

import qca_b200
import qca_oracle as oracle
import qca_oracle_c as oracle_c
from conftest import golden_names, load_golden
from qca_b200 import _lib

pytestmark = pytest.mark.gpu
TOL = 1e-10


def make_rules(spec):
    return qca_b200.Rules(spec["ncells"], range(spec["lo"], spec["hi"]), spec["distance"])


@pytest.mark.parametrize("name", golden_names("hpsi"))
def test_apply_h_matches_reference_matrix(name):
    spec, g = load_golden(name)
    n = spec["ncells"]
    rng = np.random.default_rng(spec["seed"])
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    eng = _lib.ExactEngine(make_rules(spec))
    assert np.abs(eng.apply_h(v) - g["hv"]).max() < 1e-12
    eng.close()


@pytest.mark.parametrize("name", golden_names("exact"))
def test_exact_plugin_matches_reference_run(name):
    """quantum_game.py:82-119 replayed with the B200 Exact in place of the reference's."""
    spec, g = load_golden(name)
    rules = make_rules(spec)
    args = qca_b200.Args(rules=rules, step_size=float(g["effective_step_size"]))
    algo = qca_b200.Exact(qca_b200.states.make(spec["state"], rules), qca_b200.MPO.hamiltonian_from_rules(rules), args)
    steps, n = g["population"].shape
    pop, dpop, sse = np.zeros((steps, n)), np.zeros((steps, n)), np.zeros((steps, n))
    bond = np.zeros((steps, n + 1))
    for k in range(steps):
        algo.measure(pop[k, :], dpop[k, :], sse[k, :], bond[k, :])
        algo.do_time_step()
    assert np.abs(pop - g["population"]).max() < TOL
    assert np.abs(sse - g["single_site_entropy"]).max() < TOL
    assert np.array_equal(bond, g["bond_dims"])
    clear = np.abs(g["population"] - 0.5) > 1e-9
    assert np.array_equal(dpop[clear], g["d_population"][clear])
    assert np.abs(algo.state_vector() - g["psi_final"]).max() < 1e-11
    # the psi property hands back an MPS of the same state (exact.py:19-20)
    mps = algo.psi
    assert mps.bond_dims == [int(b) for b in g["bond_dims"][0]]
    assert np.abs(mps.as_vector() - g["psi_final"]).max() < 1e-11


@pytest.mark.parametrize("n,d,lo,hi", [(1, 1, 1, 2), (2, 1, 1, 2), (3, 2, 1, 3), (12, 1, 1, 2), (13, 2, 2, 4),
                                       (14, 1, 1, 3), (16, 3, 2, 5), (18, 2, 2, 4), (20, 1, 1, 2)])
def test_apply_h_seeded_vs_oracle(n, d, lo, hi):
    """Random complex vectors, all pass-count regimes (1, 2 and 3 tile passes)."""
    rng = np.random.default_rng(100 + n)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    eng = _lib.ExactEngine(qca_b200.Rules(n, range(lo, hi), d))
    got = eng.apply_h(v)
    if n <= 11:
        want = oracle.rule_hamiltonian_direct(n, d, lo, hi) @ v
    else:  # the oracle's matrix-free restatements (pinned to the reference's H @ v fixtures in the CPU suite)
        want = oracle_c.apply_h(v, n, d, lo, hi)
        if n <= 16:
            assert np.abs(oracle.apply_h(v, n, d, lo, hi) - want).max() < 1e-12
    assert np.abs(got - want).max() < 1e-11
    eng.close()


@pytest.mark.parametrize("n,d,lo,hi,state,tau", [
    (13, 1, 1, 2, "single", 1.0), (13, 2, 2, 4, "equal_superposition", 0.37), (14, 2, 2, 4, "triple_blinker", 1.0),
    (15, 1, 1, 3, "gradient", -0.6), (17, 3, 2, 5, "blinker", 1.0), (18, 2, 2, 4, "triple_blinker", 1.0),
    (20, 1, 1, 2, "blinker", 1.0), (22, 2, 1, 3, "full_blinker", 0.5)])
def test_fast_kernel_step_and_measure_vs_c_oracle(n, d, lo, hi, state, tau):
    """pass_kernel_v2 (>= 13 qubits: one, two and three tile passes) + Clenshaw stepper + fused measurement,
    stepped against the C oracle (matrix-free Hermitian H, forward Chebyshev series in complex arithmetic;
    pinned to the reference's N <= 14 runs in tests/test_oracle_golden.py): populations and entropies before
    every step at 1e-10, the final vector at 1e-11."""
    rules = qca_b200.Rules(n, range(lo, hi), d)
    plist = qca_b200.states.plist(state, rules)
    eng = _lib.ExactEngine(rules)
    eng.set_product_state(plist)
    ref = oracle_c.Stepper(n, d, lo, hi)
    ref.set_product_state(plist)
    for k in range(3):
        pop, _, ent, _ = eng.measure()
        pop_o, ent_o = ref.measure()
        assert np.abs(pop - pop_o).max() < TOL and np.abs(ent - ent_o).max() < TOL, (k, n)
        eng.step(tau, 1)
        ref.step(tau)
    assert eng.stats()["passes_per_apply"] == len(_lib.plan_passes_v3(n, 0) or _lib.plan_passes(n))
    assert np.abs(eng.get_state() - ref.psi).max() < 1e-11
    eng.close()


@pytest.mark.parametrize("name", golden_names("cexact"))
def test_exact_plugin_matches_c_oracle_rows(name):
    """Registers of 20..26 qubits (BASELINE configs[1] and the three-pass geometry of the N = 30 bench)
    through the Exact plug-in against rows the C oracle wrote (tests/golden/make_golden_c.py)."""
    spec, g = load_golden(name)
    rules = make_rules(spec)
    args = qca_b200.Args(rules=rules, step_size=float(g["effective_step_size"]))
    algo = qca_b200.Exact(qca_b200.states.make(spec["state"], rules), qca_b200.MPO.hamiltonian_from_rules(rules), args)
    rows, n = g["population"].shape
    for k in range(rows):
        pop, dpop, sse, bond = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n + 1)
        algo.measure(pop, dpop, sse, bond)
        assert np.abs(pop - g["population"][k]).max() < TOL, k
        assert np.abs(sse - g["single_site_entropy"][k]).max() < TOL, k
        if k + 1 < rows:
            algo.do_time_step()


@pytest.mark.parametrize("state", ["single", "blinker", "equal_superposition", "gradient", "all_ket_1"])
@pytest.mark.parametrize("tau", [1.0, 0.37])
def test_step_and_measure_vs_oracle(state, tau):
    n, d, lo, hi = 10, 1, 1, 2
    rules = qca_b200.Rules(n, range(lo, hi), d)
    algo = qca_b200.Exact(qca_b200.states.make(state, rules), None, qca_b200.Args(rules=rules, step_size=tau))
    pop_o, dpop_o, ent_o, bond_o, psi_o = oracle.run_exact(state, n, d, lo, hi, tau, 4)
    for k in range(4):
        pop, dpop, ent, bond = (np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n + 1))
        algo.measure(pop, dpop, ent, bond)
        assert np.abs(pop - pop_o[k]).max() < TOL and np.abs(ent - ent_o[k]).max() < TOL
        algo.do_time_step()
    assert np.abs(algo.state_vector() - psi_o).max() < 1e-11


def test_general_complex_state_and_forced_complex_agree():
    n, d, lo, hi = 11, 2, 2, 4
    rules = qca_b200.Rules(n, range(lo, hi), d)
    rng = np.random.default_rng(5)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    u = oracle.calculate_U(oracle.rule_hamiltonian_direct(n, d, lo, hi), 1.0)
    eng = _lib.ExactEngine(rules)
    eng.set_state(psi)
    assert eng.stats()["planes"] == 2
    eng.step(1.0, 2)
    assert np.abs(eng.get_state() - u @ (u @ psi)).max() < 1e-11
    pop, _, ent, _ = eng.measure()
    pop_o, _, ent_o, _ = oracle.measure_vector(u @ (u @ psi), n)
    assert np.abs(pop - pop_o).max() < TOL and np.abs(ent - ent_o).max() < TOL
    # a basis state runs on one real plane; forcing two planes gives the same answer
    basis = np.zeros(1 << n, dtype=complex)
    basis[0b00101010100] = 1.0
    one = _lib.ExactEngine(rules)
    two = _lib.ExactEngine(rules, flags=_lib.QCA_FLAG_FORCE_COMPLEX)
    one.set_state(basis), two.set_state(basis)
    assert one.stats()["planes"] == 1 and two.stats()["planes"] == 2
    one.step(1.0, 3), two.step(1.0, 3)
    want = u @ (u @ (u @ basis))
    assert np.abs(one.get_state() - want).max() < 1e-11
    assert np.abs(two.get_state() - want).max() < 1e-11
    # purely imaginary rotated state (odd popcount basis state times i^k) also takes one plane
    odd = np.zeros(1 << n, dtype=complex)
    odd[0b00000000111] = -1j
    one.set_state(odd)
    assert one.stats()["planes"] == 1
    assert np.abs(one.get_state() - odd).max() == 0.0
    one.step(0.5, 1)
    assert np.abs(one.get_state() - oracle.calculate_U(oracle.rule_hamiltonian_direct(n, d, lo, hi), 0.5) @ odd).max() < 1e-11


def test_upload_download_roundtrip_is_exact():
    rng = np.random.default_rng(9)
    for n in (1, 4, 13, 17):
        psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        eng = _lib.ExactEngine(qca_b200.Rules(n, range(1, 2), 1))
        eng.set_state(psi)
        assert np.array_equal(eng.get_state(), psi)
        assert abs(eng.norm2() - np.vdot(psi, psi).real) < 1e-9 * (1 << n)
        eng.close()


@pytest.mark.parametrize("n,d,lo,hi", [(22, 1, 1, 2), (24, 2, 2, 4)])
def test_properties_beyond_oracle_reach(n, d, lo, hi):
    """Sizes where no dense reference exists: unitarity, time reversal, step composition,
    mirror symmetry of a symmetric initial state, and H-moment conservation."""
    rules = qca_b200.Rules(n, range(lo, hi), d)
    eng = _lib.ExactEngine(rules)
    eng.set_product_state(qca_b200.states.plist("triple_blinker", rules))
    assert eng.stats()["planes"] == 1 and eng.stats()["passes_per_apply"] >= 2
    pop0, _, ent0, _ = eng.measure()
    assert np.abs(ent0).max() < 1e-12 and np.array_equal(pop0, np.array(qca_b200.states.plist("triple_blinker", rules)))
    eng.step(1.0, 2)
    assert abs(eng.norm2() - 1.0) < 1e-12
    pop2, _, ent2, _ = eng.measure()
    if n % 2 == 1:
        assert np.abs(pop2 - pop2[::-1]).max() < 1e-12
    # one step of 2.0 == two steps of 1.0
    eng2 = _lib.ExactEngine(rules)
    eng2.set_product_state(qca_b200.states.plist("triple_blinker", rules))
    eng2.step(2.0, 1)
    popb, _, entb, _ = eng2.measure()
    assert np.abs(pop2 - popb).max() < TOL and np.abs(ent2 - entb).max() < TOL
    # time reversal returns to the basis state
    eng.step(-1.0, 2)
    popr, _, entr, _ = eng.measure()
    assert np.abs(popr - pop0).max() < TOL and np.abs(entr).max() < 1e-9
    eng.close(), eng2.close()


def test_symmetric_state_stays_symmetric_odd_chain():
    rules = qca_b200.Rules(21, range(1, 2), 1)
    eng = _lib.ExactEngine(rules)
    eng.set_product_state(qca_b200.states.plist("single", rules))
    eng.step(1.0, 3)
    pop, _, ent, _ = eng.measure()
    assert np.abs(pop - pop[::-1]).max() < 1e-12 and np.abs(ent - ent[::-1]).max() < 1e-11


def test_error_behaviour():
    rules = qca_b200.Rules(6, range(1, 2), 1)
    eng = _lib.ExactEngine(rules)
    with pytest.raises(qca_b200.QcaError) as e:
        eng.step(1.0)
    assert e.value.code == _lib.QCA_ERR_STATE
    with pytest.raises(qca_b200.QcaError) as e:
        eng.set_state(np.zeros(32, dtype=complex))
    assert e.value.code == _lib.QCA_ERR_ARG
    with pytest.raises(qca_b200.QcaError):
        eng.set_product_state([0.5] * 5)
    with pytest.raises(qca_b200.QcaError):
        eng.set_product_state([1.5] + [0.0] * 5)
    other = qca_b200.MPO.hamiltonian_from_rules(qca_b200.Rules(6, range(1, 3), 1))
    with pytest.raises(ValueError):
        qca_b200.Exact(qca_b200.states.make("single", rules), other, qca_b200.Args(rules=rules))


def test_zero_step_is_identity_and_stats_count_launches():
    rules = qca_b200.Rules(15, range(1, 2), 1)
    eng = _lib.ExactEngine(rules, flags=_lib.QCA_FLAG_PROFILE)
    eng.set_product_state(qca_b200.states.plist("blinker", rules))
    before = eng.get_state()
    eng.step(0.0, 3)
    assert np.array_equal(eng.get_state(), before)
    eng.reset_stats()
    eng.step(1.0, 1)
    st = eng.stats()
    assert st["passes_per_apply"] == len(_lib.plan_passes_v3(15, 0)) and st["last_terms"] > 10
    assert st["pass_launches"] == (st["last_terms"] - 1) * st["passes_per_apply"]
    assert st["profiled_pass_launches"] == st["pass_launches"] and st["profiled_pass_ms"] > 0.0


@pytest.mark.parametrize("n,d,lo,hi", [(9, 1, 1, 2), (11, 2, 2, 4), (10, 1, 1, 3), (8, 3, 2, 5), (3, 1, 1, 2)])
def test_tight_spectral_bound_is_a_bound(n, d, lo, hi):
    """The block-Lanczos bound the engine scales H by is above the true spectral radius (dense
    oracle), below Gershgorin, and within a few per cent of the truth when one block covers the chain."""
    rules = qca_b200.Rules(n, range(lo, hi), d)
    true_radius = np.linalg.eigvalsh(oracle.rule_hamiltonian_direct(n, d, lo, hi)).max()
    eng = _lib.ExactEngine(rules)
    loose = _lib.ExactEngine(rules, flags=_lib.QCA_FLAG_LOOSE_BOUND)
    tight_r, loose_r = eng.stats()["spectral_bound"], loose.stats()["spectral_bound"]
    assert loose_r == _lib.spectral_bound(rules)
    assert true_radius <= tight_r <= loose_r + 1e-12
    assert tight_r <= true_radius * 1.002 + 1e-9
    # both scalings evolve to the same state
    plist = qca_b200.states.plist("single", rules)
    eng.set_product_state(plist), loose.set_product_state(plist)
    eng.step(1.0, 2), loose.step(1.0, 2)
    assert np.abs(eng.get_state() - loose.get_state()).max() < 1e-12
    assert eng.stats()["last_terms"] <= loose.stats()["last_terms"]


def test_tight_bound_large_chain_blocks():
    rules = qca_b200.Rules(26, range(2, 4), 2)
    eng = _lib.ExactEngine(rules)
    r = eng.stats()["spectral_bound"]
    assert 0.70 * 26 < r < 0.80 * 26  # ~0.76 per cell for this rule (two end blocks of 13)
    eng.set_product_state(qca_b200.states.plist("triple_blinker", rules))
    eng.step(1.0, 1)
    assert abs(eng.norm2() - 1.0) < 1e-12


@pytest.mark.parametrize("n,d,lo,hi,state", [(13, 1, 1, 2, "blinker"), (14, 2, 2, 4, "triple_blinker"), (15, 1, 1, 2, "single"),
                                             (16, 2, 1, 3, "blinker"), (17, 2, 2, 4, "triple_blinker"), (18, 1, 1, 2, "blinker"),
                                             (21, 2, 2, 4, "triple_blinker"), (25, 1, 1, 2, "triple_blinker")])
def test_fused_measurement_matches_per_cell_kernels_and_oracle(n, d, lo, hi, state):
    """csrc/qca_measure.cu (one read of the state per tile pass; the default on one GPU for single-plane
    states) against the per-cell kernels on the same state, for every tile geometry (13 | 0..12 strided
    bits), and against the oracle's MPS.measure restatement on the downloaded vector."""
    rules = qca_b200.Rules(n, range(lo, hi), d)
    plist = qca_b200.states.plist(state, rules)
    outs = []
    for flag in (_lib.QCA_FLAG_PERCELL_MEASURE, _lib.QCA_FLAG_FUSED_MEASURE):
        eng = _lib.ExactEngine(rules, flags=flag | _lib.QCA_FLAG_LOOSE_BOUND)
        eng.set_product_state(plist)
        eng.step(1.0, 2)
        assert eng.stats()["planes"] == 1
        before = eng.stats()["kernel_launches"]
        outs.append(eng.measure())
        launches = eng.stats()["kernel_launches"] - before
        # (the fused measurement keeps its own 13-bit tile plan, whatever kernels apply the operator)
        assert launches == (2 * n if flag == _lib.QCA_FLAG_PERCELL_MEASURE else 2 * len(_lib.plan_passes(n)))
        if flag == _lib.QCA_FLAG_FUSED_MEASURE and n <= 18:
            pop_o, dpop_o, ent_o, bond_o = oracle.measure_vector(eng.get_state(), n)
            assert np.abs(outs[-1][0] - pop_o).max() < 1e-12 and np.abs(outs[-1][2] - ent_o).max() < 1e-10
        eng.close()
    for a, b in zip(*outs):
        assert np.abs(a - b).max() < 1e-12


@pytest.mark.parametrize("n,d,lo,hi,state", [(1, 1, 1, 2, "single"), (2, 1, 1, 2, "single"), (5, 2, 1, 3, "blinker"),
                                             (9, 1, 1, 2, "single"), (12, 2, 2, 4, "triple_blinker"), (13, 1, 1, 3, "gradient")])
def test_one_kernel_step_equals_tile_pass_path(n, d, lo, hi, state):
    """Registers of <= 13 qubits take a whole step (and a whole measurement) in one kernel (csrc/qca_small.cu);
    QCA_FLAG_TILE_PATH_ONLY keeps them on the tile-pass kernels.  Same recurrence: the states agree to
    round-off, one plane and two, both time directions, and the launch counts show which path ran."""
    rules = qca_b200.Rules(n, range(lo, hi), d)
    plist = qca_b200.states.plist(state, rules)
    rng = np.random.default_rng(n)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    for upload in (None, psi):
        small, tiles = _lib.ExactEngine(rules), _lib.ExactEngine(rules, flags=_lib.QCA_FLAG_TILE_PATH_ONLY)
        for eng in (small, tiles):
            eng.set_product_state(plist) if upload is None else eng.set_state(upload)
            eng.reset_stats()
        for tau in (1.0, -0.4, 2.5):
            small.step(tau, 2), tiles.step(tau, 2)
            assert np.abs(small.get_state() - tiles.get_state()).max() < 2e-13
            for a, b in zip(small.measure(), tiles.measure()):
                assert np.abs(a - b).max() < 1e-12
        assert small.stats()["planes"] == tiles.stats()["planes"] == (1 if upload is None and state != "gradient" else small.stats()["planes"])
        # 6 steps + 3 measurements: one launch each (plus the pack kernels of get_state)
        assert small.stats()["pass_launches"] == 0
        assert tiles.stats()["pass_launches"] > 0 or small.stats()["spectral_bound"] == 0.0   # (one cell: H == 0, no launches at all)
        assert abs(small.norm2() - 1.0) < 1e-12
        small.close(), tiles.close()


@pytest.mark.parametrize("name", golden_names("exact"))
def test_tile_pass_path_matches_reference_run_on_small_registers(name):
    """The reference's runs again, with the one-kernel step switched off: the generic / fast tile-pass kernels
    at N <= 14 (the default path of these sizes is covered by test_exact_plugin_matches_reference_run)."""
    spec, g = load_golden(name)
    rules = make_rules(spec)
    eng = _lib.ExactEngine(rules, flags=_lib.QCA_FLAG_TILE_PATH_ONLY | _lib.QCA_FLAG_NO_GRAPH)
    eng.set_product_state(qca_b200.states.plist(spec["state"], rules))
    for k in range(g["population"].shape[0]):
        pop, _, ent, _ = eng.measure()
        assert np.abs(pop - g["population"][k]).max() < TOL and np.abs(ent - g["single_site_entropy"][k]).max() < TOL
        eng.step(float(g["effective_step_size"]), 1)
    assert np.abs(eng.get_state() - g["psi_final"]).max() < 1e-11
    eng.close()


@pytest.mark.parametrize("n,d,lo,hi,state", [(14, 1, 1, 2, "blinker"), (16, 2, 2, 4, "triple_blinker"), (20, 1, 1, 2, "blinker")])
def test_graph_replay_equals_eager_launches(n, d, lo, hi, state):
    """Registers of 14..24 qubits replay a captured CUDA graph of the step (one per resident-vector index and step
    size).  Same kernels, same arguments: bit-identical states, identical launch accounting, across enough steps
    to cycle through all three vector rotations, a change of step size, and a new upload in between."""
    rules = qca_b200.Rules(n, range(lo, hi), d)
    plist = qca_b200.states.plist(state, rules)
    graph, eager = _lib.ExactEngine(rules), _lib.ExactEngine(rules, flags=_lib.QCA_FLAG_NO_GRAPH)
    for eng in (graph, eager):
        eng.set_product_state(plist)
        eng.reset_stats()
    for tau, count in ((1.0, 7), (0.5, 4), (1.0, 3), (-1.0, 2)):
        for _ in range(count):
            graph.step(tau, 1), eager.step(tau, 1)
        assert np.array_equal(graph.get_state(), eager.get_state())
        for a, b in zip(graph.measure(), eager.measure()):
            assert np.array_equal(a, b)
    sg, se = graph.stats(), eager.stats()
    for key in ("kernel_launches", "pass_launches", "pass_bytes", "last_terms"):
        assert sg[key] == se[key], key
    # a general complex state (two planes) invalidates nothing silently: new graphs, same answers
    rng = np.random.default_rng(3)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    graph.set_state(psi), eager.set_state(psi)
    for _ in range(4):
        graph.step(1.0, 1), eager.step(1.0, 1)
    assert np.array_equal(graph.get_state(), eager.get_state())
    graph.set_product_state(plist), eager.set_product_state(plist)   # back to one plane: plane 1 is released
    for _ in range(4):
        graph.step(1.0, 1), eager.step(1.0, 1)
    assert np.array_equal(graph.get_state(), eager.get_state())
    graph.close(), eager.close()


@pytest.mark.parametrize("n,d,lo,hi,env", [
    (14, 1, 1, 2, {}), (15, 2, 2, 4, {"QCA_V3_CLUSTER_BITS": "3"}), (16, 2, 1, 3, {"QCA_V3_CLUSTER_BITS": "3"}),
    (17, 2, 2, 4, {"QCA_V3_CLUSTER_BITS": "3"}), (18, 1, 1, 3, {"QCA_V3_CLUSTER_BITS": "3"}), (20, 3, 2, 5, {"QCA_V3_CLUSTER_BITS": "3"}),
    (21, 2, 2, 4, {}), (21, 2, 2, 4, {"QCA_V3_CLUSTER_BITS": "1"}), (19, 3, 2, 5, {}),
    (22, 2, 2, 4, {"QCA_V3_CLUSTER_BITS": "2"}), (22, 1, 1, 2, {"QCA_V3_MIN_LOW": "10"}), (23, 4, 3, 6, {}),
    (24, 2, 2, 4, {"QCA_V3_MIN_LOW": "7", "QCA_V3_CLUSTER_BITS": "3"}), (25, 2, 2, 4, {"QCA_V3_CLUSTER_BITS": "3"}), (26, 2, 2, 4, {})])
def test_cluster_kernels_equal_13_bit_kernels(n, d, lo, hi, env, monkeypatch):
    """pass_kernel_v3 (14-bit CTA tiles joined through distributed shared memory, the default on one GPU from 14
    qubits) against pass_kernel_v2 (QCA_FLAG_V2_KERNELS) on the same inputs: H psi on a seeded complex vector, and two
    steps from a product state (both planes and one).  All cluster sizes and strided-tile geometries via the planning
    knobs; the v2 path is itself pinned to the oracle above."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rules = qca_b200.Rules(n, range(lo, hi), d)
    v3 = _lib.ExactEngine(rules, flags=_lib.QCA_FLAG_LOOSE_BOUND)
    v2 = _lib.ExactEngine(rules, flags=_lib.QCA_FLAG_LOOSE_BOUND | _lib.QCA_FLAG_V2_KERNELS)
    cb = int(env.get("QCA_V3_CLUSTER_BITS", 0)) if d <= 3 else 0
    assert v3.stats()["passes_per_apply"] == len(_lib.plan_passes_v3(n, cb, int(env.get("QCA_V3_MIN_LOW", 4))))
    assert v2.stats()["passes_per_apply"] == len(_lib.plan_passes(n))
    rng = np.random.default_rng(n)
    vec = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    a, b = v3.apply_h(vec), v2.apply_h(vec)
    assert np.abs(a - b).max() < 1e-12
    if n <= 22:
        assert np.abs(a - oracle_c.apply_h(vec, n, d, lo, hi)).max() < 1e-11
    plist = qca_b200.states.plist("gradient" if n % 2 else "triple_blinker", rules)
    v3.set_product_state(plist), v2.set_product_state(plist)
    v3.step(1.0, 2), v2.step(1.0, 2)
    assert abs(v3.norm2() - 1.0) < 1e-12
    for x, y in zip(v3.measure(), v2.measure()):
        assert np.abs(x - y).max() < 1e-12
    if n <= 22:
        assert np.abs(v3.get_state() - v2.get_state()).max() < 1e-12
    v3.close(), v2.close()


@pytest.mark.parametrize("n,world,rank,d,lo,hi,env", [
    (17, 2, 1, 2, 2, 4, {"QCA_PERSISTENT_CTAS": "3"}), (18, 2, 0, 1, 1, 2, {"QCA_PERSISTENT_CTAS": "5", "QCA_FORCE_SLOTS2": "1"}),
    (19, 4, 2, 2, 2, 4, {"QCA_PERSISTENT_CTAS": "7"}), (20, 4, 3, 2, 1, 3, {"QCA_PERSISTENT_CTAS": "2", "QCA_REMOTE_RING": "4"}),
    (21, 8, 5, 2, 2, 4, {"QCA_PERSISTENT_CTAS": "3"}), (22, 8, 0, 2, 2, 4, {"QCA_PERSISTENT_CTAS": "11"}),
    (20, 8, 7, 1, 1, 2, {"QCA_PERSISTENT_CTAS": "1"}), (24, 8, 6, 2, 2, 4, {}), (23, 2, 1, 2, 2, 4, {})])
def test_persistent_sharded_kernel_equals_one_cta_per_tile(n, world, rank, d, lo, hi, env, monkeypatch):
    """pass_kernel_v2p (persistent CTAs, operand rings running across tile boundaries, carried mbarrier phases) against
    pass_kernel_v2 (one CTA per tile) on ONE GPU: a sharded engine whose partners are looped back to its own planes
    (qca_exact_loopback_peers).  The physics of such an engine is meaningless, the arithmetic is not: both kernels must
    produce bit-identical vectors, for every rank/world geometry, with few CTAs walking many tiles each."""
    rules = qca_b200.Rules(n, range(lo, hi), d)
    plist = list(np.random.default_rng(100 + n).uniform(0.1, 0.9, n))   # every rank's slice is populated, one real plane
    results = []
    for persistent in (True, False):
        for k, v in env.items():
            if persistent or k != "QCA_PERSISTENT_CTAS":
                monkeypatch.setenv(k, v)
        flags = _lib.QCA_FLAG_LOOSE_BOUND | (0 if persistent else _lib.QCA_FLAG_NO_PERSISTENT)
        eng = _lib.ExactEngine(rules, world_size=world, rank=rank, flags=flags)
        eng.loopback_peers()
        eng.set_product_state(plist)
        eng.resolve_planes(*eng.plane_flags())
        eng.step(1.0, 2)
        rng = np.random.default_rng(n)
        results.append((eng.get_state(), eng.measure_partial(), eng.stats()["pass_launches"]))
        # two planes as well
        psi = rng.standard_normal(eng.local_amps) + 1j * rng.standard_normal(eng.local_amps)
        eng.set_state(psi)
        eng.resolve_planes(*eng.plane_flags())
        eng.step(0.5, 1)
        results[-1] += (eng.get_state(),)
        eng.close()
    (a_state, a_sums, a_launches, a_c), (b_state, b_sums, b_launches, b_c) = results
    assert a_launches == b_launches > 0
    assert np.array_equal(a_state, b_state) and np.array_equal(a_sums, b_sums) and np.array_equal(a_c, b_c)
    assert np.isfinite(a_state).all() and np.abs(a_state).max() > 0
