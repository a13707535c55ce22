"""Pins the CPU oracle (oracle/qca_oracle.py) to the fixtures generated from the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import qca_oracle as oracle
import qca_oracle_c as oracle_c
from conftest import golden_names, load_golden

DENSE_MAX = 11   # above this the dense U of the reference costs minutes: the matrix-free C oracle steps instead


@pytest.mark.parametrize("name", golden_names("hpsi"))
def test_hamiltonian_matches_reference(name):
    spec, g = load_golden(name)
    n, d, lo, hi = spec["ncells"], spec["distance"], spec["lo"], spec["hi"]
    rng = np.random.default_rng(spec["seed"])
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    w = oracle.mpo_tensors(n, d, lo, hi)
    # same tensors and bond numbering as MPO.hamiltonian_from_rules
    assert np.array_equal(w[0], g["w_first"])
    assert np.array_equal(w[-1], g["w_last"])
    assert np.array_equal(w[1] if n > 2 else w[0], g["w_bulk"])
    hm = oracle.mpo_as_matrix(w)
    assert np.abs(hm @ v - g["hv"]).max() < 1e-12
    # the closed form the CUDA kernel evaluates is the same matrix, exactly
    hd = oracle.rule_hamiltonian_direct(n, d, lo, hi)
    assert np.array_equal(hm.real, hd) and not hm.imag.any()
    assert abs(np.linalg.eigvalsh(hd).max() - float(g["eig_max"])) < 1e-10
    # the matrix-free restatements (numpy and C) are the same operator
    assert np.abs(oracle.apply_h(v, n, d, lo, hi) - g["hv"]).max() < 1e-12
    assert np.abs(oracle_c.apply_h(v, n, d, lo, hi) - g["hv"]).max() < 1e-12


@pytest.mark.parametrize("name", golden_names("exact"))
def test_exact_run_matches_reference(name):
    spec, g = load_golden(name)
    n, d, lo, hi = spec["ncells"], spec["distance"], spec["lo"], spec["hi"]
    steps = g["population"].shape[0]
    psi0 = oracle.product_state_vector(oracle.initial_plist(spec["state"], n, d))
    assert np.abs(psi0 - g["psi0"]).max() < 1e-15
    if n > DENSE_MAX:
        pytest.skip("dense route too slow here; covered by test_c_oracle_run_matches_reference")
    pop, dpop, ent, bond, psi = oracle.run_exact(spec["state"], n, d, lo, hi,
                                                 float(g["effective_step_size"]), steps)
    assert np.abs(pop - g["population"]).max() < 1e-12
    assert np.abs(ent - g["single_site_entropy"]).max() < 1e-11
    assert np.array_equal(bond, g["bond_dims"])
    assert np.abs(psi - g["psi_final"]).max() < 1e-12
    clear = np.abs(g["population"] - 0.5) > 1e-9  # np.round at an exact tie is rounding noise
    assert np.array_equal(dpop[clear], g["d_population"][clear])
    cls = oracle.classical_evolution(g["d_population"][0], d, lo, hi, steps)
    assert np.array_equal(cls, g["classical"])


def test_unitarity_and_step_composition():
    h = oracle.rule_hamiltonian_direct(8, 1, 1, 2)
    u = oracle.calculate_U(h, 0.5)
    assert np.abs(u @ u.conj().T - np.eye(256)).max() < 1e-12
    assert np.abs(u @ u - oracle.calculate_U(h, 1.0)).max() < 1e-12


@pytest.mark.parametrize("name", golden_names("exact"))
def test_c_oracle_run_matches_reference(name):
    """The C oracle (matrix-free H, Chebyshev series) replays every reference run, including the
    N = 13 / 14 ones whose dense U took the reference minutes to build."""
    spec, g = load_golden(name)
    n, d, lo, hi = spec["ncells"], spec["distance"], spec["lo"], spec["hi"]
    steps = g["population"].shape[0]
    s = oracle_c.Stepper(n, d, lo, hi)
    s.set_product_state(oracle.initial_plist(spec["state"], n, d))
    assert np.abs(s.psi - g["psi0"]).max() < 1e-15
    for k in range(steps):
        pop, ent = s.measure()
        assert np.abs(pop - g["population"][k]).max() < 1e-12
        assert np.abs(ent - g["single_site_entropy"][k]).max() < 1e-11
        if k == steps - 1:   # the numpy measurement (and the reference's own MPS route) on the same vector
            p2, _, e2, b2 = oracle.measure_vector(s.psi, n)
            assert np.abs(p2 - pop).max() < 1e-13 and np.abs(e2 - ent).max() < 1e-12
            assert np.array_equal(b2, g["bond_dims"][k])
        s.step(float(g["effective_step_size"]))
    assert np.abs(s.psi - g["psi_final"]).max() < 1e-12


@pytest.mark.parametrize("name", ["exact_single9", "exact_gradient9_half", "exact_eqsup8"])
def test_matrix_free_numpy_step_matches_reference(name):
    spec, g = load_golden(name)
    n, d, lo, hi = spec["ncells"], spec["distance"], spec["lo"], spec["hi"]
    psi = oracle.product_state_vector(oracle.initial_plist(spec["state"], n, d))
    for k in range(3):
        pop, _, ent, _ = oracle.measure_vector(psi, n)
        assert np.abs(pop - g["population"][k]).max() < 1e-12
        psi = oracle.exact_step_matrix_free(psi, n, d, lo, hi, float(g["effective_step_size"]))


@pytest.mark.parametrize("name", ["exact_eqsup8", "exact_gradient9_half", "exact_blinker10"])
def test_reference_measure_route_equals_direct_reduction(name):
    """MPS.from_vector + QR sweeps + logm (the route the reference's CPU time goes into, restated in
    measure_via_mps) gives the numbers of the direct reduced-density-matrix sums."""
    spec, g = load_golden(name)
    n = spec["ncells"]
    psi = np.asarray(g["psi_final"])
    a, b = oracle.measure_vector(psi, n), oracle.measure_via_mps(psi, n)
    assert np.abs(a[0] - b[0]).max() < 1e-12 and np.abs(a[2] - b[2]).max() < 1e-10
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
    k = g["population"].shape[0] - 1   # psi_final is the state AFTER the last measured row's step
    assert k >= 0


@pytest.mark.parametrize("name", golden_names("cexact"))
def test_c_golden_rows_are_normalised_and_start_from_the_named_state(name):
    """Fixtures written by the C oracle at N = 20..26 (tests/golden/make_golden_c.py): sanity only --
    the oracle itself is pinned above; the GPU suite compares against these rows."""
    spec, g = load_golden(name)
    n, d = spec["ncells"], spec["distance"]
    assert np.allclose(g["population"][0], oracle.initial_plist(spec["state"], n, d))
    assert abs(float(g["norm2"]) - 1.0) < 1e-12
    assert (g["population"] > -1e-14).all() and (g["population"] < 1 + 1e-14).all()
    assert (g["single_site_entropy"] > -1e-12).all() and (g["single_site_entropy"] < 1 + 1e-12).all()
