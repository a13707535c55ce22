"""Pins the CPU oracle (oracle/qca_oracle.py) to the fixtures generated from the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import qca_oracle as oracle
from conftest import golden_names, load_golden


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


@pytest.mark.parametrize("name", golden_names("exact"))
def test_exact_run_matches_reference(name):
    spec, g = load_golden(name)
    n, d, lo, hi = spec["ncells"], spec["distance"], spec["lo"], spec["hi"]
    steps = g["population"].shape[0]
    psi0 = oracle.product_state_vector(oracle.initial_plist(spec["state"], n, d))
    assert np.abs(psi0 - g["psi0"]).max() < 1e-15
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
