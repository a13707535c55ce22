"""CPU tests of the host-side mirror of the reference interface (data formats either side of
the evolution path): Rules/Args, MPO tensors, MPS constructors and files, named states."""
import numpy as np
import pytest

import qca_b200
import qca_oracle as oracle
from conftest import golden_names, load_golden


@pytest.mark.parametrize("name", golden_names("hpsi"))
def test_mpo_tensors_equal_reference(name):
    spec, g = load_golden(name)
    rules = qca_b200.Rules(spec["ncells"], range(spec["lo"], spec["hi"]), spec["distance"])
    mpo = qca_b200.MPO.hamiltonian_from_rules(rules)
    assert np.array_equal(mpo.W[0], g["w_first"])
    assert np.array_equal(mpo.W[-1], g["w_last"])
    assert np.array_equal(mpo.W[1] if rules.ncells > 2 else mpo.W[0], g["w_bulk"])
    rng = np.random.default_rng(spec["seed"])
    v = rng.standard_normal(1 << rules.ncells) + 1j * rng.standard_normal(1 << rules.ncells)
    assert np.abs(mpo.as_matrix() @ v - g["hv"]).max() < 1e-12


@pytest.mark.parametrize("name", golden_names("exact"))
def test_named_states_equal_reference(name):
    spec, g = load_golden(name)
    rules = qca_b200.Rules(spec["ncells"], range(spec["lo"], spec["hi"]), spec["distance"])
    mps = qca_b200.states.make(spec["state"], rules)
    assert mps.is_product_state() and mps.is_valid_mps()
    assert np.abs(mps.as_vector() - g["psi0"]).max() < 1e-15
    assert qca_b200.states.plist(spec["state"], rules) == oracle.initial_plist(spec["state"], spec["ncells"], spec["distance"])


def test_mps_vector_roundtrip_and_file(tmp_path):
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    mps = qca_b200.MPS.from_vector(psi)
    assert mps.bond_dims == [1, 2, 4, 8, 4, 2, 1]
    assert np.abs(mps.as_vector() - psi).max() < 1e-13
    path = tmp_path / "state.npz"
    mps.write_to_file(str(path))
    back = qca_b200.MPS.from_file(str(path))
    assert all(np.array_equal(a, b) for a, b in zip(mps.A, back.A))


def test_args_match_reference_parser_defaults():
    a = qca_b200.Args.from_argv(["--num-cells", "11", "--distance", "2", "--activation-interval", "2", "4",
                                 "--algorithm", "2tdvp", "--num-steps", "1000", "--plotting-frequency", "10"])
    assert (a.rules.ncells, a.rules.distance, a.rules.activation_interval) == (11, 2, range(2, 4))
    assert a.algorithm == "2tdvp" and a.step_size == 0.005 and a.max_bond_dim == 32 and a.svd_epsilon == 5e-5
    assert a.plot_step_interval == 20 and a.plot_steps == 50
    d = qca_b200.Args.from_argv([])
    assert d.rules.ncells == 9 and d.plot_step_interval == 200 and d.num_steps == 10000


@pytest.mark.parametrize("name", golden_names("exact"))
def test_classical_evolution_equals_reference(name):
    spec, g = load_golden(name)
    rules = qca_b200.Rules(spec["ncells"], range(spec["lo"], spec["hi"]), spec["distance"])
    got = qca_b200.Algorithm.classical_evolution(g["d_population"][0], rules, g["classical"].shape[0])
    assert np.array_equal(got, g["classical"])


def test_periodic_rules_rejected_like_reference():
    with pytest.raises(NotImplementedError):
        qca_b200.Rules(9, range(1, 2), 1, periodic=True)
