"""CPU tests of the host side of the library: the C ABI loads and exports every symbol of
include/qca_b200.h, and its planning functions (spectral bound, Chebyshev plan, tile-pass plan)
are right.  No compute call needs a GPU here."""
import os
import re

import numpy as np
import pytest
from scipy.special import jv

import pass_model
import qca_b200
import qca_oracle as oracle
from conftest import ROOT, RuleNS
from qca_b200 import _lib


def test_library_exports_every_header_symbol():
    header = open(os.path.join(ROOT, "include", "qca_b200.h")).read()
    declared = set(re.findall(r"\b(qca_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(qca_b200.lib, name)
    assert b"sm_100a" in qca_b200.lib.qca_version()


def test_no_cpu_fallback_without_device():
    if qca_b200.lib.qca_device_count() > 0:
        pytest.skip("a device is present")
    rules = qca_b200.Rules(5, range(1, 2), 1)
    with pytest.raises(qca_b200.QcaError) as err:
        qca_b200.Exact(qca_b200.states.make("single", rules), None, qca_b200.Args(rules=rules))
    assert err.value.code == _lib.QCA_ERR_CUDA


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "quantum-cellular-automaton_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "qca_oracle" not in text and "oracle/" not in text, f


@pytest.mark.parametrize("n,d,lo,hi", [(6, 1, 1, 2), (8, 1, 1, 3), (9, 2, 2, 4), (8, 2, 1, 5), (7, 3, 2, 5),
                                       (3, 1, 1, 2), (1, 1, 1, 2), (2, 2, 0, 1), (8, 4, 3, 6)])
def test_spectral_bound_is_max_row_sum(n, d, lo, hi):
    h = oracle.rule_hamiltonian_direct(n, d, lo, hi)
    bound = _lib.spectral_bound(RuleNS(n, d, lo, hi))
    assert bound == np.abs(h).sum(axis=1).max()
    assert bound >= np.linalg.eigvalsh(h).max() - 1e-12


def test_spectral_bound_large_chain_is_cheap():
    assert _lib.spectral_bound(RuleNS(30, 2, 2, 4)) <= 30
    assert _lib.spectral_bound(RuleNS(33, 1, 1, 2)) <= 33


@pytest.mark.parametrize("z", [0.0, 0.3, 3.0, 14.1, 47.1, 120.0])
def test_chebyshev_plan_is_bessel(z):
    a = _lib.chebyshev_plan(z, 1e-15)
    k = np.arange(len(a))
    want = np.where(k == 0, 1.0, 2.0) * jv(k, z)
    assert np.abs(a - want).max() < 5e-14  # scipy jv is itself ~1e-14 at large z; Jacobi-Anger below is exact
    # truncation: what is left out is below the tolerance
    tail = 2.0 * np.abs(jv(np.arange(len(a), len(a) + 60), z)).sum()
    assert tail < 2e-15
    # Jacobi-Anger at a few points: exp(-i z x) = sum_k a_k (-i)^k T_k(x)
    for x in (-1.0, -0.3, 0.0, 0.77, 1.0):
        s = sum(a[j] * (-1j) ** j * np.cos(j * np.arccos(x)) for j in range(len(a)))
        assert abs(s - np.exp(-1j * z * x)) < 1e-13


def test_error_codes():
    with pytest.raises(qca_b200.QcaError) as e:
        _lib.spectral_bound(RuleNS(0, 1, 1, 2))
    assert e.value.code == _lib.QCA_ERR_ARG
    with pytest.raises(qca_b200.QcaError) as e:
        _lib.spectral_bound(RuleNS(9, 8, 1, 2))
    assert e.value.code == _lib.QCA_ERR_UNSUPPORTED
    with pytest.raises(qca_b200.QcaError):
        _lib.chebyshev_plan(-1.0)
    with pytest.raises(qca_b200.QcaError):
        _lib.plan_passes(0)


@pytest.mark.parametrize("nbits", [1, 2, 5, 12, 13, 14, 17, 22, 23, 27, 30, 31, 33])
def test_pass_plan_geometry(nbits):
    passes = _lib.plan_passes(nbits)
    covered = 0
    for ps in passes:
        L, H0, M = ps["low_bits"], ps["high_start"], ps["high_bits"]
        assert 1 <= L + M <= 13 and H0 >= L and H0 + M <= nbits
        tile_bits = ((1 << L) - 1) | (((1 << M) - 1) << H0)
        assert ps["flip_mask"] & ~tile_bits == 0, "a flipped qubit must be inside the tile"
        assert ps["flip_mask"] & covered == 0, "each qubit is flipped by exactly one pass"
        covered |= ps["flip_mask"]
        if M:
            assert L >= 4, "rows of a strided tile stay >= 128 bytes"
    assert covered == (1 << nbits) - 1
    assert len(passes) == (1 if nbits <= 13 else 1 + -(-(nbits - 13) // 9))


@pytest.mark.parametrize("n,d,lo,hi", [(9, 1, 1, 2), (15, 1, 1, 2), (16, 2, 2, 4), (15, 3, 2, 5)])
def test_tiled_operator_equals_reference_hamiltonian(n, d, lo, hi):
    """The tile decomposition (as planned by the C library) reproduces MPO.as_matrix() @ v."""
    rng = np.random.default_rng(n)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    passes = _lib.plan_passes(n)
    for ps in passes:  # tiles of one pass partition the index space
        T = ps["low_bits"] + ps["high_bits"]
        seen = np.concatenate([pass_model.tile_indices(ps, t) for t in range(1 << (n - T))])
        assert np.array_equal(np.sort(seen), np.arange(1 << n))
    xs = np.arange(1 << n, dtype=np.int64)
    rot = 1j ** (pass_model.popcount(xs) % 4)
    phi = psi / rot
    kphi = (pass_model.apply_k_by_tiles(phi.real.copy(), passes, n, n, d, lo, hi)
            + 1j * pass_model.apply_k_by_tiles(phi.imag.copy(), passes, n, n, d, lo, hi))
    got = 1j * rot * kphi  # H = D (iK) D^-1
    # reference operator, matrix-free in numpy (validated against the dense one at small n)
    want = np.zeros_like(psi)
    act = pass_model.activity(xs, n, d, lo, hi)
    for g in range(n):
        on = ((act >> g) & 1).astype(bool)
        want[on] += psi[xs[on] ^ (1 << g)]
    if n <= 10:
        assert np.abs(want - oracle.rule_hamiltonian_direct(n, d, lo, hi) @ psi).max() < 1e-12
    assert np.abs(got - want).max() < 1e-12


@pytest.mark.parametrize("n,d,lo,hi,tau", [(8, 1, 1, 2, 1.0), (7, 2, 2, 4, 1.0), (8, 1, 1, 3, -0.5), (6, 1, 1, 2, 0.005)])
def test_clenshaw_in_rotated_frame_equals_calculate_U(n, d, lo, hi, tau):
    """exp(-i pi/2 tau H) psi == D * Clenshaw(K) * D^-1 psi with the library's coefficients."""
    rng = np.random.default_rng(7)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    u = oracle.calculate_U(oracle.rule_hamiltonian_direct(n, d, lo, hi), tau)
    bound = _lib.spectral_bound(RuleNS(n, d, lo, hi))
    a = _lib.chebyshev_plan(bound * abs(tau) * np.pi / 2, 1e-15)
    kmat = pass_model.k_matrix(n, d, lo, hi)
    assert np.array_equal(kmat, -kmat.T)
    rot = 1j ** (pass_model.popcount(np.arange(1 << n, dtype=np.int64)) % 4)
    phi = psi / rot
    out = rot * pass_model.clenshaw_exp(kmat, phi, a, bound, np.sign(tau))
    assert np.abs(out - u @ psi).max() < 1e-13
