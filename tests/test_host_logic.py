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


def test_site_operator_csr_equals_the_einsum_contraction():
    """linalg.site_operator_csr (the sparse channel mix between the two DMMA contractions of H_eff) against
    the reference's dense contraction of the MPO tensors (tdvp.py:280-283, 350-365), plus its masks of
    structurally empty channels."""
    from qca_b200.linalg import site_operator_csr
    rng = np.random.default_rng(0)
    for (n, d, lo, hi) in [(8, 1, 1, 2), (9, 2, 2, 4), (10, 3, 2, 5)]:
        w = qca_b200.MPO.hamiltonian_from_rules(qca_b200.Rules(n, range(lo, hi), d)).W
        w1, w2 = np.asarray(w[n // 2 - 1]), np.asarray(w[n // 2])
        wl, wr = w1.shape[2], w2.shape[3]
        # two sites
        rowptr, col, val, (g, a, b) = site_operator_csr(w1, w2)
        assert (g, a, b) == (4, wl, wr) and rowptr.shape == (4 * wr + 1,)
        dense = np.zeros((4 * wr, 4 * wl), dtype=complex)
        for r in range(4 * wr):
            dense[r, col[rowptr[r]:rowptr[r + 1]]] = val[rowptr[r]:rowptr[r + 1]]
        t1 = rng.standard_normal((2, 2, wl, 3, 5)) + 1j * rng.standard_normal((2, 2, wl, 3, 5))
        want = np.einsum("cdmn,bcmyu->bdnyu", w2, np.einsum("abwm,acwyu->bcmyu", w1, t1))
        got = (dense @ t1.reshape(4 * wl, -1)).reshape(2, 2, wr, 3, 5)
        assert np.abs(got - want).max() < 1e-13
        assert int(rowptr[-1]) < dense.size // 4          # the automaton's MPO is very sparse
        # one site
        rowptr, col, val, (g, a, b) = site_operator_csr(w1)
        dense = np.zeros((2 * w1.shape[3], 2 * wl), dtype=complex)
        for r in range(dense.shape[0]):
            dense[r, col[rowptr[r]:rowptr[r + 1]]] = val[rowptr[r]:rowptr[r + 1]]
        t1 = rng.standard_normal((2, wl, 3, 5)) + 1j * rng.standard_normal((2, wl, 3, 5))
        want = np.einsum("abwm,awyu->bmyu", w1, t1)
        assert np.abs((dense @ t1.reshape(2 * wl, -1)).reshape(want.shape) - want).max() < 1e-13
    # bond matrix: identity over the MPO channels
    rowptr, col, val, (g, a, b) = site_operator_csr(None, 6)
    assert g == 1 and list(col) == list(range(6)) and np.all(val == 1)


def test_heff_workspace_is_sized_on_the_host():
    """qca_heff_workspace_bytes is pure host logic: grows with the Krylov dimension and rejects bad shapes."""
    import ctypes as C
    h = _lib.HeffStruct(1, 1, 1, 1, 1, 64, 48, 6, 6, 4, 1, (C.c_uint32 * 4)(0x17, 0x2b, 0x17, 0x2b), (C.c_uint32 * 4)(0x35, 0x35, 0x1b, 0x1b))
    sizes = []
    for m in (0, 1, 8, 64):
        n = C.c_uint64()
        _lib.check(_lib.lib.qca_heff_workspace_bytes(C.byref(h), m, C.byref(n)))
        sizes.append(n.value)
    dim = 4 * 64 * 48 * 16
    assert sizes[0] < sizes[2] < sizes[3] and sizes[3] - sizes[2] == 56 * dim
    n = C.c_uint64()
    assert _lib.lib.qca_heff_workspace_bytes(C.byref(h), 65, C.byref(n)) == _lib.QCA_ERR_ARG
    h.g = 3
    assert _lib.lib.qca_heff_workspace_bytes(C.byref(h), 4, C.byref(n)) == _lib.QCA_ERR_ARG


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the oracle's restatement of the reference's dense algorithm on host cores)."""
    import json
    import subprocess
    import sys
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-num-cells", "8"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


@pytest.mark.parametrize("nbits", [13, 15, 18])
def test_fused_measurement_decomposition(nbits):
    """The thread/row decomposition of csrc/qca_measure.cu (pass_model.measure_tiles_model) yields, over the
    tile passes of a register, the sums (s0, s1, w) of every index bit exactly once."""
    rng = np.random.default_rng(nbits)
    vec = rng.standard_normal(1 << nbits)
    seen = {}
    for ps in _lib.plan_passes(nbits):
        for g, s0, s1, w in pass_model.measure_tiles_model(vec, ps):
            assert g not in seen
            seen[g] = (s0, s1, w)
    assert sorted(seen) == list(range(nbits))
    for g, (s0, s1, w) in seen.items():
        t = vec.reshape(-1, 2, 1 << g)
        assert abs(s0 - (t[:, 0, :] ** 2).sum()) < 1e-9 and abs(s1 - (t[:, 1, :] ** 2).sum()) < 1e-9
        assert abs(w - (t[:, 0, :] * t[:, 1, :]).sum()) < 1e-9


@pytest.mark.parametrize("n,d,lo,hi,cbmax,minlow", [(14, 1, 1, 2, 3, 4), (15, 2, 2, 4, 3, 4), (17, 2, 2, 4, 3, 4), (18, 1, 1, 3, 3, 4),
                                                   (19, 2, 2, 4, 1, 4), (18, 3, 2, 5, 3, 4), (19, 4, 3, 6, 0, 4), (20, 2, 2, 4, 2, 9),
                                                   (19, 2, 1, 3, 0, 4)])
def test_cluster_kernel_model_equals_rule_operator(n, d, lo, hi, cbmax, minlow):
    """The cluster tile-pass kernel's arithmetic (tests/pass_model.apply_k_v3: plan from qca_plan_passes_v3, index
    expansion, thread-constant + per-row window tables, partner-CTA exchange) reproduces K for every amplitude."""
    passes = _lib.plan_passes_v3(n, cbmax, minlow)
    covered = 0
    for ps in passes:
        assert ps["low_bits"] + ps["high_bits"] == 14 and ps["low_bits"] >= 4 and 0 <= ps["cluster_bits"] <= cbmax
        assert covered & ps["flip_mask"] == 0
        covered |= ps["flip_mask"]
        m = ps["high_bits"] + ps["cluster_bits"]
        want = ((1 << m) - 1) << ps["high_start"] if ps["high_bits"] else (1 << (14 + ps["cluster_bits"])) - 1
        assert ps["flip_mask"] == want
    assert covered == (1 << n) - 1
    rng = np.random.default_rng(n)
    phi = rng.standard_normal(1 << n)
    got = pass_model.apply_k_v3(phi, passes, n, d, lo, hi)
    xs = np.arange(1 << n, dtype=np.int64)
    act = pass_model.activity(xs, n, d, lo, hi)
    want = np.zeros_like(phi)
    for g in range(n):
        on = ((act >> g) & 1).astype(bool)
        sign = np.where((xs >> g) & 1, -1.0, 1.0)
        want[on] += sign[on] * phi[xs[on] ^ (1 << g)]
    assert np.abs(got - want).max() < 1e-12


def test_cluster_plan_of_the_benchmark_sizes():
    assert [(p["low_bits"], p["high_bits"], p["cluster_bits"]) for p in _lib.plan_passes_v3(30)] == [(14, 0, 3), (4, 10, 3)]
    assert [(p["low_bits"], p["high_bits"], p["cluster_bits"]) for p in _lib.plan_passes_v3(27)] == [(14, 0, 3), (4, 10, 0)]
    assert len(_lib.plan_passes_v3(33)) == 3 and _lib.plan_passes_v3(13) == []
