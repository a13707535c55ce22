"""gram_svd (multi-level Gram-matrix SVD used by the 2TDVP split) against numpy's SVD.  The routine is
pure tensor algebra, so its numerics are checked on CPU tensors here and on the device in
tests/test_tdvp_gpu.py."""
import numpy as np
import pytest
import torch

from qca_b200.linalg import gram_svd


def matrix_with_spectrum(m, n, sigma, rng):
    k = min(m, n)
    u = np.linalg.qr(rng.standard_normal((m, k)) + 1j * rng.standard_normal((m, k)))[0]
    v = np.linalg.qr(rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k)))[0]
    return (u * sigma) @ v.conj().T


@pytest.mark.parametrize("m,n,decay", [(64, 64, 30.0), (128, 64, 36.0), (64, 128, 20.0), (256, 256, 5.0), (2, 8, 1.0), (1, 1, 0.0)])
def test_gram_svd_matches_lapack(m, n, decay):
    rng = np.random.default_rng(m + n)
    k = min(m, n)
    sigma = np.exp(-np.arange(k) * decay / k)
    a = matrix_with_spectrum(m, n, sigma, rng)
    u, s, vh, rest = gram_svd(torch.as_tensor(a))
    u, s, vh = u.resolve_conj().numpy(), s.numpy(), vh.resolve_conj().numpy()
    assert float(rest) == 0.0 and len(s) == k
    s_np = np.linalg.svd(a, compute_uv=False)
    assert np.abs(s - s_np).max() < 1e-13                       # backward-stable accuracy, like LAPACK
    big = s_np > 1e-9
    assert (np.abs(s[big] - s_np[big]) / s_np[big]).max() < 1e-7
    assert np.abs((u * s) @ vh - a).max() < 1e-13
    assert np.abs(u[:, big].conj().T @ u[:, big] - np.eye(big.sum())).max() < 1e-4   # worst in directions of weight < 1e-17
    assert np.abs(vh[big] @ vh[big].conj().T - np.eye(big.sum())).max() < 1e-4
    well = s_np > 1e-5
    assert np.abs(u[:, well].conj().T @ u[:, well] - np.eye(well.sum())).max() < 1e-9


def test_gram_svd_stops_below_the_truncation_floor():
    rng = np.random.default_rng(1)
    sigma = np.exp(-np.arange(128) * 36.0 / 128)
    a = matrix_with_spectrum(128, 128, sigma, rng)
    u, s, vh, rest = gram_svd(torch.as_tensor(a), stop_below=1e-6)
    s = s.numpy()
    r = len(s)
    assert r < 128 and s.min() < 1e-3      # stopped early, but resolved well below any 2TDVP cut-off
    assert abs(float(rest) - np.sqrt((sigma[r:] ** 2).sum())) < 1e-12
    assert np.abs(s - sigma[:r]).max() < 1e-13


def test_gram_svd_skips_levels_once_the_bond_cap_decides_the_truncation():
    """At the bond cap 2TDVP keeps exactly `need` singular values as long as everything beyond them
    still weighs >= svd_epsilon (tdvp.py:289-293): the deeper levels need not be resolved."""
    rng = np.random.default_rng(2)
    sigma = np.exp(-np.arange(128) * 36.0 / 128)
    a = matrix_with_spectrum(128, 128, sigma, rng)
    need = 20                                        # sigma[20] ~ 4e-3: inside the first level
    u, s, vh, rest = gram_svd(torch.as_tensor(a), stop_below=1e-16, need=need, tail_floor=1e-14)
    s, r = s.numpy(), len(s)
    assert need <= r < 64
    assert np.abs(s[:need] - sigma[:need]).max() < 1e-13
    assert abs(float(rest) - np.sqrt((sigma[r:] ** 2).sum())) < 1e-12
    # when the tail beyond `need` is lighter than the floor the spectrum is resolved as before
    u2, s2, vh2, rest2 = gram_svd(torch.as_tensor(a), stop_below=1e-16, need=need, tail_floor=1.0)
    assert len(s2) == 128


def test_gram_svd_reports_a_decided_truncation():
    rng = np.random.default_rng(4)
    sigma = np.exp(-np.arange(96) * 20.0 / 96)
    a = matrix_with_spectrum(96, 96, sigma, rng)
    info = {}
    u, s, vh, rest = gram_svd(torch.as_tensor(a), stop_below=1e-16, need=10, tail_floor=1e-14, info=info)
    assert info.get("decided") and len(s) >= 10 and np.abs(s.numpy()[:10] - sigma[:10]).max() < 1e-13
    # exact rank below `need`: nothing is decided by the cap, the caller looks at the tail itself
    low = matrix_with_spectrum(96, 96, np.concatenate([sigma[:6], np.zeros(90)]), rng)
    info = {}
    u, s, vh, rest = gram_svd(torch.as_tensor(low), stop_below=1e-16, need=10, tail_floor=1e-14, info=info)
    assert not info.get("decided")
    tail = np.sqrt(np.cumsum((s.numpy() ** 2)[::-1])[::-1] + float(rest) ** 2)
    assert (tail < 1e-12).nonzero()[0][0] == 6


# -- statement-level model of the multi-CTA Householder QR (csrc/qca_linalg.cu) ---------------------------------
def _zlarfg(col, k):
    """What every CTA of qr_factor_grid_kernel derives from column k: (tau, beta or None, reflector with v[k] = 1)."""
    alpha = col[k]
    xnorm2 = float(np.sum(np.abs(col[k + 1:]) ** 2))
    v = col.copy()
    v[:k] = 0.0
    v[k] = 1.0
    if xnorm2 == 0.0 and alpha.imag == 0.0:
        return 0.0j, None, v                      # H = I: the column (and its diagonal) stay as they are
    beta = -np.copysign(np.sqrt(alpha.real ** 2 + alpha.imag ** 2 + xnorm2), alpha.real)
    tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
    v[k + 1:] = col[k + 1:] / (alpha - beta)
    return tau, beta, v


def grid_qr_model(a, complete=False):
    """qr_factor_grid_kernel (right-looking: one reflector, then every trailing column on its own) followed by
    qr_form_cols_reg_kernel (column j of Q = H_0 ... H_min(j, kmax-1) e_j, every column on its own)."""
    a = np.array(a, dtype=np.complex128)
    m, n = a.shape
    kmax = min(m, n)
    taus, vs = [], []
    for k in range(kmax):
        tau, beta, v = _zlarfg(a[:, k], k)
        if tau != 0:
            for j in range(k + 1, n):             # one warp per column in the kernel
                w = np.vdot(v[k:], a[k:, j])
                a[k:, j] -= v[k:] * (np.conj(tau) * w)
            a[k + 1:, k] = v[k + 1:]              # written by CTA 0 after the grid barrier
            a[k, k] = beta
        taus.append(tau)
        vs.append(v)
    kq = m if complete else kmax
    q = np.zeros((m, kq), dtype=np.complex128)
    for j in range(kq):                           # one warp per column, the column in registers
        c = np.zeros(m, dtype=np.complex128)
        c[j] = 1.0
        for k in range(min(j, kmax - 1), -1, -1):
            if taus[k] != 0:
                c[k:] -= vs[k][k:] * (taus[k] * np.vdot(vs[k][k:], c[k:]))
        q[:, j] = c
    r = np.triu(a)[:kq, :]
    return q, r


@pytest.mark.parametrize("m,n", [(1, 1), (2, 1), (4, 2), (2, 4), (16, 8), (8, 16), (33, 7), (64, 32), (40, 40)])
@pytest.mark.parametrize("complete", [False, True])
def test_grid_qr_model_is_numpy_qr(m, n, complete):
    """The algorithm of the multi-CTA kernels reproduces numpy.linalg.qr including the signs of R's diagonal, the
    tau = 0 branch of a zero column and the completion of Q."""
    rng = np.random.default_rng(100 * m + n)
    mat = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
    cases = [mat]
    if n > 1:
        deficient = mat.copy(); deficient[:, 1] = 0.0
        cases.append(deficient)
    cases.append(np.eye(m, n, dtype=np.complex128))            # every reflector is the identity or a sign flip
    for a in cases:
        q_np, r_np = np.linalg.qr(a, mode="complete" if complete else "reduced")
        q, r = grid_qr_model(a, complete)
        assert q.shape == q_np.shape and r.shape == r_np.shape
        assert np.abs(r - r_np).max() < 1e-12 * max(1.0, np.abs(r_np).max())
        assert np.abs(q - q_np).max() < 1e-11
        assert np.abs(q @ r - a).max() < 1e-12
