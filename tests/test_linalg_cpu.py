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
