"""CPU oracle for the TDVP path (``--algorithm 1tdvp | 2tdvp``).

TEST INFRASTRUCTURE ONLY (see qca_oracle.py).  numpy restatement of the reference's
``algorithms/tdvp.py`` and of the MPS sweeps it relies on (``tensor_networks/mps.py``), including
its dense effective Hamiltonians and exact exponentials (``lautils.timestep``), so it is only
usable at small bond dimension.  Pinned to the ``tdvp*`` fixtures in tests/golden/.

Tensor conventions of the reference: MPS ``A[i][p, l, r]``, MPO ``W[i][a, b, wl, wr]``,
environments ``L[x, w, y]`` / ``R[u, w, v]`` with x/u on the ket (``A``) side and y/v on the bra
(``A.conj()``) side.
"""
from __future__ import annotations

import numpy as np

import qca_oracle as base


# ----------------------------------------------------------------------------
# MPS helpers (tensor_networks/mps.py)
# ----------------------------------------------------------------------------
def product_mps(plist) -> list[np.ndarray]:
    """mps.py:35-52."""
    out = []
    for p in plist:
        t = np.zeros((2, 1, 1), dtype=complex)
        t[:, 0, 0] = [(1.0 - p) ** 0.5, p ** 0.5]
        out.append(t)
    return out


def fit(t: np.ndarray, shape) -> np.ndarray:
    """truncate_and_pad_into_shape, mps.py:227-236."""
    cut = tuple(slice(0, min(a, b)) for a, b in zip(shape, t.shape))
    if all(b >= a for a, b in zip(shape, t.shape)):
        return t[cut]
    out = np.zeros(shape, dtype=t.dtype)
    out[cut] = t[cut]
    return out


def left_qr(a: np.ndarray, reduced=True):
    """mps.py:84-88."""
    s = a.shape
    q, r = np.linalg.qr(a.reshape(s[0] * s[1], s[2]), mode="reduced" if reduced else "complete")
    return q.reshape(s[0], s[1], -1), r


def right_qr(a: np.ndarray, reduced=True):
    """mps.py:91-98."""
    q, r = left_qr(a.transpose(0, 2, 1), reduced)
    return q.transpose(0, 2, 1), r.T


def shift_centre_right(mps: list, i: int) -> None:
    """orthonormalize_left_qr, mps.py:146-164."""
    a = mps[i]
    q, r = left_qr(a)
    mps[i] = fit(q, a.shape)
    r = fit(r, (a.shape[2], mps[i + 1].shape[1]))
    mps[i + 1] = np.einsum("xl,plr->pxr", r, mps[i + 1])


def shift_centre_left(mps: list, i: int) -> None:
    """orthonormalize_right_qr, mps.py:166-181."""
    a = mps[i]
    q, r = right_qr(a)
    mps[i] = fit(q, a.shape)
    r = fit(r, (mps[i - 1].shape[2], a.shape[1]))
    mps[i - 1] = np.einsum("plr,rx->plx", mps[i - 1], r)


def make_site_canonical(mps: list, i: int) -> None:
    """mps.py:184-192."""
    for j in range(i):
        shift_centre_right(mps, j)
    for j in reversed(range(i + 1, len(mps))):
        shift_centre_left(mps, j)


def mps_vector(mps: list) -> np.ndarray:
    """as_vector, mps.py:194-208."""
    acc = mps[0]
    for a in mps[1:]:
        acc = np.einsum("slm,tmr->stlr", acc, a)
        acc = acc.reshape(acc.shape[0] * acc.shape[1], acc.shape[2], acc.shape[3])
    return np.trace(acc, axis1=1, axis2=2)


def measure_mps(mps: list):
    """MPS.measure, mps.py:100-140 (sweeps the orthogonality centre through the chain; the
    tensors are modified exactly as the reference modifies them)."""
    n = len(mps)
    bonds = np.array([a.shape[1] for a in mps] + [mps[-1].shape[2]], dtype=float)
    make_site_canonical(mps, 0)
    pop, ent = np.zeros(n), np.zeros(n)
    for site in range(n):
        if site > 0:
            shift_centre_right(mps, site - 1)
        a = mps[site]
        rho = np.einsum("alr,blr->ab", a, a.conj())
        pop[site] = rho[1, 1].real
        ent[site] = base.entropy_bits(rho)
    return pop, np.round(pop), ent, bonds


# ----------------------------------------------------------------------------
# Effective Hamiltonians (algorithms/tdvp.py:299-365)
# ----------------------------------------------------------------------------
def grow_left(prev, a, w):
    """_assemble_new_layer_H_eff(side='left'), tdvp.py:329-347."""
    return np.einsum("axr,abwm,bys,xwy->rms", a, w, a.conj(), prev)


def grow_right(prev, a, w):
    """_assemble_new_layer_H_eff(side='right'), tdvp.py:329-347."""
    return np.einsum("alu,abmw,bkv,uwv->lmk", a, w, a.conj(), prev)


def heff_matrix(left, right, w):
    """_assemble_H_eff + reshape, tdvp.py:299-310, 352-359: rows (a,x,u), columns (b,y,v)."""
    h = np.einsum("abwm,xwy,umv->axubyv", w, left, right)
    d = h.shape[0] * h.shape[1] * h.shape[2]
    return h.reshape(d, d)


def keff_matrix(left, right):
    """_assemble_K_eff + reshape, tdvp.py:312-326, 176-182: rows (x,u), columns (y,v)."""
    k = np.einsum("xwy,uwv->xuyv", left, right)
    d = k.shape[0] * k.shape[1]
    return k.reshape(d, d)


def timestep(h: np.ndarray, psi: np.ndarray, delta: float) -> np.ndarray:
    """lautils.timestep, lautils.py:58-82: psi contracted with the FIRST index of U."""
    return np.tensordot(psi, base.calculate_U(h, delta), (0, 0))


def svd_with_random_phases(seed: int):
    """An SVD as valid as LAPACK's: U diag(ph), s, diag(ph)^* Vh with random unit phases."""
    rng = np.random.default_rng(seed)

    def svd(a, full_matrices=False):
        u, s, vh = np.linalg.svd(a, full_matrices=full_matrices)
        ph = np.exp(2j * np.pi * rng.random(len(s)))
        u, vh = u.copy(), vh.copy()
        u[:, :len(s)] *= ph
        vh[:len(s), :] *= ph.conj()[:, None]
        return u, s, vh
    return svd


def merge_w(w0, w1):
    """MPO.merge_mpo_tensor_pair / tdvp.py:281-284."""
    w = np.einsum("abwm,cdmv->acbdwv", w0, w1)
    s = w.shape
    return w.reshape(s[0] * s[1], s[2] * s[3], s[4], s[5])


class TDVPOracle:
    """algorithms/tdvp.py:9-146 for algorithm in {'1tdvp', '2tdvp'}."""

    def __init__(self, mps: list, wlist: list, algorithm: str, step_size: float, max_bond_dim: int,
                 svd_epsilon: float, consistent: bool = False, svd=np.linalg.svd):
        """consistent=False restates the reference literally.  The reference re-canonicalises the
        MPS at the start of every step (tdvp.py:54) and inside every measurement (mps.py:110) but
        keeps the right environments it built earlier; they are then stale by the gauge signs of
        the QR, and the 2TDVP result depends on the (arbitrary) phases of the SVD's singular
        vectors at the 1e-4 level (tests/golden/*: ``gauge_spread``).  consistent=True rebuilds the
        right environments after the re-canonicalisation -- the gauge-invariant algorithm the B200
        2TDVP implements.  `svd` lets the tests inject an equivalent SVD with other phases."""
        self.a, self.w = mps, wlist
        self.consistent, self.svd = consistent, svd
        self.algorithm, self.dt, self.chi, self.eps = algorithm, step_size, max_bond_dim, svd_epsilon
        n = len(mps)
        make_site_canonical(self.a, n - 1)  # tdvp.py:23-26
        make_site_canonical(self.a, 0)
        self.left = [None] * n
        self.right = [None] * n
        self.max_bond = [min(2 ** i, 2 ** (n - i), max_bond_dim) for i in range(n + 1)]  # tdvp.py:31-33
        self.target = list(self.max_bond)
        for site in reversed(range(1, n)):  # tdvp.py:37-39
            make_site_canonical(self.a, site - 1)
            self.right[site] = grow_right(self._right(site + 1), self.a[site], self.w[site])

    def _left(self, site):
        return np.ones((1, 1, 1)) if site < 0 else self.left[site]

    def _right(self, site):
        return np.ones((1, 1, 1)) if site >= len(self.a) else self.right[site]

    def _evolve_tensor(self, a, left, right, w, delta):
        """_evolve_A, tdvp.py:350-365."""
        return timestep(heff_matrix(left, right, w), a.reshape(-1), delta).reshape(a.shape)

    def _evolve_site(self, site, delta):
        return self._evolve_tensor(self.a[site], self._left(site - 1), self._right(site + 1), self.w[site], delta)

    def _two_site(self, i, j):
        """_evolve_split_and_truncate, tdvp.py:271-296."""
        al, ar = self.a[i], self.a[j]
        theta = np.einsum("alm,bmr->ablr", al, ar).reshape(4, al.shape[1], ar.shape[2])
        new = self._evolve_tensor(theta, self._left(i - 1), self._right(j + 1), merge_w(self.w[i], self.w[j]), self.dt / 2)
        mat = new.reshape(2, 2, al.shape[1], ar.shape[2]).transpose(0, 2, 1, 3).reshape(2 * al.shape[1], 2 * ar.shape[2])
        u, s, vh = self.svd(mat, full_matrices=False)
        cap = min(self.chi, len(s))
        keep = next((k for k in range(len(s)) if np.linalg.norm(s[k:]) < self.eps), cap)
        keep = min(keep, cap)
        ul = u.reshape(2, al.shape[1], -1)[:, :, :keep]
        vr = vh.reshape(-1, 2, ar.shape[2]).transpose(1, 0, 2)[:, :keep, :]
        sk = s[:keep] / np.linalg.norm(s[:keep])
        return ul, sk, vr

    def step(self):
        """do_time_step, tdvp.py:50-63."""
        n = len(self.a)
        make_site_canonical(self.a, 0)
        if self.consistent:
            for site in reversed(range(1, n)):
                self.right[site] = grow_right(self._right(site + 1), self.a[site], self.w[site])
        if self.algorithm == "2tdvp":
            for site in range(n - 1):  # tdvp.py:107-125
                ul, s, vr = self._two_site(site, site + 1)
                self.a[site] = ul
                self.a[site + 1] = np.einsum("k,pkr->pkr", s, vr)
                if site < n - 2:
                    self.left[site] = grow_left(self._left(site - 1), self.a[site], self.w[site])
                    self.a[site + 1] = self._evolve_site(site + 1, -self.dt / 2)
            for site in reversed(range(1, n)):  # tdvp.py:127-145
                ul, s, vr = self._two_site(site - 1, site)
                self.a[site] = vr
                self.a[site - 1] = np.einsum("plk,k->plk", ul, s)
                if site > 1:
                    self.right[site] = grow_right(self._right(site + 1), self.a[site], self.w[site])
                    self.a[site - 1] = self._evolve_site(site - 1, -self.dt / 2)
            return
        # 1tdvp: tdvp.py:65-105
        for site in range(n):
            shape = (2, self.target[site], self.target[site + 1])
            new = self._evolve_site(site, self.dt / 2)
            if site == n - 1:
                self.a[site] = new
                continue
            q, c = left_qr(new, reduced=False)
            self.a[site] = fit(q, shape)
            self.left[site] = grow_left(self._left(site - 1), self.a[site], self.w[site])
            c = fit(c, (shape[2], c.shape[1]))
            k = keff_matrix(self._left(site), self._right(site + 1))
            c = timestep(k, c.reshape(-1), -self.dt / 2).reshape(c.shape)
            self.a[site + 1] = np.einsum("xl,plr->pxr", c, self.a[site + 1])
        for site in reversed(range(n)):
            shape = (2, self.target[site], self.target[site + 1])
            new = self._evolve_site(site, self.dt / 2)
            if site == 0:
                self.a[site] = new
                continue
            q, c = right_qr(new, reduced=False)
            self.a[site] = fit(q, shape)
            self.right[site] = grow_right(self._right(site + 1), self.a[site], self.w[site])
            c = fit(c, (c.shape[0], shape[1]))
            k = keff_matrix(self._left(site - 1), self._right(site))
            c = timestep(k, c.reshape(-1), -self.dt / 2).reshape(c.shape)
            self.a[site - 1] = np.einsum("plr,rx->plx", self.a[site - 1], c)


def run_tdvp(state: str, ncells: int, distance: int, lo: int, hi: int, algorithm: str, step_size: float,
             num_steps: int, plot_step_interval: int, max_bond_dim: int, svd_epsilon: float,
             consistent: bool = False, svd=np.linalg.svd):
    """quantum_game.py:82-119 for a TDVP algorithm: measure every plot_step_interval steps."""
    mps = product_mps(base.initial_plist(state, ncells, distance))
    algo = TDVPOracle(mps, base.mpo_tensors(ncells, distance, lo, hi), algorithm, step_size, max_bond_dim, svd_epsilon,
                      consistent=consistent, svd=svd)
    pops, ents, bonds = [], [], []
    for step in range(num_steps):
        if step % plot_step_interval == 0:
            p, _, e, b = measure_mps(algo.a)
            pops.append(p), ents.append(e), bonds.append(b)
        algo.step()
    return np.array(pops), np.array(ents), np.array(bonds), mps_vector(algo.a)
