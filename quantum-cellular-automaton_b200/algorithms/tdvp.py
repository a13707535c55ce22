"""``TDVP``: one- and two-site time-dependent variational principle on one GPU, drop-in for
algorithms/tdvp.py:9-146 (``--algorithm 1tdvp | 2tdvp``).

The reference assembles every effective Hamiltonian as a dense (d*Dl*Dr)^2 matrix
(tdvp.py:299-310, 352-359) and exponentiates it by ``eigh`` (lautils.py:58-82), which limits it
to bond dimension ~32.  Here the MPS, the MPO and the environments live on the device and

* H_eff is applied matrix-free:  out[b,y,v] = sum W[a,b,wl,wr] L[x,wl,y] R[u,wr,v] psi[a,x,u]
  in the L.psi -> W -> .R order (8 w chi^3 complex MACs for a two-site tensor instead of 16 chi^4),
* exp(-i pi/2 delta H_eff) psi comes from a Lanczos recursion that never leaves the device
  (csrc/qca_heff.cu: two DMMA GEMMs + a sparse site-operator mix per application, fused
  re-orthogonalisation kernels, the small tridiagonal exponential by Jacobi on the device; Krylov
  dimension fixed from |delta| * spectral bound, no host synchronisation inside a sweep); tensors
  small enough for a dense solve (dimension <= DENSE_LIMIT) take the reference's exact route,
* QR runs on the device with LAPACK's Householder conventions (csrc/qca_linalg.cu -- the
  reference's numbers depend on them), SVD / truncation on the device through torch.linalg.

Gauge consistency (2tdvp).  The reference re-canonicalises the MPS at the start of every step
(tdvp.py:54) and inside every measurement (mps.py:110) but keeps the right environments it built
before; they are then stale by the QR's gauge signs and its 2TDVP numbers depend on the arbitrary
phases of the singular vectors np.linalg.svd returns: fed an equivalent SVD with other phases, the
unmodified reference moves its own populations by up to 4e-4 and entropies by 7e-3
(``gauge_spread_*`` in tests/golden/tdvp2_*.npz).  No implementation on another SVD can reproduce
those digits, so 2tdvp here is the gauge-INVARIANT algorithm: whenever the gauge of the tensors has
changed (construction, measurement, a new psi) the right environments are rebuilt.  It agrees with
the reference within the reference's own spread and with the oracle's ``consistent=True`` variant
to 1e-8.  1tdvp involves no SVD and mirrors the reference literally (1e-8 against its fixtures).

Index conventions are the reference's: ``A[p,l,r]``, ``W[a,b,wl,wr]``, environments ``L[x,w,y]`` /
``R[u,w,v]`` with x/u on the ket side.  Same sweep order, same truncation rule
(tdvp.py:289-296), same gauge handling (mps.py:146-192) as the reference.
"""
from __future__ import annotations

import math
import os

import numpy as np

from .algorithm import Algorithm
from .. import _lib
from ..linalg import SiteOperator, env_grow, gram_svd, gram_svd_at_cap, heff_expm, householder_qr
from ..tensor_networks import MPS, MPO

DENSE_LIMIT = 64          # effective dimension up to which H_eff is exponentiated densely
KRYLOV_TOL = 1e-16
KRYLOV_MAX = 64           # HE_MAXK of csrc/qca_heff.cu
SVD_EPSILON_FLOOR = 1e-14  # the device SVD resolves singular values to ~1e-15 sigma_1 ABSOLUTE: a cut-off below this keeps noise


def _torch():
    import torch
    return torch


def krylov_dimension(rho: float) -> int:
    """Smallest m with (rho/2)^m / m! below KRYLOV_TOL (error bound of the m-step Lanczos
    approximation of exp(-i t H) v for |t| * ||H|| = rho)."""
    m, term = 1, rho / 2.0
    while term > KRYLOV_TOL and m < 4 * KRYLOV_MAX:
        m += 1
        term *= (rho / 2.0) / m
    return max(m + 1, 4)


def krylov_substeps(rho: float) -> int:
    """Number of equal pieces a step of |t| * ||H|| = rho is cut into so that every piece meets
    KRYLOV_TOL with at most KRYLOV_MAX Lanczos vectors (rho up to ~26 needs one piece)."""
    k = 1
    while krylov_dimension(rho / k) > KRYLOV_MAX:
        k += 1
    return k


def mpo_norm_bound(tensors) -> float:
    """Upper bound of ||H|| for an arbitrary MPO W[i][a,b,wl,wr]: triangle inequality over the
    operator strings, i.e. the same chain of matrices with every 2x2 block replaced by its operator
    norm.  Loose (every string counts separately) but safe; used when H is not the rule Hamiltonian,
    whose tight bound comes from qca_spectral_bound."""
    vec = None
    for w in tensors:
        w = np.asarray(w)
        m = np.linalg.norm(w.transpose(2, 3, 0, 1), ord=2, axis=(2, 3))    # (wl, wr)
        vec = m if vec is None else vec @ m
    return float(np.trace(vec)) if vec.shape[0] == vec.shape[1] else float(vec.sum())


class TDVP(Algorithm):

    def __init__(self, psi_0: MPS, H: MPO, args, *, device: int = 0) -> None:
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.QcaError(_lib.QCA_ERR_CUDA, "TDVP needs a CUDA device; there is no CPU fallback")
        self.dev = torch.device("cuda", device)
        self.ct = torch.complex128
        if args.algorithm not in ("1tdvp", "2tdvp"):
            raise NotImplementedError(f"algorithm {args.algorithm!r}: only 1tdvp and 2tdvp are implemented on the GPU")
        self._W = [torch.as_tensor(np.ascontiguousarray(w), dtype=self.ct, device=self.dev) for w in H.W]
        self._W_host = [np.ascontiguousarray(w, dtype=np.complex128) for w in H.W]
        self._ops: dict = {}      # site operators of the native H_eff, built on first use (H is constant)
        # ||H_eff|| <= ||H|| <= R: the Krylov dimension follows from |delta| * pi/2 * R.  R of the rule
        # Hamiltonian is tight (qca_spectral_bound); any other MPO gets the string-count bound
        is_rule_h = MPO.hamiltonian_from_rules(args.rules).same_operator_as(H if isinstance(H, MPO) else MPO(list(H.W)))
        self._bound = _lib.spectral_bound(args.rules) if is_rule_h else mpo_norm_bound(H.W)
        if args.algorithm == "2tdvp" and not args.svd_epsilon >= SVD_EPSILON_FLOOR:
            raise ValueError(
                f"--svd-epsilon {args.svd_epsilon} is below the resolution of the device SVD "
                f"({SVD_EPSILON_FLOOR}: singular values are accurate to ~1e-15 of the largest, absolutely, not relatively)")
        super().__init__(psi_0, H, args)
        n = len(self._A)
        self._canonicalize(n - 1)  # tdvp.py:23-26: fix the bond dimensions, then right-orthonormal form
        self._canonicalize(0)
        self._left: list = [None] * n
        self._right: list = [None] * n
        self._max_bond_dims = [min(2 ** i, 2 ** (n - i), args.max_bond_dim) for i in range(n + 1)]
        self._target_bond_dims = list(self._max_bond_dims)
        self._one = torch.ones((1, 1, 1), dtype=self.ct, device=self.dev)
        for site in reversed(range(1, n)):  # tdvp.py:37-39
            if args.algorithm == "1tdvp":
                # literal: the reference moves the centre to site-1 with a full O(N) sweep each time
                self._canonicalize(site - 1)
            # (2tdvp is gauge invariant: every tensor right of site 0 already is right-orthonormal,
            #  so the O(N^2) re-canonicalisations of the reference would only change gauge signs)
            self._right[site] = self._grow_right(self._env_right(site + 1), site)
        self.heff_applications = 0
        self.heff_flops = 0.0      # real FP64 operations of the H_eff contractions (8 per complex MAC)
        self._gauge_dirty = False  # tensors and right environments are in the same gauge right now
        # 2tdvp at the bond cap: a split whose truncation was decided by the cap in the previous time step is computed
        # speculatively (no host read); the step's flags are read once at its end and the step is repeated with the
        # synchronising split if any of them is false
        self._speculate = os.environ.get("QCA_TDVP_SYNC_SPLIT") is None
        self._decided: dict = {}   # (left site, sweep direction) -> the last split there kept exactly the cap
        self._spec_flags: list = []
        self._speculating = False
        self.speculative_splits = 0
        self.repeated_steps = 0

    # -- Algorithm interface ----------------------------------------------------------------
    @property
    def psi(self) -> MPS:
        return MPS([a.cpu().numpy() for a in self._A])

    @psi.setter
    def psi(self, value: MPS) -> None:
        torch = _torch()
        self._A = [torch.as_tensor(np.ascontiguousarray(a), dtype=self.ct, device=self.dev) for a in value.A]
        self._gauge_dirty = True

    def measure(self, population, d_population, single_site_entropy, bond_dims) -> None:
        """MPS.measure (mps.py:100-140): sweep the orthogonality centre through the chain on the
        device, one D2H copy of the N reduced density matrices at the end.

        1tdvp mirrors the reference literally: the stored tensors are re-gauged in place with the
        LAPACK-convention QR (its later numbers depend on that gauge).  2tdvp is gauge invariant, so
        the sweep runs on a working copy with the library QR and the stored tensors -- and with them
        the right environments -- stay valid: no re-canonicalisation after a measurement."""
        torch = _torch()
        n = len(self._A)
        assert n + 1 == len(bond_dims)
        bond_dims[:n] = [a.shape[1] for a in self._A]
        bond_dims[n] = self._A[-1].shape[2]
        rhos = []
        if self.args.algorithm == "1tdvp" or self._gauge_dirty:
            self._canonicalize(0)
            self._gauge_dirty = True
            for site in range(n):
                if site > 0:
                    self._shift_right(site - 1)
                a = self._A[site].reshape(2, -1)
                rhos.append(a @ a.conj().T)
        else:
            # right-canonical with centre 0 (state after a completed step): carry the centre matrix
            carry = None
            for site in range(n):
                a = self._A[site]
                if carry is not None:
                    a = torch.einsum("xl,plr->pxr", carry, a)
                flat = a.reshape(2, -1)
                rhos.append(flat @ flat.conj().T)
                if site < n - 1:
                    s = a.shape
                    _, carry = torch.linalg.qr(a.reshape(s[0] * s[1], s[2]), mode="r")
        rho = torch.stack(rhos).cpu().numpy()
        pop = rho[:, 1, 1].real
        population[...] = pop
        d_population[...] = np.round(pop)
        lam = np.linalg.eigvalsh(rho)
        with np.errstate(divide="ignore", invalid="ignore"):
            single_site_entropy[...] = -np.where(lam > 0, lam * np.log2(np.where(lam > 0, lam, 1.0)), 0.0).sum(axis=1)

    def do_time_step(self) -> None:
        """tdvp.py:50-63."""
        if self.args.algorithm == "2tdvp":
            if self._gauge_dirty:
                # (after a completed step the MPS already is right-canonical with centre 0 and the
                #  right environments match it: re-canonicalising would be a pure gauge change)
                self._canonicalize(0)
                for site in reversed(range(1, len(self._A))):
                    self._right[site] = self._grow_right(self._env_right(site + 1), site)
                self._gauge_dirty = False
            snapshot = (list(self._A), list(self._left), list(self._right))   # tensors are replaced, never written in place
            self._spec_flags = []
            self._speculating = self._speculate
            self._sweep_right_two_site()
            self._sweep_left_two_site()
            if self._spec_flags:
                torch = _torch()
                if not bool(torch.stack(self._spec_flags).all().item()):   # the one host read of a step at the cap
                    self._A, self._left, self._right = (list(x) for x in snapshot)
                    self._decided.clear()
                    self._speculating = False
                    self.repeated_steps += 1
                    self._sweep_right_two_site()
                    self._sweep_left_two_site()
        else:
            self._canonicalize(0)
            self._sweep_right_one_site()
            self._sweep_left_one_site()

    # -- gauge (mps.py:84-98, 146-192, 227-236) ---------------------------------------------------
    def _fit(self, t, shape):
        torch = _torch()
        if all(b >= a for a, b in zip(shape, t.shape)):
            return t[tuple(slice(0, a) for a in shape)]
        out = torch.zeros(shape, dtype=t.dtype, device=t.device)
        cut = tuple(slice(0, min(a, b)) for a, b in zip(shape, t.shape))
        out[cut] = t[cut]
        return out

    def _left_qr(self, a, reduced=True):
        # LAPACK-convention Householder QR: the reference's results depend on the signs of R's
        # diagonal (the environments are not refreshed after re-canonicalisation, tdvp.py:54)
        s = a.shape
        q, r = householder_qr(a.reshape(s[0] * s[1], s[2]), complete=not reduced)
        return q.reshape(s[0], s[1], -1), r

    def _right_qr(self, a, reduced=True):
        q, r = self._left_qr(a.permute(0, 2, 1), reduced)
        return q.permute(0, 2, 1), r.T

    def _shift_right(self, i):
        torch = _torch()
        a = self._A[i]
        q, r = self._left_qr(a)
        self._A[i] = self._fit(q, a.shape)
        r = self._fit(r, (a.shape[2], self._A[i + 1].shape[1]))
        self._A[i + 1] = torch.einsum("xl,plr->pxr", r, self._A[i + 1])

    def _shift_left(self, i):
        torch = _torch()
        a = self._A[i]
        q, r = self._right_qr(a)
        self._A[i] = self._fit(q, a.shape)
        r = self._fit(r, (self._A[i - 1].shape[2], a.shape[1]))
        self._A[i - 1] = torch.einsum("plr,rx->plx", self._A[i - 1], r)

    def _canonicalize(self, i):
        for j in range(i):
            self._shift_right(j)
        for j in reversed(range(i + 1, len(self._A))):
            self._shift_left(j)

    # -- environments (tdvp.py:329-347) ---------------------------------------------------------------
    def _env_left(self, site):
        return self._one if site < 0 else self._left[site]

    def _env_right(self, site):
        return self._one if site >= len(self._A) else self._right[site]

    def _grow_left(self, prev, site):
        """Left environment including `site` (tdvp.py:329-337) on the DMMA kernel (csrc/qca_heff.cu, qca_env_grow)."""
        return env_grow(prev, self._A[site], self._site_operator("one", site))

    def _grow_right(self, prev, site):
        """Right environment including `site` (tdvp.py:339-347): the left update of the mirrored chain."""
        return env_grow(prev, self._A[site].transpose(1, 2).contiguous(), self._site_operator("one_mirrored", site))

    # -- effective Hamiltonians, matrix-free ------------------------------------------------------------
    # (torch.einsum forms: only the dense route of tiny tensors uses them; everything else runs in
    #  csrc/qca_heff.cu)
    def _apply_one_site(self, left, right, w, psi):
        torch = _torch()
        t = torch.einsum("xwy,axu->awyu", left, psi)
        t = torch.einsum("abwm,awyu->bmyu", w, t)
        return torch.einsum("bmyu,umv->byv", t, right)

    def _apply_two_site(self, left, right, w1, w2, theta):
        torch = _torch()
        t = torch.einsum("xwy,acxu->acwyu", left, theta)
        t = torch.einsum("abwm,acwyu->bcmyu", w1, t)
        t = torch.einsum("cdmn,bcmyu->bdnyu", w2, t)
        return torch.einsum("bdnyu,unv->bdyv", t, right)

    def _apply_bond(self, left, right, c):
        torch = _torch()
        t = torch.einsum("xwy,xu->wyu", left, c)
        return torch.einsum("wyu,uwv->yv", t, right)

    def _site_operator(self, kind, site):
        """Device CSR of the site operator(s): ('one', i) -> W_i, ('two', i) -> W_i W_{i+1},
        ('bond', w) -> identity on w MPO channels."""
        key = (kind, site)
        op = self._ops.get(key)
        if op is None:
            if kind == "one":
                op = SiteOperator(self._W_host[site], device=self.dev)
            elif kind == "one_mirrored":   # bond indices swapped: right-environment updates
                op = SiteOperator(np.ascontiguousarray(self._W_host[site].transpose(0, 1, 3, 2)), device=self.dev)
            elif kind == "two":
                op = SiteOperator(self._W_host[site], self._W_host[site + 1], device=self.dev)
            else:
                op = SiteOperator(None, site, device=self.dev)
            self._ops[key] = op
        return op

    # -- exponentials (lautils.py:58-82) ------------------------------------------------------------------
    def _expm_apply(self, apply, left, right, op, psi, delta):
        """exp(-i pi/2 delta H_eff)^T psi in the reference's index convention: the contraction already
        is the transposed action (psi contracted with the row index of the reference's H_eff).
        `apply` is the einsum form of the same operator, used by the dense route only."""
        torch = _torch()
        dim = psi.numel()
        t = (math.pi / 2.0) * delta
        if dim <= DENSE_LIMIT:
            eye = torch.eye(dim, dtype=self.ct, device=self.dev).reshape((dim,) + tuple(psi.shape))
            h = torch.stack([apply(eye[k]).reshape(-1) for k in range(dim)], dim=1)  # column k = H e_k
            h = 0.5 * (h + h.conj().T)
            lam, vec = torch.linalg.eigh(h)
            phase = torch.exp(-1j * t * lam)
            return ((vec * phase) @ (vec.conj().T @ psi.reshape(-1))).reshape(psi.shape)
        pieces = krylov_substeps(abs(t) * self._bound)   # 1 unless |t| * R > ~26 (large --step-size on long chains)
        m = min(krylov_dimension(abs(t) * self._bound / pieces), KRYLOV_MAX, dim)
        dl, dr = left.shape[0], right.shape[0]
        self.heff_applications += m * pieces
        # FP64 operations of the two tensor-core contractions L.psi and T.R (8 per complex MAC); MPO channels
        # that are structurally zero are neither computed nor counted
        self.heff_flops += pieces * m * 8.0 * (op.cols_used * dl * dl * dr + op.rows_used * dl * dr * dr)
        for _ in range(pieces):
            psi = heff_expm(left, right, op, psi, m, t / pieces, spectral_bound=self._bound).reshape(psi.shape)
        return psi

    def _evolve_site(self, site, delta):
        left, right, w = self._env_left(site - 1), self._env_right(site + 1), self._W[site]
        return self._expm_apply(lambda v: self._apply_one_site(left, right, w, v), left, right,
                                self._site_operator("one", site), self._A[site], delta)

    # -- two-site TDVP (tdvp.py:107-145, 271-296) ---------------------------------------------------------
    def _two_site(self, i, j):
        torch = _torch()
        al, ar = self._A[i], self._A[j]
        left, right = self._env_left(i - 1), self._env_right(j + 1)
        w1, w2 = self._W[i], self._W[j]
        theta = torch.einsum("alm,bmr->ablr", al, ar)
        new = self._expm_apply(lambda v: self._apply_two_site(left, right, w1, w2, v), left, right,
                               self._site_operator("two", i), theta, self.args.step_size / 2)
        dl, dr = al.shape[1], ar.shape[2]
        mat = new.permute(0, 2, 1, 3).reshape(2 * dl, 2 * dr)
        # singular values that truncation can never keep are not resolved one by one: with
        # stop_below = eps / (4 sqrt(n)) the unresolved rest has norm < eps, so the first index whose
        # tail norm is under eps (tdvp.py:290-292) always lies inside the resolved part
        eps = self.args.svd_epsilon
        cap = min(self.args.max_bond_dim, 2 * min(dl, dr))
        key = (i, self._sweep_dir)
        if self._speculating and self._decided.get(key) and cap <= min(mat.shape):
            u, s, vh, ok = gram_svd_at_cap(mat, cap, eps)
            self._spec_flags.append(ok)
            self.speculative_splits += 1
            keep = cap
        else:
            info = {}
            u, s, vh, rest = gram_svd(mat, stop_below=eps / (4.0 * math.sqrt(min(mat.shape))), need=cap, tail_floor=eps, info=info)
            if info.get("decided"):
                keep = cap       # everything beyond the cap still weighs >= eps: no need to look at the tail (no host sync)
            else:
                tail = torch.sqrt(torch.flip(torch.cumsum(torch.flip(s * s, [0]), 0), [0]) + rest * rest)
                below = (tail < eps).nonzero()
                keep = min(int(below[0, 0]) if below.numel() else cap, cap, s.shape[0])
            self._decided[key] = bool(info.get("decided_level0")) and keep == cap
        ul = u.reshape(2, dl, -1)[:, :, :keep]
        vr = vh.reshape(-1, 2, dr).permute(1, 0, 2)[:, :keep, :]
        sk = s[:keep] / torch.linalg.vector_norm(s[:keep])
        return ul, sk.to(self.ct), vr

    def _sweep_right_two_site(self):
        torch = _torch()
        n = len(self._A)
        self._sweep_dir = 0
        for site in range(n - 1):
            ul, s, vr = self._two_site(site, site + 1)
            self._A[site] = ul.contiguous()
            self._A[site + 1] = (s[None, :, None] * vr).resolve_conj().contiguous()
            if site < n - 2:
                self._left[site] = self._grow_left(self._env_left(site - 1), site)
                self._A[site + 1] = self._evolve_site(site + 1, -self.args.step_size / 2)

    def _sweep_left_two_site(self):
        n = len(self._A)
        self._sweep_dir = 1
        for site in reversed(range(1, n)):
            ul, s, vr = self._two_site(site - 1, site)
            self._A[site] = vr.resolve_conj().contiguous()   # (vh comes as a lazily conjugated view)
            self._A[site - 1] = (ul * s[None, None, :]).contiguous()
            if site > 1:
                self._right[site] = self._grow_right(self._env_right(site + 1), site)
                self._A[site - 1] = self._evolve_site(site - 1, -self.args.step_size / 2)

    # -- one-site TDVP (tdvp.py:65-105, 164-188) -----------------------------------------------------------
    def _evolve_bond(self, left, right, c):
        return self._expm_apply(lambda v: self._apply_bond(left, right, v), left, right,
                                self._site_operator("bond", left.shape[1]), c, -self.args.step_size / 2)

    def _sweep_right_one_site(self):
        torch = _torch()
        n = len(self._A)
        for site in range(n):
            shape = (2, self._target_bond_dims[site], self._target_bond_dims[site + 1])
            new = self._evolve_site(site, self.args.step_size / 2)
            if site == n - 1:
                self._A[site] = new
                continue
            q, c = self._left_qr(new, reduced=False)
            self._A[site] = self._fit(q, shape).contiguous()
            self._left[site] = self._grow_left(self._env_left(site - 1), site)
            c = self._fit(c, (shape[2], c.shape[1]))
            c = self._evolve_bond(self._env_left(site), self._env_right(site + 1), c)
            self._A[site + 1] = torch.einsum("xl,plr->pxr", c, self._A[site + 1])

    def _sweep_left_one_site(self):
        torch = _torch()
        n = len(self._A)
        for site in reversed(range(n)):
            shape = (2, self._target_bond_dims[site], self._target_bond_dims[site + 1])
            new = self._evolve_site(site, self.args.step_size / 2)
            if site == 0:
                self._A[site] = new
                continue
            q, c = self._right_qr(new, reduced=False)
            self._A[site] = self._fit(q, shape).contiguous()
            self._right[site] = self._grow_right(self._env_right(site + 1), site)
            c = self._fit(c, (c.shape[0], shape[1]))
            c = self._evolve_bond(self._env_left(site - 1), self._env_right(site), c)
            self._A[site - 1] = torch.einsum("plr,rx->plx", self._A[site - 1], c)
