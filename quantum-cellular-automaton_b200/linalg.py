"""Device linear algebra with the reference's (numpy/LAPACK) conventions."""
from __future__ import annotations

import ctypes as C
import math

from . import _lib


def _plain(t):
    """Contiguous tensor whose MEMORY holds its values: torch's lazy conjugation (`x.conj()` only sets a flag, and
    `.contiguous()` keeps it) must be materialised before a raw pointer goes to the C library."""
    return t.resolve_conj().resolve_neg().contiguous()


def householder_qr(mat, complete: bool = False):
    """QR of a complex128 CUDA matrix with LAPACK's Householder sign convention -- the factors
    numpy.linalg.qr(mat, mode='reduced' | 'complete') returns (csrc/qca_linalg.cu).  Returns (q, r)
    as torch tensors on the same device."""
    import torch
    assert mat.is_cuda and mat.dtype == torch.complex128 and mat.dim() == 2
    m, n = mat.shape
    kq = m if complete else min(m, n)
    a = mat.resolve_conj().T.clone(memory_format=torch.contiguous_format)     # row-major (n, m) == column-major (m, n)
    tau = torch.empty(min(m, n), dtype=mat.dtype, device=mat.device)
    q = torch.empty((kq, m), dtype=mat.dtype, device=mat.device)   # column-major m x kq
    r = torch.empty((n, kq), dtype=mat.dtype, device=mat.device)   # column-major kq x n
    stream = torch.cuda.current_stream(mat.device).cuda_stream
    _lib.check(_lib.lib.qca_qr_householder(C.c_void_p(a.data_ptr()), m, n, C.c_void_p(tau.data_ptr()),
                                           C.c_void_p(q.data_ptr()), kq, C.c_void_p(r.data_ptr()),
                                           C.c_void_p(stream)))
    return q.T, r.T


def gram_svd(mat, stop_below: float = 0.0, level_ratio: float = 1e-3, max_levels: int = 6, need: int | None = None,
             tail_floor: float = 0.0, info: dict | None = None):
    """Thin SVD of a complex128 CUDA matrix from Hermitian eigendecompositions of Gram matrices.

    cuSOLVER's SVD of a 512 x 512 complex128 matrix takes 60-130 ms on a B200 and was 88 % of a
    chi = 256 2TDVP sweep; ``eigh`` of the Gram matrix takes 6.5 ms.  One Gram step resolves
    singular values only down to ~1e-8 of the largest, so the spectrum is peeled in levels: the
    eigenvectors with sigma >= level_ratio * (largest of the level) are accepted (relative error at
    most eps / level_ratio^2 = 1e-10, at the bottom of a level), the matrix is deflated onto the
    remaining right-singular subspace and the procedure repeats there, where the small singular
    values are the large ones.  Absolute accuracy is eps * sigma_1 / level_ratio = 1e-13 sigma_1,
    the same order as LAPACK's backward-stable SVD.

    stop_below: singular values below it are not resolved individually (2TDVP never keeps them:
    tdvp.py:290-292 truncates where the tail norm drops under svd_epsilon); they are returned only
    through `rest_norm`, the Frobenius norm of the unresolved part.
    need / tail_floor: once `need` singular values are resolved and everything beyond them (resolved
    or not) still has Frobenius norm >= tail_floor, the caller's truncation is decided -- it keeps
    exactly `need` (the bond cap; tdvp.py:289-293) -- and the deeper levels are skipped; `info["decided"]` tells
    the caller so (it can then skip its own look at the tail).

    Returns (u, s, vh, rest_norm): s descending, u[:, i] = mat @ v_i / s_i, mat ~= u diag(s) vh
    up to rest_norm.
    """
    import torch
    m, n = mat.shape
    if m < n:  # work on the side with the smaller Gram matrix
        u, s, vh, rest = gram_svd(mat.conj().T, stop_below, level_ratio, max_levels, need, tail_floor, info)
        return vh.conj().T, s, u.conj().T, rest
    basis = None          # right-singular subspace still to be resolved (n x k), None = everything
    us, ss, vs = [], [], []
    rest_norm = torch.zeros((), dtype=torch.float64, device=mat.device)
    resolved, sigma1 = 0, 0.0
    for level in range(max_levels):
        work = mat if basis is None else mat @ basis
        gram = work.conj().T @ work
        gram = 0.5 * (gram + gram.conj().T)
        lam, vec = torch.linalg.eigh(gram)
        lam, vec = lam.flip(0).clamp_min(0.0), vec.flip(1)
        sig = torch.sqrt(lam)
        # everything the host needs from this level in ONE read: the top value, how many values the level
        # accepts, and (from the Gram eigenvalues, i.e. only down to ~1e-8 sigma_1) the weight of
        # whatever lies beyond the first `need` values
        accept = sig >= level_ratio * sig[0]
        beyond_from = max((need if need is not None else 0) - resolved, 0)
        idx = torch.arange(sig.shape[0], device=mat.device)
        host_vals = torch.stack([sig[0], accept.sum().to(torch.float64), (lam * (~accept | (idx >= beyond_from))).sum()])
        top, keep_f, tail2_cheap = host_vals.tolist()                # the one host sync of a level
        if level == 0:
            sigma1 = top
        if level > 0 and top < stop_below:
            rest_norm = torch.linalg.vector_norm(work)
            break
        last = level == max_levels - 1
        keep = sig.shape[0] if (last or top == 0.0) else int(keep_f)
        v_here = vec if basis is None else basis @ vec
        b = work @ vec[:, :keep]
        s_here = torch.linalg.vector_norm(b, dim=0)           # more accurate than sqrt(lam) at the level's bottom
        safe = torch.where(s_here > 0, s_here, torch.ones_like(s_here))
        us.append(b / safe)
        ss.append(s_here)
        vs.append(v_here[:, :keep])
        resolved += keep
        if keep == sig.shape[0]:
            break
        basis = v_here[:, keep:]
        if need is not None and resolved >= need:
            if tail2_cheap > 0.0 and math.sqrt(tail2_cheap) >= max(tail_floor, 1e-6 * sigma1):
                # far above the resolution of the Gram eigenvalues: the truncation is decided
                rest_norm = torch.sqrt((lam[keep:]).sum())
                if info is not None:
                    info["decided"] = True
                    info["decided_level0"] = level == 0   # what gram_svd_at_cap reproduces without a host read
                break
            rest = torch.linalg.vector_norm(mat @ basis)      # accurate route (one more host sync)
            beyond = torch.cat(ss)[need:]
            if float(torch.sqrt((beyond * beyond).sum() + rest * rest)) >= tail_floor:
                rest_norm = rest
                if info is not None:
                    info["decided"] = True
                break
    u, s, v = torch.cat(us, dim=1), torch.cat(ss), torch.cat(vs, dim=1)
    order = torch.argsort(s, descending=True, stable=True)    # levels are ordered; ties inside noise only
    return u[:, order], s[order], v[:, order].conj().T, rest_norm


class _CusolverEigh:
    """Hermitian eigendecomposition WITHOUT a host synchronisation: cuSOLVER's ``zheevd`` called directly (ctypes, the
    library torch has already loaded) on the caller's stream, `devInfo` left on the device.  ``torch.linalg.eigh`` runs
    the same routine but reads `info` back after every call, which drains the stream twice per 2TDVP split (that and
    the truncation decision: 15 % of a chi = 256 time step was idle GPU).  Library code, like torch.linalg.eigh: the
    eigensolver is the one step of the 2TDVP split that is not native ("the split")."""
    _solver = None
    _handles = {}
    _work = {}

    @classmethod
    def _load(cls):
        if cls._solver is None:
            import torch  # noqa: F401  (loads libcusolver.so.11; dlopen by soname then returns that copy)
            lib = C.CDLL("libcusolver.so.11")
            lib.cusolverDnCreate.argtypes = [C.POINTER(C.c_void_p)]
            lib.cusolverDnSetStream.argtypes = [C.c_void_p, C.c_void_p]
            lib.cusolverDnZheevd_bufferSize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                        C.POINTER(C.c_int)]
            lib.cusolverDnZheevd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_int, C.c_void_p]
            for f in (lib.cusolverDnCreate, lib.cusolverDnSetStream, lib.cusolverDnZheevd_bufferSize, lib.cusolverDnZheevd):
                f.restype = C.c_int
            cls._solver = lib
        return cls._solver

    @classmethod
    def eigh(cls, gram):
        """gram: Hermitian complex128 CUDA matrix (n, n).  Returns (lam ascending (n,), vec (n, n) with eigenvectors in
        its columns, info (int32 device tensor, 0 = converged)).  Enqueues on torch's current stream, never synchronises."""
        import torch
        lib = cls._load()
        n = gram.shape[0]
        dev = gram.device
        key = (dev.index if dev.index is not None else torch.cuda.current_device())
        handle = cls._handles.get(key)
        if handle is None:
            handle = C.c_void_p()
            rc = lib.cusolverDnCreate(C.byref(handle))
            if rc != 0:
                raise _lib.QcaError(_lib.QCA_ERR_CUDA, f"cusolverDnCreate failed ({rc})")
            cls._handles[key] = handle
        stream = torch.cuda.current_stream(dev).cuda_stream
        lib.cusolverDnSetStream(handle, C.c_void_p(stream))
        # row-major Hermitian G is column-major conj(G): its eigenvectors are conj(v), written over the input in
        # column-major order, i.e. the ROWS of the row-major tensor -> vec = a^H
        a = _plain(gram).clone()
        lam = torch.empty(n, dtype=torch.float64, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        lwork = C.c_int()
        rc = lib.cusolverDnZheevd_bufferSize(handle, 1, 0, n, C.c_void_p(a.data_ptr()), n, C.c_void_p(lam.data_ptr()), C.byref(lwork))
        if rc != 0:
            raise _lib.QcaError(_lib.QCA_ERR_CUDA, f"cusolverDnZheevd_bufferSize failed ({rc})")
        work = cls._work.get(key)
        if work is None or work.numel() < lwork.value:
            work = torch.empty(max(lwork.value, 1), dtype=torch.complex128, device=dev)
            cls._work[key] = work
        rc = lib.cusolverDnZheevd(handle, 1, 0, n, C.c_void_p(a.data_ptr()), n, C.c_void_p(lam.data_ptr()),
                                  C.c_void_p(work.data_ptr()), lwork.value, C.c_void_p(info.data_ptr()))
        if rc != 0:
            raise _lib.QcaError(_lib.QCA_ERR_CUDA, f"cusolverDnZheevd failed ({rc})")
        return lam, a.conj().T, info


def gram_svd_at_cap(mat, need: int, tail_floor: float, level_ratio: float = 1e-3):
    """The `decided` branch of ``gram_svd`` (one Gram level, exactly `need` singular triplets kept because everything
    beyond them still weighs >= tail_floor) computed SPECULATIVELY, without any host synchronisation: returns
    (u, s, vh, ok) with `ok` a device bool that is true iff ``gram_svd`` would have taken that branch and returned these
    numbers.  The caller checks `ok` later (one read per time step) and repeats the step with ``gram_svd`` if it is false."""
    import torch
    m, n = mat.shape
    if m < n:
        u, s, vh, ok = gram_svd_at_cap(mat.conj().T, need, tail_floor, level_ratio)
        return vh.conj().T, s, u.conj().T, ok
    gram = mat.conj().T @ mat
    gram = 0.5 * (gram + gram.conj().T)
    lam, vec, info = _CusolverEigh.eigh(gram)
    lam, vec = lam.flip(0).clamp_min(0.0), vec.flip(1)
    sig = torch.sqrt(lam)
    accepted = (sig >= level_ratio * sig[0]).sum()
    tail2 = lam[need:].sum()
    floor = torch.clamp_min(1e-6 * sig[0], tail_floor)
    ok = (info[0] == 0) & (accepted >= need) & (tail2 > 0) & (torch.sqrt(tail2) >= floor)
    b = mat @ vec[:, :need]
    s = torch.linalg.vector_norm(b, dim=0)
    safe = torch.where(s > 0, s, torch.ones_like(s))
    return b / safe, s, vec[:, :need].conj().T, ok


_SM_COUNT = {}


def _sm_count(device) -> int:
    import torch
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SM_COUNT[idx]


def zgemm_batched(a, b, out_shape, M, N, K, S, G, a_strides, b_strides, conj_a=False):
    """C_g[m,n] = sum_s sum_k A_{g,s}[m,k] B_{g,s}[k,n] on the FP64 tensor cores (csrc/qca_zgemm.cu).
    a, b: complex128 CUDA tensors (only their storage is used); a_strides = (sg, ss, sm, sk),
    b_strides = (sg, ss, sk) in elements.  Returns C as a contiguous tensor of `out_shape`
    (G * M * N elements).  When the tile grid would leave most SMs idle the reduction is split over
    several CTAs per tile and the partial sums are added in a fixed order."""
    import torch
    tiles = -(-M // 64) * -(-N // 64) * G
    steps = S * -(-K // 16)
    nsplit = 1
    target = 2 * _sm_count(a.device)
    if tiles < target:
        nsplit = max(1, min(target // tiles, steps // 4, 16))
    out = torch.empty((nsplit,) + tuple(out_shape), dtype=a.dtype, device=a.device)
    stream = torch.cuda.current_stream(a.device).cuda_stream
    _lib.check(_lib.lib.qca_zgemm_batched(
        C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), M, N, K, S, G,
        *[int(x) for x in a_strides], *[int(x) for x in b_strides], M * N, N,
        int(conj_a), nsplit, G * M * N, C.c_void_p(stream)))
    return out[0] if nsplit == 1 else out.sum(dim=0)


def env_times_tensor(left, theta):
    """T[g, w, y, u] = sum_x left[x, w, y] * theta[g, x, u]   (first step of H_eff: L . theta)."""
    dx, w, dy = left.shape
    g, _, du = theta.shape
    left, theta = _plain(left), _plain(theta)
    return zgemm_batched(left, theta, (g, w, dy, du), M=w * dy, N=du, K=dx, S=1, G=g,
                         a_strides=(0, 0, 1, w * dy), b_strides=(dx * du, 0, du))


def tensor_times_env(t, right):
    """out[g, y, v] = sum_{n,u} t[g, n, y, u] * right[u, n, v]   (last step of H_eff: T . R)."""
    g, w, dy, du = t.shape
    _, _, dv = right.shape
    t, right = _plain(t), _plain(right)
    return zgemm_batched(t, right, (g, dy, dv), M=dy, N=dv, K=du, S=w, G=g,
                         a_strides=(w * dy * du, dy * du, du, 1), b_strides=(0, dv, w * dv))


# -- matrix-free H_eff and its Krylov exponential (csrc/qca_heff.cu) -------------------------------------
def site_operator_csr(w1, w2=None):
    """CSR form of the site operator(s) between the two environment contractions of H_eff, as NUMPY
    arrays (rowptr int32, col int32, val complex128) plus (g, wl, wr).

    one site  (tdvp.py:350-365):  Mx[(b, m), (a, w)]       = W[a, b, w, m]
    two sites (tdvp.py:280-283):  Mx[(b, d, n), (a, c, w)] = sum_m W1[a, b, w, m] W2[c, d, m, n]
    w1 is None: bond matrix (tdvp.py:312-327), identity on the w2 = (wl) channels given as an int."""
    import numpy as np
    if w1 is None:
        wl = wr = int(w2)
        dense = np.eye(wl, dtype=np.complex128)
        g = 1
    elif w2 is None:
        w1 = np.asarray(w1)
        wl, wr = w1.shape[2], w1.shape[3]
        dense = np.transpose(w1, (1, 3, 0, 2)).reshape(2 * wr, 2 * wl)
        g = 2
    else:
        w1, w2 = np.asarray(w1), np.asarray(w2)
        wl, wr = w1.shape[2], w2.shape[3]
        full = np.einsum("abwm,cdmn->bdnacw", w1, w2)
        dense = full.reshape(4 * wr, 4 * wl)
        g = 4
    rowptr = np.zeros(dense.shape[0] + 1, dtype=np.int32)
    cols, vals = [], []
    for r in range(dense.shape[0]):
        nz = np.nonzero(dense[r])[0]
        cols.append(nz.astype(np.int32))
        vals.append(dense[r, nz].astype(np.complex128))
        rowptr[r + 1] = rowptr[r] + nz.size
    col = np.concatenate(cols) if cols else np.zeros(0, np.int32)
    val = np.concatenate(vals) if vals else np.zeros(0, np.complex128)
    if col.size == 0:   # keep the device arrays non-empty
        col, val = np.zeros(1, np.int32), np.zeros(1, np.complex128)
    return rowptr, col, val, (g, wl, wr)


class SiteOperator:
    """Device copy of ``site_operator_csr`` (built once per site / bond: H is constant)."""

    def __init__(self, w1, w2=None, device=None):
        import torch
        rowptr, col, val, (self.g, self.wl, self.wr) = site_operator_csr(w1, w2)
        self.rowptr = torch.as_tensor(rowptr, device=device)
        self.col = torch.as_tensor(col, device=device)
        self.val = torch.as_tensor(val, device=device)
        # structural zeros: used columns (g, w) and non-empty rows (g', n), as bit masks per g
        nnz = int(rowptr[-1])
        self.col_mask, self.row_mask = [0] * 4, [0] * 4
        for c in col[:nnz]:
            self.col_mask[int(c) // self.wl] |= 1 << (int(c) % self.wl)
        for r in range(self.g * self.wr):
            if rowptr[r + 1] > rowptr[r]:
                self.row_mask[r // self.wr] |= 1 << (r % self.wr)
        self.use_masks = self.wl <= 32 and self.wr <= 32
        self.cols_used = sum(bin(m).count("1") for m in self.col_mask) if self.use_masks else self.g * self.wl
        self.rows_used = sum(bin(m).count("1") for m in self.row_mask) if self.use_masks else self.g * self.wr


class _Workspace:
    """One growing device buffer per device: H_eff calls are stream-ordered, so it can be shared."""
    _buf = {}

    @classmethod
    def get(cls, nbytes: int, device):
        import torch
        key = (device.type, device.index)
        buf = cls._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
            cls._buf[key] = buf
        return buf


def _heff_struct(left, right, op: SiteOperator):
    import torch
    assert left.is_cuda and left.dtype == torch.complex128 and right.dtype == torch.complex128
    dl, wl, dl2 = left.shape
    dr, wr, dr2 = right.shape
    assert dl == dl2 and dr == dr2 and wl == op.wl and wr == op.wr, (left.shape, right.shape, op.wl, op.wr)
    left, right = _plain(left), _plain(right)
    h = _lib.HeffStruct(left.data_ptr(), right.data_ptr(), op.rowptr.data_ptr(), op.col.data_ptr(), op.val.data_ptr(),
                        dl, dr, wl, wr, op.g, int(op.use_masks), (C.c_uint32 * 4)(*op.col_mask),
                        (C.c_uint32 * 4)(*op.row_mask))
    return h, (left, right)   # keep the contiguous copies alive until the launches are enqueued


def heff_apply(left, right, op: SiteOperator, psi):
    """H_eff psi (psi: (g, dl, dr) or any shape with g*dl*dr elements); returns a tensor like psi."""
    import torch
    h, keep = _heff_struct(left, right, op)
    psi_c = _plain(psi)
    assert psi_c.numel() == op.g * h.dl * h.dr
    nbytes = C.c_uint64()
    _lib.check(_lib.lib.qca_heff_workspace_bytes(C.byref(h), 0, C.byref(nbytes)))
    ws = _Workspace.get(nbytes.value, psi.device)
    out = torch.empty_like(psi_c)
    stream = torch.cuda.current_stream(psi.device).cuda_stream
    _lib.check(_lib.lib.qca_heff_apply(C.byref(h), C.c_void_p(psi_c.data_ptr()), C.c_void_p(out.data_ptr()),
                                       C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(stream)))
    return out


def env_grow(prev, site, op: SiteOperator):
    """Environment update (tdvp.py:329-347) on the DMMA kernel:
    out[r, m, s] = sum  W[a, b, w, m] prev[x, w, y] site[a, x, r] conj(site[b, y, s])
    with `op` the one-site operator of W.  A right environment is the same call on the mirrored tensors
    (site[a, u, l], W with its bond indices swapped)."""
    import torch
    assert prev.is_cuda and prev.dtype == torch.complex128 and site.dtype == torch.complex128
    dl, wl, dl2 = prev.shape
    g, dx, dr = site.shape
    assert g == 2 and op.g == 2 and dx == dl == dl2 and wl == op.wl, (prev.shape, site.shape, op.g, op.wl)
    prev, site = _plain(prev), _plain(site)
    h = _lib.HeffStruct(prev.data_ptr(), prev.data_ptr(), op.rowptr.data_ptr(), op.col.data_ptr(), op.val.data_ptr(),
                        dl, dr, wl, op.wr, 2, int(op.use_masks), (C.c_uint32 * 4)(*op.col_mask), (C.c_uint32 * 4)(*op.row_mask))
    nbytes = C.c_uint64()
    _lib.check(_lib.lib.qca_env_grow_workspace_bytes(C.byref(h), C.byref(nbytes)))
    ws = _Workspace.get(nbytes.value, site.device)
    out = torch.empty((dr, op.wr, dr), dtype=site.dtype, device=site.device)
    stream = torch.cuda.current_stream(site.device).cuda_stream
    _lib.check(_lib.lib.qca_env_grow(C.byref(h), C.c_void_p(site.data_ptr()), C.c_void_p(out.data_ptr()),
                                     C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(stream)))
    return out


def heff_expm(left, right, op: SiteOperator, psi, krylov_dim: int, t: float, spectral_bound: float = 0.0):
    """exp(-i t H_eff) psi by `krylov_dim` Lanczos steps, all on the device, no host synchronisation.
    spectral_bound: a bound of ||H_eff|| if known (the small exponential then is a Chebyshev series)."""
    import torch
    h, keep = _heff_struct(left, right, op)
    psi_c = _plain(psi)
    assert psi_c.numel() == op.g * h.dl * h.dr
    nbytes = C.c_uint64()
    _lib.check(_lib.lib.qca_heff_workspace_bytes(C.byref(h), int(krylov_dim), C.byref(nbytes)))
    ws = _Workspace.get(nbytes.value, psi.device)
    out = torch.empty_like(psi_c)
    stream = torch.cuda.current_stream(psi.device).cuda_stream
    _lib.check(_lib.lib.qca_heff_expm(C.byref(h), C.c_void_p(psi_c.data_ptr()), C.c_void_p(out.data_ptr()),
                                      int(krylov_dim), float(t), float(spectral_bound), C.c_void_p(ws.data_ptr()), ws.numel(),
                                      C.c_void_p(stream)))
    return out
