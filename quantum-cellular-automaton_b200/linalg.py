"""Device linear algebra with the reference's (numpy/LAPACK) conventions."""
from __future__ import annotations

import ctypes as C

from . import _lib


def householder_qr(mat, complete: bool = False):
    """QR of a complex128 CUDA matrix with LAPACK's Householder sign convention -- the factors
    numpy.linalg.qr(mat, mode='reduced' | 'complete') returns (csrc/qca_linalg.cu).  Returns (q, r)
    as torch tensors on the same device."""
    import torch
    assert mat.is_cuda and mat.dtype == torch.complex128 and mat.dim() == 2
    m, n = mat.shape
    kq = m if complete else min(m, n)
    a = mat.T.clone(memory_format=torch.contiguous_format)     # row-major (n, m) == column-major (m, n)
    tau = torch.empty(min(m, n), dtype=mat.dtype, device=mat.device)
    q = torch.empty((kq, m), dtype=mat.dtype, device=mat.device)   # column-major m x kq
    r = torch.empty((n, kq), dtype=mat.dtype, device=mat.device)   # column-major kq x n
    stream = torch.cuda.current_stream(mat.device).cuda_stream
    _lib.check(_lib.lib.qca_qr_householder(C.c_void_p(a.data_ptr()), m, n, C.c_void_p(tau.data_ptr()),
                                           C.c_void_p(q.data_ptr()), kq, C.c_void_p(r.data_ptr()),
                                           C.c_void_p(stream)))
    return q.T, r.T
