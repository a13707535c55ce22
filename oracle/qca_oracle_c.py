"""ctypes front end of ``qca_oracle_c.c`` -- the matrix-free C restatement of the reference's exact
path for registers beyond the reach of its dense matrices.

TEST INFRASTRUCTURE ONLY (same rule as ``qca_oracle.py``): imported by ``tests/``,
``__graft_entry__.smoke()`` and the CPU legs of ``bench.py``; never by the product package.
Built by ``make -C oracle`` (``__graft_entry__.build()`` does that).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libqca_oracle.so")
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise RuntimeError(f"{_PATH} is missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(_PATH)
        dp = C.POINTER(C.c_double)
        _lib.qo_set_threads.argtypes = [C.c_int]
        _lib.qo_apply_h.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.qo_apply_h_rows.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64]
        _lib.qo_step.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
        _lib.qo_measure.argtypes = [dp, C.c_int, dp, dp]
        _lib.qo_product_state.argtypes = [dp, C.c_int, dp]
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def set_threads(n: int = 0) -> int:
    """Pin the OpenMP thread count (0: leave it) and return the count in effect.  torchrun exports
    OMP_NUM_THREADS=1; callers that want all cores pass len(os.sched_getaffinity(0))."""
    return int(lib().qo_set_threads(int(n)))


def apply_h(v: np.ndarray, ncells: int, distance: int, lo: int, hi: int) -> np.ndarray:
    v = np.ascontiguousarray(v, dtype=np.complex128)
    out = np.empty_like(v)
    assert v.size == 1 << ncells
    if lib().qo_apply_h(_ptr(v), _ptr(out), ncells, distance, lo, hi):
        raise ValueError("qo_apply_h: bad arguments")
    return out


def apply_h_rows(v: np.ndarray, ncells: int, distance: int, lo: int, hi: int, x0: int, count: int,
                 out: np.ndarray | None = None) -> np.ndarray:
    """Rows [x0, x0 + count) of H @ v (bounded timing sample of one application)."""
    assert v.dtype == np.complex128 and v.flags.c_contiguous and v.size == 1 << ncells
    if out is None:
        out = np.empty(count, dtype=np.complex128)
    if lib().qo_apply_h_rows(_ptr(v), _ptr(out), ncells, distance, lo, hi, int(x0), int(count)):
        raise ValueError("qo_apply_h_rows: bad arguments")
    return out


def product_state(ncells: int, plist) -> np.ndarray:
    psi = np.empty(1 << ncells, dtype=np.complex128)
    p = np.ascontiguousarray(plist, dtype=np.float64)
    lib().qo_product_state(_ptr(psi), ncells, _ptr(p))
    return psi


def chebyshev_terms(ncells: int, step_size: float) -> int:
    """Number of terms qo_step sums for this register and step (one application of H each)."""
    z = ncells * abs(np.pi / 2 * step_size)
    # same truncation rule as qo_step, evaluated through scipy's Bessel functions
    from scipy.special import jv
    kmax = int(z + 14.0 * np.cbrt(z + 1.0) + 40.0)
    j = np.abs(jv(np.arange(kmax + 1), z))
    n, tail = kmax + 1, 0.0
    while n > 2 and tail + 2.0 * j[n - 1] < 1e-17:
        tail += 2.0 * j[n - 1]
        n -= 1
    return n


class Stepper:
    """State + work space of the C oracle; ``step`` is exact.py:26-27, ``measure`` mps.py:100-140."""

    def __init__(self, ncells: int, distance: int, lo: int, hi: int):
        self.n, self.d, self.lo, self.hi = ncells, distance, lo, hi
        self.psi = np.zeros(1 << ncells, dtype=np.complex128)
        self.work = np.empty(2 << ncells, dtype=np.complex128)
        self.terms = 0

    def set_product_state(self, plist) -> None:
        p = np.ascontiguousarray(plist, dtype=np.float64)
        assert p.size == self.n
        lib().qo_product_state(_ptr(self.psi), self.n, _ptr(p))

    def set_state(self, psi) -> None:
        self.psi[:] = psi

    def step(self, step_size: float, first_terms: int = 0) -> int:
        self.terms = int(lib().qo_step(_ptr(self.psi), _ptr(self.work), self.n, self.d, self.lo, self.hi,
                                       float(step_size), int(first_terms)))
        return self.terms

    def measure(self):
        pop, ent = np.zeros(self.n), np.zeros(self.n)
        lib().qo_measure(_ptr(self.psi), self.n, _ptr(pop), _ptr(ent))
        return pop, ent
