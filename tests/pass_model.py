"""numpy model of what the CUDA tile-pass kernel computes, driven by the plan the C library
returns.  Used by the CPU tests to prove the host-side planning (tile geometry, flip masks,
Clenshaw coefficients) without a GPU, and by the GPU tests as an independent mid-size check.
Test infrastructure only."""
import numpy as np


def activity(xs: np.ndarray, ncells: int, distance: int, lo: int, hi: int) -> np.ndarray:
    """bit g of the result: cell (ncells-1-g) has its alive-neighbour count in [lo,hi)."""
    act = np.zeros_like(xs)
    for g in range(ncells):
        cnt = np.zeros(xs.shape, dtype=np.int64)
        for o in range(1, distance + 1):
            for nb in (g - o, g + o):
                if 0 <= nb < ncells:
                    cnt += (xs >> nb) & 1
        act |= (((cnt >= lo) & (cnt < hi)).astype(xs.dtype)) << g
    return act


def tile_indices(ps: dict, tile: int) -> np.ndarray:
    """global index of every element y of tile `tile`, with the kernel's formulas."""
    L, H0, M = ps["low_bits"], ps["high_start"], ps["high_bits"]
    gap = H0 - L
    t_lo, t_hi = tile & ((1 << gap) - 1), tile >> gap
    base = (t_lo << L) | (t_hi << (H0 + M))
    y = np.arange(1 << (L + M), dtype=np.int64)
    return base | (y & ((1 << L) - 1)) | ((y >> L) << H0)


def expand(xs: np.ndarray, positions, rank: int) -> np.ndarray:
    """local index -> global index: insert the sharded qubits (ascending global positions)."""
    x = xs.copy()
    for j, p in enumerate(positions):
        x = ((x >> p) << (p + 1)) | (x & ((1 << p) - 1)) | (((rank >> j) & 1) << p)
    return x


def global_pos(g: int, positions) -> int:
    for p in positions:
        if g >= p:
            g += 1
    return g


def apply_k_by_tiles(vec: np.ndarray, passes: list, nbits: int, ncells: int, distance: int, lo: int, hi: int,
                     positions=(), rank: int = 0) -> np.ndarray:
    """K vec, tile by tile exactly as the kernel stages it (real vector, one plane).  `positions`,
    `rank`: sharded register (the vector is this rank's local part)."""
    out = np.zeros_like(vec)
    for ps in passes:
        L, H0, M = ps["low_bits"], ps["high_start"], ps["high_bits"]
        T = L + M
        for tile in range(1 << (nbits - T)):
            xs = tile_indices(ps, tile)
            stage = vec[xs]
            act = activity(expand(xs, positions, rank), ncells, distance, lo, hi)
            y = np.arange(1 << T)
            acc = np.zeros(1 << T)
            for q in range(T):
                g = q if q < L else H0 + (q - L)
                if not (ps["flip_mask"] >> g) & 1:
                    continue
                on = ((act >> global_pos(g, positions)) & 1).astype(bool)
                partner = stage[y ^ (1 << q)]
                sign = np.where((y >> q) & 1, -1.0, 1.0)
                acc += np.where(on, sign * partner, 0.0)
            out[xs] += acc
    return out


def k_matrix(ncells: int, distance: int, lo: int, hi: int) -> np.ndarray:
    """Dense K with H = D (iK) D^-1, D = diag(i^popcount)."""
    dim = 1 << ncells
    xs = np.arange(dim, dtype=np.int64)
    act = activity(xs, ncells, distance, lo, hi)
    k = np.zeros((dim, dim))
    for g in range(ncells):
        on = ((act >> g) & 1).astype(bool)
        sign = np.where((xs >> g) & 1, -1.0, 1.0)
        k[xs[on], xs[on] ^ (1 << g)] += sign[on]
    return k


def popcount(xs: np.ndarray) -> np.ndarray:
    c = np.zeros_like(xs)
    v = xs.copy()
    while v.any():
        c += v & 1
        v >>= 1
    return c


def clenshaw_exp(kmat: np.ndarray, phi: np.ndarray, a: np.ndarray, bound: float, sign: float = 1.0):
    """phi' = sum_k a_k U_k(K/R) phi with the library's recurrence (qca_exact.cu step_once)."""
    K = len(a) - 1
    b1 = a[K] * phi
    b2 = np.zeros_like(phi)
    for k in range(K - 1, 0, -1):
        b1, b2 = a[k] * phi + sign * (2.0 / bound) * (kmat @ b1) + b2, b1
    return a[0] * phi + sign * (1.0 / bound) * (kmat @ b1) + b2
