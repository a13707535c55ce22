"""ctypes binding of ``libqca_b200.so`` (C ABI in ``include/qca_b200.h``).

There is no fallback: if the shared library is missing the import fails with
build instructions, and every compute call raises ``QcaError`` when no sm_100
device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))


def library_path() -> str:
    # QCA_B200_LIBRARY: another build of the SAME library (A/B timing of kernel variants, scratch/variants/); never a fallback
    return os.environ.get("QCA_B200_LIBRARY") or os.path.join(_PKG, "libqca_b200.so")


class QcaError(RuntimeError):
    """A C-ABI call returned a QCA_ERR_* code."""

    def __init__(self, code: int, message: str):
        super().__init__(f"qca_b200 error {code}: {message}")
        self.code = code


QCA_OK, QCA_ERR_ARG, QCA_ERR_CUDA, QCA_ERR_NOMEM, QCA_ERR_STATE, QCA_ERR_UNSUPPORTED = range(6)
QCA_FLAG_FORCE_COMPLEX = 1
QCA_FLAG_PROFILE = 2
QCA_FLAG_LOOSE_BOUND = 4
QCA_FLAG_FUSED_MEASURE = 8
QCA_FLAG_PERCELL_MEASURE = 16
QCA_FLAG_TILE_PATH_ONLY = 32
QCA_FLAG_NO_GRAPH = 64
QCA_FLAG_V2_KERNELS = 128
QCA_FLAG_NO_PERSISTENT = 256
QCA_IPC_HANDLE_BYTES = 64


class RuleStruct(C.Structure):
    _fields_ = [("ncells", C.c_int32), ("distance", C.c_int32), ("act_lo", C.c_int32), ("act_hi", C.c_int32)]


class PassStruct(C.Structure):
    _fields_ = [("low_bits", C.c_int32), ("high_start", C.c_int32), ("high_bits", C.c_int32),
                ("reserved", C.c_int32), ("flip_mask", C.c_uint64)]


class RemoteOpStruct(C.Structure):
    _fields_ = [("pass_", C.c_int32), ("partner", C.c_int32), ("qubit", C.c_int32), ("sign", C.c_int32),
                ("mask", C.c_uint32), ("shift", C.c_int32), ("window_bits", C.c_int32), ("reserved", C.c_int32)]


class ExactStats(C.Structure):
    _fields_ = [("spectral_bound", C.c_double), ("planes", C.c_int32), ("passes_per_apply", C.c_int32),
                ("last_terms", C.c_int32), ("local_bits", C.c_int32), ("kernel_launches", C.c_uint64),
                ("pass_launches", C.c_uint64), ("pass_bytes", C.c_double), ("profiled_pass_ms", C.c_double),
                ("profiled_pass_launches", C.c_uint64), ("device_bytes", C.c_double), ("remote_bytes", C.c_double),
                ("profiled_ms_by_pass", C.c_double * 4)]

    def as_dict(self) -> dict:
        out = {name: getattr(self, name) for name, _ in self._fields_}
        out["profiled_ms_by_pass"] = list(self.profiled_ms_by_pass)
        return out


class RotationStruct(C.Structure):
    _fields_ = [("nslots", C.c_int32), ("npasses", C.c_int32), ("rot_shift", C.c_int32), ("rot_word", C.c_uint32),
                ("op_of", C.c_int32 * 32)]


class HeffStruct(C.Structure):
    _fields_ = [("left", C.c_void_p), ("right", C.c_void_p), ("mix_rowptr", C.c_void_p), ("mix_col", C.c_void_p),
                ("mix_val", C.c_void_p), ("dl", C.c_int32), ("dr", C.c_int32), ("wl", C.c_int32), ("wr", C.c_int32),
                ("g", C.c_int32), ("use_masks", C.c_int32), ("col_mask", C.c_uint32 * 4), ("row_mask", C.c_uint32 * 4)]


_dp = C.POINTER(C.c_double)

# every symbol include/qca_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "qca_version": (C.c_char_p, []),
    "qca_last_error": (C.c_char_p, []),
    "qca_device_count": (C.c_int32, []),
    "qca_spectral_bound": (C.c_int32, [C.POINTER(RuleStruct), _dp]),
    "qca_chebyshev_plan": (C.c_int32, [C.c_double, C.c_double, _dp, C.c_int32, C.POINTER(C.c_int32)]),
    "qca_plan_passes": (C.c_int32, [C.c_int32, C.POINTER(PassStruct), C.c_int32, C.POINTER(C.c_int32)]),
    "qca_plan_passes_v3": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(PassStruct), C.c_int32, C.POINTER(C.c_int32)]),
    "qca_plan_shard": (C.c_int32, [C.POINTER(RuleStruct), C.c_int32, C.POINTER(C.c_int32)]),
    "qca_plan_remote": (C.c_int32, [C.POINTER(RuleStruct), C.c_int32, C.c_int32, C.POINTER(RemoteOpStruct), C.c_int32,
                                    C.POINTER(C.c_int32)]),
    "qca_plan_rotation": (C.c_int32, [C.POINTER(RuleStruct), C.c_int32, C.c_int32, C.POINTER(RotationStruct)]),
    "qca_exact_plane_flags": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "qca_exact_resolve_planes": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32]),
    "qca_exact_create": (C.c_int32, [C.POINTER(C.c_void_p), C.POINTER(RuleStruct), C.c_int32, C.c_int32,
                                     C.c_int32, C.c_uint32, C.c_void_p]),
    "qca_exact_destroy": (C.c_int32, [C.c_void_p]),
    "qca_exact_local_amps": (C.c_uint64, [C.c_void_p]),
    "qca_exact_set_state": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "qca_exact_set_product_state": (C.c_int32, [C.c_void_p, _dp, C.c_int32]),
    "qca_exact_get_state": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "qca_exact_step": (C.c_int32, [C.c_void_p, C.c_double, C.c_int32]),
    "qca_exact_measure": (C.c_int32, [C.c_void_p, _dp, _dp, _dp, _dp]),
    "qca_exact_measure_partial": (C.c_int32, [C.c_void_p, _dp]),
    "qca_measure_finish": (C.c_int32, [_dp, C.c_int32, _dp, _dp, _dp, _dp]),
    "qca_exact_apply_h": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "qca_exact_norm2": (C.c_int32, [C.c_void_p, _dp]),
    "qca_exact_set_spectral_bound": (C.c_int32, [C.c_void_p, C.c_double]),
    "qca_exact_get_stats": (C.c_int32, [C.c_void_p, C.POINTER(ExactStats)]),
    "qca_exact_reset_stats": (C.c_int32, [C.c_void_p]),
    "qca_qr_householder": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_void_p]),
    "qca_zgemm_batched": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 5 + [C.c_int64] * 9
                          + [C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "qca_zgemm_profile": (C.c_int32, [C.c_int32, _dp, _dp, C.POINTER(C.c_uint64)]),
    "qca_heff_workspace_bytes": (C.c_int32, [C.POINTER(HeffStruct), C.c_int32, C.POINTER(C.c_uint64)]),
    "qca_heff_apply": (C.c_int32, [C.POINTER(HeffStruct), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qca_env_grow_workspace_bytes": (C.c_int32, [C.POINTER(HeffStruct), C.POINTER(C.c_uint64)]),
    "qca_env_grow": (C.c_int32, [C.POINTER(HeffStruct), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "qca_heff_expm": (C.c_int32, [C.POINTER(HeffStruct), C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double,
                                  C.c_void_p, C.c_uint64, C.c_void_p]),
    "qca_exact_ipc_count": (C.c_int32, [C.c_void_p]),
    "qca_exact_ipc_export": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p]),
    "qca_exact_ipc_import": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "qca_exact_loopback_peers": (C.c_int32, [C.c_void_p]),
}


def _load() -> C.CDLL:
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the CUDA extension is not built and there is no CPU fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C quantum-cellular-automaton_b200/csrc` (needs nvcc, targets sm_100a).")
    dll = C.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(dll, name)  # AttributeError if the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    return dll


lib = _load()


def check(code: int) -> None:
    if code != QCA_OK:
        raise QcaError(code, lib.qca_last_error().decode("utf-8", "replace"))


def as_double_ptr(arr: np.ndarray):
    assert arr.dtype == np.float64 and arr.flags.c_contiguous
    return arr.ctypes.data_as(_dp)


def rule_struct(rules) -> RuleStruct:
    iv = rules.activation_interval
    return RuleStruct(int(rules.ncells), int(rules.distance), int(iv.start), int(iv.stop))


# ---------------------------------------------------------------------------
# Host-only planning helpers
# ---------------------------------------------------------------------------
def spectral_bound(rules) -> float:
    out = C.c_double()
    rs = rule_struct(rules)
    check(lib.qca_spectral_bound(C.byref(rs), C.byref(out)))
    return out.value


def chebyshev_plan(z: float, tol: float = 1e-15) -> np.ndarray:
    n = C.c_int32()
    check(lib.qca_chebyshev_plan(z, tol, None, 0, C.byref(n)))
    a = np.empty(n.value)
    check(lib.qca_chebyshev_plan(z, tol, as_double_ptr(a), n.value, C.byref(n)))
    return a


def plan_passes(local_bits: int) -> list[dict]:
    n = C.c_int32()
    check(lib.qca_plan_passes(local_bits, None, 0, C.byref(n)))
    buf = (PassStruct * n.value)()
    check(lib.qca_plan_passes(local_bits, buf, n.value, C.byref(n)))
    return [dict(low_bits=p.low_bits, high_start=p.high_start, high_bits=p.high_bits, flip_mask=p.flip_mask)
            for p in buf]


def plan_passes_v3(local_bits: int, max_cluster_bits: int = 3, min_low: int = 4) -> list[dict]:
    """Plan of the cluster tile-pass kernels (qca_plan_passes_v3); cluster_bits = the pass's `reserved` field."""
    n = C.c_int32()
    check(lib.qca_plan_passes_v3(local_bits, max_cluster_bits, min_low, None, 0, C.byref(n)))
    buf = (PassStruct * max(n.value, 1))()
    check(lib.qca_plan_passes_v3(local_bits, max_cluster_bits, min_low, buf, n.value, C.byref(n)))
    return [dict(low_bits=p.low_bits, high_start=p.high_start, high_bits=p.high_bits, cluster_bits=p.reserved,
                 flip_mask=p.flip_mask) for p in buf[:n.value]]


def plan_remote(rules, world_size: int, rank: int) -> list[dict]:
    n = C.c_int32()
    rs = rule_struct(rules)
    check(lib.qca_plan_remote(C.byref(rs), world_size, rank, None, 0, C.byref(n)))
    buf = (RemoteOpStruct * max(n.value, 1))()
    check(lib.qca_plan_remote(C.byref(rs), world_size, rank, buf, max(n.value, 1), C.byref(n)))
    return [dict(pass_index=o.pass_, partner=o.partner, qubit=o.qubit, sign=o.sign, mask=o.mask, shift=o.shift,
                 window_bits=o.window_bits) for o in buf[:n.value]]


def zgemm_profile(enable: bool) -> dict:
    """Collect (and reset) the per-launch timings of the DMMA contraction kernel; see qca_zgemm_profile."""
    ms, flops, n = C.c_double(), C.c_double(), C.c_uint64()
    check(lib.qca_zgemm_profile(int(enable), C.byref(ms), C.byref(flops), C.byref(n)))
    return {"ms": ms.value, "flops": flops.value, "launches": n.value}


def plan_rotation(rules, world_size: int, rank: int) -> dict:
    """How the fast kernel rotates the remote terms over the passes (qca_plan_rotation)."""
    out = RotationStruct()
    rs = rule_struct(rules)
    check(lib.qca_plan_rotation(C.byref(rs), world_size, rank, C.byref(out)))
    op_of = np.array(list(out.op_of), dtype=np.int64).reshape(4, 2, 4)
    return dict(nslots=out.nslots, npasses=out.npasses, rot_shift=out.rot_shift, rot_word=out.rot_word, op_of=op_of)


def plan_shard(rules, world_size: int) -> list[int]:
    """Global index-bit positions of the sharded qubits, ascending (rank bit j <-> positions[j])."""
    nbits = world_size.bit_length() - 1
    buf = (C.c_int32 * max(nbits, 1))()
    rs = rule_struct(rules)
    check(lib.qca_plan_shard(C.byref(rs), world_size, buf))
    return [int(buf[j]) for j in range(nbits)]


def measure_finish(sums: np.ndarray, ncells: int):
    sums = np.ascontiguousarray(sums, dtype=np.float64)
    pop, dpop, ent = np.empty(ncells), np.empty(ncells), np.empty(ncells)
    bonds = np.empty(ncells + 1)
    check(lib.qca_measure_finish(as_double_ptr(sums), ncells, as_double_ptr(pop), as_double_ptr(dpop),
                                 as_double_ptr(ent), as_double_ptr(bonds)))
    return pop, dpop, ent, bonds


# ---------------------------------------------------------------------------
# Exact engine handle
# ---------------------------------------------------------------------------
class ExactEngine:
    """Thin owner of a ``qca_exact_t`` (see include/qca_b200.h)."""

    def __init__(self, rules, device: int = 0, world_size: int = 1, rank: int = 0, flags: int = 0,
                 stream: int | None = None):
        self._h = C.c_void_p()
        self.ncells = int(rules.ncells)
        rs = rule_struct(rules)
        check(lib.qca_exact_create(C.byref(self._h), C.byref(rs), device, world_size, rank, flags,
                                   C.c_void_p(stream) if stream else None))
        self.local_amps = int(lib.qca_exact_local_amps(self._h))

    # -- sharding -----------------------------------------------------------------------------
    def ipc_handles(self) -> np.ndarray:
        """This rank's exported buffers, uint8[count, 64] (empty when not sharded)."""
        count = lib.qca_exact_ipc_count(self._h)
        out = np.zeros((count, QCA_IPC_HANDLE_BYTES), dtype=np.uint8)
        for i in range(count):
            check(lib.qca_exact_ipc_export(self._h, i, C.c_void_p(out[i].ctypes.data)))
        return out

    def ipc_import(self, table: np.ndarray) -> None:
        """table: uint8[world, count, 64] gathered from all ranks."""
        table = np.ascontiguousarray(table, dtype=np.uint8)
        check(lib.qca_exact_ipc_import(self._h, C.c_void_p(table.ctypes.data), table.shape[0], table.shape[1]))

    def loopback_peers(self) -> None:
        """Profiling aid: see qca_exact_loopback_peers."""
        check(lib.qca_exact_loopback_peers(self._h))

    def plane_flags(self) -> tuple[bool, bool]:
        re, im = C.c_int32(), C.c_int32()
        check(lib.qca_exact_plane_flags(self._h, C.byref(re), C.byref(im)))
        return bool(re.value), bool(im.value)

    def resolve_planes(self, has_re: bool, has_im: bool) -> None:
        check(lib.qca_exact_resolve_planes(self._h, int(has_re), int(has_im)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            lib.qca_exact_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _host_c128(buf) -> np.ndarray:
        arr = np.ascontiguousarray(buf, dtype=np.complex128)
        return arr

    def set_state(self, psi) -> None:
        arr = self._host_c128(psi).reshape(-1)
        check(lib.qca_exact_set_state(self._h, C.c_void_p(arr.ctypes.data), arr.size))

    def set_state_ptr(self, ptr: int, namps: int) -> None:
        """Upload from a raw host pointer (e.g. pinned memory), interleaved complex128."""
        check(lib.qca_exact_set_state(self._h, C.c_void_p(ptr), namps))

    def set_product_state(self, plist) -> None:
        p = np.ascontiguousarray(plist, dtype=np.float64)
        check(lib.qca_exact_set_product_state(self._h, as_double_ptr(p), p.size))

    def get_state(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.local_amps, dtype=np.complex128)
        assert out.dtype == np.complex128 and out.flags.c_contiguous and out.size == self.local_amps
        check(lib.qca_exact_get_state(self._h, C.c_void_p(out.ctypes.data), out.size))
        return out

    def get_state_ptr(self, ptr: int, namps: int) -> None:
        check(lib.qca_exact_get_state(self._h, C.c_void_p(ptr), namps))

    def step(self, step_size: float, nsteps: int = 1) -> None:
        check(lib.qca_exact_step(self._h, float(step_size), int(nsteps)))

    def measure(self):
        n = self.ncells
        pop, dpop, ent, bonds = np.empty(n), np.empty(n), np.empty(n), np.empty(n + 1)
        check(lib.qca_exact_measure(self._h, as_double_ptr(pop), as_double_ptr(dpop), as_double_ptr(ent),
                                    as_double_ptr(bonds)))
        return pop, dpop, ent, bonds

    def measure_partial(self) -> np.ndarray:
        sums = np.empty(4 * self.ncells)
        check(lib.qca_exact_measure_partial(self._h, as_double_ptr(sums)))
        return sums

    def apply_h(self, vec) -> np.ndarray:
        arr = self._host_c128(vec).reshape(-1)
        out = np.empty_like(arr)
        check(lib.qca_exact_apply_h(self._h, C.c_void_p(arr.ctypes.data), C.c_void_p(out.ctypes.data), arr.size))
        return out

    def norm2(self) -> float:
        v = C.c_double()
        check(lib.qca_exact_norm2(self._h, C.byref(v)))
        return v.value

    def set_spectral_bound(self, bound: float) -> None:
        check(lib.qca_exact_set_spectral_bound(self._h, float(bound)))

    def stats(self) -> dict:
        st = ExactStats()
        check(lib.qca_exact_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def reset_stats(self) -> None:
        check(lib.qca_exact_reset_stats(self._h))
