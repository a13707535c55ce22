"""One process per GPU: the state vector is sharded over log2(P) qubits (``_lib.plan_shard``): cells
{0, d+1, 2(d+1)} when the register is large enough -- every rank then pulls the same amount over
NVLink -- else the top cells; rank r holds the amplitudes whose sharded qubits spell r.

``torch.distributed`` is plumbing only: it moves the 64-byte CUDA-IPC handles once at start-up,
two booleans after a state upload and the 4*N measurement sums per measure (one all_gather of a tensor).  The data path (terms
that flip a sharded qubit) is peer-memory loads inside the tile-pass kernel plus a device-side flag
barrier (csrc/qca_exact.cu); no collective runs per step.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _dist():
    import torch.distributed as dist
    return dist


def world_and_rank(group=None) -> tuple[int, int]:
    try:
        dist = _dist()
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(group), dist.get_rank(group)
    except ImportError:
        pass
    return 1, 0


def gather_objects(obj, group=None) -> list:
    """Every rank's object, in rank order, on every rank."""
    world, _ = world_and_rank(group)
    if world == 1:
        return [obj]
    out = [None] * world
    _dist().all_gather_object(out, obj, group=group)
    return out


def gather_rows(vec: np.ndarray, device: int, group=None) -> np.ndarray:
    """Every rank's float64 vector as the rows of one array, on every rank: ONE all_gather of a device tensor
    (NCCL) or a host tensor (gloo) instead of a pickled all_gather_object (two collectives plus
    serialisation; it was a measurable part of a sharded step at 8 GPUs)."""
    world, _ = world_and_rank(group)
    vec = np.ascontiguousarray(vec, dtype=np.float64)
    if world == 1:
        return vec[None, :]
    import torch
    dist = _dist()
    on_gpu = dist.get_backend(group) == "nccl"
    src = torch.from_numpy(vec)
    if on_gpu:
        src = src.to(torch.device("cuda", device))
    out = torch.empty((world, vec.size), dtype=torch.float64, device=src.device)
    dist.all_gather(list(out.unbind(0)), src, group=group)   # (supported by gloo and NCCL alike)
    return out.cpu().numpy()


def local_indices(nbits_total: int, positions: list[int], rank: int) -> np.ndarray:
    """Global basis-state index of every local amplitude of `rank`: the local index with the sharded
    qubits (global bit positions `positions`, ascending, rank bit j <-> positions[j]) inserted."""
    x = np.arange(1 << (nbits_total - len(positions)), dtype=np.int64)
    for j, p in enumerate(positions):
        x = ((x >> p) << (p + 1)) | (x & ((1 << p) - 1)) | (((rank >> j) & 1) << p)
    return x


def local_slice(psi: np.ndarray, positions: list[int], rank: int) -> np.ndarray:
    """This rank's amplitudes of a full 2^N vector (a contiguous block when the top qubits are
    sharded, a strided gather when the sharded cells are spread out)."""
    nbits = int(psi.shape[0]).bit_length() - 1
    return np.ascontiguousarray(psi[local_indices(nbits, positions, rank)])


def assemble(slices: list[np.ndarray], positions: list[int]) -> np.ndarray:
    """Inverse of local_slice over all ranks."""
    nbits = int(sum(s.shape[0] for s in slices)).bit_length() - 1
    full = np.empty(1 << nbits, dtype=slices[0].dtype)
    for rank, part in enumerate(slices):
        full[local_indices(nbits, positions, rank)] = part
    return full


def combine_measurements(partials: list[np.ndarray], ncells: int):
    """Sum the per-rank partial sums in rank order (deterministic, identical on all ranks) and
    finish them into population / rounded population / entropy / bond dimensions."""
    total = np.zeros(4 * ncells)
    for part in partials:
        total += np.asarray(part, dtype=np.float64)
    return _lib.measure_finish(total, ncells)


class ShardedExactEngine:
    """Same surface as ``_lib.ExactEngine`` for a register sharded over the ranks of `group`."""

    def __init__(self, rules, device: int, flags: int = 0, stream: int | None = None, group=None):
        self.group = group
        self.world, self.rank = world_and_rank(group)
        self.ncells = int(rules.ncells)
        self.device = device
        self._eng = _lib.ExactEngine(rules, device=device, world_size=self.world, rank=self.rank, flags=flags,
                                     stream=stream)
        self.local_amps = self._eng.local_amps
        self.positions = _lib.plan_shard(rules, self.world)
        if self.world > 1:
            table = np.stack(gather_objects(self._eng.ipc_handles(), group))
            self._eng.ipc_import(table)
            # every rank must scale H by the same bound (they computed it independently)
            self._eng.set_spectral_bound(max(gather_objects(self._eng.stats()["spectral_bound"], group)))

    # -- state ---------------------------------------------------------------------------------
    def _resolve(self) -> None:
        if self.world == 1:
            return
        flags = gather_objects(self._eng.plane_flags(), self.group)
        self._eng.resolve_planes(any(f[0] for f in flags), any(f[1] for f in flags))

    def set_product_state(self, plist) -> None:
        self._eng.set_product_state(plist)
        self._resolve()

    def set_state(self, psi) -> None:
        """psi: the full 2^N vector (each rank takes its slice) or this rank's slice."""
        arr = np.asarray(psi).reshape(-1)
        if arr.size == self.local_amps * self.world and self.world > 1:
            arr = local_slice(arr, self.positions, self.rank)
        self._eng.set_state(arr)
        self._resolve()

    def get_local_state(self) -> np.ndarray:
        return self._eng.get_state()

    def get_state(self) -> np.ndarray:
        """The full vector on every rank (host gather; for small registers and tests)."""
        return assemble(gather_objects(self._eng.get_state(), self.group), self.positions)

    # -- evolution -------------------------------------------------------------------------------
    def step(self, step_size: float, nsteps: int = 1) -> None:
        self._eng.step(step_size, nsteps)

    def measure(self):
        if self.world == 1:
            return self._eng.measure()
        return combine_measurements(list(gather_rows(self._eng.measure_partial(), self.device, self.group)), self.ncells)

    def apply_h(self, vec) -> np.ndarray:
        arr = np.asarray(vec).reshape(-1)
        if arr.size == self.local_amps * self.world and self.world > 1:
            arr = local_slice(arr, self.positions, self.rank)
        return assemble(gather_objects(self._eng.apply_h(arr), self.group), self.positions)

    def norm2(self) -> float:
        return float(sum(gather_objects(self._eng.norm2(), self.group)))

    def stats(self) -> dict:
        return self._eng.stats()

    def reset_stats(self) -> None:
        self._eng.reset_stats()

    def close(self) -> None:
        self._eng.close()
