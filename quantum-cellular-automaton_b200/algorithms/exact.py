"""``Exact``: full state-vector evolution on the GPU, drop-in for algorithms/exact.py:9-27.

The reference builds the dense 2^N x 2^N Hamiltonian (``MPO.as_matrix``), exponentiates it by
an eigendecomposition (``calculate_U``) and multiplies.  Here the state lives on the device and
one step is a Clenshaw-Chebyshev sum of matrix-free applications of the rule operator
(``csrc/qca_exact.cu``); nothing of size 4^N ever exists.
"""
from __future__ import annotations

import numpy as np

from .algorithm import Algorithm
from .. import _lib, sharding
from ..tensor_networks import MPS, MPO


def _product_plist(mps, ncells: int):
    """Alive probabilities of a bond-dimension-1 MPS whose tensors are real, non-negative amplitudes
    (sqrt(1-p), sqrt(p)) -- what MPS.from_density_distribution (mps.py:35-52) builds -- read from the
    tensors THEMSELVES (``MPS.A`` is a public list that callers may modify after construction; a cached
    ``plist`` could be stale).  None for anything else: the caller then merges the tensors on the host."""
    tensors = getattr(mps, "A", None)
    if tensors is None or len(tensors) != ncells:
        return None
    plist = []
    for t in tensors:
        t = np.asarray(t)
        if t.shape != (2, 1, 1):
            return None
        a0, a1 = complex(t[0, 0, 0]), complex(t[1, 0, 0])
        if a0.imag != 0.0 or a1.imag != 0.0 or a0.real < 0.0 or a1.real < 0.0:
            return None
        if abs(a0.real ** 2 + a1.real ** 2 - 1.0) > 1e-15:
            return None
        plist.append(min(1.0, a1.real ** 2) if a1.real ** 2 >= 0.5 else 1.0 - min(1.0, a0.real ** 2))
    return plist


class Exact(Algorithm):

    def __init__(self, psi_0: MPS, H: MPO, args, *, device: int = 0, stream: int | None = None,
                 force_complex: bool = False, profile: bool = False, group=None) -> None:
        """Extra keyword arguments (all optional): CUDA `device` index, `stream` handle to launch
        on, `force_complex` to keep both real planes, `profile` for per-launch event timing, and a
        torch.distributed `group`: when torch.distributed is initialised with more than one rank
        the register is sharded over the ranks (one process per GPU)."""
        rules = args.rules
        if H is not None:
            expected = MPO.hamiltonian_from_rules(rules)
            theirs = H if isinstance(H, MPO) else MPO(list(H.W))
            if not expected.same_operator_as(theirs):
                raise ValueError(
                    "Exact evaluates the rule Hamiltonian MPO.hamiltonian_from_rules(args.rules) "
                    "matrix-free; the MPO passed in is a different operator")
        flags = (_lib.QCA_FLAG_FORCE_COMPLEX if force_complex else 0) | (_lib.QCA_FLAG_PROFILE if profile else 0)
        if sharding.world_and_rank(group)[0] > 1:
            self._engine = sharding.ShardedExactEngine(rules, device=device, flags=flags, stream=stream, group=group)
        else:
            self._engine = _lib.ExactEngine(rules, device=device, flags=flags, stream=stream)
        super().__init__(psi_0, H, args)

    # -- Algorithm interface --------------------------------------------------------------
    @property
    def psi(self) -> MPS:
        """exact.py:19-20: the state as an (exact) MPS.  Downloads the vector."""
        return MPS.from_vector(self.state_vector())

    @psi.setter
    def psi(self, value: MPS) -> None:
        """exact.py:22-24."""
        plist = _product_plist(value, self._engine.ncells)
        if plist is not None:
            self._engine.set_product_state(plist)  # built on the device, no 2^N host vector
        else:
            self._engine.set_state(value.as_vector())

    def do_time_step(self) -> None:
        """exact.py:26-27 with U = exp(-i pi/2 step_size H) (lautils.py:45-55)."""
        self._engine.step(self.args.step_size, 1)

    def measure(self, population, d_population, single_site_entropy, bond_dims) -> None:
        """algorithm.py:58-63 -> mps.py:100-140, as fused reductions over the resident state."""
        pop, dpop, ent, bonds = self._engine.measure()
        population[...] = pop
        d_population[...] = dpop
        single_site_entropy[...] = ent
        bond_dims[...] = bonds

    # -- extras ------------------------------------------------------------------------------
    def state_vector(self) -> np.ndarray:
        """The 2^N complex128 state (``Exact._psi`` of the reference)."""
        return self._engine.get_state()

    def set_state_vector(self, psi) -> None:
        self._engine.set_state(psi)

    def do_time_steps(self, nsteps: int) -> None:
        self._engine.step(self.args.step_size, nsteps)

    @property
    def engine(self) -> "_lib.ExactEngine":
        return self._engine
