from .mpo import MPO
from .mps import MPS

__all__ = ["MPO", "MPS"]
