from .algorithm import Algorithm
from .exact import Exact
from .tdvp import TDVP

__all__ = ["Algorithm", "Exact", "TDVP"]
