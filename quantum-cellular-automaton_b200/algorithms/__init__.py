from .algorithm import Algorithm
from .exact import Exact

__all__ = ["Algorithm", "Exact"]
