"""B200-native time-evolution path of the quantum cellular automaton simulator.

Host-side mirror of the reference's operator interface for this path
(``parameters.Rules``, ``tensor_networks.MPS/MPO``, ``algorithms.Algorithm/Exact/TDVP``)
driving the C-ABI CUDA library ``libqca_b200.so`` (``include/qca_b200.h``).
Import as ``qca_b200`` (see ``qca_b200.py`` at the repository root).
"""
from .parameters import Rules, Args
from .tensor_networks import MPS, MPO
from .algorithms import Algorithm, Exact, TDVP
from . import states
from ._lib import lib, QcaError, library_path

__all__ = ["Rules", "Args", "MPS", "MPO", "Algorithm", "Exact", "TDVP", "states", "lib", "QcaError", "library_path"]
