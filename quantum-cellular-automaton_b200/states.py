"""Named initial states of the reference (states.py:11-88) as alive-probability lists."""
from __future__ import annotations

from random import random

import numpy as np

from .tensor_networks import MPS

NAMES = ["blinker", "triple_blinker", "full_blinker", "single", "single_bottom", "all_ket_0", "all_ket_1",
         "only_outer", "all_ket_1_but_outer", "equal_superposition", "equal_superposition_but_outer",
         "gradient", "rand"]


def plist(name: str, rules) -> list[float]:
    n, d = rules.ncells, rules.distance
    mid = int(n / 2)
    p = [0.] * n
    if name == "blinker":
        p[mid - 1] = p[mid + 1] = 1.
    elif name == "triple_blinker":
        p[mid - 2] = p[mid] = p[mid + 2] = 1.
    elif name == "full_blinker":
        p = [float(i & 1) for i in range(n)]
    elif name == "single":
        p[mid] = 1.
    elif name == "single_bottom":
        p[0] = 1.
    elif name == "all_ket_0":
        pass
    elif name == "all_ket_1":
        p = [1.] * n
    elif name == "only_outer":
        p[0] = p[-1] = 1.
    elif name == "all_ket_1_but_outer":
        p = [0.] * d + [1.] * (n - 2 * d) + [0.] * d
    elif name == "equal_superposition":
        p = [.5] * n
    elif name == "equal_superposition_but_outer":
        p = [0.] * d + [.5] * (n - 2 * d) + [0.] * d
    elif name == "gradient":
        p = [float(np.sin(np.pi * i / (n - 1) / 2)) for i in range(n)]
    elif name == "rand":
        p = [0. if random() > .5 else 1. for _ in range(n)]
    else:
        raise ValueError(f"unknown initial state {name!r}; choose from {NAMES}")
    return p


def make(name: str, rules) -> MPS:
    return MPS.from_density_distribution(plist(name, rules))
