"""The reference's ``Algorithm`` plug-in interface (algorithms/algorithm.py:10-63):
``__init__(psi_0, H, args)``, ``do_time_step()``, the ``psi`` property and
``measure(population, d_population, single_site_entropy, bond_dims)``."""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

from ..tensor_networks import MPS, MPO


class Algorithm(ABC):
    _H: MPO

    @abstractmethod
    def __init__(self, psi_0: MPS, H: MPO, args) -> None:
        self._H = H
        self.args = args
        self.psi = psi_0

    @abstractmethod
    def do_time_step(self) -> None:
        ...

    @property
    @abstractmethod
    def psi(self) -> MPS:
        ...

    @psi.setter
    @abstractmethod
    def psi(self, value: MPS) -> None:
        ...

    @abstractmethod
    def measure(self, population, d_population, single_site_entropy, bond_dims) -> None:
        """Fill the four output rows in place (algorithm.py:58-63)."""

    @classmethod
    def classical_evolution(cls, first_column: np.ndarray, rules, plot_steps: int) -> np.ndarray:
        """Classical comparison heat map (algorithm.py:34-56).  Host-side plotting aid, kept
        for interface completeness; border columns are seeded with their initial value for all
        steps before the sweep, exactly as the reference does."""
        ncells = len(first_column)
        heat = np.zeros([plot_steps, ncells])
        heat[0, :] = first_column
        heat[:, 0] = first_column[0]
        heat[:, -1] = first_column[-1]
        window = range(-rules.distance, rules.distance + 1)
        for step in range(1, plot_steps):
            before = heat[step - 1]
            for site in range(ncells):
                alive = sum(before[site + o] for o in window if o != 0 and 0 <= site + o < ncells)
                flip = alive in rules.activation_interval
                heat[step, site] = 1. - before[site] if flip else before[site]
        return heat
