"""Host container for the rule Hamiltonian as a matrix product operator.

Same tensors, index order (``W[i].shape == (2, 2, Dl, Dr)``) and bond numbering as the
reference's ``MPO.hamiltonian_from_rules`` (tensor_networks/mpo.py:161-202), so a reference
``MPO`` and this one are interchangeable.  The exact path never contracts these tensors (the
CUDA kernel evaluates the rule directly); they are the operand format of the TDVP path and the
means by which ``Exact`` verifies that the Hamiltonian it was handed is the rule Hamiltonian.
"""
from __future__ import annotations

import numpy as np

_OPS = {
    "I": np.eye(2),
    "P0": np.array([[1., 0.], [0., 0.]]),
    "P1": np.array([[0., 0.], [0., 1.]]),
    "X": np.array([[0., 1.], [1., 0.]]),
}


def rule_automaton(distance: int, lo: int, hi: int) -> list[list[tuple[str, int]]]:
    """Transition lists of the operator-string automaton (mpo.py:56-151).

    State = (projectors emitted, |1>-projectors among them, sigma-x emitted).  A string is
    ``distance`` projectors, sigma-x, ``distance`` projectors, with the number of
    |1>-projectors inside [lo, hi).  States are numbered in depth-first discovery order,
    |1> branch first, final state last -- the reference's numbering.
    """
    total = 2 * distance
    number: dict[tuple[int, int, bool], int] = {}
    table: list[list[tuple[str, int]]] = []
    FINAL = -1

    def explore(nops: int, ones: int, flipped: bool) -> int:
        key = (nops, ones, flipped)
        if key in number:
            return number[key]
        if nops == total and flipped:
            return FINAL
        number[key] = me = len(table)
        table.append([])
        if nops == distance and not flipped:
            table[me].append(("X", explore(nops, ones, True)))
            return me
        if nops == 0 and not flipped:
            table[me].append(("I", me))
        if ones + 1 < hi:
            table[me].append(("P1", explore(nops + 1, ones + 1, flipped)))
        if ones + (total - nops - 1) >= lo:
            table[me].append(("P0", explore(nops + 1, ones, flipped)))
        return me

    explore(0, 0, False)
    table.append([("I", FINAL)])
    return table


class MPO(object):
    """List of tensors ``W[i]`` with axes (phys_out, phys_in, left bond, right bond)."""

    def __init__(self, Wlist: list[np.ndarray]) -> None:
        self.W = Wlist

    @classmethod
    def hamiltonian_from_rules(cls, rules) -> "MPO":
        lo, hi = rules.activation_interval.start, rules.activation_interval.stop
        table = rule_automaton(rules.distance, lo, hi)
        nb = len(table)
        bulk = np.zeros((2, 2, nb, nb), dtype=complex)
        for src, row in enumerate(table):
            for op, dst in row:
                bulk[:, :, dst, src] += _OPS[op]
        # `distance` dead cells beyond each end close the strings that overhang the chain
        dead = bulk[0, 0].real
        close_left = np.linalg.matrix_power(dead, rules.distance)[-1:, :]
        close_right = np.linalg.matrix_power(dead, rules.distance)[:, :1]
        wlist = [bulk for _ in range(rules.ncells)]
        wlist[0] = np.einsum("xl,ablr->abxr", close_left, bulk)
        wlist[-1] = np.einsum("ablr,ry->ably", bulk, close_right)
        return cls(wlist)

    @property
    def bond_dims(self) -> list[int]:
        return [self.W[0].shape[2]] + [w.shape[3] for w in self.W]

    def same_operator_as(self, other: "MPO", atol: float = 1e-12) -> bool:
        """True when both MPOs hold the same tensors (same bond numbering)."""
        if len(self.W) != len(other.W):
            return False
        return all(a.shape == b.shape and np.allclose(a, b, atol=atol, rtol=0.0)
                   for a, b in zip(self.W, other.W))

    def as_matrix(self) -> np.ndarray:
        """Dense matrix on the full Hilbert space (mpo.py:221-230).  Host-side format
        conversion for small chains (tests, debugging); never on the evolution path."""
        acc = self.W[0]
        for w in self.W[1:]:
            acc = np.einsum("ablm,cdmr->acbdlr", acc, w)
            s = acc.shape
            acc = acc.reshape(s[0] * s[1], s[2] * s[3], s[4], s[5])
        return np.trace(acc, axis1=2, axis2=3)
