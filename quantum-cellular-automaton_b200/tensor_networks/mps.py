"""Host container for matrix product states: ``A[i].shape == (2, D[i], D[i+1])``.

Data format on either side of the evolution path, compatible with the reference's
``tensor_networks.MPS`` (mps.py:14-236): same tensor layout, same ``.npz`` files
(``write_to_file`` / ``from_file``), same constructors.  Numerical work on the state happens
on the device; what is here is construction and format conversion.
"""
from __future__ import annotations

import numpy as np


class MPS(object):

    def __init__(self, Alist: list[np.ndarray], plist=None) -> None:
        self.A = Alist
        # alive probabilities when built by from_density_distribution: lets the exact engine
        # build the 2^N product state on the device instead of merging tensors on the host
        self.plist = None if plist is None else [float(p) for p in plist]

    @classmethod
    def from_tensors(cls, Alist) -> "MPS":
        return cls([np.array(A, dtype=complex) for A in Alist])

    @classmethod
    def from_density_distribution(cls, plist) -> "MPS":
        """Bond-dimension-1 state with P(cell i alive) = plist[i] (mps.py:35-52)."""
        tensors = []
        for p in plist:
            t = np.zeros((2, 1, 1), dtype=complex)
            t[0, 0, 0] = (1. - p) ** .5
            t[1, 0, 0] = p ** .5
            tensors.append(t)
        return cls(tensors, plist=plist)

    @classmethod
    def from_vector(cls, psi) -> "MPS":
        """Exact MPS of a state vector by successive reduced QRs, site 0 = top bit
        (mps.py:55-73); bond dimensions min(2^i, 2^(N-i))."""
        rest = np.array(psi, dtype=complex).reshape(2, -1)
        tensors = []
        while rest.shape[1] > 1:
            q, r = np.linalg.qr(rest)
            tensors.append(q.reshape(-1, 2, r.shape[0]).transpose(1, 0, 2))
            rest = r.reshape(r.shape[0] * 2, -1)
        tensors.append(rest.reshape(-1, 2, 1).transpose(1, 0, 2))
        return cls(tensors)

    @classmethod
    def from_file(cls, path: str) -> "MPS":
        with open(path, "rb") as f:
            data = np.load(f)
            return cls([data[f"arr_{i}"] for i in range(len(data.files))])

    def write_to_file(self, path: str) -> None:
        with open(path, "wb") as f:
            np.savez(f, *self.A)

    @property
    def bond_dims(self) -> list[int]:
        return [a.shape[1] for a in self.A] + [self.A[-1].shape[2]]

    def is_product_state(self) -> bool:
        return all(a.shape[1] == 1 and a.shape[2] == 1 for a in self.A)

    def is_valid_mps(self) -> bool:
        dims = self.bond_dims
        return dims[0] == 1 and dims[-1] == 1 and all(
            self.A[i].shape[2] == self.A[i + 1].shape[1] for i in range(len(self.A) - 1))

    def as_vector(self) -> np.ndarray:
        """Merge all tensors into the 2^N vector, site 0 = top bit (mps.py:194-208)."""
        acc = self.A[0]
        for a in self.A[1:]:
            acc = np.einsum("slm,tmr->stlr", acc, a)
            acc = acc.reshape(acc.shape[0] * acc.shape[1], acc.shape[2], acc.shape[3])
        return np.trace(acc, axis1=1, axis2=2)
