"""CPU oracle for the quantum-cellular-automaton time-evolution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; it is the checker used by ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

It restates, in plain numpy, the algorithm of the reference
(BenjaminDecker/quantum-cellular-automaton) for the path
``MPO.hamiltonian_from_rules -> MPO.as_matrix -> calculate_U -> Exact.do_time_step
-> MPS.measure``.  Every function cites the reference file:line it follows.

Parity pin: the fixtures under ``tests/golden/`` were produced by importing the
unmodified reference from /root/reference (``tests/golden/make_golden.py``);
``tests/test_oracle_golden.py`` checks this module against every one of them.

Conventions (same as the reference): cell 0 is the most significant bit of the
basis-state index (``MPS.as_vector`` / ``MPO.as_matrix`` merge site 0 first), the
evolution operator of one step is ``exp(-i*pi/2*step_size*H)``.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# Rule -> MPO (reference: tensor_networks/mpo.py:12-202)
# ----------------------------------------------------------------------------

_ID = np.eye(2)
_P0 = np.diag([1.0, 0.0])
_P1 = np.diag([0.0, 1.0])
_SX = np.array([[0.0, 1.0], [1.0, 0.0]])  # LOWERING + RISING, constants.py:15-19


def automaton(distance: int, lo: int, hi: int):
    """Finite-state automaton whose paths are the operator strings of H.

    Follows ``StateAutomaton`` (mpo.py:56-151).  A node is the triple
    (zeros, ones, flipped): how many |0>- and |1>-projectors were emitted so far
    and whether the sigma-x was emitted.  Nodes are numbered in the reference's
    depth-first discovery order (the |1>-projector child is explored before the
    |0>-projector child, mpo.py:126-149) so the bond index of the tensors is the
    same as the reference's.  Returns ``(nodes, edges)`` where ``edges[k]`` is a
    list of ``(operator, target)`` and target ``-1`` is the final node.
    """
    width = 2 * distance  # projectors per string, mpo.py:76-77
    index: dict[tuple[int, int, bool], int] = {}
    edges: list[list[tuple[np.ndarray, int]]] = []

    def visit(zeros: int, ones: int, flipped: bool) -> int:
        key = (zeros, ones, flipped)
        if key in index:
            return index[key]
        if zeros + ones + int(flipped) == width + 1:  # mpo.py:104-105
            return -1
        me = len(edges)
        index[key] = me
        edges.append([])
        out = edges[me]
        if zeros + ones == distance and not flipped:  # mpo.py:112-119
            out.append((_SX, visit(zeros, ones, True)))
            return me
        if zeros + ones == 0 and not flipped:  # mpo.py:122-123
            out.append((_ID, 0))
        if ones + 1 < hi:  # mpo.py:126-135
            out.append((_P1, visit(zeros, ones + 1, flipped)))
        left = width - (zeros + ones)
        if ones + (left - 1) >= lo:  # mpo.py:139-149
            out.append((_P0, visit(zeros + 1, ones, flipped)))
        return me

    visit(0, 0, False)
    edges.append([(_ID, -1)])  # final node loops onto itself, mpo.py:66-73
    nodes = sorted(index, key=index.get) + [(-1, -1, True)]
    return nodes, edges


def mpo_tensors(ncells: int, distance: int, lo: int, hi: int) -> list[np.ndarray]:
    """MPO tensors W[i] of shape (2, 2, Dl, Dr); mpo.py:161-202 (open boundary)."""
    _, edges = automaton(distance, lo, hi)
    nb = len(edges)
    bulk = np.zeros((2, 2, nb, nb), dtype=complex)
    for src, outs in enumerate(edges):
        for op, dst in outs:
            bulk[:, :, dst, src] += op  # dst == -1 addresses the final node
    tensors = [bulk] * ncells
    # Dead (|0>) virtual cells beyond both ends, mpo.py:181-200
    dead = bulk[0, 0]
    lvec = np.zeros((1, nb))
    lvec[0, -1] = 1.0
    rvec = np.zeros((nb, 1))
    rvec[0, 0] = 1.0
    for _ in range(distance):
        lvec = lvec @ dead
        rvec = dead @ rvec
    tensors = list(tensors)
    tensors[0] = np.einsum("xl,ablr->abxr", lvec, bulk)
    # (for ncells == 1 the reference overwrites W[0] here as well, mpo.py:198-200)
    tensors[-1] = np.einsum("ablr,ry->ably", bulk, rvec)
    return tensors


def mpo_as_matrix(tensors: list[np.ndarray]) -> np.ndarray:
    """Dense 2^N x 2^N matrix of an MPO; mpo.py:205-230."""
    acc = tensors[0]
    for w in tensors[1:]:
        acc = np.einsum("ablm,cdmr->acbdlr", acc, w)
        s = acc.shape
        acc = acc.reshape(s[0] * s[1], s[2] * s[3], s[4], s[5])
    return np.trace(acc, axis1=2, axis2=3)


def rule_hamiltonian_direct(ncells: int, distance: int, lo: int, hi: int) -> np.ndarray:
    """H = sum_i sigma^x_i * [ #alive neighbours of i within `distance` in [lo,hi) ].

    Not in the reference as such: this is the closed form the automaton encodes
    (cells beyond the ends count as dead).  ``tests/test_oracle_golden.py`` proves
    it equal to ``mpo_as_matrix(mpo_tensors(...))`` and to the reference's own
    ``MPO.as_matrix`` output, which is what licenses the matrix-free CUDA kernel.
    """
    dim = 1 << ncells
    h = np.zeros((dim, dim))
    xs = np.arange(dim)
    for cell in range(ncells):
        bit = ncells - 1 - cell
        count = np.zeros(dim, dtype=np.int64)
        for off in range(1, distance + 1):
            for nb in (cell - off, cell + off):
                if 0 <= nb < ncells:
                    count += (xs >> (ncells - 1 - nb)) & 1
        act = (count >= lo) & (count < hi)
        h[xs[act], xs[act] ^ (1 << bit)] += 1.0
    return h


# ----------------------------------------------------------------------------
# Initial states (reference: states.py:11-88, mps.py:35-52, mps.py:194-208)
# ----------------------------------------------------------------------------

def initial_plist(name: str, ncells: int, distance: int = 1) -> list[float]:
    """Alive-probability per cell for the named initial state; states.py:11-88."""
    n = ncells
    mid = int(n / 2)
    p = [0.0] * n
    if name == "blinker":
        p[mid - 1] = p[mid + 1] = 1.0
    elif name == "triple_blinker":
        p[mid - 2] = p[mid] = p[mid + 2] = 1.0
    elif name == "full_blinker":
        p = [float(i % 2) for i in range(n)]
    elif name == "single":
        p[mid] = 1.0
    elif name == "single_bottom":
        p[0] = 1.0
    elif name == "all_ket_0":
        pass
    elif name == "all_ket_1":
        p = [1.0] * n
    elif name == "only_outer":
        p[0] = p[-1] = 1.0
    elif name == "all_ket_1_but_outer":
        p = [0.0] * distance + [1.0] * (n - 2 * distance) + [0.0] * distance
    elif name == "equal_superposition":
        p = [0.5] * n
    elif name == "equal_superposition_but_outer":
        p = [0.0] * distance + [0.5] * (n - 2 * distance) + [0.0] * distance
    elif name == "gradient":
        p = [float(np.sin(np.pi * i / (n - 1) / 2)) for i in range(n)]
    else:
        raise ValueError(f"unknown initial state {name!r}")
    return p


def product_state_vector(plist) -> np.ndarray:
    """Bond-dimension-1 MPS of amplitudes (sqrt(1-p), sqrt(p)) merged to a vector.

    mps.py:35-52 (``from_density_distribution``) followed by mps.py:194-208
    (``as_vector``): site 0 is the most significant bit.
    """
    psi = np.ones(1, dtype=complex)
    for p in plist:
        psi = np.kron(psi, np.array([(1.0 - p) ** 0.5, p ** 0.5], dtype=complex))
    return psi


# ----------------------------------------------------------------------------
# Exact evolution (reference: lautils/lautils.py:45-55, algorithms/exact.py:15-27)
# ----------------------------------------------------------------------------

def calculate_U(h: np.ndarray, step_size: float) -> np.ndarray:
    """exp(-i*pi/2*step_size*H) through the eigendecomposition; lautils.py:45-55."""
    w, v = np.linalg.eigh(h)
    phase = np.exp(-1j * (np.pi / 2) * step_size * w)
    return (v * phase) @ v.conj().T


def exact_step(u: np.ndarray, psi: np.ndarray) -> np.ndarray:
    """One time step psi <- U psi; exact.py:26-27."""
    return u @ psi


# ----------------------------------------------------------------------------
# Measurement (reference: tensor_networks/mps.py:100-140 via mps.py:55-73)
# ----------------------------------------------------------------------------

def site_density_matrices(psi: np.ndarray, ncells: int) -> np.ndarray:
    """Single-site reduced density matrices rho[i] (2x2), index order as the
    reference's ``tensordot(A, A.conj(), ((1,2),(1,2)))`` at the orthogonality
    centre (mps.py:122-126): rho[a, b] = sum_rest psi[a, rest] conj(psi[b, rest])."""
    rho = np.empty((ncells, 2, 2), dtype=complex)
    for cell in range(ncells):
        t = psi.reshape(1 << cell, 2, -1)
        rho[cell] = np.einsum("lar,lbr->ab", t, t.conj())
    return rho


def entropy_bits(rho: np.ndarray) -> float:
    """-Tr(rho log2 rho); mps.py:135-140 (scipy ``logm`` there; 0*log 0 := 0 here,
    which is what the reference prints for product states)."""
    lam = np.linalg.eigvalsh(rho)
    lam = lam[lam > 0.0]
    return float(-(lam * np.log2(lam)).sum())


def measure_vector(psi: np.ndarray, ncells: int):
    """population, rounded population, single-site entropy, bond dimensions.

    mps.py:100-140 applied to ``MPS.from_vector(psi)`` (mps.py:55-73), which is what
    ``Exact.psi`` hands to ``Algorithm.measure`` (exact.py:19-20, algorithm.py:58-63).
    The bond dimensions of ``from_vector`` are those of successive reduced QRs:
    min(2^i, 2^(N-i)).
    """
    rho = site_density_matrices(psi, ncells)
    pop = rho[:, 1, 1].real.copy()
    dpop = np.round(pop)
    ent = np.array([entropy_bits(r) for r in rho])
    bonds = np.array([float(min(1 << i, 1 << (ncells - i))) for i in range(ncells + 1)])
    return pop, dpop, ent, bonds


def classical_evolution(first_column: np.ndarray, distance: int, lo: int, hi: int,
                        plot_steps: int) -> np.ndarray:
    """Classical (Wolfram-like) evolution used for the comparison heat map;
    algorithm.py:34-56, including its quirk that the two border columns are
    pinned to their initial values before the sweep and then updated anyway."""
    n = len(first_column)
    heat = np.zeros((plot_steps, n))
    heat[0] = first_column
    heat[:, 0] = first_column[0]
    heat[:, -1] = first_column[-1]
    for step in range(1, plot_steps):
        prev = heat[step - 1]
        for site in range(n):
            alive = 0.0
            for off in range(-distance, distance + 1):
                j = site + off
                if off != 0 and 0 <= j < n:
                    alive += prev[j]
            heat[step, site] = 1.0 - prev[site] if alive in range(lo, hi) else prev[site]
    return heat


def run_exact(name_or_plist, ncells: int, distance: int, lo: int, hi: int,
              step_size: float, nsteps: int):
    """The reference's exact loop (quantum_game.py:82-119 with algorithm == 'exact'):
    measure, then step, ``nsteps`` times.  Returns the stacked measurements and the
    final vector."""
    plist = (initial_plist(name_or_plist, ncells, distance)
             if isinstance(name_or_plist, str) else list(name_or_plist))
    psi = product_state_vector(plist)
    u = calculate_U(mpo_as_matrix(mpo_tensors(ncells, distance, lo, hi)), step_size)
    pops, dpops, ents, bonds = [], [], [], []
    for _ in range(nsteps):
        p, d, e, b = measure_vector(psi, ncells)
        pops.append(p), dpops.append(d), ents.append(e), bonds.append(b)
        psi = exact_step(u, psi)
    return (np.array(pops), np.array(dpops), np.array(ents), np.array(bonds), psi)


# ----------------------------------------------------------------------------
# Matrix-free restatement of H (for registers too large for a dense matrix)
# ----------------------------------------------------------------------------

def rule_activity(xs: np.ndarray, ncells: int, distance: int, lo: int, hi: int) -> np.ndarray:
    """act[x] bit (ncells-1-cell) = [# alive neighbours of `cell` within `distance` in [lo, hi)]
    for every basis state in ``xs``: the predicate of ``rule_hamiltonian_direct`` (the count test the
    automaton of mpo.py:126-149 encodes, dead cells beyond both ends mpo.py:181-200) as bit words."""
    xs = np.asarray(xs, dtype=np.int64)
    act = np.zeros_like(xs)
    for cell in range(ncells):
        count = np.zeros_like(xs)
        for off in range(1, distance + 1):
            for nb in (cell - off, cell + off):
                if 0 <= nb < ncells:
                    count += (xs >> (ncells - 1 - nb)) & 1
        act |= ((count >= lo) & (count < hi)).astype(np.int64) << (ncells - 1 - cell)
    return act


def apply_h(v: np.ndarray, ncells: int, distance: int, lo: int, hi: int) -> np.ndarray:
    """H @ v without the matrix: (H v)[x] = sum_cells [P_cell(x)] v[x ^ bit_cell].

    Same operator as ``mpo_as_matrix(mpo_tensors(...))`` (= the reference's ``MPO.as_matrix``,
    mpo.py:221-230); ``tests/test_oracle_golden.py`` checks it against the reference's own ``H @ v``
    fixtures and against the dense matrix for every N <= 11 case.  P_cell does not depend on the
    cell's own bit, so the predicate at x and at the flipped partner agree (H is symmetric)."""
    v = np.asarray(v)
    xs = np.arange(1 << ncells, dtype=np.int64)
    act = rule_activity(xs, ncells, distance, lo, hi)
    out = np.zeros_like(v)
    for bit in range(ncells):
        on = ((act >> bit) & 1).astype(bool)
        out[on] += v[xs[on] ^ (1 << bit)]
    return out


def exact_step_matrix_free(psi: np.ndarray, ncells: int, distance: int, lo: int, hi: int,
                           step_size: float, tol: float = 1e-16) -> np.ndarray:
    """exp(-i pi/2 step_size H) psi by the Taylor series with ``apply_h`` (scaling by halving so
    that every sub-step has |t| * N <= 1: plain series, no cancellation).  Used where the dense
    ``calculate_U`` (lautils.py:45-55) does not fit; checked against it in the CPU suite."""
    t = (np.pi / 2) * step_size
    pieces = max(1, int(np.ceil(abs(t) * ncells)))
    dt = t / pieces
    out = np.array(psi, dtype=complex)
    for _ in range(pieces):
        term = out.copy()
        acc = out.copy()
        for k in range(1, 200):
            term = (-1j * dt / k) * apply_h(term, ncells, distance, lo, hi)
            acc += term
            if np.abs(term).max() < tol:
                break
        out = acc
    return out


# ----------------------------------------------------------------------------
# The reference's actual measurement route (what its CPU time goes into)
# ----------------------------------------------------------------------------

def vector_to_mps(psi: np.ndarray) -> list[np.ndarray]:
    """Successive reduced QRs from the left; mps.py:55-73 (``MPS.from_vector``).
    Tensors are (physical, left bond, right bond)."""
    tensors = []
    rest = np.array(psi, dtype=complex).reshape(2, -1)
    while rest.shape[1] > 1:
        q, r = np.linalg.qr(rest)
        tensors.append(q.reshape(-1, 2, r.shape[0]).transpose(1, 0, 2))
        rest = r.reshape(r.shape[0] * 2, -1)
    tensors.append(rest.reshape(-1, 2, 1).transpose(1, 0, 2))
    return tensors


def _fit(a: np.ndarray, shape) -> np.ndarray:
    """Zero-pad / cut to ``shape`` (MPS.truncate_and_pad_into_shape as used by mps.py:146-181)."""
    out = np.zeros(shape, dtype=a.dtype)
    cut = tuple(slice(0, min(s, t)) for s, t in zip(a.shape, shape))
    out[cut] = a[cut]
    return out


def measure_via_mps(psi: np.ndarray, ncells: int):
    """``Exact.psi`` -> ``MPS.measure`` exactly as the reference does it (exact.py:19-20,
    mps.py:100-140): build the MPS by QR, move the orthogonality centre to site 0 by right-QRs
    (mps.py:183-192), then per site the 2x2 density matrix, the population, the entropy through
    ``scipy.linalg.logm`` and one left-QR to shift the centre.  Slower than ``measure_vector`` and
    equal to it to round-off; this is the routine the CPU baseline times."""
    from scipy.linalg import logm
    import warnings
    a = vector_to_mps(psi)
    n = len(a)
    assert n == ncells
    bonds = np.array([float(t.shape[1]) for t in a] + [float(a[-1].shape[2])])
    for i in range(n - 1, 0, -1):  # make_site_canonical(0): orthonormalize_right_qr, mps.py:164-181
        t = a[i]
        s = t.shape
        q, r = np.linalg.qr(t.transpose(0, 2, 1).reshape(s[0] * s[2], s[1]))
        q = _fit(q.reshape(s[0], s[2], -1).transpose(0, 2, 1), s)
        r = _fit(r.T, (a[i - 1].shape[2], s[1]))
        a[i - 1] = np.tensordot(a[i - 1], r, (2, 0))
        a[i] = q
    pop, dpop, ent = np.zeros(n), np.zeros(n), np.zeros(n)
    for site in range(n):
        if site > 0:  # orthonormalize_left_qr(site - 1), mps.py:146-162
            t = a[site - 1]
            s = t.shape
            q, r = np.linalg.qr(t.reshape(s[0] * s[1], s[2]))
            a[site - 1] = _fit(q.reshape(s[0], s[1], -1), s)
            r = _fit(r, (s[2], a[site].shape[1]))
            a[site] = np.tensordot(r, a[site], (1, 1)).transpose(1, 0, 2)
        t = a[site]
        rho = np.tensordot(t, t.conj(), ((1, 2), (1, 2)))
        pop[site] = rho[1, 1].real
        dpop[site] = np.round(pop[site])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ent[site] = (-np.trace(rho @ (logm(rho) / np.log(2)))).real
    return pop, dpop, ent, bonds
