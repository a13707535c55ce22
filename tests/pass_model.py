"""numpy model of what the CUDA tile-pass kernel computes, driven by the plan the C library
returns.  Used by the CPU tests to prove the host-side planning (tile geometry, flip masks,
Clenshaw coefficients) without a GPU, and by the GPU tests as an independent mid-size check.
Test infrastructure only."""
import numpy as np


def activity(xs: np.ndarray, ncells: int, distance: int, lo: int, hi: int) -> np.ndarray:
    """bit g of the result: cell (ncells-1-g) has its alive-neighbour count in [lo,hi)."""
    act = np.zeros_like(xs)
    for g in range(ncells):
        cnt = np.zeros(xs.shape, dtype=np.int64)
        for o in range(1, distance + 1):
            for nb in (g - o, g + o):
                if 0 <= nb < ncells:
                    cnt += (xs >> nb) & 1
        act |= (((cnt >= lo) & (cnt < hi)).astype(xs.dtype)) << g
    return act


def tile_indices(ps: dict, tile: int) -> np.ndarray:
    """global index of every element y of tile `tile`, with the kernel's formulas."""
    L, H0, M = ps["low_bits"], ps["high_start"], ps["high_bits"]
    gap = H0 - L
    t_lo, t_hi = tile & ((1 << gap) - 1), tile >> gap
    base = (t_lo << L) | (t_hi << (H0 + M))
    y = np.arange(1 << (L + M), dtype=np.int64)
    return base | (y & ((1 << L) - 1)) | ((y >> L) << H0)


def expand(xs: np.ndarray, positions, rank: int) -> np.ndarray:
    """local index -> global index: insert the sharded qubits (ascending global positions)."""
    x = xs.copy()
    for j, p in enumerate(positions):
        x = ((x >> p) << (p + 1)) | (x & ((1 << p) - 1)) | (((rank >> j) & 1) << p)
    return x


def global_pos(g: int, positions) -> int:
    for p in positions:
        if g >= p:
            g += 1
    return g


def apply_k_by_tiles(vec: np.ndarray, passes: list, nbits: int, ncells: int, distance: int, lo: int, hi: int,
                     positions=(), rank: int = 0) -> np.ndarray:
    """K vec, tile by tile exactly as the kernel stages it (real vector, one plane).  `positions`,
    `rank`: sharded register (the vector is this rank's local part)."""
    out = np.zeros_like(vec)
    for ps in passes:
        L, H0, M = ps["low_bits"], ps["high_start"], ps["high_bits"]
        T = L + M
        for tile in range(1 << (nbits - T)):
            xs = tile_indices(ps, tile)
            stage = vec[xs]
            act = activity(expand(xs, positions, rank), ncells, distance, lo, hi)
            y = np.arange(1 << T)
            acc = np.zeros(1 << T)
            for q in range(T):
                g = q if q < L else H0 + (q - L)
                if not (ps["flip_mask"] >> g) & 1:
                    continue
                on = ((act >> global_pos(g, positions)) & 1).astype(bool)
                partner = stage[y ^ (1 << q)]
                sign = np.where((y >> q) & 1, -1.0, 1.0)
                acc += np.where(on, sign * partner, 0.0)
            out[xs] += acc
    return out


def k_matrix(ncells: int, distance: int, lo: int, hi: int) -> np.ndarray:
    """Dense K with H = D (iK) D^-1, D = diag(i^popcount)."""
    dim = 1 << ncells
    xs = np.arange(dim, dtype=np.int64)
    act = activity(xs, ncells, distance, lo, hi)
    k = np.zeros((dim, dim))
    for g in range(ncells):
        on = ((act >> g) & 1).astype(bool)
        sign = np.where((xs >> g) & 1, -1.0, 1.0)
        k[xs[on], xs[on] ^ (1 << g)] += sign[on]
    return k


def popcount(xs: np.ndarray) -> np.ndarray:
    c = np.zeros_like(xs)
    v = xs.copy()
    while v.any():
        c += v & 1
        v >>= 1
    return c


def clenshaw_exp(kmat: np.ndarray, phi: np.ndarray, a: np.ndarray, bound: float, sign: float = 1.0):
    """phi' = sum_k a_k U_k(K/R) phi with the library's recurrence (qca_exact.cu step_once)."""
    K = len(a) - 1
    b1 = a[K] * phi
    b2 = np.zeros_like(phi)
    for k in range(K - 1, 0, -1):
        b1, b2 = a[k] * phi + sign * (2.0 / bound) * (kmat @ b1) + b2, b1
    return a[0] * phi + sign * (1.0 / bound) * (kmat @ b1) + b2


# ---------------------------------------------------------------------------------------------
# Model of the REMOTE OPERAND SLOTS of the fast tile-pass kernel (csrc/qca_pass.cuh, NREM > 0):
# per-tile activity decided by warp votes, one bulk copy per contiguous piece into a ring slot, one
# mbarrier per (piece, slot, ring row) whose phase is the parity of its ACTIVE earlier uses.
# Follows the kernel statement by statement; the CPU tests use it to prove the bookkeeping (which row
# sits in which ring slot when it is read, which parity is waited for) for every tile geometry.
# ---------------------------------------------------------------------------------------------
ROW_SHIFT = 9      # register rows are tile bits 9..12
ROWS = 16


def ring_rows_remote(nrem: int, rd: int) -> int:
    """ring_rem() of the kernel."""
    return 0 if nrem == 0 else (rd if nrem == 1 else 4)


def remote_slot_trace(ps: dict, tile: int, ops: list, rot: dict, pass_index: int, rd: int = 6):
    """For one CTA (tile of pass `ps`): returns, for every slot k, a dict (tid, row) -> index of the
    remote term applied to that thread's pair on that row (as the kernel would), after checking the ring
    and mbarrier bookkeeping of every warp.  ops / rot: qca_plan_remote / qca_plan_rotation."""
    L, H0 = ps["low_bits"], ps["high_start"]
    low_mask = (1 << L) - 1
    gap = H0 - L
    t_lo, t_hi = tile & ((1 << gap) - 1), tile >> gap
    M = 13 - L
    base = (t_lo << L) | (t_hi << (H0 + M))
    nrem = rot["nslots"]
    dc = ring_rows_remote(nrem, rd)
    chunk_lanes = 32 if L >= 6 else 1 << (L - 1)
    issuers = 32 // chunk_lanes
    assert issuers * nrem * dc <= 32                      # mbarrier budget of a warp

    def row_x(x_thr, e):
        ye = e << ROW_SHIFT
        return x_thr | (ye & low_mask) | ((ye >> L) << H0)

    def alt(k, r):                                        # a.rs[k].alt[r]
        j = int(rot["op_of"][pass_index, k, r]) if r < rot["npasses"] else -1
        return (j, ops[j]["mask"], ops[j]["shift"]) if j >= 0 else (-1, 0, 0)

    applied = [dict() for _ in range(nrem)]
    for warp in range(8):
        x_thr = [base | ((tid << 1) & low_mask) | (((tid << 1) >> L) << H0) for tid in range(32 * warp, 32 * warp + 32)]
        # --- per-tile decisions through votes -------------------------------------------------
        rows_per_lane = 1 if chunk_lanes >= 16 else 16 // chunk_lanes
        field = 16 if chunk_lanes >= 16 else chunk_lanes
        act = [[0] * 32 for _ in range(nrem)]
        rot0, rot1 = [0] * 32, [0] * 32
        for q in range(rows_per_lane):
            e_of = [((lane & (chunk_lanes - 1)) + q * chunk_lanes) & 15 for lane in range(32)]
            x_of = [row_x(x_thr[lane], e_of[lane]) for lane in range(32)]
            r_of = [(rot["rot_word"] >> (2 * ((x_of[lane] >> rot["rot_shift"]) & 15))) & 3 for lane in range(32)]
            v0 = sum(((r_of[lane] & 1) != 0) << lane for lane in range(32))
            v1 = sum(((r_of[lane] & 2) != 0) << lane for lane in range(32))
            vk = []
            for k in range(nrem):
                bits = 0
                for lane in range(32):
                    _, mask, shift = alt(k, r_of[lane])
                    bits |= ((mask >> ((x_of[lane] >> shift) & 15)) & 1) << lane
                vk.append(bits)
            for lane in range(32):
                fs = 0 if chunk_lanes >= 32 else (lane // chunk_lanes) * chunk_lanes
                rot0[lane] |= ((v0 >> fs) & ((1 << field) - 1)) << (q * field)
                rot1[lane] |= ((v1 >> fs) & ((1 << field) - 1)) << (q * field)
                for k in range(nrem):
                    act[k][lane] |= ((vk[k] >> fs) & ((1 << field) - 1)) << (q * field)
        # the votes must agree with what every lane would find for ITS OWN pair on every row
        for lane in range(32):
            for e in range(ROWS):
                x = row_x(x_thr[lane], e)
                r = (rot["rot_word"] >> (2 * ((x >> rot["rot_shift"]) & 15))) & 3
                assert r == ((rot0[lane] >> e) & 1) | (((rot1[lane] >> e) & 1) << 1)
                for k in range(nrem):
                    _, mask, shift = alt(k, r)
                    assert ((mask >> ((x >> shift) & 15)) & 1) == (act[k][lane] >> e) & 1
        # --- ring and mbarrier bookkeeping, piece by piece -------------------------------------
        for piece in range(issuers):
            lead = piece * chunk_lanes                      # issuing lane
            for k in range(nrem):
                a_bits = act[k][lead]
                assert all(act[k][lane] == a_bits for lane in range(lead, lead + chunk_lanes))   # constant over a piece
                slot_row = [None] * dc                      # which row's data a ring slot holds
                completed = [0] * dc                        # completed phases of the slot's mbarrier
                def issue(e):
                    if (a_bits >> e) & 1:
                        assert slot_row[e % dc] is None     # the previous occupant has been consumed
                        slot_row[e % dc] = e
                        completed[e % dc] += 1              # arrive.expect_tx + complete_tx: the phase completes
                for e in range(dc):
                    issue(e)
                for e in range(ROWS):
                    if (a_bits >> e) & 1:
                        uses_before = sum(1 << q for q in range(e % dc, e, dc))
                        parity = bin(a_bits & uses_before).count("1") & 1
                        # try_wait.parity(p) succeeds once the phase with that parity has completed:
                        # phases 0..completed-1 are done, the waited one is number (completed - 1)
                        assert completed[e % dc] >= 1 and (completed[e % dc] - 1) & 1 == parity
                        assert slot_row[e % dc] == e
                        slot_row[e % dc] = None
                        for lane in range(lead, lead + chunk_lanes):
                            r = ((rot0[lane] >> e) & 1) | (((rot1[lane] >> e) & 1) << 1)
                            applied[k][(32 * warp + lane, e)] = alt(k, r)[0]
                    if e + dc < ROWS:
                        issue(e + dc)
                assert all(s is None for s in slot_row)
    return applied


def measure_tiles_model(vec: np.ndarray, ps: dict):
    """What csrc/qca_measure.cu computes for one tile pass on a real vector: for every tile bit t that the
    pass measures, (local index bit, s0, s1, w); follows the kernel's thread/row decomposition."""
    L, H0, M = ps["low_bits"], ps["high_start"], ps["high_bits"]
    assert L + M == 13
    nbits = int(vec.shape[0]).bit_length() - 1
    first_bit = 0 if M == 0 else L
    low_mask = (1 << L) - 1
    tid = np.arange(256, dtype=np.int64)
    w = np.zeros((13, 256)); s1row = np.zeros((4, 256)); s1b0 = np.zeros(256); tot = np.zeros(256)
    for tile in range(1 << (nbits - 13)):
        gap = H0 - L
        base = ((tile & ((1 << gap) - 1)) << L) | ((tile >> gap) << (H0 + M))
        y_thr = tid << 1
        x_thr = base | (y_thr & low_mask) | ((y_thr >> L) << H0)
        v = np.zeros((16, 256, 2))
        for e in range(16):
            ye = e << 9
            x = x_thr | (ye & low_mask) | ((ye >> L) << H0)
            v[e, :, 0], v[e, :, 1] = vec[x], vec[x + 1]
        for e in range(16):
            nx, ny = v[e, :, 0] ** 2, v[e, :, 1] ** 2
            tot += nx + ny
            s1b0 += ny
            w[0] += v[e, :, 0] * v[e, :, 1]
            for k in range(4):
                if (e >> k) & 1:
                    s1row[k] += nx + ny
                else:
                    p = v[e | (1 << k)]
                    w[9 + k] += v[e, :, 0] * p[:, 0] + v[e, :, 1] * p[:, 1]
        for b in range(8):
            if b + 1 >= first_bit:
                for e in range(16):
                    p = v[e][tid ^ (1 << b)]
                    w[b + 1] += v[e, :, 0] * p[:, 0] + v[e, :, 1] * p[:, 1]
    total = tot.sum()
    out = []
    for t in range(first_bit, 13):
        if t == 0:
            s1, wt = s1b0.sum(), w[0].sum()
        elif t <= 8:
            s1, wt = tot[((tid >> (t - 1)) & 1) == 1].sum(), 0.5 * w[t].sum()
        else:
            s1, wt = s1row[t - 9].sum(), w[t].sum()
        g = t if t < L else H0 + (t - L)
        out.append((g, total - s1, s1, wt))
    return out


# ---------------------------------------------------------------------------------------------
# Model of the CLUSTER tile-pass kernel (csrc/qca_pass3.cuh): CTA tiles of 14 index bits, 2^CB CTAs per
# cluster, predicates from one thread-constant and one per-row window-table lookup (build_tables_v3 in
# csrc/qca_exact.cu).  Follows the kernel's index arithmetic statement by statement, vectorised over the
# 512 threads x 16 rows of a CTA; `tile_bits` / `row_shift` are parameters so the CPU tests can also run
# scaled-down geometries.
# ---------------------------------------------------------------------------------------------
def window_table(K: int, d: int, lo: int, hi: int) -> np.ndarray:
    """entry[w] = activity of the middle K bits of the (K + 2d)-bit window w (qca_exact.cu window_table)."""
    w = np.arange(1 << (K + 2 * d), dtype=np.int64)
    return (activity(w, K + 2 * d, d, lo, hi) >> d) & ((1 << K) - 1)


def v3_tables(ps: dict, d: int, lo: int, hi: int) -> dict:
    """build_tables_v3 for one pass."""
    ROWSHIFT, REGHIGH = 10, 4
    cb = ps["cluster_bits"]
    t = dict(thr=None, thr_shift=0, thr_pos=0, thr_mask=0)
    if ps["high_bits"] == 0:
        t["thr"] = window_table(ROWSHIFT - d, d, lo, hi)
        K = REGHIGH + d + cb
        t.update(row=window_table(K, d, lo, hi), row_shift=ROWSHIFT - 2 * d, row_mask=(1 << (K + 2 * d)) - 1, row_pos=ROWSHIFT - d)
    else:
        L, H0, M = ps["low_bits"], ps["high_start"], ps["high_bits"]
        MT = max(0, ROWSHIFT - L)
        KA = max(0, MT - d)
        if KA > 0:
            t.update(thr=window_table(KA, d, lo, hi), thr_shift=H0 - d, thr_mask=(1 << (KA + 2 * d)) - 1, thr_pos=L)
        K = M + cb - KA
        t.update(row=window_table(K, d, lo, hi), row_shift=H0 + KA - d, row_mask=(1 << (K + 2 * d)) - 1, row_pos=L + KA)
    return t


def apply_k_v3(vec: np.ndarray, passes: list, nbits: int, d: int, lo: int, hi: int) -> np.ndarray:
    """K vec through the cluster kernel's arithmetic (one plane)."""
    TILE, ROWSHIFT, THR = 14, 10, 9
    out = np.zeros_like(vec)
    tid = np.arange(512, dtype=np.int64)[None, :]
    e = np.arange(16, dtype=np.int64)[:, None]
    for ps in passes:
        L, H0, M, CB = ps["low_bits"], ps["high_start"], ps["high_bits"], ps["cluster_bits"]
        flip_low = M == 0
        tb = v3_tables(ps, d, lo, hi)
        low_mask = (1 << L) - 1
        gap = H0 - L
        ye = e << ROWSHIFT
        row_xg = (ye & low_mask) | ((ye >> L) << H0)
        for t in range(1 << (nbits - TILE - CB)):
            t_lo, t_hi = t & ((1 << gap) - 1), t >> gap
            staged = {}
            xs_of = {}
            for rank in range(1 << CB):
                base = (t_lo << L) | (t_hi << (H0 + M + CB)) | (rank << (H0 + M))
                y_thr = tid << 1
                x_thr = base | (y_thr & low_mask) | ((y_thr >> L) << H0)
                x = x_thr | row_xg                      # element 0 of the pair (row e, thread tid); element 1 = x | 1
                xs_of[rank] = (x, x_thr)
                staged[rank] = (vec[x], vec[x | 1])
            for rank in range(1 << CB):
                x, x_thr = xs_of[rank]
                if flip_low:
                    w = (x_thr & 0x3FF) << d
                    thr0, thr1 = tb["thr"][w], tb["thr"][w | (1 << d)]
                elif tb["thr"] is not None:
                    thr0 = thr1 = tb["thr"][(x_thr >> tb["thr_shift"]) & tb["thr_mask"]] << tb["thr_pos"]
                else:
                    thr0 = thr1 = np.zeros_like(x_thr)
                row_act = tb["row"][(x >> tb["row_shift"]) & tb["row_mask"]] << tb["row_pos"]
                la0, la1 = thr0 | row_act, thr1 | row_act
                v0, v1 = staged[rank]
                acc0, acc1 = np.zeros(v0.shape), np.zeros(v0.shape)
                qlo = 0 if flip_low else L
                for k in range(CB):
                    p0, p1 = staged[rank ^ (1 << k)]
                    sg = -1.0 if (rank >> k) & 1 else 1.0
                    on = ((la0 >> (TILE + k)) & 1).astype(bool)
                    acc0 += np.where(on, sg * p0, 0.0)
                    acc1 += np.where(on, sg * p1, 0.0)
                if qlo == 0:
                    acc0 += np.where(la0 & 1, v1, 0.0)
                    acc1 -= np.where(la1 & 1, v0, 0.0)
                for b in range(THR):
                    if b + 1 >= qlo:
                        partner = (tid ^ (1 << b))[0]
                        sg = np.where((tid >> b) & 1, -1.0, 1.0)
                        acc0 += np.where((la0 >> (b + 1)) & 1, sg * v0[:, partner], 0.0)
                        acc1 += np.where((la1 >> (b + 1)) & 1, sg * v1[:, partner], 0.0)
                for k in range(4):
                    if ROWSHIFT + k >= qlo:
                        prow = (e ^ (1 << k))[:, 0]
                        sg = np.where((e >> k) & 1, -1.0, 1.0)
                        acc0 += np.where((la0 >> (ROWSHIFT + k)) & 1, sg * v0[prow, :], 0.0)
                        acc1 += np.where((la1 >> (ROWSHIFT + k)) & 1, sg * v1[prow, :], 0.0)
                out[x] += acc0
                out[x | 1] += acc1
    return out
