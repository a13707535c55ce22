// Tile-pass kernel, third generation: 2^14-amplitude CTA tiles joined into thread-block CLUSTERS.
//
//   out[x] = alpha * a[x] + beta * c[x] + gamma * sum_{q in pass} P_q(x) * s_q(x) * in[x ^ bit_q]
//
// Why: a tile of T index bits applies the rule terms of T qubits with one read and one write of the state;
// at 13 bits per CTA a 30-qubit register needs three tile passes per operator application (13 + 9 + 8;
// 80 bytes per amplitude per Chebyshev term).  Shared memory cannot hold more than 2^14 doubles per SM, but
// the 2^CB CTAs of a cluster can read each other's staged tiles through distributed shared memory:
// 14 + 3 = 17 qubits per pass, 30 = 17 + 13 in TWO passes (56 bytes per amplitude per term).
//
//   * one CTA = 512 threads, one per SM (128 KiB tile + 96 KiB operand ring);
//   * per thread 32 amplitudes in registers: tile bit 0 (the halves of a 16-byte load) and tile bits 10..13
//     (16 rows `e`): their flips never leave the register file;
//   * tile bits 1..9 index the thread: their flips read the partner thread's pair with one LDS.128;
//   * cluster bits (the CB index bits above the CTA tile, = the CTA's rank in the cluster): their flips read
//     the partner CTA's staged tile at the same offset with one 16-byte ld.shared::cluster per row, skipped
//     for the whole warp when the predicate fails (it depends on row, rank and tile only, never on the lane);
//   * the rule predicate comes from two window-table lookups: one per thread (the bits whose neighbourhood
//     lies inside the thread/pair bits) and one per row (the rest, including the cluster bits);
//   * the local epilogue operands (a, c) stream through a per-thread cp.async ring as in pass_kernel_v2.
//
// Sharded registers (remote operand slots) stay on pass_kernel_v2 (qca_pass.cuh).
#pragma once
#include "qca_pass.cuh"

namespace qca {

constexpr int kPass3Threads = 512;
constexpr int kTile3 = 14;                       // index bits of one CTA tile
constexpr int kRowShift3 = kTile3 - kRegHigh;    // 10: tile bits 10..13 are register rows
constexpr int kThrBits3 = kRowShift3 - 1;        // 9 thread bits (tile bits 1..9)
constexpr int kMaxClusterBits = 3;               // portable cluster size 8
constexpr int kRing3Rows = 12;                   // 8 KiB each
// Request order and ring depth (A/B switches; the defaults are the measured winners, see the staging block of the kernel):
#ifndef QCA_V3_RING_FIRST
#define QCA_V3_RING_FIRST 0   // operand ring requested before the tile: 0 never (default), 1 in pass 0 only, 2 always (round-2 start)
#endif
#ifndef QCA_V3_DU1
#define QCA_V3_DU1 12  // ring rows in flight when a launch streams ONE operand (later passes); <= kRing3Rows
#endif
constexpr int kPass3TileRingBytes = (8 << kTile3) + kRing3Rows * kPass3Threads * 16;   // 224 KiB
constexpr int kPass3SmemBytes = kPass3TileRingBytes + 16;   // + the mbarrier of the bulk-copied tile

struct Pass3Args {
    const double* in[2];
    double* out[2];
    double gamma;
    const double* opnd[2][2];   // local epilogue operand k, plane p
    double coef[2];
    unsigned long long row_xg[16];   // index bits of register row e (tile bits 10..13 at their positions)
    unsigned long long ntiles;       // cluster tiles
    int low_bits, high_start, high_bits;   // CTA tile = bits [0,low) U [high_start, high_start+high_bits); cluster bits follow
    int distance;
    // predicate tables: act = tab_thr[w_thr] << thr_pos  |  tab_row[w_row] << row_pos   (tile-local bit numbering,
    // cluster bits at kTile3..)
    const unsigned short* tab_thr;   // nullptr: no thread-constant part
    const unsigned short* tab_row;
    int thr_shift, thr_pos;          // later passes: w_thr = (x >> thr_shift) & thr_mask
    unsigned thr_mask;
    int row_shift, row_pos;          // w_row = (x >> row_shift) & row_mask
    unsigned row_mask;
};

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
// arrival that publishes nothing (the end-of-tile "I have stopped reading your shared memory"): no fence
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned map_to_rank(unsigned smem_addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ double2 ld_dsmem(unsigned addr) {
    double2 v;
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
// v with its sign flipped when m == 0x80000000 (m is 0 otherwise): the (sigma^- - sigma^+) sign of a flip without a
// register per thread bit (a table of +-1.0 doubles cost 18 registers of a 128-register budget)
__device__ __forceinline__ double sign_flip(double v, unsigned m) {
    return __hiloint2double(__double2hiint(v) ^ (int)m, __double2loint(v));
}

// L: contiguous low bits of the CTA tile, M = 14 - L strided bits at H0; CB cluster bits directly above them.
// FLIP_LOW: pass 0 (L == 14): every tile bit is flipped; otherwise only the M high bits (and the cluster bits).
template <typename I, int L, bool FLIP_LOW, int NUNC, int CB>
__global__ void __launch_bounds__(kPass3Threads, 1) pass_kernel_v3(const Pass3Args a) {
    constexpr int M = kTile3 - L;
    constexpr int QLO = FLIP_LOW ? 0 : L;
    constexpr int DU = NUNC == 0 ? 0 : (NUNC == 1 ? QCA_V3_DU1 : 6);
    static_assert(NUNC * DU <= kRing3Rows, "ring budget");
    static_assert(FLIP_LOW ? (M == 0) : (M >= 1), "geometry");
    static_assert(L >= 4 && CB >= 0 && CB <= kMaxClusterBits, "geometry");
    extern __shared__ double tile[];
    const int plane = blockIdx.y;
    const double* __restrict__ in = a.in[plane];
    double* out = a.out[plane];
    const int H0 = a.high_start;
    const int d = a.distance;
    const unsigned tid = threadIdx.x;
    constexpr unsigned low_mask = (1u << L) - 1u;
    const unsigned rank = CB ? cluster_ctarank() : 0u;

    // cluster tile t, this CTA's part of it (rank -> the CB bits above the CTA tile), this thread's part
    const unsigned long long t = blockIdx.x >> CB;
    const int gap = H0 - L;
    const I t_lo = (I)(t & ((1ull << gap) - 1ull));
    const I t_hi = (I)(t >> gap);
    const I base = (t_lo << L) | (t_hi << (H0 + M + CB)) | ((I)rank << (H0 + M));
    const unsigned y_thr = tid << 1;  // tile bits 1..9
    const I x_thr = base | (I)(y_thr & low_mask) | ((I)(y_thr >> L) << H0);

    double2* tile2 = reinterpret_cast<double2*>(tile);
    double2* ring = tile2 + (1 << (kTile3 - 1));
    auto row_x = [&](int e) -> I {
        const unsigned ye = (unsigned)e << kRowShift3;
        return x_thr | (I)(ye & low_mask) | ((I)(ye >> L) << H0);
    };
    auto local_slot = [&](int k, int e) -> double2* { return ring + ((k * DU + (e % (DU ? DU : 1))) * kPass3Threads + tid); };
    auto local_fetch = [&](int e) {
#pragma unroll
        for (int k = 0; k < NUNC; ++k) cp_async16(local_slot(k, e), a.opnd[k][plane] + row_x(e));
    };
    auto ring_prologue = [&]() {
        if (NUNC) {
#pragma unroll
            for (int e = 0; e < DU; ++e) {
                local_fetch(e);
                cp_async_commit();
            }
        }
    };
#if defined(QCA_THREAD_ISSUE) || defined(QCA_V3_LDG_TILE)
    constexpr bool RING_FIRST = true;
#else
    constexpr bool RING_FIRST = (QCA_V3_RING_FIRST == 2) || (QCA_V3_RING_FIRST == 1 && FLIP_LOW);
#endif
    if (RING_FIRST) ring_prologue();
    // ---- stage the tile --------------------------------------------------------------------------------
    double2 v[kRows];
#ifndef QCA_V3_LDG_TILE
    // The tile goes global -> shared by bulk copies (TMA engine), one per contiguous piece (2^L doubles, at most 8 KiB),
    // all completing on one mbarrier; the threads then read their 16 pairs from shared memory.  Instead of 16 LDG.128
    // per thread (ncu, round 2: lg_throttle 3.1-3.4 stalled warps per issue next to long_scoreboard 4.4 -- the
    // load/store unit's instruction queue, not DRAM, paced the staging) the LSU sees no global load for the tile at all.
    // Measured at N = 30 on one B200, same box, interleaved runs (profiles/r02_ab_v3_*.json): pass 0 6.59 -> 5.49 ms
    // (6.26 TB/s, 0.96 of the measured copy peak), later passes 5.03 -> 4.79 and 5.08 -> 4.94 ms, 0.831 -> 0.911 steps/s.
    {
        const unsigned tile_bar = smem_u32(tile) + (unsigned)kPass3TileRingBytes;
        if (tid == 0) {
            mbar_init(tile_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) mbar_arrive_expect_tx(tile_bar, 8u << kTile3);
        constexpr int PIECE_BITS = L < 10 ? L : 10;
        constexpr int NPIECES = 1 << (kTile3 - PIECE_BITS);
#ifndef QCA_THREAD_ISSUE
        // THE TILE IS REQUESTED BEFORE THE OPERAND RING, by one elected lane per warp in a warp-uniform loop.
        // A CTA cannot compute before its tile has landed but needs only the first ring row at the end of its first
        // row, and one CTA per SM means nothing else runs while it waits: with the ring's 64-96 KiB queued ahead of
        // the tile (round-2 start) ncu put 34 % of all warp samples of a later pass behind the tile's mbarrier, plus
        // 11 % into the ELECT / R2UR / BRA.U.ANY loop in which the compiler serialises the (uniform-datapath) UBLKCP
        // of the 32 lanes of a warp when every thread issues one piece.  Same-box interleaved A/B at N = 30
        // (profiles/r02_ab_staging_order.txt): later passes 4.68 / 4.74 -> 4.09 / 4.14 ms (6.3 TB/s, 0.96 of the
        // measured copy peak), 0.954 -> 1.021 steps/s; the issue scheme alone, with the ring still first, gains
        // nothing (4.91 / 5.02 ms), the order is what pays; a 12-row ring is worth +0.4 % once the tile goes first and
        // costs 5 % when it does not.  Pass 0 slows by the clock the power cap takes back (5.17 -> 5.41 ms at
        // 1777 -> 1695 MHz): with every pass at the DRAM roofline the step is power-limited.
        {
            constexpr int PER_WARP = NPIECES / (kPass3Threads / 32);
            static_assert(PER_WARP >= 1, "at least one piece per warp");
            const unsigned wu = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp index, known to be warp-uniform
            if (elect_one()) {
#pragma unroll 4
                for (int i = 0; i < PER_WARP; ++i) {
                    const unsigned r = wu * PER_WARP + (unsigned)i;
                    const unsigned y = r << PIECE_BITS;
                    const I x = base | (I)(y & low_mask) | ((I)(y >> L) << H0);
                    bulk_g2s(smem_u32(tile + y), in + x, 8u << PIECE_BITS, tile_bar);
                }
            }
            __syncwarp();
        }
        if (!RING_FIRST) ring_prologue();
#else
#pragma unroll 1
        for (int r = (int)tid; r < NPIECES; r += kPass3Threads) {
            const unsigned y = (unsigned)r << PIECE_BITS;   // tile-local index of the first amplitude of the piece
            const I x = base | (I)(y & low_mask) | ((I)(y >> L) << H0);
            bulk_g2s(smem_u32(tile + y), in + x, 8u << PIECE_BITS, tile_bar);
        }
#endif
        mbar_wait(tile_bar, 0);
#pragma unroll
        for (int e = 0; e < kRows; ++e) v[e] = tile2[(e << kThrBits3) | tid];
    }
#else
#pragma unroll
    for (int e = 0; e < kRows; ++e) v[e] = ldg_stream(in + row_x(e));
#pragma unroll
    for (int e = 0; e < kRows; ++e) tile2[(e << kThrBits3) | tid] = v[e];
#endif

    // ---- thread-constant predicate words ---------------------------------------------------------------
    unsigned thr_act0 = 0, thr_act1 = 0;
    if (FLIP_LOW) {
        // tile bits [0, 10-d): window = x bits [0,10) shifted up by d (dead cells below the chain end)
        const unsigned w = ((unsigned)x_thr & 0x3ffu) << d;
        thr_act0 = a.tab_thr[w];
        thr_act1 = a.tab_thr[w | (1u << d)];
    } else if (a.tab_thr) {
        thr_act0 = thr_act1 = (unsigned)a.tab_thr[(unsigned)((unsigned long long)x_thr >> a.thr_shift) & a.thr_mask] << a.thr_pos;
    }
    // partner CTAs' tiles (same offset in their shared memory)
    unsigned peer[CB ? CB : 1];
    if (CB) {
        const unsigned mine = smem_u32(tile2 + tid);
#pragma unroll
        for (int k = 0; k < CB; ++k) peer[k] = map_to_rank(mine, rank ^ (1u << k));
        cluster_arrive();   // my tile is staged ...
        cluster_wait();     // ... and so is everybody else's (also orders this CTA's own stores: no __syncthreads needed)
    } else {
#ifdef QCA_V3_LDG_TILE
        __syncthreads();    // (bulk-copied tile: every thread has waited for the tile's mbarrier itself)
#endif
    }
    // Row predicates and the partner CTAs' pairs are fetched ONE ROW AHEAD: a distributed-shared-memory load takes
    // ~200+ cycles, and issued where it is consumed it was 45 % of the row time (ncu source page, round 2).
    auto row_lookup = [&](int e) -> unsigned {
        const unsigned long long xr = (unsigned long long)x_thr | a.row_xg[e];
        return (unsigned)a.tab_row[(unsigned)(xr >> a.row_shift) & a.row_mask] << a.row_pos;
    };
    unsigned ract[2];
    double2 rp[2][CB ? CB : 1];
    auto fetch_remote = [&](int e, unsigned act, double2* dst) {   // (the predicate is uniform over the warp)
#pragma unroll
        for (int k = 0; k < CB; ++k)
            if (act & (1u << (kTile3 + k))) dst[k] = ld_dsmem(peer[k] + (unsigned)(e << kThrBits3) * 16u);
    };
    ract[0] = row_lookup(0);
    fetch_remote(0, ract[0], rp[0]);
    // (sigma^- - sigma^+) sign of the flip of thread bit b as seen by this thread; only the flipped bits cost registers
#ifdef QCA_V3_P0_SIGN_TABLE
    constexpr bool SIGN_TABLE = true;
#else
    constexpr bool SIGN_TABLE = !FLIP_LOW;
#endif
    double sgn[kThrBits3];
#pragma unroll
    for (int b = 0; b < kThrBits3; ++b) sgn[b] = ((tid >> b) & 1u) ? -1.0 : 1.0;

#pragma unroll
    for (int e = 0; e < kRows; ++e) {
        if (e + 1 < kRows) {
            ract[(e + 1) & 1] = row_lookup(e + 1);
            fetch_remote(e + 1, ract[(e + 1) & 1], rp[(e + 1) & 1]);
        }
        const unsigned row_act = ract[e & 1];
        const unsigned la0 = thr_act0 | row_act, la1 = thr_act1 | row_act;
        // Later passes (SIGN_TABLE): written the way nvcc turns into ONE DFMA/DADD + two FSELs per amplitude and flip,
        // with the predicates of a row moved into predicate registers wholesale (R2P).  Applying the thread's sign by
        // XOR on the loaded pair instead (saves the +-1.0 table) makes nvcc wrap every partner LDS in a divergent
        // branch: 6.8 instructions per amplitude and flip, 1736 instead of 1296 per thread and tile.  Pass 0 flips all
        // nine thread bits; there the 18 registers of the table do not fit and the XOR form stays.
        // Two accumulator pairs (partner threads / partner rows) halve the FP64 dependency chain.
        double acc0 = 0.0, acc1 = 0.0, accr0 = 0.0, accr1 = 0.0;
        if (QLO == 0) {   // tile bit 0: the other half of the pair
            if (la0 & 1u) accr0 += v[e].y;
            if (la1 & 1u) accr1 -= v[e].x;
        }
        // tile bits 1..9: partner thread, same row
#pragma unroll
        for (int b = 0; b < kThrBits3; ++b) {
            if (b + 1 >= QLO) {
                double2 p = tile2[(e << kThrBits3) | (tid ^ (1u << b))];
#ifndef QCA_V3_LDG_TILE
                asm volatile("" : "+d"(p.x), "+d"(p.y));   // pins the LDS here: never sunk into a (divergent) branch
#endif
                if (SIGN_TABLE) {
                    if (la0 & (2u << b)) acc0 = fma(p.x, sgn[b], acc0);
                    if (la1 & (2u << b)) acc1 = fma(p.y, sgn[b], acc1);
                } else {
                    const unsigned m = (tid << (31 - b)) & 0x80000000u;
                    if (la0 & (2u << b)) acc0 += sign_flip(p.x, m);
                    if (la1 & (2u << b)) acc1 += sign_flip(p.y, m);
                }
            }
        }
        // tile bits 10..13: partner row, same thread (registers)
#pragma unroll
        for (int k = 0; k < kRegHigh; ++k) {
            if (kRowShift3 + k >= QLO) {
                const double2 p = v[e ^ (1 << k)];
                if ((e >> k) & 1) {
                    if (la0 & (1u << (kRowShift3 + k))) accr0 -= p.x;
                    if (la1 & (1u << (kRowShift3 + k))) accr1 -= p.y;
                } else {
                    if (la0 & (1u << (kRowShift3 + k))) accr0 += p.x;
                    if (la1 & (1u << (kRowShift3 + k))) accr1 += p.y;
                }
            }
        }
        acc0 += accr0;
        acc1 += accr1;
        // cluster bits: the pairs fetched one row ago
        if (CB) {
#pragma unroll
            for (int k = 0; k < CB; ++k) {
                if (row_act & (1u << (kTile3 + k))) {
                    const unsigned m = (rank << (31 - k)) & 0x80000000u;
                    acc0 += sign_flip(rp[e & 1][k].x, m);
                    acc1 += sign_flip(rp[e & 1][k].y, m);
                }
            }
        }
        double2 r;
        r.x = a.gamma * acc0;
        r.y = a.gamma * acc1;
        if (NUNC) {
            cp_async_wait<(NUNC ? DU : 1) - 1>();
#pragma unroll
            for (int k = 0; k < NUNC; ++k) {
                const double2 sv = *local_slot(k, e);
                r.x = fma(a.coef[k], sv.x, r.x);
                r.y = fma(a.coef[k], sv.y, r.y);
            }
        }
        stg_stream(out + row_x(e), r);
        if (NUNC) {
            if (e + DU < kRows) local_fetch(e + DU);
            cp_async_commit();
        }
    }
    if (CB) {   // nobody may leave (and hand its shared memory to the next CTA) while a partner still reads its tile
        cluster_arrive_relaxed();
        cluster_wait();
    }
}

typedef void (*Pass3Kernel)(const Pass3Args);
Pass3Kernel pass3_kernel_u32(int low_bits, int nunc, int cluster_bits);
Pass3Kernel pass3_kernel_u64(int low_bits, int nunc, int cluster_bits);

}  // namespace qca
