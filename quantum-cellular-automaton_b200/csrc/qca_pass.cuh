// Tile-pass kernels of the rule operator K (see qca_exact.cu header for the rotated frame).
//
//   out[x] = alpha * a[x] + beta * c[x] + gamma * sum_{q in pass} P_q(x) * s_q(x) * in[x ^ bit_q]
//
// A CTA owns one tile: the 2^13 amplitudes whose index bits [0,L) and [H0,H0+M) vary
// (L + M = 13; 64 KiB of shared memory, two CTAs per SM).
//
// Fast kernel (pass_kernel_v2), per thread (256 threads):
//   * 32 amplitudes live in registers: local bit 0 (the two halves of a 16-byte load) and
//     local bits 9..12 (16 rows `e`), so 5 of the 13 possible flips never leave the
//     register file;
//   * local bits 1..8 index the thread; their flips read the partner thread's pair with
//     one LDS.128 from the staged tile;
//   * the rule predicate P_q(x) comes from window lookup tables (a few KiB, L1-resident)
//     instead of being recomputed per amplitude;
//   * global traffic is 16-byte vector loads/stores, 128-byte rows at least.
#pragma once
#include "qca_common.cuh"

namespace qca {

constexpr int kPassThreads = 256;
constexpr int kTile = 13;      // == kTileBits
constexpr int kRegHigh = 4;    // local bits 9..12 are register rows
constexpr int kRows = 1 << kRegHigh;
constexpr int kRowShift = kTile - kRegHigh;  // 9

constexpr int kMaxStreams = 6;       // generic kernel; the fast kernel takes at most two local operands
constexpr int kMaxRemoteSlots = 2;   // fast kernel: remote (partner-rank) operand slots per launch
constexpr int kMaxRotations = 4;     // == largest number of tile passes the rotation spreads terms over

// One epilogue operand:  out[x] += coef * [active(x)] * ptr[x].
// Local recurrence operands (a, c) are unconditional; a *remote* operand is the partner rank's
// copy of `in` (peer memory over NVLink) for a term that flips a sharded qubit: it is active where
// the rule predicate of that qubit holds, a function of the top `distance` local bits only.
struct EpiStream {
    const double* ptr[2];  // per plane
    double coef;
    unsigned mask;         // fast kernel: active iff (mask >> ((x_local >> shift) & 15)) & 1; ~0u = always
    int shift;
    int bit;               // generic kernel: active iff activity bit `bit` of the GLOBAL index is set; -1 = always
    int pad;
};

// Fast kernel, sharded register.  A remote slot carries, for every amplitude x, at most one of the
// terms that flip a sharded qubit: which one depends on the rotation r(x) (a function of four local
// index bits above the first tile), so that over the passes of one operator application every term
// is applied exactly once for every x while every launch pulls the same share of the NVLink traffic.
struct RemoteAlt {
    const double* ptr[2];  // partner rank's copy of `in`, per plane (peer mapped)
    double coef;           // gamma * sign
    unsigned mask;         // active iff (mask >> ((x_local >> shift) & 15)) & 1; 0 = no term
    int shift;
};
struct RemoteSlot {
    RemoteAlt alt[kMaxRotations];
};

struct PassArgs {
    const double* in[2];   // vector the operator is applied to (plane 0/1)
    double* out[2];
    double gamma;
    EpiStream s[kMaxStreams];
    int nstreams;
    unsigned long long row_xg[16];    // fast kernel: GLOBAL index bits of register row e (sharded qubits inserted, no rank bits)
    RemoteSlot rs[kMaxRemoteSlots];   // fast kernel only
    int nrem;
    unsigned rot_word;     // r(x) = (rot_word >> 2 * ((x_local >> rot_shift) & 15)) & 3
    int rot_shift;
    unsigned long long flip_mask;  // qubits (local index bits) whose terms this pass applies
    ShardMap shard;                // local index -> global basis-state index (sharded qubits inserted)
    int win_shift;                 // later passes: window of tab_hi = (global index >> win_shift) & win_mask
    unsigned win_mask;
    unsigned long long ntiles;
    int low_bits, high_start, high_bits;  // tile = bits [0,low) U [high_start, high_start+high_bits)
    int distance;
    unsigned interval_mask;
    // window tables (fast kernel): entry[w] = activity of the middle K bits of the (K+2d)-bit window w
    const unsigned short* tab_lo;  // pass 0: K = 9-d, window = x bits [0,9) << d
    const unsigned short* tab_hi;  // pass 0: K = 4+d, window = global bits [9-2d, 13+d); pass>=1: the pass's
                                   // flipped qubits' global span +-d (entries already compressed to local bits)
};

// ---------------------------------------------------------------------------
// Generic kernel: any tile geometry, any distance <= 7.  Used for registers of fewer
// than 13 qubits and for distance > 4; one amplitude at a time from shared memory.
// ---------------------------------------------------------------------------
template <typename I>
__global__ void __launch_bounds__(kPassThreads) pass_kernel_generic(const PassArgs a) {
    extern __shared__ double tile[];
    const int plane = blockIdx.y;
    const double* __restrict__ in = a.in[plane];
    double* out = a.out[plane];
    const int L = a.low_bits, H0 = a.high_start, M = a.high_bits;
    const int T = L + M;
    const unsigned tile_elems = 1u << T;
    const unsigned low_mask = (1u << L) - 1u;
    const int gap = H0 - L;

    for (unsigned long long t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const I t_lo = (I)(t & ((1ull << gap) - 1ull));
        const I t_hi = (I)(t >> gap);
        const I base = (t_lo << L) | (t_hi << (H0 + M));
        for (unsigned y = threadIdx.x; y < tile_elems; y += kPassThreads) {
            const I x = base | (I)(y & low_mask) | ((I)(y >> L) << H0);
            tile[y] = in[x];
        }
        __syncthreads();
        for (unsigned y = threadIdx.x; y < tile_elems; y += kPassThreads) {
            const I x = base | (I)(y & low_mask) | ((I)(y >> L) << H0);
            const unsigned long long act_all =
                activity_word<unsigned long long>(expand_index((unsigned long long)x, a.shard), a.distance, a.interval_mask);
            double acc = 0.0;
            for (int q = 0; q < T; ++q) {
                const int g = q < L ? q : H0 + (q - L);            // local bit of tile bit q
                if (((a.flip_mask >> g) & 1ull) && ((act_all >> global_pos(g, a.shard)) & 1ull)) {
                    const double v = tile[y ^ (1u << q)];
                    acc += ((y >> q) & 1u) ? -v : v;  // K = sum_c P_c (sigma^-  -  sigma^+)_c
                }
            }
            double r = a.gamma * acc;
            for (int k = 0; k < a.nstreams; ++k) {
                const EpiStream& st = a.s[k];
                if (st.bit < 0 || ((act_all >> st.bit) & 1ull)) r += st.coef * st.ptr[plane][x];
            }
            out[x] = r;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Fast kernel
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 ldg_stream(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ldg_rw(const double* p) {
    double2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(double* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- mbarrier / bulk-copy (TMA) primitives for the remote operand ring ----------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
static __device__ __noinline__ void mbar_wait_slow(unsigned bar, unsigned parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 40000000000ll) {  // ~20 s: a partner's memory never answered; fail instead of hanging
            printf("qca_b200: remote operand ring timed out (block %u)\n", blockIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
// true in exactly one lane of the (converged) warp; tells the compiler so (no per-lane serialisation of the
// uniform-datapath instructions in the branch it guards)
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
// global -> shared bulk copy (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Stage a tile global -> shared with bulk copies (TMA engine): one per contiguous piece of 2^L doubles (at most 8 KiB),
// all completing on one mbarrier, which every thread then waits for.  Replaces one LDG.128 per thread and register
// row: on one GPU this took the 14-bit kernel from 0.79 to 0.86 of the measured HBM peak (ncu had shown the LSU's
// instruction queue -- lg_throttle -- pacing the staging, not DRAM; profiles/r02_ab_v3_*.json).
// `bar` must have been initialised with count 1; `parity` is the phase this use completes (0, 1, 0, ... per use).
template <typename I, int L, int TILE_BITS, int THREADS>
__device__ __forceinline__ void stage_tile_bulk(double* tile, const double* in, I base, int H0, unsigned tid, unsigned bar,
                                                unsigned parity) {
    constexpr int PIECE_BITS = L < 10 ? L : 10;
    constexpr int NPIECES = 1 << (TILE_BITS - PIECE_BITS);
    constexpr unsigned low_mask = (1u << L) - 1u;
    if (tid == 0) mbar_arrive_expect_tx(bar, 8u << TILE_BITS);
#ifndef QCA_THREAD_ISSUE
    // one elected lane per warp issues the warp's share in a warp-uniform loop (see pass_kernel_v3)
    constexpr int WARPS = THREADS / 32;
    constexpr int PER_WARP = NPIECES >= WARPS ? NPIECES / WARPS : 1;
    const unsigned wu = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if ((NPIECES >= WARPS || wu < (unsigned)NPIECES) && elect_one()) {
#pragma unroll 4
        for (int i = 0; i < PER_WARP; ++i) {
            const unsigned y = (wu * PER_WARP + (unsigned)i) << PIECE_BITS;
            const I x = base | (I)(y & low_mask) | ((I)(y >> L) << H0);
            bulk_g2s(smem_u32(tile + y), in + x, 8u << PIECE_BITS, bar);
        }
    }
    __syncwarp();
#else
#pragma unroll 1
    for (int r = (int)tid; r < NPIECES; r += THREADS) {
        const unsigned y = (unsigned)r << PIECE_BITS;   // tile-local index of the first amplitude of the piece
        const I x = base | (I)(y & low_mask) | ((I)(y >> L) << H0);
        bulk_g2s(smem_u32(tile + y), in + x, 8u << PIECE_BITS, bar);
    }
#endif
    mbar_wait(bar, parity);
}

// Epilogue operands are streamed through rings of shared memory, kRingRowsTotal rows of 4 KiB
// (256 threads x 16 B), filled from kernel entry on: deep memory-level parallelism at no register cost.
//   * local operands (HBM, ~1 us): per-thread cp.async, one commit group per row;
//   * remote operands (partner rank over NVLink, several us): bulk copies issued per warp -- the 32
//     pairs of a warp's row are 512 contiguous bytes (128-byte pieces when the tile keeps fewer than 6
//     low bits) -- each row completing on its own mbarrier.  The two mechanisms are independent: a
//     slow NVLink row never holds back the HBM rows behind it (with cp.async groups, which retire in
//     order, it did), and the remote ring runs a full DC rows ahead.  Which rows a slot is active on
//     (and for which term) is decided once per tile in a compact loop, two words per thread: the
//     per-row cost of an idle slot is one predicate (the first version re-derived everything per
//     row and tripled the instruction count of the launch: ncu, profiles/r01_pass_v4_*).
constexpr int kRingRowsTotal = 12;   // 64 KiB tile + 48 KiB ring = 112 KiB: two CTAs per SM, not a byte to spare
constexpr int kPassWarps = kPassThreads / 32;
constexpr int kPassMbarBytes = kPassWarps * 8 * 8;  // up to 8 mbarriers per warp (slots x rows)
constexpr int kPassSmemBytes = (8 << kTile) + kRingRowsTotal * kPassThreads * 16;
// with remote slots the last ring row is given up for the mbarriers (occupancy would halve otherwise)
__host__ __device__ constexpr int ring_rows(int nrem) { return nrem == 0 ? kRingRowsTotal : kRingRowsTotal - 1; }

// ring depth (rows of look-ahead) of a remote / a local operand: one remote slot gets RD rows (6 by
// default: 4.40 against 4.32 steps/s with RD = 4 at N = 30 on 8 GPUs; QCA_REMOTE_RING=4 selects the
// latter), two slots 4 rows each; the local operands share the rest.
__host__ __device__ constexpr int ring_rem(int nrem, int rd) {
    return nrem == 0 ? 0 : (nrem == 1 ? rd : 4);
}
__host__ __device__ constexpr int ring_unc(int nunc, int nrem, int rd) {
    return nunc == 0 ? 0 : (ring_rows(nrem) - nrem * ring_rem(nrem, rd)) / nunc;
}

// L: contiguous low bits of the tile, M = 13 - L strided bits at H0.
// FLIP_LOW: pass 0 (M == 0, L == 13): every local bit is flipped.  Otherwise only the M high bits are.
// NUNC local epilogue operands (recurrence vectors), NREM remote slots (terms of sharded qubits).
template <typename I, int L, bool FLIP_LOW, int NUNC, int NREM, int RD = 4>
__global__ void __launch_bounds__(kPassThreads, 2) pass_kernel_v2(const PassArgs a) {
    constexpr int M = kTile - L;
    constexpr int QLO = FLIP_LOW ? 0 : L;  // first flipped local bit
    constexpr int DU = ring_unc(NUNC, NREM, RD), DC = ring_rem(NREM, RD);   // look-ahead per operand kind
    constexpr int CHUNK_LANES = (L >= 6) ? 32 : (1 << (L - 1));  // lanes whose pairs are contiguous in global memory
    constexpr int ISSUERS = 32 / CHUNK_LANES;                    // bulk copies per warp and row
    static_assert(NUNC * DU + NREM * DC <= ring_rows(NREM), "ring budget");
    static_assert(NUNC == 0 || DU >= 1, "local ring");
    static_assert((32 / ((L >= 6) ? 32 : (1 << (L - 1)))) * NREM * DC <= 32, "mbarrier budget: 32 per warp");
    static_assert(FLIP_LOW ? (M == 0) : (M >= 1), "geometry");
    static_assert(NUNC >= 0 && NUNC <= 2 && NREM >= 0 && NREM <= kMaxRemoteSlots, "operands");
    static_assert(L >= 4, "rows of at least 128 bytes");
    extern __shared__ double tile[];
    const int plane = blockIdx.y;
    const double* __restrict__ in = a.in[plane];
    double* out = a.out[plane];
    const int H0 = a.high_start;
    const int d = a.distance;
    const unsigned tid = threadIdx.x;
    const unsigned lane = tid & 31u, warp = tid >> 5;
    constexpr unsigned low_mask = (1u << L) - 1u;

    // tile base and this thread's part of the index
    const unsigned long long t = blockIdx.x;
    const int gap = H0 - L;
    const I t_lo = (I)(t & ((1ull << gap) - 1ull));
    const I t_hi = (I)(t >> gap);
    const I base = (t_lo << L) | (t_hi << (H0 + M));
    const unsigned y_thr = tid << 1;  // local bits 1..8
    const I x_thr = base | (I)(y_thr & low_mask) | ((I)(y_thr >> L) << H0);

    double2* tile2 = reinterpret_cast<double2*>(tile);
    double2* ring = tile2 + (1 << (kTile - 1));            // local operand k: rows [k*DU, (k+1)*DU)
    double2* rring = ring + NUNC * DU * kPassThreads;      // remote slot k: rows [k*DC, (k+1)*DC)
    const unsigned mbar0 = smem_u32(ring + (kRingRowsTotal - 1) * kPassThreads) + warp * 256u;  // this warp's mbarriers (last ring row)
    auto row_x = [&](int e) -> I {  // index of the pair (row e, this thread); folds to x_thr | const << H0
        const unsigned ye = (unsigned)e << kRowShift;
        return x_thr | (I)(ye & low_mask) | ((I)(ye >> L) << H0);
    };
    auto local_slot = [&](int k, int e) -> double2* {   // k, e are compile-time after unrolling
        return ring + ((k * DU + (e % DU)) * kPassThreads + tid);
    };
    auto remote_slot = [&](int k, int e) -> double2* {
        return rring + ((k * DC + (e % DC)) * kPassThreads + tid);
    };
    constexpr int NBAR = NREM * DC;                      // mbarriers per piece
    const unsigned piece = lane / CHUNK_LANES;           // which contiguous piece of the warp's row this lane reads
    auto remote_bar = [&](int k, int e) -> unsigned { return mbar0 + (piece * NBAR + (unsigned)(k * DC + (e % DC))) * 8u; };
    // Per tile: on which of its 16 rows each slot is active for this thread (act[k], bit e) and with which
    // rotation (rot0/rot1: low/high bit per row).  Both read index bits >= 13 only, so they are constant over
    // a piece: the lanes of a piece evaluate one row each (two for 8-lane pieces) and share the answers
    // through warp votes -- ~25 instructions per tile instead of ~20 per row.
    unsigned act[NREM ? NREM : 1] = {0};
    unsigned rot0 = 0, rot1 = 0;
    if (NREM) {
        constexpr int ROWS_PER_LANE = (CHUNK_LANES >= 16) ? 1 : 16 / CHUNK_LANES;
        constexpr int FIELD = (CHUNK_LANES >= 16) ? 16 : CHUNK_LANES;      // rows answered by one vote
        const unsigned field_shift = (CHUNK_LANES >= 32) ? 0u : piece * CHUNK_LANES;
#pragma unroll
        for (int q = 0; q < ROWS_PER_LANE; ++q) {
            const unsigned e = ((lane & (CHUNK_LANES - 1)) + q * CHUNK_LANES) & 15u;
            const unsigned ye = e << kRowShift;
            const I x = x_thr | (I)(ye & low_mask) | ((I)(ye >> L) << H0);
            const unsigned rot = (a.rot_word >> (2u * ((unsigned)(x >> a.rot_shift) & 15u))) & 3u;
            const unsigned v0 = __ballot_sync(0xffffffffu, rot & 1u), v1 = __ballot_sync(0xffffffffu, rot & 2u);
            rot0 |= ((v0 >> field_shift) & ((1u << FIELD) - 1u)) << (q * FIELD);
            rot1 |= ((v1 >> field_shift) & ((1u << FIELD) - 1u)) << (q * FIELD);
#pragma unroll
            for (int k = 0; k < NREM; ++k) {
                const RemoteAlt* al = &a.rs[k].alt[rot];
                const unsigned vk = __ballot_sync(0xffffffffu, (al->mask >> ((unsigned)(x >> al->shift) & 15u)) & 1u);
                act[k] |= ((vk >> field_shift) & ((1u << FIELD) - 1u)) << (q * FIELD);
            }
        }
    }
    auto row_rot = [&](int e) -> unsigned { return ((rot0 >> e) & 1u) | (((rot1 >> e) & 1u) << 1); };
    // earlier ACTIVE uses of the ring slot of row e decide the phase of its mbarrier (idle rows never arrive)
    auto uses_before = [](int e) -> unsigned {   // bits e' < e with e' == e (mod DC); folds to a constant
        unsigned m = 0;
        for (int q = e % (DC ? DC : 1); q < e; q += (DC ? DC : 1)) m |= 1u << q;
        return m;
    };
    // row e of every remote slot: the first lane of each contiguous piece issues its bulk copy
    auto remote_issue = [&](int e) {
#pragma unroll
        for (int k = 0; k < NREM; ++k) {
            if (((act[k] >> e) & 1u) && (lane & (CHUNK_LANES - 1)) == 0) {
                const RemoteAlt* al = &a.rs[k].alt[row_rot(e)];
                const unsigned bar = remote_bar(k, e);
                mbar_arrive_expect_tx(bar, CHUNK_LANES * 16);
                bulk_g2s(smem_u32(remote_slot(k, e)), al->ptr[plane] + row_x(e), CHUNK_LANES * 16, bar);
            }
        }
    };
    auto local_fetch = [&](int e) {
#pragma unroll
        for (int k = 0; k < NUNC; ++k) cp_async16(local_slot(k, e), a.s[k].ptr[plane] + row_x(e));
    };
    // ---- start streaming the epilogue operands: the far ones first ---------------------------------
    if (NREM) {
        if (lane < ISSUERS * NBAR) mbar_init(mbar0 + lane * 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int e = 0; e < DC; ++e) remote_issue(e);
    }
    if (NUNC) {
#pragma unroll
        for (int e = 0; e < DU; ++e) {
            local_fetch(e);
            cp_async_commit();
        }
    }
    // ---- stage the tile: 16 rows x one 16-byte pair per thread -------------------------------
    double2 v[kRows];
    if (NREM) {   // sharded launches: bulk copies (the mbarrier sits in the unused upper half of the mbarrier row)
        const unsigned tile_bar = smem_u32(ring + (kRingRowsTotal - 1) * kPassThreads) + 2048u;
        if (tid == 0) {
            mbar_init(tile_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        stage_tile_bulk<I, L, kTile, kPassThreads>(tile, in, base, H0, tid, tile_bar, 0u);
#pragma unroll
        for (int e = 0; e < kRows; ++e) v[e] = tile2[(e << (kRowShift - 1)) | tid];
    } else {
#pragma unroll
        for (int e = 0; e < kRows; ++e) v[e] = ldg_stream(in + row_x(e));
#pragma unroll
        for (int e = 0; e < kRows; ++e) tile2[(e << (kRowShift - 1)) | tid] = v[e];
    }

    // ---- rule predicate words ------------------------------------------------------------------
    // la[j] bit q: local bit q of element (e, j) is flipped by an active term
    unsigned lo_act0 = 0, lo_act1 = 0;
    if (FLIP_LOW) {
        // bits [0, 9-d): window = x bits [0,9) shifted up by d (dead cells below the chain end)
        const unsigned w = ((unsigned)x_thr & 0x1ffu) << d;
        lo_act0 = a.tab_lo[w];
        lo_act1 = a.tab_lo[w | (1u << d)];
    }
    // thread-bit signs: -1 where this thread's bit is set (sigma^- - sigma^+)
    double sgn[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) sgn[b] = ((tid >> b) & 1u) ? -1.0 : 1.0;

    // global index of this thread's pairs without the row bits (sharded qubits of this rank inserted)
    const unsigned long long xg_thr = expand_index((unsigned long long)x_thr, a.shard);

    __syncthreads();

#pragma unroll
    for (int e = 0; e < kRows; ++e) {
        unsigned la0, la1;
        // windows are cut out of the GLOBAL index (sharded qubits of this rank inserted); every
        // sharded position is >= 13, so local bits 0..12 are global bits 0..12
        const unsigned long long xg = xg_thr | a.row_xg[e];
        if (FLIP_LOW) {
            const int S = 9 - d;
            const unsigned w = (unsigned)(xg >> (9 - 2 * d)) & ((1u << (4 + 3 * d)) - 1u);
            const unsigned hi_act = a.tab_hi[w];
            la0 = lo_act0 | (hi_act << S);
            la1 = lo_act1 | (hi_act << S);
        } else {
            const unsigned w = (unsigned)(xg >> a.win_shift) & a.win_mask;
            la0 = la1 = (unsigned)a.tab_hi[w] << L;
        }
        double acc0 = 0.0, acc1 = 0.0;
        // local bit 0: the other half of the pair
        if (QLO == 0) {
            if (la0 & 1u) acc0 += v[e].y;
            if (la1 & 1u) acc1 -= v[e].x;
        }
        // local bits 1..8: partner thread, same row
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (b + 1 >= QLO) {
                const double2 p = tile2[(e << (kRowShift - 1)) | (tid ^ (1u << b))];
                if (la0 & (2u << b)) acc0 = fma(p.x, sgn[b], acc0);
                if (la1 & (2u << b)) acc1 = fma(p.y, sgn[b], acc1);
            }
        }
        // local bits 9..12: partner row, same thread (registers)
#pragma unroll
        for (int k = 0; k < kRegHigh; ++k) {
            if (kRowShift + k >= QLO) {
                const double2 p = v[e ^ (1 << k)];
                if ((e >> k) & 1) {
                    if (la0 & (1u << (kRowShift + k))) acc0 -= p.x;
                    if (la1 & (1u << (kRowShift + k))) acc1 -= p.y;
                } else {
                    if (la0 & (1u << (kRowShift + k))) acc0 += p.x;
                    if (la1 & (1u << (kRowShift + k))) acc1 += p.y;
                }
            }
        }
        double2 r;
        r.x = a.gamma * acc0;
        r.y = a.gamma * acc1;
        if (NUNC) {
            cp_async_wait<(NUNC ? DU : 1) - 1>();  // row e of every local operand has landed (own data: no barrier needed)
#pragma unroll
            for (int k = 0; k < NUNC; ++k) {
                const double2 sv = *local_slot(k, e);
                r.x = fma(a.s[k].coef, sv.x, r.x);
                r.y = fma(a.s[k].coef, sv.y, r.y);
            }
        }
        if (NREM) {
#pragma unroll
            for (int k = 0; k < NREM; ++k) {
                if ((act[k] >> e) & 1u) {
                    mbar_wait(remote_bar(k, e), __popc(act[k] & uses_before(e)) & 1u);
                    const double2 sv = *remote_slot(k, e);
                    const double coef = a.rs[k].alt[row_rot(e)].coef;
                    r.x = fma(coef, sv.x, r.x);
                    r.y = fma(coef, sv.y, r.y);
                }
            }
        }
        stg_stream(out + row_x(e), r);
        if (NUNC) {
            if (e + DU < kRows) local_fetch(e + DU);  // refill the slots just consumed
            cp_async_commit();
        }
        if (NREM) {
            if (e + DC < kRows) {
                __syncwarp();  // every lane has read row e of the remote ring: its slots may be overwritten
                remote_issue(e + DC);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Fast kernel, sharded registers, PERSISTENT variant.
//
// pass_kernel_v2 gives every tile its own CTA: the partner rows of a tile are requested at CTA entry and the
// first of them is needed 2-3 us later, sooner than NVLink answers, so every tile starts with a stall (about a
// quarter of the sharded launch at 8 GPUs).  Here a CTA walks over tiles t, t + gridDim.x, ... and its
// operand rings run on across the tile boundary: while the last DC (DU) rows of a tile are processed, the first
// rows of the NEXT tile are already being fetched -- remote rows by bulk copies over NVLink, local rows by
// cp.async -- so in steady state a remote row has DC row times (5-6 us) to arrive.  Ring depths divide 16 so the
// slot of a row is the same in every tile; the phase of an mbarrier is the parity of the active uses of its slot
// so far: the uses inside the tile (a popcount against a compile-time mask) plus a carried bit per slot.
// ---------------------------------------------------------------------------
template <typename I, int L, bool FLIP_LOW, int NUNC, int NREM, int DC>
__global__ void __launch_bounds__(kPassThreads, 2) pass_kernel_v2p(const PassArgs a) {
    constexpr int M = kTile - L;
    constexpr int QLO = FLIP_LOW ? 0 : L;
    constexpr int DU_FREE = NUNC == 0 ? 0 : (ring_rows(NREM) - NREM * DC) / (NUNC ? NUNC : 1);
    constexpr int DU = DU_FREE >= 8 ? 8 : (DU_FREE >= 4 ? 4 : (DU_FREE >= 2 ? 2 : (DU_FREE >= 1 ? 1 : 0)));
    constexpr int CHUNK_LANES = (L >= 6) ? 32 : (1 << (L - 1));
    constexpr int ISSUERS = 32 / CHUNK_LANES;
    static_assert(NREM >= 1 && NREM <= kMaxRemoteSlots && (DC == 4 || DC == 8), "remote ring");
    static_assert(NUNC == 0 || DU >= 1, "local ring");
    static_assert(NUNC * DU + NREM * DC <= ring_rows(NREM), "ring budget");
    static_assert(ISSUERS * NREM * DC <= 32, "mbarrier budget: 32 per warp");
    static_assert(FLIP_LOW ? (M == 0) : (M >= 1), "geometry");
    static_assert(L >= 4, "rows of at least 128 bytes");
    extern __shared__ double tile[];
    const int plane = blockIdx.y;
    const double* __restrict__ in = a.in[plane];
    double* out = a.out[plane];
    const int H0 = a.high_start;
    const int d = a.distance;
    const unsigned tid = threadIdx.x;
    const unsigned lane = tid & 31u, warp = tid >> 5;
    constexpr unsigned low_mask = (1u << L) - 1u;
    const int gap = H0 - L;
    const unsigned y_thr = tid << 1;

    double2* tile2 = reinterpret_cast<double2*>(tile);
    double2* ring = tile2 + (1 << (kTile - 1));
    double2* rring = ring + NUNC * DU * kPassThreads;
    const unsigned mbar0 = smem_u32(ring + (kRingRowsTotal - 1) * kPassThreads) + warp * 256u;
    constexpr int NBAR = NREM * DC;
    const unsigned piece = lane / CHUNK_LANES;

    // everything that depends on the tile
    struct TileCtx { I base, x_thr; unsigned act[NREM]; unsigned rot0, rot1; };
    auto setup = [&](unsigned long long t, TileCtx& c) {
        const I t_lo = (I)(t & ((1ull << gap) - 1ull));
        const I t_hi = (I)(t >> gap);
        const I base = (t_lo << L) | (t_hi << (H0 + M));
        c.base = base;
        c.x_thr = base | (I)(y_thr & low_mask) | ((I)(y_thr >> L) << H0);
        c.rot0 = c.rot1 = 0;
#pragma unroll
        for (int k = 0; k < NREM; ++k) c.act[k] = 0;
        constexpr int ROWS_PER_LANE = (CHUNK_LANES >= 16) ? 1 : 16 / CHUNK_LANES;
        constexpr int FIELD = (CHUNK_LANES >= 16) ? 16 : CHUNK_LANES;
        const unsigned field_shift = (CHUNK_LANES >= 32) ? 0u : piece * CHUNK_LANES;
#pragma unroll
        for (int q = 0; q < ROWS_PER_LANE; ++q) {
            const unsigned e = ((lane & (CHUNK_LANES - 1)) + q * CHUNK_LANES) & 15u;
            const unsigned ye = e << kRowShift;
            const I x = c.x_thr | (I)(ye & low_mask) | ((I)(ye >> L) << H0);
            const unsigned rot = (a.rot_word >> (2u * ((unsigned)(x >> a.rot_shift) & 15u))) & 3u;
            const unsigned v0 = __ballot_sync(0xffffffffu, rot & 1u), v1 = __ballot_sync(0xffffffffu, rot & 2u);
            c.rot0 |= ((v0 >> field_shift) & ((1u << FIELD) - 1u)) << (q * FIELD);
            c.rot1 |= ((v1 >> field_shift) & ((1u << FIELD) - 1u)) << (q * FIELD);
#pragma unroll
            for (int k = 0; k < NREM; ++k) {
                const RemoteAlt* al = &a.rs[k].alt[rot];
                const unsigned vk = __ballot_sync(0xffffffffu, (al->mask >> ((unsigned)(x >> al->shift) & 15u)) & 1u);
                c.act[k] |= ((vk >> field_shift) & ((1u << FIELD) - 1u)) << (q * FIELD);
            }
        }
    };
    auto row_off = [&](int e) -> I {   // index bits of register row e (compile-time after unrolling)
        const unsigned ye = (unsigned)e << kRowShift;
        return (I)(ye & low_mask) | ((I)(ye >> L) << H0);
    };
    auto local_slot = [&](int k, int e) -> double2* { return ring + ((k * DU + (e % (DU ? DU : 1))) * kPassThreads + tid); };
    auto remote_slot = [&](int k, int e) -> double2* { return rring + ((k * DC + (e % DC)) * kPassThreads + tid); };
    auto remote_bar = [&](int k, int e) -> unsigned { return mbar0 + (piece * NBAR + (unsigned)(k * DC + (e % DC))) * 8u; };
    auto uses_before = [](int e) -> unsigned {   // rows e' < e of the same tile in the same ring slot
        unsigned m = 0;
        for (int q = e % DC; q < e; q += DC) m |= 1u << q;
        return m;
    };
    auto slot_class = [](int s) -> unsigned {    // all rows of a tile that use ring slot s
        unsigned m = 0;
        for (int q = s; q < kRows; q += DC) m |= 1u << q;
        return m;
    };
    auto remote_issue = [&](const TileCtx& c, int e) {
#pragma unroll
        for (int k = 0; k < NREM; ++k) {
            if (((c.act[k] >> e) & 1u) && (lane & (CHUNK_LANES - 1)) == 0) {
                const unsigned rot = ((c.rot0 >> e) & 1u) | (((c.rot1 >> e) & 1u) << 1);
                const RemoteAlt* al = &a.rs[k].alt[rot];
                const unsigned bar = remote_bar(k, e);
                mbar_arrive_expect_tx(bar, CHUNK_LANES * 16);
                bulk_g2s(smem_u32(remote_slot(k, e)), al->ptr[plane] + (c.x_thr | row_off(e)), CHUNK_LANES * 16, bar);
            }
        }
    };
    auto local_fetch = [&](const TileCtx& c, int e) {
#pragma unroll
        for (int k = 0; k < NUNC; ++k) cp_async16(local_slot(k, e), a.s[k].ptr[plane] + (c.x_thr | row_off(e)));
    };

    unsigned long long t = blockIdx.x;
    if (t >= a.ntiles) return;
    TileCtx cur, nxt;
    setup(t, cur);
    unsigned carry[NREM];   // bit s: parity of the active uses of ring slot s in the tiles done so far
#pragma unroll
    for (int k = 0; k < NREM; ++k) carry[k] = 0;
    // ---- prologue: barriers, first rows of the first tile ----------------------------------------------
    const unsigned tile_bar = smem_u32(ring + (kRingRowsTotal - 1) * kPassThreads) + 2048u;   // upper half of the mbarrier row
    if (lane < ISSUERS * NBAR) mbar_init(mbar0 + lane * 8u, 1);
    if (tid == 0) mbar_init(tile_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    unsigned tile_phase = 0;
#pragma unroll
    for (int e = 0; e < DC; ++e) remote_issue(cur, e);
    if (NUNC) {
#pragma unroll
        for (int e = 0; e < DU; ++e) {
            local_fetch(cur, e);
            cp_async_commit();
        }
    }
    double sgn[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) sgn[b] = ((tid >> b) & 1u) ? -1.0 : 1.0;

    for (;;) {
        const unsigned long long t_next = t + gridDim.x;
        const bool has_next = t_next < a.ntiles;   // uniform over the CTA
        if (has_next) setup(t_next, nxt);
        // ---- stage the tile: bulk copies onto one mbarrier (stage_tile_bulk), then the thread's 16 pairs ------------
        double2 v[kRows];
        stage_tile_bulk<I, L, kTile, kPassThreads>(tile, in, cur.base, H0, tid, tile_bar, tile_phase);
        tile_phase ^= 1u;
#pragma unroll
        for (int e = 0; e < kRows; ++e) v[e] = tile2[(e << (kRowShift - 1)) | tid];
        unsigned lo_act0 = 0, lo_act1 = 0;
        if (FLIP_LOW) {
            const unsigned w = ((unsigned)cur.x_thr & 0x1ffu) << d;
            lo_act0 = a.tab_lo[w];
            lo_act1 = a.tab_lo[w | (1u << d)];
        }
        const unsigned long long xg_thr = expand_index((unsigned long long)cur.x_thr, a.shard);
        __syncthreads();

#pragma unroll
        for (int e = 0; e < kRows; ++e) {
            unsigned la0, la1;
            const unsigned long long xg = xg_thr | a.row_xg[e];
            if (FLIP_LOW) {
                const int S = 9 - d;
                const unsigned w = (unsigned)(xg >> (9 - 2 * d)) & ((1u << (4 + 3 * d)) - 1u);
                const unsigned hi_act = a.tab_hi[w];
                la0 = lo_act0 | (hi_act << S);
                la1 = lo_act1 | (hi_act << S);
            } else {
                const unsigned w = (unsigned)(xg >> a.win_shift) & a.win_mask;
                la0 = la1 = (unsigned)a.tab_hi[w] << L;
            }
            double acc0 = 0.0, acc1 = 0.0;
            if (QLO == 0) {
                if (la0 & 1u) acc0 += v[e].y;
                if (la1 & 1u) acc1 -= v[e].x;
            }
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (b + 1 >= QLO) {
                    const double2 p = tile2[(e << (kRowShift - 1)) | (tid ^ (1u << b))];
                    if (la0 & (2u << b)) acc0 = fma(p.x, sgn[b], acc0);
                    if (la1 & (2u << b)) acc1 = fma(p.y, sgn[b], acc1);
                }
            }
#pragma unroll
            for (int k = 0; k < kRegHigh; ++k) {
                if (kRowShift + k >= QLO) {
                    const double2 p = v[e ^ (1 << k)];
                    if ((e >> k) & 1) {
                        if (la0 & (1u << (kRowShift + k))) acc0 -= p.x;
                        if (la1 & (1u << (kRowShift + k))) acc1 -= p.y;
                    } else {
                        if (la0 & (1u << (kRowShift + k))) acc0 += p.x;
                        if (la1 & (1u << (kRowShift + k))) acc1 += p.y;
                    }
                }
            }
            double2 r;
            r.x = a.gamma * acc0;
            r.y = a.gamma * acc1;
            if (NUNC) {
                cp_async_wait<(NUNC ? DU : 1) - 1>();   // one commit group per row, in every tile: row e has landed
#pragma unroll
                for (int k = 0; k < NUNC; ++k) {
                    const double2 sv = *local_slot(k, e);
                    r.x = fma(a.s[k].coef, sv.x, r.x);
                    r.y = fma(a.s[k].coef, sv.y, r.y);
                }
            }
#pragma unroll
            for (int k = 0; k < NREM; ++k) {
                if ((cur.act[k] >> e) & 1u) {
                    const unsigned parity = (__popc(cur.act[k] & uses_before(e)) + ((carry[k] >> (e % DC)) & 1u)) & 1u;
                    mbar_wait(remote_bar(k, e), parity);
                    const double2 sv = *remote_slot(k, e);
                    const unsigned rot = ((cur.rot0 >> e) & 1u) | (((cur.rot1 >> e) & 1u) << 1);
                    const double coef = a.rs[k].alt[rot].coef;
                    r.x = fma(coef, sv.x, r.x);
                    r.y = fma(coef, sv.y, r.y);
                }
            }
            stg_stream(out + (cur.x_thr | row_off(e)), r);
            // refill the slots just consumed: later rows of this tile, then the first rows of the next one
            if (NUNC) {
                if (e + DU < kRows) local_fetch(cur, e + DU);
                else if (has_next) local_fetch(nxt, e + DU - kRows);
                cp_async_commit();
            }
            __syncwarp();   // every lane has read row e of the remote ring: its slots may be overwritten
            if (e + DC < kRows) remote_issue(cur, e + DC);
            else if (has_next) remote_issue(nxt, e + DC - kRows);
        }
        if (!has_next) break;
#pragma unroll
        for (int k = 0; k < NREM; ++k) {
            unsigned flip = 0;
#pragma unroll
            for (int s = 0; s < DC; ++s) flip |= (unsigned)(__popc(cur.act[k] & slot_class(s)) & 1) << s;
            carry[k] ^= flip;
        }
        t = t_next;
        cur = nxt;
        __syncthreads();   // everybody has read the staged tile: the next one may overwrite it
    }
}

// kernel tables, one translation unit per index type (qca_pass_u32.cu / qca_pass_u64.cu)
typedef void (*PassKernel)(const PassArgs);
PassKernel fast_pass_kernel_u32(int low_bits, int nunc, int nrem, int remote_rows);
PassKernel fast_pass_kernel_u64(int low_bits, int nunc, int nrem, int remote_rows);
PassKernel persistent_pass_kernel_u32(int low_bits, int nunc, int nrem, int remote_rows);
PassKernel persistent_pass_kernel_u64(int low_bits, int nunc, int nrem, int remote_rows);
PassKernel generic_pass_kernel(bool wide);

}  // namespace qca
