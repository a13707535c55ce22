// Fused measurement of a single-plane resident state (MPS.measure, tensor_networks/mps.py:100-140).
//
// For cell c (index bit q) the reduced density matrix needs  s0 = sum_{x: bit q = 0} |phi|^2,
// s1 = sum_{x: bit q = 1} |phi|^2  and  w = sum_{pairs} phi[x0] phi[x1]  (x1 = x0 | 1 << q; real because the
// rotated state is real).  The per-cell kernel (measure_pairs_kernel) reads the whole vector once per cell:
// N reads.  Here one launch per TILE PASS (the same 2^13-amplitude tiles as the rule operator: bits
// [0, L) U [H0, H0 + 13 - L)) produces the sums of every cell whose bit varies inside the tile: 3 reads of
// the vector at N = 30 instead of 30.
//
// Thread mapping as in pass_kernel_v2: 256 threads, 32 amplitudes per thread in registers (bit 0 = the two
// halves of a 16-byte load, tile bits 9..12 = 16 register rows), tile bits 1..8 index the thread and their
// partners are read from the staged tile with one LDS.128.  Persistent grid (accumulators live in
// registers across tiles), per-block partial sums, fixed-order final reduction: deterministic.
#include <cuda_runtime.h>

#include <algorithm>

#include "qca_common.cuh"
#include "qca_measure.h"

namespace qca {

constexpr int kMsThreads = 256;
constexpr int kMsRows = 16;
constexpr int kMsRowShift = 9;
constexpr int kMsTileBits = 13;
constexpr int kMsVals = 27;   // s1 of tile bit t (0..12), w of tile bit t (13..25), total norm (26)

struct MeasureTileArgs {
    const double* re;
    unsigned long long ntiles;
    int low_bits, high_start;
    int first_bit;      // tile bits >= first_bit are measured by this launch (0: all 13, L: the strided high bits)
    double* partials;   // [gridDim.x][kMsVals]
};

__global__ void __launch_bounds__(kMsThreads, 2) measure_tile_kernel(const MeasureTileArgs a) {
    extern __shared__ double2 ms_tile[];   // 16 rows x 256 pairs = 64 KiB
    const unsigned tid = threadIdx.x;
    const int L = a.low_bits, H0 = a.high_start, M = kMsTileBits - L;
    const unsigned long long low_mask = (1ull << L) - 1ull;
    const int gap = H0 - L;
    const int fb = a.first_bit;
    double w[kMsTileBits];
#pragma unroll
    for (int t = 0; t < kMsTileBits; ++t) w[t] = 0.0;
    double s1row[4] = {0.0, 0.0, 0.0, 0.0};
    double s1b0 = 0.0, tot = 0.0;

    for (unsigned long long t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const unsigned long long t_lo = t & ((1ull << gap) - 1ull), t_hi = t >> gap;
        const unsigned long long base = (t_lo << L) | (t_hi << (H0 + M));
        const unsigned long long y_thr = (unsigned long long)tid << 1;
        const unsigned long long x_thr = base | (y_thr & low_mask) | ((y_thr >> L) << H0);
        double2 v[kMsRows];
#pragma unroll
        for (int e = 0; e < kMsRows; ++e) {
            const unsigned long long ye = (unsigned long long)e << kMsRowShift;
            const unsigned long long x = x_thr | (ye & low_mask) | ((ye >> L) << H0);
            v[e] = *reinterpret_cast<const double2*>(a.re + x);
        }
#pragma unroll
        for (int e = 0; e < kMsRows; ++e) ms_tile[(e << 8) | tid] = v[e];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kMsRows; ++e) {
            const double nx = v[e].x * v[e].x, ny = v[e].y * v[e].y;
            tot += nx + ny;
            s1b0 += ny;
            w[0] = fma(v[e].x, v[e].y, w[0]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if ((e >> k) & 1) {
                    s1row[k] += nx + ny;
                } else {
                    const double2 p = v[e | (1 << k)];
                    w[kMsRowShift + k] += v[e].x * p.x + v[e].y * p.y;
                }
            }
        }
        // thread bits: every pair is seen from both of its threads (halved in the end)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (b + 1 >= fb) {
#pragma unroll
                for (int e = 0; e < kMsRows; ++e) {
                    const double2 p = ms_tile[(e << 8) | (tid ^ (1u << b))];
                    w[b + 1] += v[e].x * p.x + v[e].y * p.y;
                }
            }
        }
        __syncthreads();
    }

    // per-block partial sums, fixed order
    double vals[kMsVals];
    vals[0] = s1b0;
#pragma unroll
    for (int b = 0; b < 8; ++b) vals[1 + b] = ((tid >> b) & 1u) ? tot : 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) vals[kMsRowShift + k] = s1row[k];
    vals[kMsTileBits + 0] = w[0];
#pragma unroll
    for (int b = 0; b < 8; ++b) vals[kMsTileBits + 1 + b] = 0.5 * w[1 + b];
#pragma unroll
    for (int k = 0; k < 4; ++k) vals[kMsTileBits + kMsRowShift + k] = w[kMsRowShift + k];
    vals[26] = tot;
    double* red = reinterpret_cast<double*>(ms_tile);   // [kMsVals][8]
    const unsigned lane = tid & 31u, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < kMsVals; ++i) {
        double s = vals[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[i * 8 + warp] = s;
    }
    __syncthreads();
    if (tid < kMsVals) {
        double s = 0.0;
        for (int q = 0; q < kMsThreads / 32; ++q) s += red[tid * 8 + q];
        a.partials[(unsigned long long)blockIdx.x * kMsVals + tid] = s;
    }
}

struct MeasureFinishArgs {
    const double* partials;
    int nblocks;
    int first_bit;
    int cell_of[kMsTileBits];   // chain cell of tile bit t
    double* sums;               // [4 * ncells]
};

// one warp per value; block partials added in a fixed order
__global__ void measure_tile_finish_kernel(const MeasureFinishArgs a) {
    __shared__ double tot[kMsVals];
    const int v = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int b = lane; b < a.nblocks; b += 32) s += a.partials[(unsigned long long)b * kMsVals + v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) tot[v] = s;
    __syncthreads();
    if (threadIdx.x < kMsTileBits && (int)threadIdx.x >= a.first_bit) {
        const int t = threadIdx.x;
        double* out = a.sums + 4 * a.cell_of[t];
        out[0] = tot[26] - tot[t];
        out[1] = tot[t];
        out[2] = tot[kMsTileBits + t];
        out[3] = 0.0;
    }
}

int32_t measure_tiles(const double* re, unsigned long long namps, const qca_pass_t& ps, const ShardMap& shard, int ncells,
                      double* d_partials, int max_blocks, double* d_sums, cudaStream_t stream) {
    const int L = ps.low_bits, H0 = ps.high_start;
    QCA_REQUIRE(L + ps.high_bits == kMsTileBits, QCA_ERR_ARG, "fused measurement needs full 13-bit tiles");
    MeasureTileArgs a{};
    a.re = re;
    a.ntiles = namps >> kMsTileBits;
    a.low_bits = L; a.high_start = H0;
    a.first_bit = ps.high_bits == 0 ? 0 : L;
    a.partials = d_partials;
    const int smem = 16 << kMsTileBits;   // 16 bytes per pair, 2^12 pairs
    QCA_CUDA(cudaFuncSetAttribute(measure_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int blocks = (int)std::max<unsigned long long>(1, std::min<unsigned long long>(a.ntiles, (unsigned long long)max_blocks));
    measure_tile_kernel<<<blocks, kMsThreads, smem, stream>>>(a);
    QCA_CUDA(cudaGetLastError());
    MeasureFinishArgs f{};
    f.partials = d_partials; f.nblocks = blocks; f.first_bit = a.first_bit; f.sums = d_sums;
    for (int t = 0; t < kMsTileBits; ++t) {
        const int g = t < L ? t : H0 + (t - L);          // local index bit of tile bit t
        f.cell_of[t] = ncells - 1 - global_pos(g, shard);
    }
    measure_tile_finish_kernel<<<1, 32 * kMsVals, 0, stream>>>(f);
    QCA_CUDA(cudaGetLastError());
    return QCA_OK;
}

}  // namespace qca
