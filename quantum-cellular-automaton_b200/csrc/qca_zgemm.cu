// Batched, segmented complex128 GEMM on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64) for the
// effective-Hamiltonian contractions of TDVP (reference: algorithms/tdvp.py:299-347, np.tensordot).
//
//   C_g[m, n] = sum_{s < S} sum_{k < K}  A_{g,s}[m, k] * B_{g,s}[k, n]        g < G
//
// with element addresses (in complex128 units)
//   A: a + g*a_sg + s*a_ss + m*a_sm + k*a_sk      (either a_sm == 1 or a_sk == 1)
//   B: b + g*b_sg + s*b_ss + k*b_sk + n
//   C: c + g*c_sg + m*c_sm + n
// which covers, without any permuted copy, the two heavy steps of  out = L . theta -> W -> . R :
//   T1[(a,c)][(w,y)][u] = sum_x L[x][(w,y)] theta[(a,c)][x][u]          (A with m contiguous, S = 1)
//   out[(b,d)][y][v]    = sum_n sum_u T3[(b,d)][n][y][u] R[u][n][v]     (A with k contiguous, S = w)
// and the environment updates.
//
// CTA tile 64 x 64 x 16 complex, 3-stage cp.async pipeline (32 KiB per stage), 8 warps in a 2 x 4
// grid, each owning 32 x 16 complex = 8 complex 8x8 tiles = 64 accumulator registers.  A complex
// tile product is four real DMMAs:  Cr += Ar Br + Ai (-Bi),  Ci += Ar Bi + Ai Br.
// Shared-memory rows are padded so that the fragment loads (one LDS.128 = (re, im) per lane) are
// conflict free: the 8 lanes of a quarter-warp touch 8 distinct 16-byte slots of a 128-byte window.
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "qca_common.cuh"
#include "qca_zgemm.h"

namespace qca {

constexpr int ZG_TM = 64, ZG_TN = 64, ZG_TK = 16, ZG_STAGES = 3, ZG_THREADS = 256;
// smem strides in complex elements
constexpr int ZG_A_KMAJOR_LD = ZG_TM + 2;   // A stored [k][m], row shift 32 B mod 128
constexpr int ZG_A_MMAJOR_LD = ZG_TK + 4;   // A stored [m][k], row shift 64 B mod 128 (20 * 16 = 320 = 2*128 + 64)
constexpr int ZG_B_LD = ZG_TN + 2;          // B stored [k][n]
constexpr int ZG_A_ELEMS = (ZG_TK * ZG_A_KMAJOR_LD > ZG_TM * ZG_A_MMAJOR_LD) ? ZG_TK * ZG_A_KMAJOR_LD : ZG_TM * ZG_A_MMAJOR_LD;
constexpr int ZG_B_ELEMS = ZG_TK * ZG_B_LD;
constexpr int ZG_STAGE_ELEMS = ZG_A_ELEMS + ZG_B_ELEMS;
constexpr int ZG_SMEM_BYTES = ZG_STAGES * ZG_STAGE_ELEMS * 16;


__device__ __forceinline__ void zg_cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;  // src-size 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <bool A_MMAJOR>   // A_MMAJOR: A has k contiguous in global memory (a_sk == 1), staged as [m][k]
__global__ void __launch_bounds__(ZG_THREADS, 2) zgemm_dmma_kernel(const ZgemmArgs p) {
    extern __shared__ double2 zsm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;          // warp grid 2 x 4
    const int m0 = blockIdx.y * ZG_TM, n0 = blockIdx.x * ZG_TN;
    const int g = blockIdx.z / p.nsplit, split = blockIdx.z % p.nsplit;
    const double2* ag = p.a + (long long)g * p.a_sg;
    const double2* bg = p.b + (long long)g * p.b_sg;
    unsigned smask = 0xffffffffu;
    int nseg = p.S;
    if (p.use_masks) {
        const unsigned cm = p.chan_mask[g];
        const int c0 = m0 / p.chan_len, c1 = min(m0 + ZG_TM - 1, p.M - 1) / p.chan_len;
        bool any = false;
        for (int c = c0; c <= c1; ++c) any |= (cm >> c) & 1u;
        if (!any) return;                              // (uniform over the CTA)
        smask = p.seg_mask[g];
        nseg = __popc(smask);
    }
    const int ktiles = (p.K + ZG_TK - 1) / ZG_TK;
    const int all_steps = ktiles * nseg;               // pipeline steps: (segment, k-tile)
    const int first = (int)((long long)all_steps * split / p.nsplit);
    const int total = (int)((long long)all_steps * (split + 1) / p.nsplit) - first;   // this CTA's share

    auto load_stage = [&](int step, int stage) {
        double2* sa = zsm + stage * ZG_STAGE_ELEMS;
        double2* sb = sa + ZG_A_ELEMS;
        int s = (first + step) / ktiles;
        const int k0 = ((first + step) % ktiles) * ZG_TK;
        if (p.use_masks) s = __fns(smask, 0, s + 1);     // s-th used segment
        const double2* as = ag + (long long)s * p.a_ss;
        const double2* bs = bg + (long long)s * p.b_ss;
        // A tile: 64 x 16 complex = 1024 elements, 4 per thread
#pragma unroll
        for (int i = 0; i < (ZG_TM * ZG_TK) / ZG_THREADS; ++i) {
            const int e = tid + i * ZG_THREADS;
            int m, k;
            if (A_MMAJOR) { k = e % ZG_TK; m = e / ZG_TK; } else { m = e % ZG_TM; k = e / ZG_TM; }
            const bool ok = (m0 + m < p.M) && (k0 + k < p.K);
            const double2* src = ok ? as + (long long)(m0 + m) * p.a_sm + (long long)(k0 + k) * p.a_sk : as;
            double2* dst = A_MMAJOR ? sa + m * ZG_A_MMAJOR_LD + k : sa + k * ZG_A_KMAJOR_LD + m;
            zg_cp_async16(dst, src, ok);
        }
        // B tile: 16 x 64
#pragma unroll
        for (int i = 0; i < (ZG_TK * ZG_TN) / ZG_THREADS; ++i) {
            const int e = tid + i * ZG_THREADS;
            const int n = e % ZG_TN, k = e / ZG_TN;
            const bool ok = (n0 + n < p.N) && (k0 + k < p.K);
            const double2* src = ok ? bs + (long long)(k0 + k) * p.b_sk + (n0 + n) : bs;
            zg_cp_async16(sb + k * ZG_B_LD + n, src, ok);
        }
    };

    double cr[4][2][2], ci[4][2][2];   // [m8 block][n8 block][2 columns of the lane]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) { cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0; }

    for (int st = 0; st < ZG_STAGES - 1; ++st) {
        if (st < total) load_stage(st, st);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int frow = lane >> 2, fk = lane & 3;   // fragment coordinates of this lane
    for (int step = 0; step < total; ++step) {
        asm volatile("cp.async.wait_group %0;" ::"n"(ZG_STAGES - 2) : "memory");
        __syncthreads();
        {   // refill the stage consumed in the previous iteration
            const int nxt = step + ZG_STAGES - 1;
            if (nxt < total) load_stage(nxt, nxt % ZG_STAGES);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const double2* sa = zsm + (step % ZG_STAGES) * ZG_STAGE_ELEMS;
        const double2* sb = sa + ZG_A_ELEMS;
#pragma unroll
        for (int kk = 0; kk < ZG_TK; kk += 4) {
            double2 af[4], bf[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = wm * 32 + i * 8 + frow, k = kk + fk;
                af[i] = A_MMAJOR ? sa[m * ZG_A_MMAJOR_LD + k] : sa[k * ZG_A_KMAJOR_LD + m];
                if (p.conj_a) af[i].y = -af[i].y;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) bf[j] = sb[(kk + fk) * ZG_B_LD + wn * 16 + j * 8 + frow];
            // four sweeps over the 8 tiles: consecutive DMMAs never touch the same accumulator
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(cr[i][j][0], cr[i][j][1], af[i].x, bf[j].x);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(ci[i][j][0], ci[i][j][1], af[i].x, bf[j].y);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(cr[i][j][0], cr[i][j][1], af[i].y, -bf[j].y);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(ci[i][j][0], ci[i][j][1], af[i].y, bf[j].x);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // epilogue: lane holds C(row = lane/4, cols 2*(lane%4) + {0,1}) of every 8x8 tile
    double2* cg = p.c + (long long)g * p.c_sg + (long long)split * p.c_ssplit;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = m0 + wm * 32 + i * 8 + frow;
            const int n = n0 + wn * 16 + j * 8 + 2 * fk;
            if (m < p.M) {
                double2* dst = cg + (long long)m * p.c_sm + n;
                if (n < p.N) dst[0] = make_double2(cr[i][j][0], ci[i][j][0]);
                if (n + 1 < p.N) dst[1] = make_double2(cr[i][j][1], ci[i][j][1]);
            }
        }
}

// Optional per-launch timing (bench.py's roofline of the contraction kernel): an event pair around
// every launch while enabled; single-threaded use.
struct ZgemmProfile {
    bool on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    double flops = 0.0;
};
static ZgemmProfile g_zprof;

}  // namespace qca

extern "C" {

int32_t qca_zgemm_profile(int32_t enable, double* ms, double* flops, uint64_t* launches) {
    qca::ZgemmProfile& pr = qca::g_zprof;
    double total = 0.0;
    for (auto& ev : pr.events) {
        QCA_CUDA(cudaEventSynchronize(ev.second));
        float t = 0.f;
        QCA_CUDA(cudaEventElapsedTime(&t, ev.first, ev.second));
        total += t;
        cudaEventDestroy(ev.first); cudaEventDestroy(ev.second);
    }
    if (ms) *ms = total;
    if (flops) *flops = pr.flops;
    if (launches) *launches = pr.events.size();
    pr.events.clear();
    pr.flops = 0.0;
    pr.on = enable != 0;
    return QCA_OK;
}

int32_t qca_zgemm_batched(const void* a, const void* b, void* c, int32_t M, int32_t N, int32_t K, int32_t S, int32_t G,
                          int64_t a_sg, int64_t a_ss, int64_t a_sm, int64_t a_sk, int64_t b_sg, int64_t b_ss,
                          int64_t b_sk, int64_t c_sg, int64_t c_sm, int32_t conj_a, int32_t nsplit, int64_t c_ssplit,
                          void* stream) {
    qca::ZgemmArgs p{};
    p.a = (const double2*)a; p.b = (const double2*)b; p.c = (double2*)c;
    p.a_sg = a_sg; p.a_ss = a_ss; p.a_sm = a_sm; p.a_sk = a_sk;
    p.b_sg = b_sg; p.b_ss = b_ss; p.b_sk = b_sk; p.c_sg = c_sg; p.c_sm = c_sm;
    p.M = M; p.N = N; p.K = K; p.S = S; p.G = G; p.conj_a = conj_a; p.nsplit = nsplit; p.c_ssplit = c_ssplit;
    return qca::zgemm_launch(p, (cudaStream_t)stream);
}

}  // extern "C"

namespace qca {

int32_t zgemm_launch(const ZgemmArgs& p, cudaStream_t stream) {
    QCA_REQUIRE(p.a && p.b && p.c, QCA_ERR_ARG, "NULL argument");
    QCA_REQUIRE(p.M >= 1 && p.N >= 1 && p.K >= 1 && p.S >= 1 && p.G >= 1 && p.G <= 65535, QCA_ERR_ARG, "bad GEMM shape");
    QCA_REQUIRE(p.a_sm == 1 || p.a_sk == 1, QCA_ERR_ARG, "A needs a unit stride in m or in k");
    QCA_REQUIRE(p.nsplit >= 1 && (long long)p.G * p.nsplit <= 65535, QCA_ERR_ARG, "bad split count %d", p.nsplit);
    if (p.use_masks)
        QCA_REQUIRE(p.G <= 4 && p.S <= 32 && p.chan_len >= 1 && (p.M + p.chan_len - 1) / p.chan_len <= 32, QCA_ERR_ARG,
                    "masks need G <= 4, S <= 32 and at most 32 channels");
    const dim3 grid((p.N + ZG_TN - 1) / ZG_TN, (p.M + ZG_TM - 1) / ZG_TM, p.G * p.nsplit);
    // k contiguous in global memory: stage A as [m][k]; otherwise (m contiguous) as [k][m]
    auto kern = (p.a_sk == 1 && p.a_sm != 1) ? zgemm_dmma_kernel<true> : zgemm_dmma_kernel<false>;
    QCA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ZG_SMEM_BYTES));
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (g_zprof.on) {
        QCA_CUDA(cudaEventCreate(&ev0)); QCA_CUDA(cudaEventCreate(&ev1));
        QCA_CUDA(cudaEventRecord(ev0, stream));
    }
    kern<<<grid, ZG_THREADS, ZG_SMEM_BYTES, stream>>>(p);
    QCA_CUDA(cudaGetLastError());
    if (g_zprof.on) {
        QCA_CUDA(cudaEventRecord(ev1, stream));
        g_zprof.events.emplace_back(ev0, ev1);
        // FP64 operations actually executed: 8 per complex multiply-add, skipped tiles / segments not counted
        double rows_segs = 0.0;
        for (int g = 0; g < p.G; ++g) {
            if (!p.use_masks) { rows_segs += (double)p.M * p.S; continue; }
            double rows = 0.0;
            for (int m0 = 0; m0 < p.M; m0 += ZG_TM) {
                const int c0 = m0 / p.chan_len, c1 = std::min(m0 + ZG_TM - 1, p.M - 1) / p.chan_len;
                bool any = false;
                for (int c = c0; c <= c1; ++c) any |= (p.chan_mask[g] >> c) & 1u;
                if (any) rows += std::min(ZG_TM, p.M - m0);
            }
            rows_segs += rows * __builtin_popcount(p.seg_mask[g]);
        }
        g_zprof.flops += 8.0 * rows_segs * p.N * p.K;
    }
    return QCA_OK;
}

}  // namespace qca
