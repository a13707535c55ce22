// Householder QR on the device with LAPACK's conventions (zgeqr2 / zlarfg / zung2r).
//
// Why not cuSOLVER: the reference re-canonicalises the MPS at the start of every TDVP step
// (tdvp.py:54) but keeps the environments it built before (tdvp.py:37-39, 122-145), so its results
// depend on the SIGNS numpy/LAPACK's QR puts on the diagonal of R.  LAPACK's zlarfg chooses
// beta = -sign(Re alpha) * ||(alpha, x)||; cuSOLVER's geqrf chooses differently, which changes the
// evolved populations at the 1e-4 level.  Being a drop-in therefore needs this exact convention.
//
// Column-major complex128 in global memory (L2 resident; the matrices are MPS tensors, at most 2*chi x chi).
//   * qr_factor_grid_kernel: right-looking zgeqr2 over a cooperative grid -- every CTA derives the reflector of
//     column k redundantly (same data, same summation order: no broadcast), the trailing columns are spread one
//     per warp over the whole grid, one grid barrier per column;
//   * qr_form_cols_kernel: zung2r column by column -- column j of Q only sees the reflectors k <= j and no other
//     column, so every column is one warp's private job and the kernel needs no barrier at all.
// The first version ran both steps in ONE CTA (qr_factor_kernel / qr_form_kernel, kept behind
// QCA_QR_SINGLE_CTA for A/B and as the reference in the tests): 7.9 + 17.3 ms at 512 x 256 on a B200, bound by one
// SM's path to L2 -- 92 % of a chi = 256 1TDVP time step (profiles/r02_tdvp1_step_kernel_breakdown_chi256.txt).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "qca_common.cuh"

namespace qca {

constexpr int kQrThreads = 1024;

struct cplx { double x, y; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { return {a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x}; }  // conj(a) * b
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    // Smith's algorithm as in LAPACK's zladiv (robust scaling is not needed at these magnitudes)
    if (fabs(b.y) < fabs(b.x)) {
        const double e = b.y / b.x, f = b.x + b.y * e;
        return {(a.x + a.y * e) / f, (a.y - a.x * e) / f};
    }
    const double e = b.x / b.y, f = b.y + b.x * e;
    return {(a.y + a.x * e) / f, (-a.x + a.y * e) / f};
}

__device__ __forceinline__ double block_sum(double v, double* scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kQrThreads / 32; ++w) t += scratch[w];  // fixed order: deterministic
    return t;
}

// A: m x n column-major (lda = m), overwritten with R (upper triangle) and the reflectors below
// the diagonal; tau[k], k < min(m, n).
__global__ void __launch_bounds__(kQrThreads) qr_factor_kernel(cplx* __restrict__ a, int m, int n, cplx* __restrict__ tau) {
    extern __shared__ double smem[];
    cplx* v = reinterpret_cast<cplx*>(smem);           // current reflector, m entries
    double* scratch = smem + 2 * (size_t)m;             // 32 doubles
    __shared__ cplx s_tau;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kQrThreads / 32;
    const int kmax = m < n ? m : n;
    for (int k = 0; k < kmax; ++k) {
        cplx* col = a + (size_t)k * m;
        // zlarfg: norm of the sub-diagonal part
        double part = 0.0;
        for (int i = k + 1 + tid; i < m; i += kQrThreads) part += col[i].x * col[i].x + col[i].y * col[i].y;
        const double xnorm2 = block_sum(part, scratch);
        if (tid == 0) {
            const cplx alpha = col[k];
            cplx t = {0.0, 0.0};
            cplx scale = {0.0, 0.0};
            double beta = alpha.x;
            if (!(xnorm2 == 0.0 && alpha.y == 0.0)) {
                const double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xnorm2);
                beta = -copysign(nrm, alpha.x);
                t = {(beta - alpha.x) / beta, -alpha.y / beta};
                scale = cdiv({1.0, 0.0}, {alpha.x - beta, alpha.y});
                col[k] = {beta, 0.0};
            }
            s_tau = t;
            tau[k] = t;
            v[k] = {1.0, 0.0};
            scratch[32] = scale.x; scratch[33] = scale.y;
        }
        __syncthreads();
        const cplx t = s_tau;
        const cplx scale = {scratch[32], scratch[33]};
        const bool identity = (t.x == 0.0 && t.y == 0.0);
        for (int i = k + 1 + tid; i < m; i += kQrThreads) {
            cplx x = col[i];
            if (!identity) x = cmul(scale, x);
            col[i] = x;
            v[i] = x;
        }
        __syncthreads();
        if (!identity) {
            // apply H^H = I - conj(tau) v v^H to the trailing columns: one warp per column
            const cplx tc = {t.x, -t.y};
            for (int j = k + 1 + warp; j < n; j += nwarps) {
                cplx* c = a + (size_t)j * m;
                cplx w = {0.0, 0.0};
                for (int i = k + lane; i < m; i += 32) {
                    const cplx p = cmulc(v[i], c[i]);
                    w.x += p.x; w.y += p.y;
                }
                for (int o = 16; o > 0; o >>= 1) {
                    w.x += __shfl_xor_sync(0xffffffffu, w.x, o);
                    w.y += __shfl_xor_sync(0xffffffffu, w.y, o);
                }
                const cplx f = cmul(tc, w);
                for (int i = k + lane; i < m; i += 32) {
                    const cplx p = cmul(v[i], f);
                    c[i].x -= p.x; c[i].y -= p.y;
                }
            }
        }
        __syncthreads();
    }
}

// Q (m x kq, column-major) = H_0 H_1 ... H_{kmax-1} applied to the first kq columns of the identity;
// R (kr x n, column-major) = upper triangle of the factored A (kr = kq).
__global__ void __launch_bounds__(kQrThreads) qr_form_kernel(const cplx* __restrict__ a, int m, int n,
                                                              const cplx* __restrict__ tau, cplx* __restrict__ q,
                                                              int kq, cplx* __restrict__ r) {
    extern __shared__ double smem[];
    cplx* v = reinterpret_cast<cplx*>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kQrThreads / 32;
    const int kmax = m < n ? m : n;
    for (size_t idx = tid; idx < (size_t)kq * n; idx += kQrThreads) {
        const int i = (int)(idx % kq), j = (int)(idx / kq);
        r[idx] = (i <= j && i < kmax) ? a[(size_t)j * m + i] : cplx{0.0, 0.0};
    }
    for (size_t idx = tid; idx < (size_t)m * kq; idx += kQrThreads) {
        const int i = (int)(idx % m), j = (int)(idx / m);
        q[idx] = (i == j) ? cplx{1.0, 0.0} : cplx{0.0, 0.0};
    }
    __syncthreads();
    for (int k = kmax - 1; k >= 0; --k) {
        const cplx t = tau[k];
        for (int i = k + tid; i < m; i += kQrThreads) v[i] = (i == k) ? cplx{1.0, 0.0} : a[(size_t)k * m + i];
        __syncthreads();
        if (!(t.x == 0.0 && t.y == 0.0)) {
            // columns j < k of Q are still unit vectors e_j with zero rows >= k: H_k leaves them alone
            for (int j = k + warp; j < kq; j += nwarps) {
                cplx* c = q + (size_t)j * m;
                cplx w = {0.0, 0.0};
                for (int i = k + lane; i < m; i += 32) {
                    const cplx p = cmulc(v[i], c[i]);
                    w.x += p.x; w.y += p.y;
                }
                for (int o = 16; o > 0; o >>= 1) {
                    w.x += __shfl_xor_sync(0xffffffffu, w.x, o);
                    w.y += __shfl_xor_sync(0xffffffffu, w.y, o);
                }
                const cplx f = cmul(t, w);
                for (int i = k + lane; i < m; i += 32) {
                    const cplx p = cmul(v[i], f);
                    c[i].x -= p.x; c[i].y -= p.y;
                }
            }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------------------
// Multi-CTA versions (cooperative launch: the grid barrier needs every CTA resident).
// ---------------------------------------------------------------------------------------------------------
constexpr int kQrGridThreads = 256;
constexpr int kQrGridWarps = kQrGridThreads / 32;

__device__ __forceinline__ double block_sum_grid(double v, double* scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kQrGridWarps; ++w) t += scratch[w];  // fixed order: the same number in every CTA
    return t;
}

__global__ void __launch_bounds__(kQrGridThreads) qr_factor_grid_kernel(cplx* __restrict__ a, int m, int n, cplx* __restrict__ tau) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smem[];
    cplx* v = reinterpret_cast<cplx*>(smem);           // current reflector, m entries
    double* scratch = smem + 2 * (size_t)m;             // kQrGridWarps partial sums + scale
    __shared__ cplx s_tau;
    __shared__ double s_beta;
    __shared__ int s_changed;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gwarp = blockIdx.x * kQrGridWarps + warp, gwarps = gridDim.x * kQrGridWarps;
    const int kmax = m < n ? m : n;
    for (int k = 0; k < kmax; ++k) {
        cplx* col = a + (size_t)k * m;
        // zlarfg on the (final) column k, redundantly in every CTA
        double part = 0.0;
        for (int i = k + 1 + tid; i < m; i += kQrGridThreads) part += col[i].x * col[i].x + col[i].y * col[i].y;
        const double xnorm2 = block_sum_grid(part, scratch);
        if (tid == 0) {
            const cplx alpha = col[k];
            cplx t = {0.0, 0.0};
            cplx scale = {0.0, 0.0};
            double beta = alpha.x;
            int changed = 0;
            if (!(xnorm2 == 0.0 && alpha.y == 0.0)) {
                const double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xnorm2);
                beta = -copysign(nrm, alpha.x);
                t = {(beta - alpha.x) / beta, -alpha.y / beta};
                scale = cdiv({1.0, 0.0}, {alpha.x - beta, alpha.y});
                changed = 1;
            }
            s_tau = t; s_beta = beta; s_changed = changed;
            v[k] = {1.0, 0.0};
            scratch[kQrGridWarps] = scale.x; scratch[kQrGridWarps + 1] = scale.y;
        }
        __syncthreads();
        const cplx t = s_tau;
        const cplx scale = {scratch[kQrGridWarps], scratch[kQrGridWarps + 1]};
        const bool identity = (t.x == 0.0 && t.y == 0.0);
        for (int i = k + 1 + tid; i < m; i += kQrGridThreads) {
            cplx x = col[i];
            if (!identity) x = cmul(scale, x);
            v[i] = x;
        }
        __syncthreads();
        if (!identity) {
            // apply H^H = I - conj(tau) v v^H to the trailing columns: one warp per column, over the whole grid
            const cplx tc = {t.x, -t.y};
            for (int j = k + 1 + gwarp; j < n; j += gwarps) {
                cplx* c = a + (size_t)j * m;
                cplx w = {0.0, 0.0};
                for (int i = k + lane; i < m; i += 32) {
                    const cplx p = cmulc(v[i], c[i]);
                    w.x += p.x; w.y += p.y;
                }
                for (int o = 16; o > 0; o >>= 1) {
                    w.x += __shfl_xor_sync(0xffffffffu, w.x, o);
                    w.y += __shfl_xor_sync(0xffffffffu, w.y, o);
                }
                const cplx f = cmul(tc, w);
                for (int i = k + lane; i < m; i += 32) {
                    const cplx p = cmul(v[i], f);
                    c[i].x -= p.x; c[i].y -= p.y;
                }
            }
        }
        // Every CTA has read column k and the trailing columns are updated: column k + 1 is final.  Only now may
        // CTA 0 overwrite column k with the scaled reflector (the others were still reading the unscaled one).
        grid.sync();
        if (blockIdx.x == 0) {
            if (!identity)
                for (int i = k + 1 + tid; i < m; i += kQrGridThreads) col[i] = v[i];
            if (tid == 0) {
                if (s_changed) col[k] = {s_beta, 0.0};
                tau[k] = t;
            }
        }
    }
}

__global__ void __launch_bounds__(kQrGridThreads) qr_form_cols_kernel(const cplx* __restrict__ a, int m, int n,
                                                                        const cplx* __restrict__ tau, cplx* q, int kq,
                                                                        cplx* __restrict__ r) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kmax = m < n ? m : n;
    const size_t gtid = (size_t)blockIdx.x * kQrGridThreads + tid, gsize = (size_t)gridDim.x * kQrGridThreads;
    for (size_t idx = gtid; idx < (size_t)kq * n; idx += gsize) {
        const int i = (int)(idx % kq), j = (int)(idx / kq);
        r[idx] = (i <= j && i < kmax) ? a[(size_t)j * m + i] : cplx{0.0, 0.0};
    }
    const int gwarps = gridDim.x * kQrGridWarps;
    for (int j = blockIdx.x * kQrGridWarps + warp; j < kq; j += gwarps) {
        cplx* c = q + (size_t)j * m;
        for (int i = lane; i < m; i += 32) c[i] = (i == j) ? cplx{1.0, 0.0} : cplx{0.0, 0.0};
        __syncwarp();
        // column j of the identity has zero rows > j: the reflectors k > j leave it alone
        for (int k = (j < kmax - 1 ? j : kmax - 1); k >= 0; --k) {
            const cplx t = tau[k];
            if (t.x == 0.0 && t.y == 0.0) continue;
            const cplx* vk = a + (size_t)k * m;
            cplx w = {0.0, 0.0};
            for (int i = k + lane; i < m; i += 32) {
                const cplx vi = (i == k) ? cplx{1.0, 0.0} : vk[i];
                const cplx p = cmulc(vi, c[i]);
                w.x += p.x; w.y += p.y;
            }
            for (int o = 16; o > 0; o >>= 1) {
                w.x += __shfl_xor_sync(0xffffffffu, w.x, o);
                w.y += __shfl_xor_sync(0xffffffffu, w.y, o);
            }
            const cplx f = cmul(t, w);
            for (int i = k + lane; i < m; i += 32) {
                const cplx vi = (i == k) ? cplx{1.0, 0.0} : vk[i];
                const cplx p = cmul(vi, f);
                c[i].x -= p.x; c[i].y -= p.y;
            }
            __syncwarp();   // the lane that owns row i changes with k
        }
    }
}


// The same, with the column in REGISTERS (m <= 32 * S): lane l owns the rows i = l + 32 s for the whole job, so nothing
// goes through memory between two reflectors.  (qr_form_cols_kernel re-reads and re-writes its column in global memory
// for every reflector, and the lane that owns a row changes with k: 0.81 ms at 512 x 256, two L2 round trips per reflector.)
template <int S>
__global__ void __launch_bounds__(kQrGridThreads) qr_form_cols_reg_kernel(const cplx* __restrict__ a, int m, int n,
                                                                            const cplx* __restrict__ tau, cplx* __restrict__ q,
                                                                            int kq, cplx* __restrict__ r) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kmax = m < n ? m : n;
    const size_t gtid = (size_t)blockIdx.x * kQrGridThreads + tid, gsize = (size_t)gridDim.x * kQrGridThreads;
    for (size_t idx = gtid; idx < (size_t)kq * n; idx += gsize) {
        const int i = (int)(idx % kq), j = (int)(idx / kq);
        r[idx] = (i <= j && i < kmax) ? a[(size_t)j * m + i] : cplx{0.0, 0.0};
    }
    const int gwarps = gridDim.x * kQrGridWarps;
    for (int j = blockIdx.x * kQrGridWarps + warp; j < kq; j += gwarps) {
        cplx c[S];
#pragma unroll
        for (int s = 0; s < S; ++s) c[s] = (lane + 32 * s == j) ? cplx{1.0, 0.0} : cplx{0.0, 0.0};
        for (int k = (j < kmax - 1 ? j : kmax - 1); k >= 0; --k) {
            const cplx t = tau[k];
            if (t.x == 0.0 && t.y == 0.0) continue;
            const cplx* vk = a + (size_t)k * m;
            cplx vv[S];
            cplx w = {0.0, 0.0};
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int i = lane + 32 * s;
                vv[s] = (i > k && i < m) ? vk[i] : ((i == k) ? cplx{1.0, 0.0} : cplx{0.0, 0.0});   // rows above k: untouched
                const cplx p = cmulc(vv[s], c[s]);
                w.x += p.x; w.y += p.y;
            }
            for (int o = 16; o > 0; o >>= 1) {
                w.x += __shfl_xor_sync(0xffffffffu, w.x, o);
                w.y += __shfl_xor_sync(0xffffffffu, w.y, o);
            }
            const cplx f = cmul(t, w);
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const cplx p = cmul(vv[s], f);
                c[s].x -= p.x; c[s].y -= p.y;
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int i = lane + 32 * s;
            if (i < m) q[(size_t)j * m + i] = c[s];
        }
    }
}

}  // namespace qca

extern "C" {

int32_t qca_qr_householder(void* a, int32_t m, int32_t n, void* tau, void* q, int32_t kq, void* r, void* stream) {
    QCA_REQUIRE(a && tau && q && r, QCA_ERR_ARG, "NULL argument");
    QCA_REQUIRE(m >= 1 && n >= 1 && m <= 8192, QCA_ERR_ARG, "matrix %d x %d out of range", m, n);
    const int kmax = m < n ? m : n;
    QCA_REQUIRE(kq == kmax || kq == m, QCA_ERR_ARG, "kq must be min(m,n) (reduced) or m (complete)");
    const size_t smem = (2 * (size_t)m + 40) * sizeof(double);
    cudaStream_t s = (cudaStream_t)stream;
    const bool single_cta = getenv("QCA_QR_SINGLE_CTA") != nullptr;   // (read per call: the tests switch it)
    if (!single_cta) {
        int dev = 0, sms = 0, coop = 0, per_sm = 0;
        QCA_CUDA(cudaGetDevice(&dev));
        QCA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        QCA_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        QCA_REQUIRE(coop != 0, QCA_ERR_CUDA, "device %d cannot launch cooperative kernels", dev);
        QCA_CUDA(cudaFuncSetAttribute(qca::qr_factor_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QCA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qca::qr_factor_grid_kernel, qca::kQrGridThreads, smem));
        QCA_REQUIRE(per_sm >= 1, QCA_ERR_CUDA, "qr_factor_grid_kernel does not fit an SM (m = %d)", m);
        // one warp per trailing column is all the parallelism a column step has; every CTA must be resident
        int grid = (n + qca::kQrGridWarps - 1) / qca::kQrGridWarps;
        if (grid > sms * per_sm) grid = sms * per_sm;
        if (grid > sms) grid = sms;
        if (grid < 1) grid = 1;
        void* fargs[] = {&a, &m, &n, &tau};
        QCA_CUDA(cudaLaunchCooperativeKernel((const void*)qca::qr_factor_grid_kernel, dim3(grid), dim3(qca::kQrGridThreads), fargs, smem, s));
        int fgrid = (kq + qca::kQrGridWarps - 1) / qca::kQrGridWarps;
        if (fgrid > 4 * sms) fgrid = 4 * sms;
        const bool in_regs = getenv("QCA_QR_FORM_GLOBAL") == nullptr;
        if (in_regs && m <= 128)
            qca::qr_form_cols_reg_kernel<4><<<fgrid, qca::kQrGridThreads, 0, s>>>((const qca::cplx*)a, m, n, (const qca::cplx*)tau,
                                                                                 (qca::cplx*)q, kq, (qca::cplx*)r);
        else if (in_regs && m <= 512)
            qca::qr_form_cols_reg_kernel<16><<<fgrid, qca::kQrGridThreads, 0, s>>>((const qca::cplx*)a, m, n, (const qca::cplx*)tau,
                                                                                  (qca::cplx*)q, kq, (qca::cplx*)r);
        else
            qca::qr_form_cols_kernel<<<fgrid, qca::kQrGridThreads, 0, s>>>((const qca::cplx*)a, m, n, (const qca::cplx*)tau,
                                                                          (qca::cplx*)q, kq, (qca::cplx*)r);
        QCA_CUDA(cudaGetLastError());
        return QCA_OK;
    }
    QCA_CUDA(cudaFuncSetAttribute(qca::qr_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QCA_CUDA(cudaFuncSetAttribute(qca::qr_form_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qca::qr_factor_kernel<<<1, qca::kQrThreads, smem, s>>>((qca::cplx*)a, m, n, (qca::cplx*)tau);
    QCA_CUDA(cudaGetLastError());
    qca::qr_form_kernel<<<1, qca::kQrThreads, smem, s>>>((const qca::cplx*)a, m, n, (const qca::cplx*)tau,
                                                       (qca::cplx*)q, kq, (qca::cplx*)r);
    QCA_CUDA(cudaGetLastError());
    return QCA_OK;
}

}  // extern "C"
