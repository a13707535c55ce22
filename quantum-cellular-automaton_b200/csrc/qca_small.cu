// Small registers (<= 13 qubits, one GPU): the sizes the reference itself runs (its dense U stops near
// N = 12, BASELINE configs[0] is N = 9).  A launch per tile pass costs more than the arithmetic there,
// so ONE kernel does a whole time step: a CTA per plane keeps the three Clenshaw vectors and the rule
// predicates of all 2^n basis states in shared memory (3 x 64 KiB + 16 KiB at n = 13) and sums every
// Chebyshev term with one __syncthreads per term.  Same recurrence, coefficients and signs as step_once /
// pass_kernel_generic in qca_exact.cu; `small_measure_kernel` is MPS.measure (mps.py:100-140) for the
// same sizes in one launch (all 4 N sums, fixed reduction order).
#include "qca_small.h"

namespace qca {

constexpr int kSmallThreads = 1024;

__global__ void __launch_bounds__(kSmallThreads, 1) small_step_kernel(const SmallStepArgs a) {
    extern __shared__ double smem[];
    const unsigned dim = 1u << a.nbits;
    double* P = smem;
    double* X = P + dim;
    double* Y = X + dim;
    unsigned short* act = reinterpret_cast<unsigned short*>(Y + dim);
    const int plane = blockIdx.x;
    const double* __restrict__ src = a.src[plane];
    const unsigned qmask = dim - 1u;
    const int K = a.nterms - 1;
    for (unsigned x = threadIdx.x; x < dim; x += kSmallThreads) {
        const double p = src[x];
        P[x] = p;
        X[x] = a.coef[K] * p;   // B_K
        act[x] = (unsigned short)(activity_word<unsigned>(x, a.distance, a.interval_mask) & qmask);
    }
    __syncthreads();
    for (int k = K - 1; k >= 0; --k) {
        const double gamma = (k == 0) ? a.gamma_last : a.gamma;
        const bool first = (k == K - 1);   // B_{K+1} = 0
        const double ak = a.coef[k];
        for (unsigned x = threadIdx.x; x < dim; x += kSmallThreads) {
            const unsigned w = act[x];
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < kSmallMaxBits; ++q) {
                if (q < a.nbits && ((w >> q) & 1u)) {
                    const double v = X[x ^ (1u << q)];
                    acc += ((x >> q) & 1u) ? -v : v;   // K = sum_c P_c (sigma^- - sigma^+)_c
                }
            }
            double r = fma(gamma, acc, ak * P[x]);
            if (!first) r += Y[x];
            Y[x] = r;
        }
        __syncthreads();
        double* t = X; X = Y; Y = t;
    }
    double* dst = a.dst[plane];
    for (unsigned x = threadIdx.x; x < dim; x += kSmallThreads) dst[x] = X[x];
}

// sums[4*cell + {0,1,2,3}] = sum|phi_0|^2, sum|phi_1|^2, Re w, Im w, w = sum phi_0 conj(phi_1) over the pairs of
// the cell's index bit (cell = ncells-1-bit).  One CTA; per cell a block reduction in a fixed order.
__global__ void __launch_bounds__(kSmallThreads, 1) small_measure_kernel(const double* __restrict__ re,
                                                                         const double* __restrict__ im, int nbits,
                                                                         double* __restrict__ sums) {
    extern __shared__ double smem[];
    const unsigned dim = 1u << nbits;
    double* sre = smem;
    double* sim = smem + dim;
    __shared__ double part[4][kSmallThreads / 32];
    for (unsigned x = threadIdx.x; x < dim; x += kSmallThreads) {
        sre[x] = re[x];
        sim[x] = im ? im[x] : 0.0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned npairs = dim >> 1;
    for (int bit = 0; bit < nbits; ++bit) {
        double s0 = 0.0, s1 = 0.0, wr = 0.0, wi = 0.0;
        for (unsigned j = threadIdx.x; j < npairs; j += kSmallThreads) {
            const unsigned x0 = ((j >> bit) << (bit + 1)) | (j & ((1u << bit) - 1u));
            const unsigned x1 = x0 | (1u << bit);
            const double a0 = sre[x0], b0 = sim[x0], a1 = sre[x1], b1 = sim[x1];
            s0 += a0 * a0 + b0 * b0;
            s1 += a1 * a1 + b1 * b1;
            wr += a0 * a1 + b0 * b1;
            wi += b0 * a1 - a0 * b1;
        }
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            wr += __shfl_xor_sync(0xffffffffu, wr, o);
            wi += __shfl_xor_sync(0xffffffffu, wi, o);
        }
        if (lane == 0) { part[0][warp] = s0; part[1][warp] = s1; part[2][warp] = wr; part[3][warp] = wi; }
        __syncthreads();
        if (threadIdx.x < 4) {
            double t = 0.0;
            for (int w = 0; w < kSmallThreads / 32; ++w) t += part[threadIdx.x][w];
            sums[4 * (nbits - 1 - bit) + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

static size_t small_step_smem(int nbits) { return ((size_t)3 * sizeof(double) + sizeof(unsigned short)) << nbits; }

// the opt-in shared-memory limit is a per-device function attribute: set it once per device
static int32_t configure_small_kernels() {
    static bool configured[64] = {};
    int dev = 0;
    QCA_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || configured[dev]) return QCA_OK;
    QCA_CUDA(cudaFuncSetAttribute(small_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)small_step_smem(kSmallMaxBits)));
    QCA_CUDA(cudaFuncSetAttribute(small_measure_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)((2 * sizeof(double)) << kSmallMaxBits)));
    configured[dev] = true;
    return QCA_OK;
}

int32_t launch_small_step(const SmallStepArgs& a, int nplanes, cudaStream_t stream) {
    const size_t smem = small_step_smem(a.nbits);
    QCA_CHECK(configure_small_kernels());
    small_step_kernel<<<nplanes, kSmallThreads, smem, stream>>>(a);
    QCA_CUDA(cudaGetLastError());
    return QCA_OK;
}

int32_t launch_small_measure(const double* re, const double* im, int nbits, double* d_sums, cudaStream_t stream) {
    QCA_CHECK(configure_small_kernels());
    small_measure_kernel<<<1, kSmallThreads, (2 * sizeof(double)) << nbits, stream>>>(re, im, nbits, d_sums);
    QCA_CUDA(cudaGetLastError());
    return QCA_OK;
}

}  // namespace qca
