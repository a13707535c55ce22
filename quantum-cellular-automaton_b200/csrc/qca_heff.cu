// Matrix-free effective Hamiltonian of TDVP and its Krylov exponential, resident on the device.
//
// Replaces, for one site tensor / two-site tensor / bond matrix psi,
//   TDVP._assemble_H_eff + _evolve_A   (algorithms/tdvp.py:299-310, 350-365: dense (g dl dr)^2 matrix)
//   lautils.timestep / calculate_U      (lautils/lautils.py:45-82: eigh of that matrix)
// by
//   H_eff psi = (L . psi) -> site operator(s) -> (. R)                  [two DMMA GEMMs + a sparse mix]
//   exp(-i t H_eff) psi = ||psi|| V exp(-i t T) e_0                     [m-step Lanczos, V and T on the device]
// launched as one fixed sequence of kernels on the caller's stream: no host synchronisation, no
// allocation (the caller passes a workspace).
//
// Index conventions are the reference's: psi[g][x][u], L[x][w][y], R[u][w][v]; g enumerates the physical
// indices (1: bond matrix, 2: one site, 4: two sites, (a, c) -> 2a + c).  The site operators enter as
// the sparse matrix  Mx[(g', n), (g, w)]  (two sites: sum_m W1[a,b,w,m] W2[c,d,m,n]) in CSR form:
//   T1[g][w][y][u]  = sum_x L[x][w][y] psi[g][x][u]
//   T3[g'][n][y][u] = sum_{g,w} Mx[(g',n),(g,w)] T1[g][w][y][u]
//   out[g'][y][v]   = sum_{n,u} T3[g'][n][y][u] R[u][n][v]
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>

#include <vector>

#include "qca_common.cuh"
#include "qca_plan.h"
#include "qca_zgemm.h"

namespace qca {

constexpr int HE_THREADS = 256;
constexpr int HE_MAXK = 64;   // largest Krylov dimension

// ---- deterministic block reductions -----------------------------------------------------------
__device__ __forceinline__ double he_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NV>
__device__ __forceinline__ void he_block_sum(double (&v)[NV], double* red /* [NV][8] shared */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
        v[u] = he_warp_sum(v[u]);
        if (lane == 0) red[u * 8 + warp] = v[u];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < HE_THREADS / 32; ++w) s += red[threadIdx.x * 8 + w];
        red[threadIdx.x * 8] = s;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < NV; ++u) v[u] = red[u * 8];
    __syncthreads();
}

__device__ __forceinline__ void he_slice(long long dim, long long* s0, long long* s1) {
    const long long chunk = (dim + gridDim.x - 1) / gridDim.x;
    *s0 = (long long)blockIdx.x * chunk;
    *s1 = (*s0 + chunk < dim) ? *s0 + chunk : dim;
}

// ---- site operator(s): sparse mix of the (g, w) channels, summing split-K partials on the way ---
__global__ void __launch_bounds__(HE_THREADS)
he_mix_kernel(const double2* __restrict__ t1, int nsplit, long long sstride, double2* __restrict__ t3,
              const int* __restrict__ rowptr, const int* __restrict__ col, const double2* __restrict__ val,
              int nrows, long long plane) {
    for (long long s = (long long)blockIdx.x * HE_THREADS + threadIdx.x; s < plane; s += (long long)gridDim.x * HE_THREADS) {
        for (int r = 0; r < nrows; ++r) {
            double ar = 0.0, ai = 0.0;
            for (int z = rowptr[r]; z < rowptr[r + 1]; ++z) {
                const double2 c = val[z];
                const double2* src = t1 + (long long)col[z] * plane + s;
                double2 x = src[0];
                for (int k = 1; k < nsplit; ++k) {
                    const double2 y = src[(long long)k * sstride];
                    x.x += y.x; x.y += y.y;
                }
                ar += c.x * x.x - c.y * x.y;
                ai += c.x * x.y + c.y * x.x;
            }
            t3[(long long)r * plane + s] = make_double2(ar, ai);
        }
    }
}

// out = sum of split-K partials
__global__ void __launch_bounds__(HE_THREADS)
he_sum_kernel(double2* __restrict__ out, const double2* __restrict__ parts, int nsplit, long long sstride, long long dim) {
    for (long long s = (long long)blockIdx.x * HE_THREADS + threadIdx.x; s < dim; s += (long long)gridDim.x * HE_THREADS) {
        double2 x = parts[s];
        for (int k = 1; k < nsplit; ++k) {
            const double2 y = parts[(long long)k * sstride + s];
            x.x += y.x; x.y += y.y;
        }
        out[s] = x;
    }
}

// ---- Lanczos vector work --------------------------------------------------------------------------
// partial[block][i] = sum over the block's slice of conj(V_i) w,  i < nvec.  parts != NULL: w is first
// formed as the sum of the split-K partials of the second GEMM and stored.
__global__ void __launch_bounds__(HE_THREADS)
he_dots_kernel(double2* __restrict__ w, const double2* __restrict__ parts, int nsplit, long long sstride,
               const double2* __restrict__ V, long long dim, int nvec, double2* __restrict__ partial) {
    __shared__ double red[8 * 8];
    long long s0, s1;
    he_slice(dim, &s0, &s1);
    if (parts) {
        for (long long s = s0 + threadIdx.x; s < s1; s += HE_THREADS) {
            double2 x = parts[s];
            for (int k = 1; k < nsplit; ++k) {
                const double2 y = parts[(long long)k * sstride + s];
                x.x += y.x; x.y += y.y;
            }
            w[s] = x;   // re-read below by the same thread only
        }
    }
    for (int i0 = 0; i0 < nvec; i0 += 4) {
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (long long s = s0 + threadIdx.x; s < s1; s += HE_THREADS) {
            const double2 x = w[s];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u < nvec) {
                    const double2 v = V[(long long)(i0 + u) * dim + s];
                    acc[2 * u] += v.x * x.x + v.y * x.y;       // conj(v) * x
                    acc[2 * u + 1] += v.x * x.y - v.y * x.x;
                }
            }
        }
        he_block_sum<8>(acc, red);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u < nvec) partial[(long long)blockIdx.x * HE_MAXK + i0 + u] = make_double2(acc[2 * u], acc[2 * u + 1]);
        }
    }
}

// w -= sum_i c_i V_i with c_i = sum over blocks (fixed order) of partial[block][i].
// alpha_out != NULL: alpha_out = Re c_{nvec-1} (Lanczos diagonal).  norm_partial != NULL: the block's
// share of ||w||^2 after the update.
__global__ void __launch_bounds__(HE_THREADS)
he_update_kernel(double2* __restrict__ w, const double2* __restrict__ V, long long dim, int nvec,
                 const double2* __restrict__ partial, int nparts, double* alpha_out, double* norm_partial) {
    __shared__ double2 c[HE_MAXK];
    __shared__ double red[8];
    {   // one warp per coefficient: lanes add the block partials in a fixed order, then a butterfly
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int i = warp; i < nvec; i += HE_THREADS / 32) {
            double cr = 0.0, ci = 0.0;
            for (int b = lane; b < nparts; b += 32) {
                const double2 p = partial[(long long)b * HE_MAXK + i];
                cr += p.x; ci += p.y;
            }
            cr = he_warp_sum(cr); ci = he_warp_sum(ci);
            if (lane == 0) {
                c[i] = make_double2(cr, ci);
                if (alpha_out && blockIdx.x == 0 && i == nvec - 1) *alpha_out = cr;
            }
        }
    }
    __syncthreads();
    long long s0, s1;
    he_slice(dim, &s0, &s1);
    double nacc[1] = {0.0};
    for (long long s = s0 + threadIdx.x; s < s1; s += HE_THREADS) {
        double2 x = w[s];
        for (int i = 0; i < nvec; ++i) {
            const double2 v = V[(long long)i * dim + s];
            const double2 ci = c[i];
            x.x -= ci.x * v.x - ci.y * v.y;
            x.y -= ci.x * v.y + ci.y * v.x;
        }
        w[s] = x;
        nacc[0] += x.x * x.x + x.y * x.y;
    }
    if (norm_partial) {
        he_block_sum<1>(nacc, red);
        if (threadIdx.x == 0) norm_partial[blockIdx.x] = nacc[0];
    }
}

// alpha = Re sum_blocks partial[block][0]   (last Lanczos step: only the diagonal entry is needed)
__global__ void he_alpha_kernel(const double2* __restrict__ partial, int nparts, double* alpha_out) {   // <<<1, 32>>>
    double cr = 0.0;   // same summation order as he_update_kernel
    for (int b = threadIdx.x; b < nparts; b += 32) cr += partial[(long long)b * HE_MAXK].x;
    cr = he_warp_sum(cr);
    if (threadIdx.x == 0) *alpha_out = cr;
}

__global__ void __launch_bounds__(HE_THREADS)
he_norm_kernel(const double2* __restrict__ src, long long dim, double* __restrict__ norm_partial) {
    __shared__ double red[8];
    long long s0, s1;
    he_slice(dim, &s0, &s1);
    double nacc[1] = {0.0};
    for (long long s = s0 + threadIdx.x; s < s1; s += HE_THREADS) {
        const double2 x = src[s];
        nacc[0] += x.x * x.x + x.y * x.y;
    }
    he_block_sum<1>(nacc, red);
    if (threadIdx.x == 0) norm_partial[blockIdx.x] = nacc[0];
}

// dst = src / b with b = sqrt(sum of the norm partials, fixed order); a vanishing b (Krylov space
// exhausted, tdvp never divides by it) gives dst = 0 and b_out = 0.
__global__ void __launch_bounds__(HE_THREADS)
he_normalize_kernel(double2* dst, const double2* src, long long dim,   // dst may alias src
                    const double* __restrict__ norm_partial, int nparts, double* b_out) {
    __shared__ double sh_inv;
    if (threadIdx.x < 32) {
        double s = 0.0;
        for (int b = threadIdx.x; b < nparts; b += 32) s += norm_partial[b];
        s = he_warp_sum(s);
        if (threadIdx.x == 0) {
            const double b = sqrt(s);
            const bool ok = b > 1e-13;
            sh_inv = ok ? 1.0 / b : 0.0;
            if (blockIdx.x == 0) *b_out = ok ? b : 0.0;
        }
    }
    __syncthreads();
    const double inv = sh_inv;
    long long s0, s1;
    he_slice(dim, &s0, &s1);
    for (long long s = s0 + threadIdx.x; s < s1; s += HE_THREADS) {
        const double2 x = src[s];
        dst[s] = make_double2(x.x * inv, x.y * inv);
    }
}

constexpr int HE_MAXCHEB = 200;
struct ChebCoefs {
    double a[HE_MAXCHEB];   // a_k = (2 - delta_k0) J_k(R |t|)
    int n;                  // 0: no plan (bound unknown or too many terms)
    double R;               // ||T|| <= R
};

// coef = norm0 * exp(-i t T) e_0 for the m x m Lanczos matrix T = tridiag(beta, alpha, beta)
// (lautils.py:45-55 applied to T).  One warp.
//  * Chebyshev path: with ||T|| <= ||H_eff|| <= R,  exp(-i t T) e_0 = sum_k a_k (-i sgn t)^k T_k(T/R) e_0
//    (three-term recurrence on a real m-vector, a_k from the host's Bessel plan): a few dozen tridiagonal
//    matrix-vector products, no divisions or square roots;
//  * Jacobi path (no bound given, too many terms, or the Gershgorin radius of T exceeds R): cyclic Jacobi
//    in shared memory, coef = norm0 * Z exp(-i t Lambda) Z^T e_0; unconditionally accurate for the small,
//    possibly decoupled (beta = 0) matrices that occur here.
__global__ void he_tridiag_expm_kernel(const double* __restrict__ alpha, const double* __restrict__ beta, int m,
                                       const double* __restrict__ norm0, double t, double2* __restrict__ coef,
                                       const ChebCoefs cheb) {
    extern __shared__ double jsm[];
    double* A = jsm;            // m x m
    double* Z = jsm + m * m;    // m x m, columns = eigenvectors
    const int lane = threadIdx.x;
    if (cheb.n > 0) {
        // rows i = lane and lane + 32 of T
        double al[2], bl[2], bu[2];   // diagonal, coupling to i-1, coupling to i+1
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = lane + 32 * h;
            al[h] = i < m ? alpha[i] : 0.0;
            bl[h] = (i < m && i > 0) ? beta[i] : 0.0;
            bu[h] = (i + 1 < m) ? beta[i + 1] : 0.0;
        }
        // Is the spectrum of T inside [-R, R]?  Exact test by two Sturm counts (every lane runs the same 2 m steps):
        // the number of negative pivots of T - x equals the number of eigenvalues below x.  (The first version used
        // the Gershgorin radius, which overshoots ||T|| by up to 3x: Lanczos matrices of a well-bounded H_eff went to
        // the one-warp Jacobi path below, 245 us a call, 4 % of a chi = 256 2TDVP step.)
        bool inside = true;
        {
            const double edge = cheb.R * (1.0 + 1e-12);
            int below_hi = 0, below_lo = 0;
            double qh = 1.0, ql = 1.0;
            for (int i = 0; i < m; ++i) {
                const double a_i = alpha[i], b2 = (i > 0) ? beta[i] * beta[i] : 0.0;
                qh = a_i - edge - (i > 0 ? b2 / qh : 0.0);
                ql = a_i + edge - (i > 0 ? b2 / ql : 0.0);
                if (qh == 0.0) qh = -1e-300;
                if (ql == 0.0) ql = 1e-300;
                below_hi += qh < 0.0;
                below_lo += ql < 0.0;
            }
            inside = (below_hi == m) && (below_lo == 0);
        }
        if (inside) {   // (uniform) the spectrum of T lies inside [-R, R]
            double* u = jsm;     // current Chebyshev vector, with one zero on either side: u[1 + i]
            const double inv = 1.0 / cheb.R;
            const double sg = t < 0.0 ? -1.0 : 1.0;
            for (int e = lane; e < m + 2; e += 32) u[e] = (e == 1) ? 1.0 : 0.0;
            __syncwarp();
            double prev[2] = {0.0, 0.0}, cur[2], re[2], im[2] = {0.0, 0.0};
#pragma unroll
            for (int h = 0; h < 2; ++h) { cur[h] = (lane + 32 * h == 0) ? 1.0 : 0.0; re[h] = cheb.a[0] * cur[h]; }
            for (int k = 1; k < cheb.n; ++k) {
                double nxt[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = lane + 32 * h;
                    double tv = 0.0;
                    if (i < m) tv = (al[h] * u[1 + i] + bl[h] * u[i] + bu[h] * u[2 + i]) * inv;
                    nxt[h] = (k == 1) ? tv : 2.0 * tv - prev[h];
                }
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = lane + 32 * h;
                    if (i < m) u[1 + i] = nxt[h];
                    prev[h] = cur[h]; cur[h] = nxt[h];
                    // (-i sg)^k: 1, -i sg, -1, +i sg
                    const double ak = cheb.a[k];
                    switch (k & 3) {
                        case 0: re[h] += ak * nxt[h]; break;
                        case 1: im[h] -= sg * ak * nxt[h]; break;
                        case 2: re[h] -= ak * nxt[h]; break;
                        default: im[h] += sg * ak * nxt[h]; break;
                    }
                }
                __syncwarp();
            }
            const double n0 = *norm0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = lane + 32 * h;
                if (i < m) coef[i] = make_double2(n0 * re[h], n0 * im[h]);
            }
            return;
        }
    }
    for (int e = lane; e < m * m; e += 32) {
        const int r = e / m, c = e % m;
        double v = 0.0;
        if (r == c) v = alpha[r];
        else if (r == c + 1) v = beta[r];
        else if (c == r + 1) v = beta[c];
        A[e] = v;
        Z[e] = (r == c) ? 1.0 : 0.0;
    }
    __syncwarp();
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int e = lane; e < m * m; e += 32) {
            const int r = e / m, c = e % m;
            if (r == c) diag += A[e] * A[e];
            else off += A[e] * A[e];
        }
        off = he_warp_sum(off);
        diag = he_warp_sum(diag);
        if (off <= 1e-31 * diag || off == 0.0) break;   // off-diagonal norm below 3e-16 of the matrix norm
        for (int p = 0; p < m - 1; ++p) {
            for (int q = p + 1; q < m; ++q) {
                const double apq = A[p * m + q];
                if (apq == 0.0) continue;       // uniform across the warp
                const double app = A[p * m + p], aqq = A[q * m + q];
                if (fabs(apq) <= 1e-19 * (fabs(app) + fabs(aqq))) continue;   // below the rounding of the diagonal
                const double theta = (aqq - app) / (2.0 * apq);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
                __syncwarp();
                for (int k = lane; k < m; k += 32) {
                    if (k != p && k != q) {
                        const double akp = A[k * m + p], akq = A[k * m + q];
                        const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
                        A[k * m + p] = np_; A[p * m + k] = np_;
                        A[k * m + q] = nq_; A[q * m + k] = nq_;
                    }
                    const double zkp = Z[k * m + p], zkq = Z[k * m + q];
                    Z[k * m + p] = c * zkp - s * zkq;
                    Z[k * m + q] = s * zkp + c * zkq;
                }
                if (lane == 0) {
                    A[p * m + p] = app - tt * apq;
                    A[q * m + q] = aqq + tt * apq;
                    A[p * m + q] = 0.0; A[q * m + p] = 0.0;
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
    const double n0 = *norm0;
    for (int k = lane; k < m; k += 32) {
        double cr = 0.0, ci = 0.0;
        for (int l = 0; l < m; ++l) {
            const double lam = A[l * m + l];
            double sn, cs;
            sincos(t * lam, &sn, &cs);
            const double f = Z[k * m + l] * Z[l];   // Z[0][l]
            cr += f * cs;
            ci -= f * sn;
        }
        coef[k] = make_double2(n0 * cr, n0 * ci);
    }
}

// out = sum_k coef_k V_k
__global__ void __launch_bounds__(HE_THREADS)
he_combine_kernel(double2* __restrict__ out, const double2* __restrict__ V, long long dim, int m,
                  const double2* __restrict__ coef) {
    __shared__ double2 c[HE_MAXK];
    if (threadIdx.x < m) c[threadIdx.x] = coef[threadIdx.x];
    __syncthreads();
    for (long long s = (long long)blockIdx.x * HE_THREADS + threadIdx.x; s < dim; s += (long long)gridDim.x * HE_THREADS) {
        double ar = 0.0, ai = 0.0;
        for (int k = 0; k < m; ++k) {
            const double2 v = V[(long long)k * dim + s];
            ar += c[k].x * v.x - c[k].y * v.y;
            ai += c[k].x * v.y + c[k].y * v.x;
        }
        out[s] = make_double2(ar, ai);
    }
}

// ---- host side ---------------------------------------------------------------------------------------
struct HeffPlan {
    long long plane, dim, t1, t3;
    int ns1, ns2, nblocks, sms;
    // workspace offsets in complex128 units
    long long off_t1, off_t3, off_p2, off_v, off_partial, off_small, total;
};

static int he_sm_count() {
    static int sms[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!sms[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

static int he_split(long long M, long long N, int G, long long steps, int sms) {
    const long long tiles = ((M + 63) / 64) * ((N + 63) / 64) * G;
    const long long target = 2ll * sms;
    if (tiles >= target) return 1;
    long long n = std::min<long long>(std::min<long long>(target / tiles, steps / 4), 16);
    return (int)std::max<long long>(1, n);
}

static int32_t he_plan(const qca_heff_t* h, int m, HeffPlan* p) {
    QCA_REQUIRE(h && h->left && h->right && h->mix_rowptr && h->mix_col && h->mix_val, QCA_ERR_ARG, "NULL argument");
    QCA_REQUIRE(h->dl >= 1 && h->dr >= 1 && h->wl >= 1 && h->wr >= 1, QCA_ERR_ARG, "bad H_eff shape");
    QCA_REQUIRE(h->g == 1 || h->g == 2 || h->g == 4, QCA_ERR_ARG, "g must be 1 (bond), 2 (one site) or 4 (two sites)");
    QCA_REQUIRE(m >= 0 && m <= HE_MAXK, QCA_ERR_ARG, "Krylov dimension %d out of range (<= %d)", m, HE_MAXK);
    p->sms = he_sm_count();
    p->plane = (long long)h->dl * h->dr;
    p->dim = p->plane * h->g;
    p->t1 = p->plane * h->g * h->wl;
    p->t3 = p->plane * h->g * h->wr;
    // (with structural zeros: tiles / segments that are actually computed)
    const bool masks = h->use_masks && h->wl <= 32 && h->wr <= 32;
    int cols = h->g * h->wl, rows_max = h->wr;
    if (masks) {
        cols = 0; rows_max = 1;
        for (int i = 0; i < h->g; ++i) {
            cols += __builtin_popcount(h->col_mask[i]);
            rows_max = std::max(rows_max, __builtin_popcount(h->row_mask[i]));
        }
        QCA_REQUIRE(cols >= 1, QCA_ERR_ARG, "site operator without any entry");
    }
    p->ns1 = he_split((long long)cols * h->dl, h->dr, 1, (h->dl + 15) / 16, p->sms);
    p->ns2 = he_split(h->dl, h->dr, h->g, (long long)rows_max * ((h->dr + 15) / 16), p->sms);
    p->nblocks = (int)std::max<long long>(1, std::min<long long>((p->dim + 2 * HE_THREADS - 1) / (2 * HE_THREADS), 2ll * p->sms));
    long long o = 0;
    p->off_t1 = o; o += p->t1 * p->ns1;
    p->off_t3 = o; o += std::max(p->t3, p->dim);
    p->off_p2 = o; o += p->dim * p->ns2;
    p->off_v = o; o += p->dim * std::max(m, 1);
    p->off_partial = o; o += (long long)p->nblocks * HE_MAXK;
    p->off_small = o; o += 4 * HE_MAXK + p->nblocks;   // alpha, beta, norm0, coef, norm partials (doubles, padded)
    p->total = o;
    return QCA_OK;
}

// T1 = L . psi and T3 = (site operator) T1: the first two steps of H_eff psi (and of an environment update)
static int32_t he_front(const qca_heff_t* h, const HeffPlan& p, const double2* psi, double2* ws, cudaStream_t st) {
    double2* t1 = ws + p.off_t1;
    double2* t3 = ws + p.off_t3;
    const int dl = h->dl, dr = h->dr, wl = h->wl, wr = h->wr, g = h->g;
    // structural zeros of the site operator(s): channels (g, w) of T1 nobody reads, channels (g', n) of T3
    // that are identically zero (two thirds of the work remain for the automaton's MPO)
    const bool masks = h->use_masks && wl <= 32 && wr <= 32;
    // T1[g][w][y][u] = sum_x L[x][(w,y)] psi[g][x][u]
    ZgemmArgs z{};
    z.a = (const double2*)h->left; z.b = psi; z.c = t1;
    z.M = wl * dl; z.N = dr; z.K = dl; z.S = 1; z.G = g;
    z.a_sg = 0; z.a_ss = 0; z.a_sm = 1; z.a_sk = (long long)wl * dl;
    z.b_sg = (long long)dl * dr; z.b_ss = 0; z.b_sk = dr;
    z.c_sg = (long long)wl * dl * dr; z.c_sm = dr;
    z.nsplit = p.ns1; z.c_ssplit = p.t1;
    if (masks) {
        z.use_masks = 1; z.chan_len = dl;
        for (int i = 0; i < g; ++i) { z.chan_mask[i] = h->col_mask[i]; z.seg_mask[i] = 1u; }
    }
    QCA_CHECK(zgemm_launch(z, st));
    const int mix_blocks = (int)std::max<long long>(1, std::min<long long>((p.plane + HE_THREADS - 1) / HE_THREADS, 8ll * p.sms));
    he_mix_kernel<<<mix_blocks, HE_THREADS, 0, st>>>(t1, p.ns1, p.t1, t3, h->mix_rowptr, h->mix_col, (const double2*)h->mix_val,
                                                     g * wr, p.plane);
    QCA_CUDA(cudaGetLastError());
    return QCA_OK;
}

// T1, T3 and the split-K partials of out = H_eff psi; the caller sums the ns2 partials at ws + off_p2
static int32_t he_apply(const qca_heff_t* h, const HeffPlan& p, const double2* psi, double2* ws, cudaStream_t st) {
    QCA_CHECK(he_front(h, p, psi, ws, st));
    double2* t3 = ws + p.off_t3;
    double2* p2 = ws + p.off_p2;
    const int dl = h->dl, dr = h->dr, wl = h->wl, wr = h->wr, g = h->g;
    const bool masks = h->use_masks && wl <= 32 && wr <= 32;
    // out[g][y][v] = sum_n sum_u T3[g][n][y][u] R[u][n][v]
    ZgemmArgs y{};
    y.a = t3; y.b = (const double2*)h->right; y.c = p2;
    y.M = dl; y.N = dr; y.K = dr; y.S = wr; y.G = g;
    y.a_sg = (long long)wr * dl * dr; y.a_ss = (long long)dl * dr; y.a_sm = dr; y.a_sk = 1;
    y.b_sg = 0; y.b_ss = dr; y.b_sk = (long long)wr * dr;
    y.c_sg = (long long)dl * dr; y.c_sm = dr;
    y.nsplit = p.ns2; y.c_ssplit = p.dim;
    if (masks) {
        y.use_masks = 1; y.chan_len = dl;
        for (int i = 0; i < g; ++i) { y.chan_mask[i] = 1u; y.seg_mask[i] = h->row_mask[i]; }
    }
    QCA_CHECK(zgemm_launch(y, st));
    return QCA_OK;
}

// ---- environment update (algorithms/tdvp.py:329-347) ----------------------------------------------------
// E'[r][m][s] = sum_{b,y} T3[b][m][y][r] conj(A[b][y][s]),  T3 = (site operator)(E . A): the front of H_eff applied to
// the site tensor A, then one more DMMA contraction with conj(A) instead of the other environment.
__global__ void __launch_bounds__(HE_THREADS)
he_conj_kernel(double2* __restrict__ dst, const double2* __restrict__ src, long long n) {
    for (long long s = (long long)blockIdx.x * HE_THREADS + threadIdx.x; s < n; s += (long long)gridDim.x * HE_THREADS) {
        const double2 v = src[s];
        dst[s] = make_double2(v.x, -v.y);
    }
}

struct GrowPlan {
    HeffPlan hp;
    int ns3;
    long long out_elems, off_conj, off_p3, total;
};

static int32_t grow_plan(const qca_heff_t* h, GrowPlan* g) {
    QCA_CHECK(he_plan(h, 0, &g->hp));
    QCA_REQUIRE(h->g == 2, QCA_ERR_ARG, "environment updates take a one-site operator (g = 2)");
    g->out_elems = (long long)h->dr * h->wr * h->dr;
    g->ns3 = he_split(h->dr, h->dr, h->wr, 2ll * ((h->dl + 15) / 16), g->hp.sms);
    g->off_conj = g->hp.total;
    g->off_p3 = g->off_conj + g->hp.dim;
    g->total = g->off_p3 + g->out_elems * g->ns3;
    return QCA_OK;
}

}  // namespace qca

extern "C" {

int32_t qca_heff_workspace_bytes(const qca_heff_t* h, int32_t krylov_dim, uint64_t* bytes) {
    QCA_REQUIRE(bytes, QCA_ERR_ARG, "NULL argument");
    qca::HeffPlan p{};
    QCA_CHECK(qca::he_plan(h, krylov_dim, &p));
    *bytes = (uint64_t)p.total * sizeof(double2);
    return QCA_OK;
}

int32_t qca_heff_apply(const qca_heff_t* h, const void* psi, void* out, void* workspace, uint64_t workspace_bytes,
                       void* stream) {
    QCA_REQUIRE(psi && out && workspace, QCA_ERR_ARG, "NULL argument");
    qca::HeffPlan p{};
    QCA_CHECK(qca::he_plan(h, 0, &p));
    QCA_REQUIRE(workspace_bytes >= (uint64_t)p.total * sizeof(double2), QCA_ERR_ARG, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double2* ws = (double2*)workspace;
    QCA_CHECK(qca::he_apply(h, p, (const double2*)psi, ws, st));
    const int blocks = (int)std::max<long long>(1, std::min<long long>((p.dim + qca::HE_THREADS - 1) / qca::HE_THREADS, 8ll * p.sms));
    qca::he_sum_kernel<<<blocks, qca::HE_THREADS, 0, st>>>((double2*)out, ws + p.off_p2, p.ns2, p.dim, p.dim);
    QCA_CUDA(cudaGetLastError());
    return QCA_OK;
}

int32_t qca_env_grow_workspace_bytes(const qca_heff_t* h, uint64_t* bytes) {
    QCA_REQUIRE(bytes, QCA_ERR_ARG, "NULL argument");
    qca::GrowPlan g{};
    QCA_CHECK(qca::grow_plan(h, &g));
    *bytes = (uint64_t)g.total * sizeof(double2);
    return QCA_OK;
}

int32_t qca_env_grow(const qca_heff_t* h, const void* site, void* out, void* workspace, uint64_t workspace_bytes,
                     void* stream) {
    using namespace qca;
    QCA_REQUIRE(site && out && workspace, QCA_ERR_ARG, "NULL argument");
    GrowPlan g{};
    QCA_CHECK(grow_plan(h, &g));
    QCA_REQUIRE(workspace_bytes >= (uint64_t)g.total * sizeof(double2), QCA_ERR_ARG, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double2* ws = (double2*)workspace;
    const HeffPlan& p = g.hp;
    const int dl = h->dl, dr = h->dr, wr = h->wr;
    const int blocks = (int)std::max<long long>(1, std::min<long long>((p.dim + HE_THREADS - 1) / HE_THREADS, 8ll * p.sms));
    he_conj_kernel<<<blocks, HE_THREADS, 0, st>>>(ws + g.off_conj, (const double2*)site, p.dim);
    QCA_CUDA(cudaGetLastError());
    QCA_CHECK(he_front(h, p, (const double2*)site, ws, st));
    // C_m[r][s] = sum_b sum_y T3[b][m][y][r] conj(A)[b][y][s]   (batch m, segments b; A operand with r contiguous)
    ZgemmArgs z{};
    z.a = ws + p.off_t3; z.b = ws + g.off_conj; z.c = ws + g.off_p3;
    z.M = dr; z.N = dr; z.K = dl; z.S = 2; z.G = wr;
    z.a_sg = (long long)dl * dr; z.a_ss = (long long)wr * dl * dr; z.a_sm = 1; z.a_sk = dr;
    z.b_sg = 0; z.b_ss = (long long)dl * dr; z.b_sk = dr;
    z.c_sg = dr; z.c_sm = (long long)wr * dr;
    z.nsplit = g.ns3; z.c_ssplit = g.out_elems;
    QCA_CHECK(zgemm_launch(z, st));
    const int ob = (int)std::max<long long>(1, std::min<long long>((g.out_elems + HE_THREADS - 1) / HE_THREADS, 8ll * p.sms));
    he_sum_kernel<<<ob, HE_THREADS, 0, st>>>((double2*)out, ws + g.off_p3, g.ns3, g.out_elems, g.out_elems);
    QCA_CUDA(cudaGetLastError());
    return QCA_OK;
}

int32_t qca_heff_expm(const qca_heff_t* h, const void* psi, void* out, int32_t m, double t, double spectral_bound,
                      void* workspace, uint64_t workspace_bytes, void* stream) {
    using namespace qca;
    QCA_REQUIRE(psi && out && workspace, QCA_ERR_ARG, "NULL argument");
    QCA_REQUIRE(m >= 1, QCA_ERR_ARG, "Krylov dimension must be >= 1");
    HeffPlan p{};
    QCA_CHECK(he_plan(h, m, &p));
    QCA_REQUIRE(workspace_bytes >= (uint64_t)p.total * sizeof(double2), QCA_ERR_ARG, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double2* ws = (double2*)workspace;
    double2* V = ws + p.off_v;
    double2* partial = ws + p.off_partial;
    double* small = (double*)(ws + p.off_small);
    double* alpha = small;                  // [HE_MAXK]
    double* beta = small + HE_MAXK;         // [HE_MAXK], beta[j] couples j-1 and j
    double* norm0 = small + 2 * HE_MAXK;    // [1] (+ padding)
    double2* coef = (double2*)(small + 4 * HE_MAXK);            // [HE_MAXK] complex
    double* normp = small + 6 * HE_MAXK;    // [nblocks]
    const int nb = p.nblocks;
    QCA_CUDA(cudaMemsetAsync(small, 0, 4 * HE_MAXK * sizeof(double), st));
    he_norm_kernel<<<nb, HE_THREADS, 0, st>>>((const double2*)psi, p.dim, normp);
    he_normalize_kernel<<<nb, HE_THREADS, 0, st>>>(V, (const double2*)psi, p.dim, normp, nb, norm0);
    QCA_CUDA(cudaGetLastError());
    for (int j = 0; j < m; ++j) {
        QCA_CHECK(he_apply(h, p, V + (long long)j * p.dim, ws, st));
        const double2* parts = ws + p.off_p2;
        if (j + 1 < m) {
            double2* w = V + (long long)(j + 1) * p.dim;
            he_dots_kernel<<<nb, HE_THREADS, 0, st>>>(w, parts, p.ns2, p.dim, V, p.dim, j + 1, partial);
            he_update_kernel<<<nb, HE_THREADS, 0, st>>>(w, V, p.dim, j + 1, partial, nb, alpha + j, nullptr);
            // second Gram-Schmidt pass ("twice is enough"), then normalise
            he_dots_kernel<<<nb, HE_THREADS, 0, st>>>(w, nullptr, 0, 0, V, p.dim, j + 1, partial);
            he_update_kernel<<<nb, HE_THREADS, 0, st>>>(w, V, p.dim, j + 1, partial, nb, nullptr, normp);
            he_normalize_kernel<<<nb, HE_THREADS, 0, st>>>(w, w, p.dim, normp, nb, beta + j + 1);
        } else {
            double2* w = ws + p.off_t3;   // scratch: T3 is dead after the second GEMM
            he_dots_kernel<<<nb, HE_THREADS, 0, st>>>(w, parts, p.ns2, p.dim, V + (long long)j * p.dim, p.dim, 1, partial);
            he_alpha_kernel<<<1, 32, 0, st>>>(partial, nb, alpha + j);
        }
        QCA_CUDA(cudaGetLastError());
    }
    ChebCoefs cheb{};
    if (spectral_bound > 0.0) {
        std::vector<double> a;
        QCA_CHECK(chebyshev_plan(spectral_bound * fabs(t), 1e-17, a));
        if ((int)a.size() <= HE_MAXCHEB) {
            cheb.n = (int)a.size();
            cheb.R = spectral_bound;
            std::copy(a.begin(), a.end(), cheb.a);
        }
    }
    const int jsm = std::max(2 * m * m, m + 2) * (int)sizeof(double);
    if (jsm > 48 * 1024)
        QCA_CUDA(cudaFuncSetAttribute(he_tridiag_expm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, jsm));
    he_tridiag_expm_kernel<<<1, 32, jsm, st>>>(alpha, beta, m, norm0, t, coef, cheb);
    const int blocks = (int)std::max<long long>(1, std::min<long long>((p.dim + HE_THREADS - 1) / HE_THREADS, 8ll * p.sms));
    he_combine_kernel<<<blocks, HE_THREADS, 0, st>>>((double2*)out, V, p.dim, m, coef);
    QCA_CUDA(cudaGetLastError());
    return QCA_OK;
}

}  // extern "C"
