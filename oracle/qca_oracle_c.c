/* CPU oracle, C part: matrix-free restatement of the reference's exact path for registers too
 * large for the dense matrices of the reference itself (N > 13).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package links or loads this file; it is the
 * checker behind tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
 * bench.py.  Parity pin: tests/test_oracle_golden.py checks every routine below against the
 * fixtures produced by the unmodified reference (tests/golden/exact_*.npz, hpsi_*.npz).
 *
 * What is restated (reference file:line):
 *   qo_apply_h    H @ v with H = MPO.hamiltonian_from_rules(rules).as_matrix()
 *                 (tensor_networks/mpo.py:161-202, 221-230): H = sum_c sigma^x_c * [the number of alive
 *                 neighbours of cell c within `distance` lies in [lo, hi)], cells beyond both ends dead
 *                 (mpo.py:181-200); cell 0 is the most significant index bit (mps.py:194-208).
 *   qo_step       psi <- exp(-i*pi/2*step_size*H) psi (lautils/lautils.py:45-55 + algorithms/exact.py:26-27).
 *                 The reference diagonalises the dense H; here the same operator is summed as a Chebyshev
 *                 series of the HERMITIAN H in plain complex arithmetic (forward recurrence, Gershgorin
 *                 scale R = N).  Deliberately a different formulation from the CUDA product (no parity
 *                 rotation, no Clenshaw, no tightened bound), so that agreement is evidence.
 *   qo_measure    population and single-site entropy of every cell (tensor_networks/mps.py:100-140):
 *                 rho_c = Tr_rest |psi><psi|, population = rho_c[1,1], entropy = -Tr rho log2 rho.
 *
 * Plain C + OpenMP; complex numbers as interleaved doubles (numpy complex128 layout).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;

int qo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* predicate of cell `cell` at basis state x: count alive neighbours one by one (no bit tricks) */
static inline int cell_active(uint64_t x, int cell, int n, int d, int lo, int hi) {
    int count = 0;
    for (int off = 1; off <= d; ++off) {
        const int a = cell - off, b = cell + off;
        if (a >= 0) count += (int)((x >> (n - 1 - a)) & 1u);
        if (b < n) count += (int)((x >> (n - 1 - b)) & 1u);
    }
    return count >= lo && count < hi;
}

/* out = scale * H v + add_coef * add   (add may be NULL or alias out; out must not alias v) */
static void apply_h_axpy(const cplx* v, cplx* out, int n, int d, int lo, int hi, double scale,
                         const cplx* add, double add_coef) {
    const uint64_t dim = 1ull << n;
#pragma omp parallel for schedule(static)
    for (uint64_t x = 0; x < dim; ++x) {
        double sr = 0.0, si = 0.0;
        for (int cell = 0; cell < n; ++cell) {
            if (cell_active(x, cell, n, d, lo, hi)) {
                const cplx p = v[x ^ (1ull << (n - 1 - cell))];
                sr += p.re; si += p.im;
            }
        }
        cplx r = {scale * sr, scale * si};
        if (add) { r.re += add_coef * add[x].re; r.im += add_coef * add[x].im; }
        out[x] = r;
    }
}

/* Rows [x0, x0 + count) of H v only: out[i] = (H v)[x0 + i].  A bounded sample of one operator application
 * on a register whose full application takes minutes on the host (bench.py, CPU legs). */
int qo_apply_h_rows(const double* v_, double* out_, int n, int d, int lo, int hi, uint64_t x0, uint64_t count) {
    if (n < 1 || n > 40 || d < 1 || x0 + count > (1ull << n)) return 1;
    const cplx* v = (const cplx*)v_;
    cplx* out = (cplx*)out_;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < count; ++i) {
        const uint64_t x = x0 + i;
        double sr = 0.0, si = 0.0;
        for (int cell = 0; cell < n; ++cell) {
            if (cell_active(x, cell, n, d, lo, hi)) {
                const cplx p = v[x ^ (1ull << (n - 1 - cell))];
                sr += p.re; si += p.im;
            }
        }
        out[i].re = sr; out[i].im = si;
    }
    return 0;
}

int qo_apply_h(const double* v, double* out, int n, int d, int lo, int hi) {
    if (n < 1 || n > 40 || d < 1) return 1;
    apply_h_axpy((const cplx*)v, (cplx*)out, n, d, lo, hi, 1.0, NULL, 0.0);
    return 0;
}

/* J_k(z), k = 0..kmax: Miller's backward recurrence, normalised by J_0 + 2 sum J_2k = 1 */
static void bessel_table(double z, int kmax, double* out) {
    if (z == 0.0) { memset(out, 0, (size_t)(kmax + 1) * sizeof(double)); out[0] = 1.0; return; }
    int start = kmax + (int)(20.0 + 12.0 * cbrt(fabs(z) + 1.0) + fabs(z));
    start += start & 1;
    long double jp = 0.0L, jc = 1e-280L, norm = 0.0L;
    long double* tmp = (long double*)calloc((size_t)start + 1, sizeof(long double));
    tmp[start] = jc;
    for (int k = start; k >= 1; --k) {
        const long double jm = (2.0L * k / (long double)z) * jc - jp;
        jp = jc; jc = jm;
        tmp[k - 1] = jc;
        if (fabsl(jc) > 1e200L) {
            for (int q = k - 1; q <= start; ++q) tmp[q] *= 1e-200L;
            jc *= 1e-200L; jp *= 1e-200L;
        }
    }
    for (int k = 0; k <= start; k += 2) norm += (k == 0 ? 1.0L : 2.0L) * tmp[k];
    for (int k = 0; k <= kmax; ++k) out[k] = (double)(tmp[k] / norm);
    free(tmp);
}

/* exp(-i t H) psi = sum_k (2 - delta_k0) (-i)^k J_k(R t) T_k(H/R) psi, T_{k+1} = 2 (H/R) T_k - T_{k-1}.
 * work: 2 vectors of 2^n complex.  first_terms > 0 stops after that many terms (bounded timing
 * samples; the state is then NOT the evolved one).  Returns the number of terms of the full series. */
int qo_step(double* psi_io, double* work, int n, int d, int lo, int hi, double step_size, int first_terms) {
    const uint64_t dim = 1ull << n;
    cplx* acc = (cplx*)psi_io;
    cplx* t_prev = (cplx*)work;
    cplx* t_cur = t_prev + dim;
    const double t = (M_PI / 2.0) * step_size;
    if (t == 0.0) return 0;
    const double R = (double)n;             /* every row of H has at most N ones */
    const double z = R * fabs(t);
    const int kmax = (int)(z + 14.0 * cbrt(z + 1.0) + 40.0);
    double* J = (double*)malloc((size_t)(kmax + 1) * sizeof(double));
    bessel_table(z, kmax, J);
    int nterms = kmax + 1;
    double tail = 0.0;
    while (nterms > 2 && tail + 2.0 * fabs(J[nterms - 1]) < 1e-17) { tail += 2.0 * fabs(J[nterms - 1]); --nterms; }
    const double sg = t < 0.0 ? -1.0 : 1.0;  /* exp(+i|t|H): (-i)^k -> (+i)^k */
    /* T_0 = psi, acc = J_0 psi */
#pragma omp parallel for schedule(static)
    for (uint64_t x = 0; x < dim; ++x) {
        t_prev[x] = acc[x];
        acc[x].re *= J[0]; acc[x].im *= J[0];
    }
    /* T_1 = (H/R) psi */
    apply_h_axpy(t_prev, t_cur, n, d, lo, hi, 1.0 / R, NULL, 0.0);
    int done = 1;
    for (int k = 1; k < nterms; ++k) {
        /* acc += 2 (-i sg)^k J_k T_k */
        const double c = 2.0 * J[k];
        const int q = k & 3;
#pragma omp parallel for schedule(static)
        for (uint64_t x = 0; x < dim; ++x) {
            const cplx v = t_cur[x];
            double a, b;             /* (-i sg)^k * v */
            switch (q) {
                case 0: a = v.re; b = v.im; break;
                case 1: a = sg * v.im; b = -sg * v.re; break;
                case 2: a = -v.re; b = -v.im; break;
                default: a = -sg * v.im; b = sg * v.re; break;
            }
            acc[x].re += c * a; acc[x].im += c * b;
        }
        ++done;
        if (first_terms > 0 && done >= first_terms) break;
        if (k + 1 < nterms) {
            /* T_{k+1} = 2/R H T_k - T_{k-1}, written over T_{k-1} (element-wise dependence only) */
            apply_h_axpy(t_cur, t_prev, n, d, lo, hi, 2.0 / R, t_prev, -1.0);
            cplx* s = t_prev; t_prev = t_cur; t_cur = s;
        }
    }
    free(J);
    return nterms;
}

/* population[c] = rho_c[1,1], entropy[c] = -Tr rho_c log2 rho_c (0 log 0 := 0) */
int qo_measure(const double* psi_, int n, double* population, double* entropy) {
    const cplx* psi = (const cplx*)psi_;
    const uint64_t half = 1ull << (n - 1);
    for (int cell = 0; cell < n; ++cell) {
        const int bit = n - 1 - cell;
        double s0 = 0.0, s1 = 0.0, wr = 0.0, wi = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s0, s1, wr, wi)
        for (uint64_t j = 0; j < half; ++j) {
            const uint64_t x0 = ((j >> bit) << (bit + 1)) | (j & ((1ull << bit) - 1ull));
            const cplx a = psi[x0], b = psi[x0 | (1ull << bit)];
            s0 += a.re * a.re + a.im * a.im;
            s1 += b.re * b.re + b.im * b.im;
            wr += a.re * b.re + a.im * b.im;      /* a * conj(b) */
            wi += a.im * b.re - a.re * b.im;
        }
        population[cell] = s1;
        const double mean = 0.5 * (s0 + s1), dif = 0.5 * (s0 - s1);
        const double rad = sqrt(dif * dif + wr * wr + wi * wi);
        const double lam[2] = {mean + rad, mean - rad};
        double e = 0.0;
        for (int q = 0; q < 2; ++q) if (lam[q] > 0.0) e -= lam[q] * log2(lam[q]);
        entropy[cell] = e;
    }
    return 0;
}

/* product state of amplitudes (sqrt(1-p), sqrt(p)) per cell, cell 0 = top bit (mps.py:35-52, 194-208) */
int qo_product_state(double* psi_, int n, const double* p_alive) {
    cplx* psi = (cplx*)psi_;
    const uint64_t dim = 1ull << n;
    double amp[2 * 64];
    for (int c = 0; c < n; ++c) { amp[2 * c] = sqrt(1.0 - p_alive[c]); amp[2 * c + 1] = sqrt(p_alive[c]); }
#pragma omp parallel for schedule(static)
    for (uint64_t x = 0; x < dim; ++x) {
        double v = 1.0;
        for (int c = 0; c < n; ++c) v *= amp[2 * c + (int)((x >> (n - 1 - c)) & 1u)];
        psi[x].re = v; psi[x].im = 0.0;
    }
    return 0;
}
