// Exact (state-vector) evolution engine: matrix-free rule operator, Clenshaw-Chebyshev
// stepper, fused measurement.  Replaces algorithms/exact.py + lautils.calculate_U +
// MPO.as_matrix + MPS.measure of the reference for the `--algorithm exact` path.
//
// Representation (DESIGN.md "parity rotation"): the resident state is phi with
//     psi[x] = g * i^popcount(x) * phi[x],   g in {1, i}
// held as separate real planes (SoA).  In this frame exp(-i t H) = exp(t K) with the real
// antisymmetric K[x, x^b_c] = P_c(x) * (+1 if cell c dead in x else -1), so re and im planes
// evolve independently under real arithmetic, and a state whose rotated form is real (every
// basis state) needs ONE plane.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "qca_common.cuh"
#include "qca_plan.h"
#include "qca_pass.cuh"
#include "qca_pass3.cuh"
#include "qca_measure.h"
#include "qca_small.h"

namespace qca {

// out = alpha * src
__global__ void scale_kernel(double* __restrict__ out, const double* __restrict__ src, double alpha,
                             unsigned long long n) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = alpha * src[i];
}

// ---------------------------------------------------------------------------
// Host <-> device state conversion
// ---------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_abs(unsigned long long* slot, double v) {
    atomicMax(slot, (unsigned long long)__double_as_longlong(fabs(v)));
}

// staged: interleaved psi.  phi = i^{-popcount} psi, exact (swaps and negations only).
__global__ void unpack_rotate_kernel(const double2* __restrict__ staged, double* __restrict__ re,
                                     double* __restrict__ im, unsigned long long offset,
                                     unsigned long long count, const ShardMap shard,
                                     unsigned long long* maxabs) {
    double mre = 0.0, mim = 0.0;
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < count; i += stride) {
        const unsigned long long x = offset + i;
        const double2 p = staged[i];
        double a, b;
        switch (__popcll(expand_index(x, shard)) & 3) {
            case 0: a = p.x; b = p.y; break;
            case 1: a = p.y; b = -p.x; break;
            case 2: a = -p.x; b = -p.y; break;
            default: a = -p.y; b = p.x; break;
        }
        re[x] = a; im[x] = b;
        mre = fmax(mre, fabs(a)); mim = fmax(mim, fabs(b));
    }
    for (int o = 16; o > 0; o >>= 1) {
        mre = fmax(mre, __shfl_xor_sync(0xffffffffu, mre, o));
        mim = fmax(mim, __shfl_xor_sync(0xffffffffu, mim, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (mre > 0.0) atomic_max_abs(maxabs + 0, mre);
        if (mim > 0.0) atomic_max_abs(maxabs + 1, mim);
    }
}

// staged[i] = (gr + i gi) * i^{popcount + extra} * (re + i im)
__global__ void pack_rotate_kernel(double2* __restrict__ staged, const double* __restrict__ re,
                                   const double* __restrict__ im, unsigned long long offset,
                                   unsigned long long count, const ShardMap shard, int extra_quarter) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < count; i += stride) {
        const unsigned long long x = offset + i;
        const double a = re[x], b = im ? im[x] : 0.0;
        double2 p;
        switch ((__popcll(expand_index(x, shard)) + extra_quarter) & 3) {
            case 0: p.x = a; p.y = b; break;
            case 1: p.x = -b; p.y = a; break;
            case 2: p.x = -a; p.y = -b; break;
            default: p.x = b; p.y = -a; break;
        }
        staged[i] = p;
    }
}

// Product state (mps.py:35-52 + 194-208) directly in the rotated frame.
// amp[2*cell + s] = sqrt(1-p) (s=0) or sqrt(p) (s=1); cell 0 is the top bit.
__global__ void product_state_kernel(double* __restrict__ re, double* __restrict__ im,
                                     const double* __restrict__ amp, int ncells, int local_bits,
                                     const ShardMap shard, unsigned long long* maxabs) {
    double mre = 0.0, mim = 0.0;
    const unsigned long long n = 1ull << local_bits;
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const unsigned long long xf = expand_index(i, shard);
        double v = 1.0;
        for (int cell = 0; cell < ncells; ++cell) {
            const int s = (int)((xf >> (ncells - 1 - cell)) & 1ull);
            v *= amp[2 * cell + s];
        }
        double a, b;
        switch (__popcll(xf) & 3) {
            case 0: a = v; b = 0.0; break;
            case 1: a = 0.0; b = -v; break;
            case 2: a = -v; b = 0.0; break;
            default: a = 0.0; b = v; break;
        }
        re[i] = a; im[i] = b;
        mre = fmax(mre, fabs(a)); mim = fmax(mim, fabs(b));
    }
    for (int o = 16; o > 0; o >>= 1) {
        mre = fmax(mre, __shfl_xor_sync(0xffffffffu, mre, o));
        mim = fmax(mim, __shfl_xor_sync(0xffffffffu, mim, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (mre > 0.0) atomic_max_abs(maxabs + 0, mre);
        if (mim > 0.0) atomic_max_abs(maxabs + 1, mim);
    }
}

// ---------------------------------------------------------------------------
// Measurement: for qubit `bit` the four sums over pairs (x0, x1 = x0 | 1<<bit)
//   s0 = sum |phi[x0]|^2, s1 = sum |phi[x1]|^2, w = sum phi[x0] conj(phi[x1]).
// (rho_01 of the reference's density matrix, mps.py:122-126, is -i*w: same modulus.)
// Per-block partials, then a fixed-order final reduction: deterministic.
// ---------------------------------------------------------------------------
constexpr int kMeasureThreads = 256;

__global__ void __launch_bounds__(kMeasureThreads)
measure_pairs_kernel(const double* __restrict__ re0, const double* __restrict__ im0,
                     const double* __restrict__ re1, const double* __restrict__ im1, int bit,
                     unsigned long long npairs, double* __restrict__ partials) {
    double s0 = 0.0, s1 = 0.0, wr = 0.0, wi = 0.0;
    unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; j < npairs; j += stride) {
        unsigned long long x0, x1;
        if (bit >= 0) {
            x0 = ((j >> bit) << (bit + 1)) | (j & ((1ull << bit) - 1ull));
            x1 = x0 | (1ull << bit);
        } else {
            x0 = x1 = j;  // partner vector is a different buffer (sharded qubit)
        }
        const double a0 = re0[x0], a1 = re1[x1];
        const double b0 = im0 ? im0[x0] : 0.0, b1 = im1 ? im1[x1] : 0.0;
        s0 += a0 * a0 + b0 * b0;
        s1 += a1 * a1 + b1 * b1;
        wr += a0 * a1 + b0 * b1;
        wi += b0 * a1 - a0 * b1;
    }
    __shared__ double sh[4][kMeasureThreads / 32];
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        wr += __shfl_xor_sync(0xffffffffu, wr, o);
        wi += __shfl_xor_sync(0xffffffffu, wi, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = s0; sh[1][warp] = s1; sh[2][warp] = wr; sh[3][warp] = wi; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < kMeasureThreads / 32; ++w) t += sh[threadIdx.x][w];
        partials[4ull * blockIdx.x + threadIdx.x] = t;
    }
}

__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nblocks,
                                       double* __restrict__ out4) {
    // one warp per component; fixed summation order
    const int comp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double t = 0.0;
    for (int b = lane; b < nblocks; b += 32) t += partials[4ull * b + comp];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) out4[comp] = t;
}

// ---------------------------------------------------------------------------
// Cross-GPU barrier over peer-mapped flags (one process per GPU, no NCCL on the data path):
// thread r publishes `epoch` into rank r's flag slot for this rank, then waits until rank r
// has published the same epoch here.  Stream order makes every earlier kernel of this rank
// complete (and visible at system scope) before the flags are written.
// ---------------------------------------------------------------------------
struct BarrierArgs {
    unsigned long long* flags[8];  // flags[r]: rank r's flag array (peer mapped), 8 slots
    int world, rank;
    unsigned long long epoch;
};

__global__ void barrier_kernel(const BarrierArgs b) {
    const int r = threadIdx.x;
    if (r >= b.world || r == b.rank) return;
    __threadfence_system();
    volatile unsigned long long* theirs = b.flags[r] + b.rank;
    *theirs = b.epoch;
    __threadfence_system();
    volatile unsigned long long* mine = b.flags[b.rank] + r;
    const long long t0 = clock64();
    while (*mine < b.epoch) {
        if (clock64() - t0 > 60000000000ll) {  // ~30 s: a peer died; fail instead of hanging the box
            printf("qca_b200: barrier timeout on rank %d waiting for rank %d (epoch %llu)\n", b.rank, r, b.epoch);
            __trap();
        }
    }
    __threadfence_system();
}

// ---------------------------------------------------------------------------
// Engine
// ---------------------------------------------------------------------------
constexpr int kMaxWorld = 8;
constexpr int kIpcBuffers = 7;  // 3 vectors x 2 planes + barrier flags

struct Engine {
    qca_rule_t rule{};
    int device = 0, world = 1, rank = 0, rank_bits = 0, local_bits = 0;
    uint32_t flags = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // host <-> device state traffic runs on its own highest-priority stream: its small pack/unpack kernels are
    // dispatched ahead of the queued CTAs of another engine's tile passes, so the PCIe copies of one state
    // overlap the kernels of another (bench.py e2e pipeline); ordered against `stream` with events
    cudaStream_t io_stream = nullptr;
    cudaEvent_t io_event = nullptr;
    unsigned long long namps = 0;
    ShardMap shard{};     // local index -> global basis-state index
    double* plane[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    double* spare_plane = nullptr;   // the upload plane of a state that turned out real: kept for the next upload (see park_plane)
    int cur = 0;          // vector index of the resident state
    int nplanes = 0;      // 0: no state yet
    bool resolved = true; // sharded: planes agreed with the other ranks
    int g_quarter = 0;    // psi = i^g_quarter * D * phi
    double bound = 0.0;   // spectral bound R
    std::vector<qca_pass_t> passes;
    std::vector<qca_remote_op_t> remote;
    qca_remote_rotation_t rotation{};  // fast kernel: how the remote terms rotate over the passes
    int persistent_ctas = -1;          // sharded fast kernel: -1 persistent with 2 CTAs per SM, 0 one CTA per tile (pass_kernel_v2), > 0 that many CTAs
    int remote_rows = 6;               // rows of the remote operand ring (QCA_REMOTE_RING=4: shallower, deeper local ring)
    double remote_fraction[64] = {};   // per sharded qubit (global bit): fraction of the plane its term reads
    double2* staging = nullptr;
    unsigned long long staging_amps = 0;
    unsigned long long* d_maxabs = nullptr;
    double* d_partials = nullptr;
    double* d_sums = nullptr;
    double* d_amp = nullptr;
    int measure_blocks = 0;
    int num_sms = 148;
    // window lookup tables of the fast pass kernel (device), one pair per pass
    std::vector<unsigned short*> d_tab_lo, d_tab_hi;
    std::vector<int> win_shift;
    std::vector<unsigned> win_mask;
    bool fast_path = false;
    // cluster kernels (qca_pass3.cuh): one GPU, >= 14 qubits, distance <= 4
    bool use_v3 = false;
    std::vector<qca_pass_t> passes3;
    struct V3Tables { unsigned short* thr = nullptr; unsigned short* row = nullptr; int thr_shift = 0, thr_pos = 0; unsigned thr_mask = 0;
                      int row_shift = 0, row_pos = 0; unsigned row_mask = 0; };
    std::vector<V3Tables> tabs3;
    // sharding
    unsigned long long* d_flags = nullptr;
    double* peer_plane[kMaxWorld][3][2] = {};
    unsigned long long* peer_flags[kMaxWorld] = {};
    bool peers_ready = false;
    bool graph_warm = false;    // an eager step has run (function attributes set, work planes allocated)
    bool bound_agreed = false;  // sharded: the ranks have been given one common spectral bound (qca_exact_set_spectral_bound)
    bool loopback = false;   // profiling aid: "partners" are this rank's own planes, no cross-rank barrier
    unsigned long long epoch = 0;
    // CUDA graphs of a whole step for launch-bound registers (14..kGraphMaxBits local qubits, one GPU)
    struct StepGraph { double step_size; int cur, nplanes; double bound; cudaGraphExec_t exec; uint64_t launches, pass_launches;
                       double pass_bytes; int terms, cur_after; };
    std::vector<StepGraph> graphs;
    std::vector<const void*> configured;   // pass kernels whose shared-memory attribute has been set
    // stats
    qca_exact_stats_t st{};
    struct ProfiledLaunch { cudaEvent_t ev0, ev1; int pass; };
    std::vector<ProfiledLaunch> prof;

    size_t plane_bytes() const { return (size_t)namps * sizeof(double); }
};

// captured steps hold the plane addresses: drop them whenever a plane comes or goes
static void invalidate_graphs(Engine* e) {
    for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
    e->graphs.clear();
}

static int32_t ensure_plane(Engine* e, int v, int p) {
    if (e->plane[v][p]) return QCA_OK;
    invalidate_graphs(e);
    if (e->spare_plane) {   // (already counted in device_bytes)
        e->plane[v][p] = e->spare_plane;
        e->spare_plane = nullptr;
        return QCA_OK;
    }
    cudaError_t err = cudaMalloc(&e->plane[v][p], e->plane_bytes());
    if (err != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of a %zu-byte state plane failed: %s", e->plane_bytes(), cudaGetErrorString(err));
        return QCA_ERR_NOMEM;
    }
    e->st.device_bytes += (double)e->plane_bytes();
    return QCA_OK;
}

static void release_plane(Engine* e, int v, int p) {
    if (!e->plane[v][p] || e->world > 1) return;  // sharded planes are exported: never freed
    invalidate_graphs(e);
    cudaFree(e->plane[v][p]);
    e->plane[v][p] = nullptr;
    e->st.device_bytes -= (double)e->plane_bytes();
}

// Take a plane out of use WITHOUT freeing it: every upload needs a second plane until the state is known to be real
// in the rotated frame, and cudaFree / cudaMalloc of 8 GiB per upload synchronise the whole device -- in the
// pipelined end-to-end path (two engines taking turns) an upload waited ~0.5 s for the OTHER engine's step to finish
// (torch.profiler timeline, round 2).  One plane is parked per engine; a second one is freed.
static void park_plane(Engine* e, int v, int p) {
    if (!e->plane[v][p] || e->world > 1) return;
    if (e->spare_plane) { release_plane(e, v, p); return; }
    invalidate_graphs(e);
    e->spare_plane = e->plane[v][p];
    e->plane[v][p] = nullptr;
}

// entry[w] = activity of the middle K bits of the (K + 2d)-bit window w
static std::vector<unsigned short> window_table(int K, int d, uint32_t imask) {
    std::vector<unsigned short> tab((size_t)1 << (K + 2 * d));
    for (size_t w = 0; w < tab.size(); ++w) {
        const unsigned long long act = activity_word<unsigned long long>((unsigned long long)w, d, imask);
        tab[w] = (unsigned short)((act >> d) & ((1ull << K) - 1ull));
    }
    return tab;
}

static int32_t upload_table(Engine* e, const std::vector<unsigned short>& tab, unsigned short** out) {
    QCA_CUDA(cudaMalloc(out, tab.size() * sizeof(unsigned short)));
    QCA_CUDA(cudaMemcpy(*out, tab.data(), tab.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
    e->st.device_bytes += (double)(tab.size() * sizeof(unsigned short));
    return QCA_OK;
}

// The fast kernel needs full 13-bit tiles, window tables of sane size (distance <= 4) and, when
// sharded, a later pass to carry the remote terms.
// Table of a later pass: its flipped local bits [H0, H0+M) occupy the global positions G0..G1 (with
// the sharded qubits that lie in between); the window is the global bits [G0-d, G1+d] and the entry
// holds the predicates of the M LOCAL bits only (sharded positions squeezed out).
static std::vector<unsigned short> span_table(Engine* e, const qca_pass_t& ps, int* win_shift, unsigned* win_mask) {
    const int d = e->rule.distance;
    const uint32_t imask = interval_mask_of(e->rule.act_lo, e->rule.act_hi);
    const int g0 = global_pos(ps.high_start, e->shard), g1 = global_pos(ps.high_start + ps.high_bits - 1, e->shard);
    const int span = g1 - g0 + 1;
    *win_shift = g0 - d;
    *win_mask = (1u << (span + 2 * d)) - 1u;
    std::vector<unsigned short> tab((size_t)1 << (span + 2 * d));
    for (size_t w = 0; w < tab.size(); ++w) {
        const unsigned long long act = activity_word<unsigned long long>((unsigned long long)w, d, imask) >> d;
        unsigned out = 0;
        for (int q = 0; q < ps.high_bits; ++q)
            out |= (unsigned)((act >> (global_pos(ps.high_start + q, e->shard) - g0)) & 1ull) << q;
        tab[w] = (unsigned short)out;
    }
    return tab;
}

// The fast kernel needs full 13-bit tiles, window tables of sane size (distance <= 4), every sharded
// qubit above the first tile and, when sharded, a later pass to carry remote terms.
static int32_t build_tables(Engine* e) {
    const int d = e->rule.distance;
    e->fast_path = (e->local_bits >= kTile) && d <= 4 && (e->world == 1 || e->passes.size() >= 2);
    for (int j = 0; j < e->shard.nins; ++j) e->fast_path = e->fast_path && e->shard.pos[j] >= kTile;
    for (const qca_remote_op_t& op : e->remote) e->fast_path = e->fast_path && op.mask != 0;
    if (!e->remote.empty() && e->rotation.nslots == 0) e->fast_path = false;   // more terms than the fast kernel has slots
    e->d_tab_lo.assign(e->passes.size(), nullptr);
    e->d_tab_hi.assign(e->passes.size(), nullptr);
    e->win_shift.assign(e->passes.size(), 0);
    e->win_mask.assign(e->passes.size(), 0);
    if (!e->fast_path) return QCA_OK;
    const uint32_t imask = interval_mask_of(e->rule.act_lo, e->rule.act_hi);
    for (size_t i = 0; i < e->passes.size(); ++i) {
        const qca_pass_t& ps = e->passes[i];
        if (ps.high_bits == 0) {
            QCA_CHECK(upload_table(e, window_table(9 - d, d, imask), &e->d_tab_lo[i]));
            QCA_CHECK(upload_table(e, window_table(4 + d, d, imask), &e->d_tab_hi[i]));
        } else {
            int shift = 0; unsigned mask = 0;
            QCA_CHECK(upload_table(e, span_table(e, ps, &shift, &mask), &e->d_tab_hi[i]));
            e->win_shift[i] = shift; e->win_mask[i] = mask;
        }
    }
    return QCA_OK;
}

// Window tables of the cluster kernels.  Pass 0: thread part = tile bits [0, 10-d) from the window x[0,10) << d,
// row part = bits [10-d, 14+CB) from the window bits [10-2d, 14+CB+d).  Later pass with strided bits at H0 and
// MT = 10-L of them on thread bits: thread part = the KA = max(0, MT-d) lowest strided bits (window
// [H0-d, H0+KA+d): thread and tile-base bits only), row part = the rest including the cluster bits.
static int32_t build_tables_v3(Engine* e) {
    const int d = e->rule.distance;
    const uint32_t imask = interval_mask_of(e->rule.act_lo, e->rule.act_hi);
    e->tabs3.assign(e->passes3.size(), Engine::V3Tables{});
    for (size_t i = 0; i < e->passes3.size(); ++i) {
        const qca_pass_t& ps = e->passes3[i];
        Engine::V3Tables& tb = e->tabs3[i];
        const int cb = ps.reserved;
        if (ps.high_bits == 0) {
            QCA_CHECK(upload_table(e, window_table(kRowShift3 - d, d, imask), &tb.thr));
            const int K = kRegHigh + d + cb;
            QCA_CHECK(upload_table(e, window_table(K, d, imask), &tb.row));
            tb.row_shift = kRowShift3 - 2 * d; tb.row_mask = (1u << (K + 2 * d)) - 1u; tb.row_pos = kRowShift3 - d;
        } else {
            const int L = ps.low_bits, H0 = ps.high_start, M = ps.high_bits;
            const int MT = std::max(0, kRowShift3 - L);
            const int KA = std::max(0, MT - d);
            if (KA > 0) {
                QCA_CHECK(upload_table(e, window_table(KA, d, imask), &tb.thr));
                tb.thr_shift = H0 - d; tb.thr_mask = (1u << (KA + 2 * d)) - 1u; tb.thr_pos = L;
            }
            const int K = M + cb - KA;
            QCA_CHECK(upload_table(e, window_table(K, d, imask), &tb.row));
            tb.row_shift = H0 + KA - d; tb.row_mask = (1u << (K + 2 * d)) - 1u; tb.row_pos = L + KA;
        }
    }
    return QCA_OK;
}

static int32_t launch_pass3(Engine* e, size_t pass_index, Pass3Args& a, int nunc) {
    const qca_pass_t& ps = e->passes3[pass_index];
    const Engine::V3Tables& tb = e->tabs3[pass_index];
    const int cb = ps.reserved;
    a.low_bits = ps.low_bits; a.high_start = ps.high_start; a.high_bits = ps.high_bits;
    a.distance = e->rule.distance;
    a.tab_thr = tb.thr; a.tab_row = tb.row;
    a.thr_shift = tb.thr_shift; a.thr_pos = tb.thr_pos; a.thr_mask = tb.thr_mask;
    a.row_shift = tb.row_shift; a.row_pos = tb.row_pos; a.row_mask = tb.row_mask;
    a.ntiles = e->namps >> (kTile3 + cb);
    for (unsigned r = 0; r < 16; ++r) {
        const unsigned long long y = (unsigned long long)r << kRowShift3;
        a.row_xg[r] = (y & ((1ull << ps.low_bits) - 1ull)) | ((y >> ps.low_bits) << ps.high_start);
    }
    const bool wide = (e->local_bits > 31);
    Pass3Kernel kern = wide ? pass3_kernel_u64(ps.low_bits, nunc, cb) : pass3_kernel_u32(ps.low_bits, nunc, cb);
    QCA_REQUIRE(kern != nullptr, QCA_ERR_UNSUPPORTED, "no cluster tile-pass kernel for %d low bits, %d operands, %d cluster bits",
                ps.low_bits, nunc, cb);
    if (std::find(e->configured.begin(), e->configured.end(), (const void*)kern) == e->configured.end()) {
        QCA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPass3SmemBytes));
        e->configured.push_back((const void*)kern);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(a.ntiles << cb), e->nplanes, 1);
    cfg.blockDim = dim3(kPass3Threads, 1, 1);
    cfg.dynamicSmemBytes = kPass3SmemBytes;
    cfg.stream = e->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1u << cb; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = cb ? 1 : 0;
    const bool profile = (e->flags & QCA_FLAG_PROFILE) != 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (profile) {
        QCA_CUDA(cudaEventCreate(&ev0)); QCA_CUDA(cudaEventCreate(&ev1));
        QCA_CUDA(cudaEventRecord(ev0, e->stream));
    }
    QCA_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    if (profile) {
        QCA_CUDA(cudaEventRecord(ev1, e->stream));
        e->prof.push_back({ev0, ev1, (int)pass_index});
    }
    e->st.pass_bytes += (2.0 + nunc) * (double)e->plane_bytes() * e->nplanes;   // in + out + local operands, once each
    e->st.pass_launches += 1;
    e->st.kernel_launches += 1;
    return QCA_OK;
}

static int32_t apply_operator_v3(Engine* e, int v_out, int v_in, int v_a, double alpha, int v_c, double beta, double gamma) {
    for (size_t i = 0; i < e->passes3.size(); ++i) {
        Pass3Args a{};
        int ns = 0;
        auto add_local = [&](int v, double coef) {
            for (int p = 0; p < 2; ++p) a.opnd[ns][p] = e->plane[v][p];
            a.coef[ns++] = coef;
        };
        for (int p = 0; p < 2; ++p) { a.in[p] = e->plane[v_in][p]; a.out[p] = e->plane[v_out][p]; }
        if (i == 0) {
            if (v_a >= 0) add_local(v_a, alpha);
            if (v_c >= 0) add_local(v_c, beta);
        } else {
            add_local(v_out, 1.0);
        }
        a.gamma = gamma;
        QCA_CHECK(launch_pass3(e, i, a, ns));
    }
    return QCA_OK;
}

static int32_t launch_barrier(Engine* e) {
    if (e->world == 1 || e->loopback) return QCA_OK;
    QCA_REQUIRE(e->peers_ready, QCA_ERR_STATE, "sharded engine used before qca_exact_ipc_import");
    BarrierArgs b{};
    for (int r = 0; r < e->world; ++r) b.flags[r] = e->peer_flags[r];
    b.world = e->world; b.rank = e->rank; b.epoch = ++e->epoch;
    barrier_kernel<<<1, 32, 0, e->stream>>>(b);
    QCA_CUDA(cudaGetLastError());
    e->st.kernel_launches += 1;
    return QCA_OK;
}

static int32_t launch_pass(Engine* e, size_t pass_index, PassArgs& a) {
    const qca_pass_t& ps = e->passes[pass_index];
    a.low_bits = ps.low_bits; a.high_start = ps.high_start; a.high_bits = ps.high_bits;
    a.flip_mask = ps.flip_mask;
    a.shard = e->shard;
    a.win_shift = e->win_shift[pass_index];
    a.win_mask = e->win_mask[pass_index];
    a.distance = e->rule.distance;
    a.interval_mask = interval_mask_of(e->rule.act_lo, e->rule.act_hi);
    a.tab_lo = e->d_tab_lo[pass_index];
    a.tab_hi = e->d_tab_hi[pass_index];
    const int T = ps.low_bits + ps.high_bits;
    a.ntiles = e->namps >> T;
    if (T == kTile) {   // fast kernel: index bits of register row r (tile bits 9..12), expanded to global positions
        ShardMap bits_only = e->shard;
        bits_only.rank_or = 0;
        for (unsigned r = 0; r < 16; ++r) {
            const unsigned long long y = (unsigned long long)r << kRowShift;
            const unsigned long long off = (y & ((1ull << ps.low_bits) - 1ull)) | ((y >> ps.low_bits) << ps.high_start);
            a.row_xg[r] = expand_index(off, bits_only);
        }
    }
    int smem = (int)(sizeof(double) << T);
    const bool wide = (e->rule.ncells > 31);
    PassKernel kern = nullptr;
    int nunc = 0;
    while (nunc < a.nstreams && a.s[nunc].bit < 0) ++nunc;  // local operands come first
    if (e->fast_path && T == kTile && nunc == a.nstreams)   // remote terms travel in a.rs
        kern = wide ? fast_pass_kernel_u64(ps.low_bits, nunc, a.nrem, e->remote_rows)
                    : fast_pass_kernel_u32(ps.low_bits, nunc, a.nrem, e->remote_rows);
    // sharded launches: persistent CTAs whose operand rings run across tile boundaries (pass_kernel_v2p)
    bool persistent = false;
    if (kern != nullptr && a.nrem > 0 && e->persistent_ctas != 0) {
        const int rd = (e->remote_rows == 4) ? 4 : 8;
        PassKernel pk = wide ? persistent_pass_kernel_u64(ps.low_bits, nunc, a.nrem, rd)
                             : persistent_pass_kernel_u32(ps.low_bits, nunc, a.nrem, rd);
        if (pk != nullptr) { kern = pk; persistent = true; }
    }
    const bool fast = kern != nullptr;
    if (fast) smem = kPassSmemBytes;
    else {
        QCA_REQUIRE(a.nrem == 0, QCA_ERR_UNSUPPORTED, "no fast tile-pass kernel for %d local operands and %d remote slots",
                    nunc, a.nrem);
        kern = generic_pass_kernel(wide);
    }
    // (once per kernel and engine: nothing but launches may happen while a step is being captured into a graph)
    if (std::find(e->configured.begin(), e->configured.end(), (const void*)kern) == e->configured.end()) {
        QCA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        e->configured.push_back((const void*)kern);
    }
    unsigned gx = (unsigned)std::min<unsigned long long>(a.ntiles, 1u << 30);
    if (persistent)   // two resident CTAs per SM (QCA_PERSISTENT_CTAS: tests walk many tiles per CTA on small registers)
        gx = (unsigned)std::min<unsigned long long>(a.ntiles, e->persistent_ctas > 0 ? (unsigned long long)e->persistent_ctas
                                                                                        : 2ull * e->num_sms / std::max(e->nplanes, 1));
    dim3 grid(gx, e->nplanes, 1);
    const bool profile = (e->flags & QCA_FLAG_PROFILE) != 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (profile) {
        QCA_CUDA(cudaEventCreate(&ev0)); QCA_CUDA(cudaEventCreate(&ev1));
        QCA_CUDA(cudaEventRecord(ev0, e->stream));
    }
    kern<<<grid, kPassThreads, smem, e->stream>>>(a);
    QCA_CUDA(cudaGetLastError());
    if (profile) {
        QCA_CUDA(cudaEventRecord(ev1, e->stream));
        e->prof.push_back({ev0, ev1, (int)pass_index});
    }
    // algorithmic bytes: every LOCAL operand vector once per plane (remote operands travel over
    // NVLink and are accounted separately)
    double vecs = 2.0;  // in + out
    double remote = 0.0;
    for (int k = 0; k < a.nstreams; ++k) {
        if (a.s[k].bit < 0) vecs += 1.0;
        else remote += e->remote_fraction[a.s[k].bit];
    }
    e->st.pass_bytes += vecs * (double)e->plane_bytes() * e->nplanes;
    e->st.remote_bytes += remote * (double)e->plane_bytes() * e->nplanes;
    e->st.pass_launches += 1;
    e->st.kernel_launches += 1;
    return QCA_OK;
}

// out = alpha*a + beta*c + gamma * K in      (a, c optional; c may be out).  Sharded: `in` must be
// complete on every rank (barrier) because remote terms read the partner's copy of it.
static int32_t apply_operator(Engine* e, int v_out, int v_in, int v_a, double alpha, int v_c, double beta,
                              double gamma) {
    if (e->use_v3) return apply_operator_v3(e, v_out, v_in, v_a, alpha, v_c, beta, gamma);
    QCA_CHECK(launch_barrier(e));
    for (size_t i = 0; i < e->passes.size(); ++i) {
        PassArgs a{};
        int ns = 0;
        auto add_local = [&](int v, double coef) {
            EpiStream& st = a.s[ns++];
            for (int p = 0; p < 2; ++p) st.ptr[p] = e->plane[v][p];
            st.coef = coef; st.mask = ~0u; st.shift = 0; st.bit = -1;
        };
        for (int p = 0; p < 2; ++p) { a.in[p] = e->plane[v_in][p]; a.out[p] = e->plane[v_out][p]; }
        if (i == 0) {
            if (v_a >= 0) add_local(v_a, alpha);
            if (v_c >= 0) add_local(v_c, beta);
        } else {
            add_local(v_out, 1.0);
        }
        if (e->fast_path && !e->remote.empty()) {
            // fast kernel: every launch has the same remote slots; which term a slot carries at x
            // follows the rotation of x (qca_plan_rotation)
            const qca_remote_rotation_t& rot = e->rotation;
            a.nrem = rot.nslots; a.rot_word = rot.rot_word; a.rot_shift = rot.rot_shift;
            for (int sl = 0; sl < rot.nslots; ++sl)
                for (int r = 0; r < rot.npasses; ++r) {
                    const int j = rot.op_of[i][sl][r];
                    if (j < 0) continue;
                    const qca_remote_op_t& op = e->remote[j];
                    RemoteAlt& al = a.rs[sl].alt[r];
                    for (int p = 0; p < 2; ++p) al.ptr[p] = e->peer_plane[op.partner][v_in][p];
                    al.coef = gamma * (double)op.sign; al.mask = op.mask; al.shift = op.shift;
                }
        } else {
            for (const qca_remote_op_t& op : e->remote) {   // generic kernel: static placement
                if (op.pass != (int)i) continue;
                QCA_REQUIRE(ns < kMaxStreams, QCA_ERR_UNSUPPORTED, "too many remote terms in one pass");
                EpiStream& st = a.s[ns++];
                for (int p = 0; p < 2; ++p) st.ptr[p] = e->peer_plane[op.partner][v_in][p];
                st.coef = gamma * (double)op.sign; st.mask = op.mask; st.shift = op.shift; st.bit = op.qubit;
            }
        }
        a.nstreams = ns;
        a.gamma = gamma;
        QCA_CHECK(launch_pass(e, i, a));
    }
    if (e->fast_path)   // (the generic path counts its remote streams per launch)
        for (const qca_remote_op_t& op : e->remote)
            e->st.remote_bytes += e->remote_fraction[op.qubit] * (double)e->plane_bytes() * e->nplanes;
    return QCA_OK;
}

static int32_t launch_scale(Engine* e, int v_out, int v_src, double alpha) {
    for (int p = 0; p < e->nplanes; ++p) {
        const unsigned blocks = (unsigned)std::min<unsigned long long>((e->namps + 255) / 256, (unsigned long long)e->num_sms * 16);
        scale_kernel<<<blocks, 256, 0, e->stream>>>(e->plane[v_out][p], e->plane[v_src][p], alpha, e->namps);
        QCA_CUDA(cudaGetLastError());
        e->st.kernel_launches += 1;
    }
    return QCA_OK;
}

static int32_t ensure_work_planes(Engine* e) {
    for (int v = 0; v < 3; ++v)
        for (int p = 0; p < e->nplanes; ++p) QCA_CHECK(ensure_plane(e, v, p));
    return QCA_OK;
}

// phi <- exp(t K) phi by Clenshaw summation of sum_k a_k U_k(K/R) phi with
// U_{k+1} = 2x U_k + U_{k-1}:  B_k = a_k phi + (2/R) K B_{k+1} + B_{k+2},
// phi' = a_0 phi + (1/R) K B_1 + B_2.
static int32_t step_once(Engine* e, double step_size) {
    const double t = (M_PI / 2.0) * step_size;
    if (t == 0.0 || e->bound == 0.0) return QCA_OK;   // R == 0: no rule term can ever fire, H == 0 (e.g. a single cell)
    const double sgn = t < 0.0 ? -1.0 : 1.0;
    const double z = e->bound * fabs(t);
    std::vector<double> a;
    QCA_CHECK(chebyshev_plan(z, 1e-15, a));
    const int K = (int)a.size() - 1;  // >= 1
    e->st.last_terms = K + 1;
    if (e->world == 1 && e->local_bits <= kSmallMaxBits && K + 1 <= kSmallMaxTerms && !(e->flags & QCA_FLAG_TILE_PATH_ONLY)) {
        // whole step in one kernel, vectors in shared memory (csrc/qca_small.cu)
        SmallStepArgs sa{};
        for (int p = 0; p < 2; ++p) { sa.src[p] = e->plane[e->cur][p]; sa.dst[p] = e->plane[e->cur][p]; }
        sa.nbits = e->local_bits; sa.distance = e->rule.distance; sa.nterms = K + 1;
        sa.interval_mask = interval_mask_of(e->rule.act_lo, e->rule.act_hi);
        sa.gamma = sgn * 2.0 / e->bound; sa.gamma_last = sgn / e->bound;
        for (int k = 0; k <= K; ++k) sa.coef[k] = a[k];
        QCA_CHECK(launch_small_step(sa, e->nplanes, e->stream));
        e->st.kernel_launches += 1;
        return QCA_OK;
    }
    QCA_CHECK(ensure_work_planes(e));
    const int P = e->cur;
    int X = (P + 1) % 3, Y = (P + 2) % 3;
    const double R = e->bound;
    // (sharded: the barrier that opens the first apply also orders this write after every
    //  partner's remote reads of X in the previous step)
    QCA_CHECK(launch_barrier(e));
    QCA_CHECK(launch_scale(e, X, P, a[K]));  // B_K
    for (int k = K - 1; k >= 1; --k) {
        const bool first = (k == K - 1);  // B_{K+1} = 0
        QCA_CHECK(apply_operator(e, Y, X, P, a[k], first ? -1 : Y, 1.0, sgn * 2.0 / R));
        std::swap(X, Y);
    }
    QCA_CHECK(apply_operator(e, Y, X, P, a[0], (K == 1) ? -1 : Y, 1.0, sgn / R));
    e->cur = Y;
    return QCA_OK;
}

// Registers of 14..kGraphMaxBits qubits on one GPU: a step is 100-200 launches of a few microseconds each, i.e.
// launch-bound from the host.  The launch sequence of step_once depends only on (step size, resident vector
// index, planes, bound): it is captured once per such key and replayed as one cudaGraphLaunch.
constexpr int kGraphMaxBits = 24;

static int32_t step_graphed(Engine* e, double step_size) {
    const bool eligible = e->world == 1 && e->local_bits > kSmallMaxBits && e->local_bits <= kGraphMaxBits &&
                          !(e->flags & (QCA_FLAG_PROFILE | QCA_FLAG_NO_GRAPH)) && step_size != 0.0 && e->bound != 0.0;
    if (!eligible) return step_once(e, step_size);
    for (const Engine::StepGraph& g : e->graphs) {
        if (g.step_size == step_size && g.cur == e->cur && g.nplanes == e->nplanes && g.bound == e->bound) {
            QCA_CUDA(cudaGraphLaunch(g.exec, e->stream));
            e->st.kernel_launches += g.launches; e->st.pass_launches += g.pass_launches; e->st.pass_bytes += g.pass_bytes;
            e->st.last_terms = g.terms; e->cur = g.cur_after;
            return QCA_OK;
        }
    }
    // first use of this key: make sure nothing inside the capture allocates or configures
    QCA_CHECK(ensure_work_planes(e));
    if (e->graphs.size() >= 12) return step_once(e, step_size);   // a caller cycling through step sizes: stay eager
    Engine::StepGraph g{};
    g.step_size = step_size; g.cur = e->cur; g.nplanes = e->nplanes; g.bound = e->bound;
    const qca_exact_stats_t before = e->st;
    // the function attributes of the pass kernels are set on their first eager launch
    if (!e->graph_warm) {
        e->graph_warm = true;
        return step_once(e, step_size);
    }
    QCA_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    const int32_t rc = step_once(e, step_size);
    cudaGraph_t graph = nullptr;
    const cudaError_t end = cudaStreamEndCapture(e->stream, &graph);
    if (rc != QCA_OK || end != cudaSuccess || graph == nullptr) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (rc == QCA_OK) set_error("CUDA graph capture of a step failed: %s", cudaGetErrorString(end));
        return rc != QCA_OK ? rc : QCA_ERR_CUDA;
    }
    const cudaError_t inst = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (inst != cudaSuccess) { set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(inst)); return QCA_ERR_CUDA; }
    g.launches = e->st.kernel_launches - before.kernel_launches;
    g.pass_launches = e->st.pass_launches - before.pass_launches;
    g.pass_bytes = e->st.pass_bytes - before.pass_bytes;
    g.terms = e->st.last_terms; g.cur_after = e->cur;
    e->graphs.push_back(g);
    QCA_CUDA(cudaGraphLaunch(g.exec, e->stream));   // the capture recorded the step without running it
    return QCA_OK;
}

static int32_t finish_profile(Engine* e) {
    if (e->prof.empty()) return QCA_OK;
    QCA_CUDA(cudaStreamSynchronize(e->stream));
    for (auto& pr : e->prof) {
        float ms = 0.f;
        QCA_CUDA(cudaEventElapsedTime(&ms, pr.ev0, pr.ev1));
        e->st.profiled_pass_ms += ms;
        e->st.profiled_pass_launches += 1;
        if (pr.pass >= 0 && pr.pass < 4) e->st.profiled_ms_by_pass[pr.pass] += ms;
        cudaEventDestroy(pr.ev0); cudaEventDestroy(pr.ev1);
    }
    e->prof.clear();
    return QCA_OK;
}

static int32_t read_plane_flags(Engine* e, bool* has_re, bool* has_im) {
    unsigned long long h[2];
    QCA_CUDA(cudaMemcpyAsync(h, e->d_maxabs, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
    QCA_CUDA(cudaStreamSynchronize(e->stream));
    *has_re = h[0] != 0; *has_im = h[1] != 0;
    return QCA_OK;
}

// Decide the number of planes from the exact maxima (of all ranks when sharded).
static int32_t settle_planes(Engine* e, bool has_re, bool has_im) {
    e->g_quarter = 0;
    if ((e->flags & QCA_FLAG_FORCE_COMPLEX) || (has_re && has_im)) {
        e->nplanes = 2;
    } else if (has_im) {  // purely imaginary rotated state: psi = i * D * (im plane)
        std::swap(e->plane[e->cur][0], e->plane[e->cur][1]);
        // every rank takes the same decision, so the views of the partners' planes swap too
        for (int r = 0; r < e->world && e->world > 1; ++r) std::swap(e->peer_plane[r][e->cur][0], e->peer_plane[r][e->cur][1]);
        e->g_quarter = 1;
        e->nplanes = 1;
    } else {
        e->nplanes = 1;
    }
    if (e->nplanes == 1) park_plane(e, e->cur, 1);
    e->resolved = true;
    return QCA_OK;
}

static int32_t prepare_upload(Engine* e) {
    QCA_CHECK(ensure_plane(e, e->cur, 0));
    QCA_CHECK(ensure_plane(e, e->cur, 1));
    QCA_CUDA(cudaMemsetAsync(e->d_maxabs, 0, 2 * sizeof(unsigned long long), e->stream));
    return QCA_OK;
}

static int32_t finish_upload(Engine* e) {
    if (e->world > 1) {  // the ranks must agree: qca_exact_resolve_planes
        e->nplanes = 2;
        e->resolved = false;
        return QCA_OK;
    }
    bool re, im;
    QCA_CHECK(read_plane_flags(e, &re, &im));
    return settle_planes(e, re, im);
}

static int32_t ensure_staging(Engine* e) {
    if (e->staging) return QCA_OK;
    e->staging_amps = std::min<unsigned long long>(e->namps, 1ull << 22);
    QCA_CUDA(cudaMalloc(&e->staging, e->staging_amps * sizeof(double2)));
    e->st.device_bytes += (double)(e->staging_amps * sizeof(double2));
    return QCA_OK;
}

static unsigned stream_blocks(Engine* e, unsigned long long n) {
    return (unsigned)std::max<unsigned long long>(1, std::min<unsigned long long>((n + 255) / 256, (unsigned long long)e->num_sms * 16));
}

// make the IO stream wait for everything queued on the compute stream so far
static int32_t io_after_compute(Engine* e) {
    QCA_CUDA(cudaEventRecord(e->io_event, e->stream));
    QCA_CUDA(cudaStreamWaitEvent(e->io_stream, e->io_event, 0));
    return QCA_OK;
}

static int32_t upload_vector(Engine* e, int v, const double* host, bool track) {
    QCA_CHECK(ensure_staging(e));
    QCA_CHECK(io_after_compute(e));
    for (unsigned long long off = 0; off < e->namps; off += e->staging_amps) {
        const unsigned long long cnt = std::min(e->staging_amps, e->namps - off);
        // (stream order on the IO stream protects the staging buffer: copy i+1 starts after unpack i)
        QCA_CUDA(cudaMemcpyAsync(e->staging, host + 2 * off, cnt * sizeof(double2), cudaMemcpyHostToDevice, e->io_stream));
        unpack_rotate_kernel<<<stream_blocks(e, cnt), 256, 0, e->io_stream>>>(
            e->staging, e->plane[v][0], e->plane[v][1], off, cnt, e->shard, e->d_maxabs + (track ? 0 : 2));
        QCA_CUDA(cudaGetLastError());
        e->st.kernel_launches += 1;
    }
    QCA_CUDA(cudaStreamSynchronize(e->io_stream));   // the host buffer is the caller's again; later compute sees the planes
    return QCA_OK;
}

static int32_t download_vector(Engine* e, int v, double* host, int extra_quarter, bool both_planes) {
    QCA_CHECK(ensure_staging(e));
    QCA_CHECK(io_after_compute(e));
    for (unsigned long long off = 0; off < e->namps; off += e->staging_amps) {
        const unsigned long long cnt = std::min(e->staging_amps, e->namps - off);
        pack_rotate_kernel<<<stream_blocks(e, cnt), 256, 0, e->io_stream>>>(
            e->staging, e->plane[v][0], both_planes ? e->plane[v][1] : nullptr, off, cnt, e->shard, extra_quarter);
        QCA_CUDA(cudaGetLastError());
        e->st.kernel_launches += 1;
        QCA_CUDA(cudaMemcpyAsync(host + 2 * off, e->staging, cnt * sizeof(double2), cudaMemcpyDeviceToHost, e->io_stream));
    }
    QCA_CUDA(cudaStreamSynchronize(e->io_stream));
    return QCA_OK;
}

static int32_t measure_partial(Engine* e, double* sums) {
    QCA_REQUIRE(e->nplanes > 0 && e->resolved, QCA_ERR_STATE, "measure before a state was set (and resolved)");
    const int n = e->rule.ncells;
    const double* re = e->plane[e->cur][0];
    const double* im = e->nplanes == 2 ? e->plane[e->cur][1] : nullptr;
    QCA_CHECK(launch_barrier(e));  // sharded: partner states are final before they are read
    QCA_CUDA(cudaMemsetAsync(e->d_sums, 0, 4 * n * sizeof(double), e->stream));
    auto blocks_for = [&](unsigned long long npairs) {
        return (int)std::max<unsigned long long>(1, std::min<unsigned long long>((npairs + kMeasureThreads - 1) / kMeasureThreads, (unsigned long long)e->measure_blocks));
    };
    // fused path (csrc/qca_measure.cu): one read of the state per tile pass instead of one per cell
    const bool fused = (e->flags & QCA_FLAG_FUSED_MEASURE) && e->nplanes == 1 && e->local_bits >= kTile;
    const bool small = e->world == 1 && e->local_bits <= kSmallMaxBits && !(e->flags & (QCA_FLAG_PERCELL_MEASURE | QCA_FLAG_TILE_PATH_ONLY)) && !fused;
    if (small) {   // all cells in one launch (csrc/qca_small.cu)
        QCA_CHECK(launch_small_measure(re, im, e->local_bits, e->d_sums, e->stream));
        e->st.kernel_launches += 1;
    }
    if (fused) {
        for (const qca_pass_t& ps : e->passes) {
            QCA_CHECK(measure_tiles(re, e->namps, ps, e->shard, n, e->d_partials, 2 * e->num_sms, e->d_sums, e->stream));
            e->st.kernel_launches += 2;
        }
    }
    for (int bit = 0; bit < e->local_bits && !fused && !small; ++bit) {
        const int cell = n - 1 - global_pos(bit, e->shard);
        const unsigned long long npairs = e->namps >> 1;
        const int blocks = blocks_for(npairs);
        measure_pairs_kernel<<<blocks, kMeasureThreads, 0, e->stream>>>(re, im, re, im, bit, npairs, e->d_partials);
        QCA_CUDA(cudaGetLastError());
        reduce_partials_kernel<<<1, 128, 0, e->stream>>>(e->d_partials, blocks, e->d_sums + 4 * cell);
        QCA_CUDA(cudaGetLastError());
        e->st.kernel_launches += 2;
    }
    // sharded qubits: the rank holding the qubit dead pairs its slice with the partner's
    for (int j = 0; j < e->rank_bits; ++j) {
        if ((e->rank >> j) & 1) continue;
        const int cell = n - 1 - e->shard.pos[j];
        const int partner = e->rank ^ (1 << j);
        const double* pre = e->peer_plane[partner][e->cur][0];
        const double* pim = e->nplanes == 2 ? e->peer_plane[partner][e->cur][1] : nullptr;
        const int blocks = blocks_for(e->namps);
        measure_pairs_kernel<<<blocks, kMeasureThreads, 0, e->stream>>>(re, im, pre, pim, -1, e->namps, e->d_partials);
        QCA_CUDA(cudaGetLastError());
        reduce_partials_kernel<<<1, 128, 0, e->stream>>>(e->d_partials, blocks, e->d_sums + 4 * cell);
        QCA_CUDA(cudaGetLastError());
        e->st.kernel_launches += 2;
    }
    QCA_CUDA(cudaMemcpyAsync(sums, e->d_sums, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    QCA_CUDA(cudaStreamSynchronize(e->stream));
    return QCA_OK;
}


// ---------------------------------------------------------------------------
// Tight spectral bound.  ||H|| <= sum over blocks of consecutive cells of ||H_block||, where
// H_block keeps only the terms whose flipped cell lies in the block (it acts on the block plus
// `distance` context cells on each inner side).  Each block norm is the top Ritz value of a
// Lanczos run on the block's antisymmetric K (alpha_j = 0: q_{j+1} b_{j+1} = K q_j + b_j q_{j-1}),
// executed on the device with the engine's own kernels on a register of at most kBoundSites
// qubits; the tridiagonal eigenvalue is found on the host by Sturm bisection.  For the
// N=30, distance 2, [2,4) workload this gives R = 22.7 instead of the Gershgorin 30.
// ---------------------------------------------------------------------------
struct Engine;
static Engine* engine_of(qca_exact_t h);

constexpr int kBoundSites = 20;
constexpr int kLanczosMax = 160;

static double tridiag_top_eigenvalue(const std::vector<double>& b) {
    // symmetric tridiagonal, zero diagonal, off-diagonals b[0..m-2]  (m = b.size() + 1)
    const int m = (int)b.size() + 1;
    if (m == 1) return 0.0;
    double hi = 0.0;
    for (int i = 0; i < m; ++i) hi = std::max(hi, (i > 0 ? fabs(b[i - 1]) : 0.0) + (i < m - 1 ? fabs(b[i]) : 0.0));
    double lo = 0.0;
    auto count_below = [&](double x) {  // number of eigenvalues < x
        int cnt = 0;
        double q = -x;
        if (q < 0) ++cnt;
        for (int i = 1; i < m; ++i) {
            if (q == 0.0) q = 1e-300;
            q = -x - b[i - 1] * b[i - 1] / q;
            if (q < 0) ++cnt;
        }
        return cnt;
    };
    for (int it = 0; it < 100; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (count_below(mid) >= m) hi = mid; else lo = mid;
    }
    return hi;
}

// Norm of the block operator on a register of `nsites` qubits with flipped qubits `centers`.
// *converged: the top Ritz value settled (relative change <= 1e-10 over 8 further Lanczos vectors, or the
// Krylov space closed).  A Ritz value approaches the norm from BELOW, so an unconverged one must not be
// used as a bound: the caller then keeps the provable Gershgorin bound.
static int32_t block_norm(const qca_rule_t& rule, int nsites, unsigned long long centers, int device,
                          double* norm_out, bool* converged, uint64_t* launches) {
    *converged = false;
    qca_rule_t sub = rule;
    sub.ncells = nsites;
    qca_exact_t h = nullptr;
    QCA_CHECK(qca_exact_create(&h, &sub, device, 1, 0, QCA_FLAG_LOOSE_BOUND, nullptr));
    Engine* e = engine_of(h);
    int32_t rc = QCA_OK;
    do {
        e->fast_path = false;  // the generic kernel honours flip_mask
        e->use_v3 = false;
        for (auto& ps : e->passes) ps.flip_mask &= centers;
        for (int v = 0; v < 3; ++v) if ((rc = ensure_plane(e, v, 0))) break;
        if (rc) break;
        e->nplanes = 1; e->resolved = true; e->cur = 0;
        std::vector<double> start((size_t)e->namps);
        unsigned long long lcg = 0x9E3779B97F4A7C15ull;
        double nrm = 0.0;
        for (auto& x : start) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            x = (double)(long long)(lcg >> 11) / 9007199254740992.0 - 0.5;
            nrm += x * x;
        }
        nrm = 1.0 / sqrt(nrm);
        for (auto& x : start) x *= nrm;
        if (cudaMemcpyAsync(e->plane[1][0], start.data(), e->plane_bytes(), cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
            cudaStreamSynchronize(e->stream) != cudaSuccess) { set_error("bound: upload failed"); rc = QCA_ERR_CUDA; break; }
        int prev = 0, cur = 1;  // q_{j-1} (overwritten by w), q_j
        std::vector<double> beta;
        double bj = 0.0, theta = 0.0, theta_old = -1.0;
        for (int j = 0; j < kLanczosMax; ++j) {
            // w = b_j q_{j-1} + K q_j, written over q_{j-1}
            if ((rc = apply_operator(e, prev, cur, -1, 0.0, j == 0 ? -1 : prev, bj, 1.0))) break;
            e->cur = prev;
            double n2 = 0.0;
            if ((rc = qca_exact_norm2(h, &n2))) break;
            const double bn = sqrt(n2);
            if (!(bn > 1e-12)) { *converged = true; break; }  // invariant subspace: the Ritz values are exact
            beta.push_back(bn);
            if ((rc = launch_scale(e, prev, prev, 1.0 / bn))) break;
            std::swap(prev, cur);      // cur = q_{j+1}, prev = q_j
            bj = bn;
            if ((j % 8) == 7 || j == kLanczosMax - 1) {
                theta = tridiag_top_eigenvalue(beta);
                if (theta_old > 0 && fabs(theta - theta_old) <= 1e-10 * theta) { *converged = true; break; }
                theta_old = theta;
            }
        }
        if (rc) break;
        theta = tridiag_top_eigenvalue(beta);
        *norm_out = theta;
        if (launches) *launches += e->st.kernel_launches;
    } while (0);
    qca_exact_destroy(h);
    return rc;
}

// *bound is left untouched (the caller's provable bound stands) unless every block norm converged.
static int32_t tight_spectral_bound(const qca_rule_t& rule, int device, double* bound) {
    const int n = rule.ncells, d = rule.distance;
    // fewest blocks such that every block register fits kBoundSites qubits
    int nb = 1;
    for (;; ++nb) {
        const int bmax = (n + nb - 1) / nb;
        const int sites = (nb == 1) ? n : (nb == 2 ? bmax + d : bmax + 2 * d);
        if (sites <= kBoundSites || bmax <= 1) break;
    }
    std::vector<int> sizes(nb, n / nb);
    // larger blocks at the two ends (they need context on one side only)
    for (int i = 0, extra = n % nb; extra > 0; --extra, ++i) sizes[(i % 2) ? nb - 1 - i / 2 : i / 2] += 1;
    double total = 0.0;
    std::vector<std::pair<long long, double>> cache;  // (kind * 64 + size) -> norm
    for (int i = 0; i < nb; ++i) {
        const bool low_end = (i == 0), high_end = (i == nb - 1);
        const int B = sizes[i];
        const int kind = (low_end && high_end) ? 0 : ((low_end || high_end) ? 1 : 2);  // mirror: both ends alike
        const long long key = kind * 64 + B;
        double val = -1.0;
        for (auto& c : cache) if (c.first == key) val = c.second;
        if (val < 0.0) {
            const int ctx_low = (kind == 2) ? d : 0;          // an end block is computed as the low end
            const int ctx_high = (kind == 0) ? 0 : d;
            const int sites = ctx_low + B + ctx_high;
            const unsigned long long centers = ((1ull << B) - 1ull) << ctx_low;
            bool converged = false;
            QCA_CHECK(block_norm(rule, sites, centers, device, &val, &converged, nullptr));
            if (!converged) return QCA_OK;   // keep the Gershgorin bound rather than trust a lower estimate
            cache.emplace_back(key, val);
        }
        total += val;
    }
    *bound = total * (1.0 + 1e-3);  // Lanczos converges from below; 0.1 % covers the residual
    return QCA_OK;
}

}  // namespace qca

using qca::Engine;

struct qca_exact {
    Engine e;
};

namespace qca {
static Engine* engine_of(qca_exact_t h) { return &h->e; }
}  // namespace qca

extern "C" {

int32_t qca_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int32_t qca_exact_create(qca_exact_t* out, const qca_rule_t* rule, int32_t device, int32_t world_size,
                         int32_t rank, uint32_t flags, void* stream) {
    QCA_REQUIRE(out != nullptr, QCA_ERR_ARG, "out handle is NULL");
    *out = nullptr;
    QCA_CHECK(qca::validate_rule(rule));
    QCA_REQUIRE(world_size == 1 || world_size == 2 || world_size == 4 || world_size == 8, QCA_ERR_ARG,
                "world_size must be 1, 2, 4 or 8 (got %d)", world_size);
    QCA_REQUIRE(rank >= 0 && rank < world_size, QCA_ERR_ARG, "rank %d outside world of %d", rank, world_size);
    int rank_bits = 0;
    while ((1 << rank_bits) < world_size) ++rank_bits;
    QCA_REQUIRE(rule->ncells - rank_bits >= 1, QCA_ERR_ARG, "ncells %d too small for %d ranks", rule->ncells, world_size);
    QCA_REQUIRE(world_size == 1 || rule->ncells - rank_bits >= rule->distance, QCA_ERR_UNSUPPORTED,
                "sharding needs at least `distance` local qubits");
    int ndev = 0;
    QCA_CUDA(cudaGetDeviceCount(&ndev));
    QCA_REQUIRE(device >= 0 && device < ndev, QCA_ERR_CUDA, "device %d not present (%d visible)", device, ndev);
    QCA_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    QCA_CUDA(cudaGetDeviceProperties(&prop, device));
    QCA_REQUIRE(prop.major >= 10, QCA_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);

    const bool dbg = getenv("QCA_DEBUG") != nullptr;
    auto peek = [&](const char* where) {
        if (dbg) fprintf(stderr, "qca_b200 debug: %s: %s\n", where, cudaGetErrorString(cudaPeekAtLastError()));
    };
    peek("create entry");
    qca_exact* h = new qca_exact();
    Engine* e = &h->e;
    e->rule = *rule; e->device = device; e->world = world_size; e->rank = rank; e->rank_bits = rank_bits;
    e->local_bits = rule->ncells - rank_bits;
    e->namps = 1ull << e->local_bits;
    qca::plan_shard(*rule, world_size, &e->shard, rank);
    e->flags = flags;
    // fused measurement (one read of the state per tile pass): the default; sharded engines add the per-cell pairing of
    // their slice with the partner's for the log2(P) sharded cells
    if (!getenv("QCA_PERCELL_MEASURE") && !(flags & QCA_FLAG_PERCELL_MEASURE)) e->flags |= QCA_FLAG_FUSED_MEASURE;
    else e->flags &= ~QCA_FLAG_FUSED_MEASURE;
    e->num_sms = prop.multiProcessorCount;
    e->bound = qca::spectral_bound(*rule);
    qca::plan_passes(e->local_bits, e->passes);
    if (int32_t rc = qca::plan_remote(*rule, world_size, rank, e->remote)) { delete h; return rc; }
    if (int32_t rc = qca::plan_rotation(*rule, world_size, rank, &e->rotation)) { delete h; return rc; }
    if (const char* env = getenv("QCA_REMOTE_RING")) e->remote_rows = (atoi(env) == 4) ? 4 : 6;
    if (flags & QCA_FLAG_NO_PERSISTENT) e->persistent_ctas = 0;
    else if (const char* env = getenv("QCA_PERSISTENT_CTAS")) e->persistent_ctas = atoi(env);
    for (const qca_remote_op_t& op : e->remote)  // fraction of the plane the term reads on this rank
        e->remote_fraction[op.qubit] = op.window_bits <= 4
            ? (double)__builtin_popcount(op.mask & 0xffffu) / 16.0 : 1.0;   // the mask is replicated over 16 entries
    e->st.spectral_bound = e->bound;
    e->st.passes_per_apply = (int32_t)e->passes.size();
    e->st.local_bits = e->local_bits;
    if (stream) { e->stream = (cudaStream_t)stream; e->own_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
            qca::set_error("cudaStreamCreate failed"); delete h; return QCA_ERR_CUDA;
        }
        e->own_stream = true;
    }
    {
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (cudaStreamCreateWithPriority(&e->io_stream, cudaStreamNonBlocking, greatest) != cudaSuccess ||
            cudaEventCreateWithFlags(&e->io_event, cudaEventDisableTiming) != cudaSuccess) {
            qca::set_error("creating the IO stream failed: %s", cudaGetErrorString(cudaGetLastError()));
            qca_exact_destroy(h); return QCA_ERR_CUDA;
        }
    }
    e->measure_blocks = e->num_sms * 8;
    bool ok = cudaMalloc(&e->d_maxabs, 4 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc(&e->d_partials, std::max<size_t>(4ull * e->measure_blocks, (size_t)qca::kMeasureTileVals * 2 * e->num_sms) * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&e->d_sums, 4ull * rule->ncells * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&e->d_amp, 2ull * rule->ncells * sizeof(double)) == cudaSuccess;
    if (!ok) { qca::set_error("cudaMalloc of scratch failed: %s", cudaGetErrorString(cudaGetLastError())); qca_exact_destroy(h); return QCA_ERR_NOMEM; }
    peek("before tables");
    if (int32_t rc = qca::build_tables(e)) { qca_exact_destroy(h); return rc; }
    peek("after tables");
    if (world_size == 1 && e->local_bits >= qca::kTile3 && rule->distance <= 4 && !(flags & QCA_FLAG_V2_KERNELS) && !getenv("QCA_V2_KERNELS")) {
        // Cluster bits: OFF by default.  Measured on a B200 at N = 30 (profiles/r02_v3_cluster_sweep.txt): every cluster
        // bit adds ~1.2 ms to a 6.5 ms pass -- distributed shared memory moves ~17 B/clk per SM, about what the SM's
        // share of HBM moves, and the partner reads do not overlap the HBM stream -- so two passes with 3 + 3 cluster
        // bits (21 ms per Chebyshev term) lose against three passes without (17 ms).  QCA_V3_CLUSTER_BITS=1..3 enables
        // them (tests; future parts with faster SM-to-SM paths).  The per-row window table has 4 + CB + 3 d index bits.
        int max_cb = 0;
        if (const char* env = getenv("QCA_V3_CLUSTER_BITS"))
            max_cb = std::max(0, std::min({atoi(env), qca::kMaxClusterBits, 16 - qca::kRegHigh - 3 * rule->distance}));
        int min_low = 4;
        if (const char* env = getenv("QCA_V3_MIN_LOW")) min_low = atoi(env);
        qca::plan_passes_v3(e->local_bits, std::max(0, max_cb), min_low, e->passes3);
        e->use_v3 = !e->passes3.empty();
        if (e->use_v3) {
            if (int32_t rc = qca::build_tables_v3(e)) { qca_exact_destroy(h); return rc; }
            e->st.passes_per_apply = (int32_t)e->passes3.size();
        }
    }
    peek("before bound");
    if (!(flags & QCA_FLAG_LOOSE_BOUND)) {
        double tight = e->bound;
        if (int32_t rc = qca::tight_spectral_bound(*rule, device, &tight)) { qca_exact_destroy(h); return rc; }
        if (cudaSetDevice(device) != cudaSuccess) {
            qca::set_error("cudaSetDevice(%d) failed after the spectral-bound computation", device);
            qca_exact_destroy(h); return QCA_ERR_CUDA;
        }
        if (tight < e->bound) e->bound = tight;
        e->st.spectral_bound = e->bound;
    }
    peek("after bound");
    if (world_size > 1) {
        // every plane exists up front so that it can be exported once
        for (int v = 0; v < 3; ++v)
            for (int p = 0; p < 2; ++p) {
                if (int32_t rc = qca::ensure_plane(e, v, p)) { qca_exact_destroy(h); return rc; }
                e->peer_plane[rank][v][p] = e->plane[v][p];
            }
        if (cudaMalloc(&e->d_flags, 8 * sizeof(unsigned long long)) != cudaSuccess ||
            cudaMemset(e->d_flags, 0, 8 * sizeof(unsigned long long)) != cudaSuccess) {
            qca::set_error("cudaMalloc of barrier flags failed"); qca_exact_destroy(h); return QCA_ERR_NOMEM;
        }
        e->peer_flags[rank] = e->d_flags;
    }
    peek("create exit");
    *out = h;
    return QCA_OK;
}

int32_t qca_exact_destroy(qca_exact_t h) {
    if (!h) return QCA_OK;
    Engine* e = &h->e;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (auto& pr : e->prof) { cudaEventDestroy(pr.ev0); cudaEventDestroy(pr.ev1); }
    for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
    if (e->peers_ready && !e->loopback) {   // (loop-back "peers" are this engine's own planes, not IPC mappings)
        for (int r = 0; r < e->world; ++r) {
            if (r == e->rank) continue;
            for (int v = 0; v < 3; ++v) for (int p = 0; p < 2; ++p) if (e->peer_plane[r][v][p]) cudaIpcCloseMemHandle(e->peer_plane[r][v][p]);
            if (e->peer_flags[r]) cudaIpcCloseMemHandle(e->peer_flags[r]);
        }
    }
    for (int v = 0; v < 3; ++v) for (int p = 0; p < 2; ++p) if (e->plane[v][p]) cudaFree(e->plane[v][p]);
    if (e->spare_plane) cudaFree(e->spare_plane);
    for (auto* p : e->d_tab_lo) if (p) cudaFree(p);
    for (auto* p : e->d_tab_hi) if (p) cudaFree(p);
    for (auto& tb : e->tabs3) { if (tb.thr) cudaFree(tb.thr); if (tb.row) cudaFree(tb.row); }
    if (e->d_flags) cudaFree(e->d_flags);
    if (e->staging) cudaFree(e->staging);
    if (e->d_maxabs) cudaFree(e->d_maxabs);
    if (e->d_partials) cudaFree(e->d_partials);
    if (e->d_sums) cudaFree(e->d_sums);
    if (e->d_amp) cudaFree(e->d_amp);
    if (e->io_stream) { cudaStreamSynchronize(e->io_stream); cudaStreamDestroy(e->io_stream); }
    if (e->io_event) cudaEventDestroy(e->io_event);
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    cudaGetLastError();   // nothing a teardown call reported may surface in a later engine's launch check
    delete h;
    return QCA_OK;
}

uint64_t qca_exact_local_amps(qca_exact_t h) { return h ? h->e.namps : 0; }

int32_t qca_exact_set_state(qca_exact_t h, const double* psi, uint64_t namps) {
    QCA_REQUIRE(h && psi, QCA_ERR_ARG, "NULL argument");
    Engine* e = &h->e;
    QCA_REQUIRE(namps == e->namps, QCA_ERR_ARG, "state has %llu amplitudes, engine holds %llu",
                (unsigned long long)namps, e->namps);
    QCA_CUDA(cudaSetDevice(e->device));
    QCA_CHECK(qca::prepare_upload(e));
    QCA_CHECK(qca::upload_vector(e, e->cur, psi, true));
    return qca::finish_upload(e);
}

int32_t qca_exact_set_product_state(qca_exact_t h, const double* p_alive, int32_t ncells) {
    QCA_REQUIRE(h && p_alive, QCA_ERR_ARG, "NULL argument");
    Engine* e = &h->e;
    QCA_REQUIRE(ncells == e->rule.ncells, QCA_ERR_ARG, "plist has %d entries, rule has %d cells", ncells, e->rule.ncells);
    std::vector<double> amp(2 * ncells);
    for (int c = 0; c < ncells; ++c) {
        QCA_REQUIRE(p_alive[c] >= 0.0 && p_alive[c] <= 1.0, QCA_ERR_ARG, "p_alive[%d] = %g outside [0,1]", c, p_alive[c]);
        amp[2 * c] = sqrt(1.0 - p_alive[c]);
        amp[2 * c + 1] = sqrt(p_alive[c]);
    }
    QCA_CUDA(cudaSetDevice(e->device));
    QCA_CHECK(qca::prepare_upload(e));
    QCA_CUDA(cudaMemcpyAsync(e->d_amp, amp.data(), amp.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    qca::product_state_kernel<<<qca::stream_blocks(e, e->namps), 256, 0, e->stream>>>(
        e->plane[e->cur][0], e->plane[e->cur][1], e->d_amp, ncells, e->local_bits, e->shard, e->d_maxabs);
    QCA_CUDA(cudaGetLastError());
    e->st.kernel_launches += 1;
    QCA_CUDA(cudaStreamSynchronize(e->stream));  // amp is a host temporary
    return qca::finish_upload(e);
}

int32_t qca_exact_plane_flags(qca_exact_t h, int32_t* has_re, int32_t* has_im) {
    QCA_REQUIRE(h && has_re && has_im, QCA_ERR_ARG, "NULL argument");
    QCA_REQUIRE(h->e.nplanes > 0, QCA_ERR_STATE, "no state uploaded");
    QCA_CUDA(cudaSetDevice(h->e.device));
    bool re, im;
    QCA_CHECK(qca::read_plane_flags(&h->e, &re, &im));
    *has_re = re; *has_im = im;
    return QCA_OK;
}

int32_t qca_exact_resolve_planes(qca_exact_t h, int32_t has_re, int32_t has_im) {
    QCA_REQUIRE(h, QCA_ERR_ARG, "NULL handle");
    QCA_REQUIRE(h->e.nplanes > 0, QCA_ERR_STATE, "no state uploaded");
    QCA_CUDA(cudaSetDevice(h->e.device));
    return qca::settle_planes(&h->e, has_re != 0, has_im != 0);
}

int32_t qca_exact_get_state(qca_exact_t h, double* psi, uint64_t namps) {
    QCA_REQUIRE(h && psi, QCA_ERR_ARG, "NULL argument");
    Engine* e = &h->e;
    QCA_REQUIRE(e->nplanes > 0 && e->resolved, QCA_ERR_STATE, "get_state before a state was set (and resolved)");
    QCA_REQUIRE(namps == e->namps, QCA_ERR_ARG, "buffer has %llu amplitudes, engine holds %llu",
                (unsigned long long)namps, e->namps);
    QCA_CUDA(cudaSetDevice(e->device));
    return qca::download_vector(e, e->cur, psi, e->g_quarter, e->nplanes == 2);
}

int32_t qca_exact_step(qca_exact_t h, double step_size, int32_t nsteps) {
    QCA_REQUIRE(h, QCA_ERR_ARG, "NULL handle");
    Engine* e = &h->e;
    QCA_REQUIRE(e->nplanes > 0 && e->resolved, QCA_ERR_STATE, "step before a state was set (and resolved)");
    QCA_REQUIRE(nsteps >= 0, QCA_ERR_ARG, "nsteps must be >= 0");
    QCA_REQUIRE(isfinite(step_size), QCA_ERR_ARG, "step_size must be finite");
    // every rank derives the number of Chebyshev terms -- hence of cross-rank barriers -- from R: ranks that
    // computed R independently could disagree in the last digits and dead-lock
    QCA_REQUIRE(e->world == 1 || e->loopback || e->bound_agreed, QCA_ERR_STATE,
                "sharded engine: call qca_exact_set_spectral_bound with the maximum over all ranks before stepping");
    QCA_CUDA(cudaSetDevice(e->device));
    for (int s = 0; s < nsteps; ++s) QCA_CHECK(qca::step_graphed(e, step_size));
    return QCA_OK;
}

int32_t qca_exact_measure_partial(qca_exact_t h, double* sums) {
    QCA_REQUIRE(h && sums, QCA_ERR_ARG, "NULL argument");
    QCA_CUDA(cudaSetDevice(h->e.device));
    return qca::measure_partial(&h->e, sums);
}

int32_t qca_measure_finish(const double* sums, int32_t ncells, double* population, double* d_population,
                           double* entropy, double* bond_dims) {
    QCA_REQUIRE(sums && ncells >= 1, QCA_ERR_ARG, "bad arguments");
    for (int c = 0; c < ncells; ++c) {
        const double s0 = sums[4 * c], s1 = sums[4 * c + 1], wr = sums[4 * c + 2], wi = sums[4 * c + 3];
        if (population) population[c] = s1;
        if (d_population) d_population[c] = nearbyint(s1);  // np.round: half to even
        if (entropy) {
            const double mean = 0.5 * (s0 + s1), half = 0.5 * (s0 - s1);
            const double rad = sqrt(half * half + wr * wr + wi * wi);
            const double lam[2] = {mean + rad, mean - rad};
            double ent = 0.0;
            for (double l : lam) if (l > 0.0) ent -= l * log2(l);
            entropy[c] = ent;
        }
    }
    if (bond_dims)
        for (int i = 0; i <= ncells; ++i) bond_dims[i] = ldexp(1.0, std::min(i, ncells - i));
    return QCA_OK;
}

int32_t qca_exact_measure(qca_exact_t h, double* population, double* d_population, double* entropy,
                          double* bond_dims) {
    QCA_REQUIRE(h, QCA_ERR_ARG, "NULL handle");
    QCA_REQUIRE(h->e.world == 1, QCA_ERR_STATE, "sharded engine: use measure_partial + all-reduce + qca_measure_finish");
    std::vector<double> sums(4 * h->e.rule.ncells);
    QCA_CHECK(qca_exact_measure_partial(h, sums.data()));
    return qca_measure_finish(sums.data(), h->e.rule.ncells, population, d_population, entropy, bond_dims);
}

int32_t qca_exact_apply_h(qca_exact_t h, const double* in, double* out, uint64_t namps) {
    QCA_REQUIRE(h && in && out, QCA_ERR_ARG, "NULL argument");
    Engine* e = &h->e;
    QCA_REQUIRE(namps == e->namps, QCA_ERR_ARG, "vector has %llu amplitudes, engine holds %llu",
                (unsigned long long)namps, e->namps);
    QCA_CUDA(cudaSetDevice(e->device));
    const int X = (e->cur + 1) % 3, Y = (e->cur + 2) % 3;
    const int saved_planes = e->nplanes;
    e->nplanes = 2;
    int32_t rc = QCA_OK;
    do {
        if ((rc = qca::ensure_plane(e, X, 0)) || (rc = qca::ensure_plane(e, X, 1)) ||
            (rc = qca::ensure_plane(e, Y, 0)) || (rc = qca::ensure_plane(e, Y, 1))) break;
        if ((rc = qca::upload_vector(e, X, in, false))) break;
        // H psi = D (i K) D^-1 psi: apply K in the rotated frame, one extra quarter turn on the way out
        if ((rc = qca::apply_operator(e, Y, X, -1, 0.0, -1, 0.0, 1.0))) break;
        if ((rc = qca::launch_barrier(e))) break;  // sharded: partners are done reading X
        rc = qca::download_vector(e, Y, out, 1, true);
    } while (0);
    e->nplanes = saved_planes;
    if (saved_planes < 2) { qca::release_plane(e, X, 1); qca::release_plane(e, Y, 1); }
    return rc;
}

int32_t qca_exact_norm2(qca_exact_t h, double* norm2) {
    QCA_REQUIRE(h && norm2, QCA_ERR_ARG, "NULL argument");
    Engine* e = &h->e;
    QCA_REQUIRE(e->nplanes > 0, QCA_ERR_STATE, "norm before a state was set");
    QCA_CUDA(cudaSetDevice(e->device));
    const double* re = e->plane[e->cur][0];
    const double* im = e->nplanes == 2 ? e->plane[e->cur][1] : nullptr;
    double hsum[4];
    const unsigned long long npairs = e->namps >> 1;
    const int blocks = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((npairs + qca::kMeasureThreads - 1) / qca::kMeasureThreads, (unsigned long long)e->measure_blocks));
    qca::measure_pairs_kernel<<<blocks, qca::kMeasureThreads, 0, e->stream>>>(re, im, re, im, 0, npairs, e->d_partials);
    QCA_CUDA(cudaGetLastError());
    qca::reduce_partials_kernel<<<1, 128, 0, e->stream>>>(e->d_partials, blocks, e->d_sums);
    QCA_CUDA(cudaGetLastError());
    e->st.kernel_launches += 2;
    QCA_CUDA(cudaMemcpyAsync(hsum, e->d_sums, sizeof(hsum), cudaMemcpyDeviceToHost, e->stream));
    QCA_CUDA(cudaStreamSynchronize(e->stream));
    *norm2 = hsum[0] + hsum[1];
    return QCA_OK;
}

int32_t qca_exact_set_spectral_bound(qca_exact_t h, double bound) {
    QCA_REQUIRE(h, QCA_ERR_ARG, "NULL handle");
    QCA_REQUIRE(bound > 0.0 && isfinite(bound), QCA_ERR_ARG, "spectral bound must be positive");
    h->e.bound = bound;
    h->e.bound_agreed = true;
    h->e.st.spectral_bound = bound;
    return QCA_OK;
}

int32_t qca_exact_get_stats(qca_exact_t h, qca_exact_stats_t* out) {
    QCA_REQUIRE(h && out, QCA_ERR_ARG, "NULL argument");
    QCA_CUDA(cudaSetDevice(h->e.device));
    QCA_CHECK(qca::finish_profile(&h->e));
    h->e.st.planes = h->e.nplanes;
    *out = h->e.st;
    return QCA_OK;
}

int32_t qca_exact_reset_stats(qca_exact_t h) {
    QCA_REQUIRE(h, QCA_ERR_ARG, "NULL handle");
    QCA_CUDA(cudaSetDevice(h->e.device));
    QCA_CHECK(qca::finish_profile(&h->e));
    Engine* e = &h->e;
    e->st.kernel_launches = 0; e->st.pass_launches = 0; e->st.pass_bytes = 0.0; e->st.remote_bytes = 0.0;
    e->st.profiled_pass_ms = 0.0; e->st.profiled_pass_launches = 0;
    for (double& v : e->st.profiled_ms_by_pass) v = 0.0;
    return QCA_OK;
}

// ---- sharding: CUDA IPC peer mapping ----------------------------------------------------------
int32_t qca_exact_ipc_count(qca_exact_t h) { return (h && h->e.world > 1) ? qca::kIpcBuffers : 0; }

static void* ipc_buffer(Engine* e, int index) {
    return index < 6 ? (void*)e->plane[index / 2][index % 2] : (void*)e->d_flags;
}

int32_t qca_exact_ipc_export(qca_exact_t h, int32_t index, uint8_t handle[QCA_IPC_HANDLE_BYTES]) {
    QCA_REQUIRE(h && handle, QCA_ERR_ARG, "NULL argument");
    Engine* e = &h->e;
    QCA_REQUIRE(e->world > 1, QCA_ERR_STATE, "engine is not sharded");
    QCA_REQUIRE(index >= 0 && index < qca::kIpcBuffers, QCA_ERR_ARG, "buffer index %d out of range", index);
    static_assert(sizeof(cudaIpcMemHandle_t) == QCA_IPC_HANDLE_BYTES, "handle size");
    QCA_CUDA(cudaSetDevice(e->device));
    cudaIpcMemHandle_t mh;
    QCA_CUDA(cudaIpcGetMemHandle(&mh, ipc_buffer(e, index)));
    memcpy(handle, &mh, sizeof(mh));
    return QCA_OK;
}

int32_t qca_exact_loopback_peers(qca_exact_t h) {
    QCA_REQUIRE(h, QCA_ERR_ARG, "NULL handle");
    Engine* e = &h->e;
    QCA_REQUIRE(e->world > 1 && !e->peers_ready, QCA_ERR_STATE, "loopback needs a sharded engine without imported peers");
    for (int r = 0; r < e->world; ++r) {
        for (int v = 0; v < 3; ++v)
            for (int p = 0; p < 2; ++p) e->peer_plane[r][v][p] = e->plane[v][p];
        e->peer_flags[r] = e->d_flags;
    }
    e->loopback = true;
    e->peers_ready = true;
    return QCA_OK;
}

int32_t qca_exact_ipc_import(qca_exact_t h, const uint8_t* handles, int32_t world_size, int32_t count) {
    QCA_REQUIRE(h && handles, QCA_ERR_ARG, "NULL argument");
    Engine* e = &h->e;
    QCA_REQUIRE(e->world > 1 && world_size == e->world && count == qca::kIpcBuffers, QCA_ERR_ARG,
                "handle table does not match the engine (world %d, count %d)", world_size, count);
    QCA_REQUIRE(!e->peers_ready, QCA_ERR_STATE, "peers already imported");
    QCA_CUDA(cudaSetDevice(e->device));
    for (int r = 0; r < e->world; ++r) {
        if (r == e->rank) continue;
        for (int i = 0; i < count; ++i) {
            cudaIpcMemHandle_t mh;
            memcpy(&mh, handles + ((size_t)r * count + i) * QCA_IPC_HANDLE_BYTES, sizeof(mh));
            void* ptr = nullptr;
            QCA_CUDA(cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
            if (i < 6) e->peer_plane[r][i / 2][i % 2] = (double*)ptr;
            else e->peer_flags[r] = (unsigned long long*)ptr;
        }
    }
    e->peers_ready = true;
    return QCA_OK;
}

}  // extern "C"
