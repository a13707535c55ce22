// Host-only planning: spectral bound, Chebyshev/Bessel coefficients, tile-pass plan.
// None of these touch the GPU, so the CPU test-suite exercises them through the C ABI.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <stdlib.h>
#include <vector>

#include "qca_common.cuh"
#include "qca_plan.h"

namespace qca {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int32_t validate_rule(const qca_rule_t* r, int max_cells) {
    QCA_REQUIRE(r != nullptr, QCA_ERR_ARG, "rule is NULL");
    QCA_REQUIRE(r->ncells >= 1, QCA_ERR_ARG, "ncells must be >= 1 (got %d)", r->ncells);
    QCA_REQUIRE(r->distance >= 1, QCA_ERR_ARG, "distance must be >= 1 (got %d)", r->distance);
    QCA_REQUIRE(r->distance <= 7, QCA_ERR_UNSUPPORTED, "distance > 7 not supported (got %d)", r->distance);
    QCA_REQUIRE(r->ncells <= max_cells, QCA_ERR_UNSUPPORTED, "ncells > %d not supported here (got %d)", max_cells, r->ncells);
    QCA_REQUIRE(r->act_lo >= 0 && r->act_hi >= r->act_lo, QCA_ERR_ARG,
                "activation interval [%d,%d) is not a range", r->act_lo, r->act_hi);
    return QCA_OK;
}

// Largest number of simultaneously active cells: sliding window of 2d+1 cells,
// dead cells beyond both ends (mpo.py:181-200).
double spectral_bound(const qca_rule_t& r) {
    const int d = r.distance, n = r.ncells;
    const int wbits = 2 * d + 1;
    const uint32_t wmask = (1u << wbits) - 1u;
    const uint32_t imask = interval_mask_of(r.act_lo, r.act_hi);
    const int NEG = -1000000;
    std::vector<int> cur(1u << wbits, NEG), nxt(1u << wbits, NEG);
    cur[0] = 0;
    for (int i = 0; i < n + d; ++i) {  // append cell i (virtual dead cell when i >= n)
        std::fill(nxt.begin(), nxt.end(), NEG);
        for (uint32_t w = 0; w <= wmask; ++w) {
            if (cur[w] == NEG) continue;
            for (int c = 0; c <= (i < n ? 1 : 0); ++c) {
                uint32_t w2 = ((w << 1) | (uint32_t)c) & wmask;
                int gain = 0;
                int centre = i - d;  // its whole neighbourhood is now inside the window
                if (centre >= 0 && centre < n) {
                    uint32_t nb = w2 & ~(1u << d);
                    int cnt = __builtin_popcount(nb);
                    gain = (imask >> cnt) & 1u;
                }
                nxt[w2] = std::max(nxt[w2], cur[w] + gain);
            }
        }
        cur.swap(nxt);
    }
    int best = 0;
    for (int v : cur) best = std::max(best, v);
    return (double)best;
}

// J_k(z), k = 0..kmax, by Miller's backward recurrence normalised with
// 1 = J_0 + 2 sum_{k>=1} J_2k.
static void bessel_j(double z, int kmax, std::vector<long double>& out) {
    out.assign(kmax + 1, 0.0L);
    if (z == 0.0) { out[0] = 1.0L; return; }
    int start = (int)(std::max((double)kmax, z) + 14.0 * cbrt(std::max(z, 1.0)) + 60.0);
    start += start & 1;
    long double jp = 0.0L, jc = 1e-300L, norm = 0.0L;
    const long double zz = z;
    for (int k = start; k >= 1; --k) {
        long double jm = (2.0L * k / zz) * jc - jp;  // J_{k-1}
        jp = jc; jc = jm;
        if (k - 1 <= kmax) out[k - 1] = jc;
        if (((k - 1) & 1) == 0) norm += (k - 1 == 0) ? jc : 2.0L * jc;
        if (fabsl(jc) > 1e250L) {  // rescale
            const long double s = 1e-250L;
            jc *= s; jp *= s; norm *= s;
            for (auto& v : out) v *= s;
        }
    }
    for (auto& v : out) v /= norm;
}

int32_t chebyshev_plan(double z, double tol, std::vector<double>& a) {
    QCA_REQUIRE(z >= 0.0 && isfinite(z), QCA_ERR_ARG, "chebyshev z must be finite and >= 0");
    QCA_REQUIRE(tol > 0.0, QCA_ERR_ARG, "chebyshev tolerance must be > 0");
    int kmax = (int)(z + 12.0 * cbrt(std::max(z, 1.0)) + 40.0);
    std::vector<long double> j;
    bessel_j(z, kmax, j);
    // keep a[0..n-1] with tail sum_{k>=n} 2|J_k| < tol
    long double tail = 0.0L;
    int n = kmax + 1;
    while (n > 2) {
        long double t = tail + 2.0L * fabsl(j[n - 1]);
        if (t >= tol) break;
        tail = t; --n;
    }
    a.resize(n);
    for (int k = 0; k < n; ++k) a[k] = (double)((k == 0 ? 1.0L : 2.0L) * j[k]);
    return QCA_OK;
}

void plan_passes(int local_bits, std::vector<qca_pass_t>& out) {
    out.clear();
    const int n = local_bits;
    int covered = std::min(n, kTileBits);
    qca_pass_t p0{};
    p0.low_bits = covered; p0.high_start = covered; p0.high_bits = 0;
    p0.flip_mask = (covered >= 64) ? ~0ull : ((1ull << covered) - 1ull);
    out.push_back(p0);
    // remaining qubits: spread evenly over the fewest passes
    int rem = n - covered;
    if (rem <= 0) return;
    const int per = kTileBits - kMinLowBits;
    int npass = (rem + per - 1) / per;
    for (int i = 0; i < npass; ++i) {
        int m = rem / (npass - i) + ((rem % (npass - i)) ? 1 : 0);
        qca_pass_t p{};
        p.low_bits = kTileBits - m;
        p.high_start = covered;
        p.high_bits = m;
        p.flip_mask = ((1ull << m) - 1ull) << covered;
        out.push_back(p);
        covered += m; rem -= m;
    }
}

// Cluster plan (pass_kernel_v3, qca_pass3.cuh): CTA tiles of 14 bits, up to max_cluster_bits more through
// distributed shared memory.  Pass 0 = the contiguous low bits (up to 17); every later pass takes up to
// (14 - min_low) CTA-local strided bits plus up to max_cluster_bits cluster bits.  `reserved` of a pass holds its
// number of cluster bits, high_bits the CTA-local strided bits only (the cluster bits follow directly above).
// N = 30: 17 + 13 -> two passes (56 instead of 80 bytes per amplitude per Chebyshev term).
void plan_passes_v3(int local_bits, int max_cluster_bits, int min_low, std::vector<qca_pass_t>& out) {
    out.clear();
    const int n = local_bits, T = 14;
    if (n < T) return;
    min_low = std::max(4, std::min(min_low, T - 1));
    max_cluster_bits = std::max(0, std::min(max_cluster_bits, 3));
    int covered = std::min(n, T + max_cluster_bits);
    qca_pass_t p0{};
    p0.low_bits = T; p0.high_start = T; p0.high_bits = 0; p0.reserved = covered - T;
    p0.flip_mask = (1ull << covered) - 1ull;
    out.push_back(p0);
    int rem = n - covered;
    if (rem <= 0) return;
    const int local_max = T - min_low;
    const int per = local_max + max_cluster_bits;
    const int npass = (rem + per - 1) / per;
    for (int i = 0; i < npass; ++i) {
        const int m = rem / (npass - i) + ((rem % (npass - i)) ? 1 : 0);
        const int cb = std::max(0, m - local_max);
        qca_pass_t p{};
        p.high_bits = m - cb;
        p.low_bits = T - p.high_bits;
        p.high_start = covered;
        p.reserved = cb;
        p.flip_mask = ((1ull << m) - 1ull) << covered;
        out.push_back(p);
        covered += m; rem -= m;
    }
}

// Which qubits are sharded.  The partner-rank traffic of a sharded qubit's term is proportional to
// how often its rule predicate holds on a rank, and for ADJACENT sharded cells that depends on the
// rank itself (the rank whose top three cells are all alive pulls 2.75 planes per application, the
// all-dead rank 0.25, and every barrier waits for the heaviest).  Cells spaced distance+1 apart do
// not see each other, so every rank pulls the same amount (1.5 planes for distance 2, [2,4)); cell 0
// (the chain end, cheapest predicate) is always one of them.  The scattered layout needs every
// sharded position above the contiguous first tile (13 qubits) plus `distance` context bits and the
// 16-entry remote mask (2*distance <= 4 window bits); otherwise the top qubits are sharded.
void plan_shard(const qca_rule_t& r, int world, ShardMap* map, int rank) {
    int rank_bits = 0;
    while ((1 << rank_bits) < world) ++rank_bits;
    const int N = r.ncells, n = N - rank_bits, d = r.distance;
    map->nins = rank_bits;
    const int lowest = N - 1 - (rank_bits - 1) * (d + 1);
    const bool scattered = rank_bits >= 2 && d <= 2 && lowest >= kTileBits + d + 1;
    for (int j = 0; j < rank_bits; ++j)
        map->pos[j] = scattered ? N - 1 - (rank_bits - 1 - j) * (d + 1) : n + j;
    map->rank_or = 0;
    for (int j = 0; j < rank_bits; ++j)
        if ((rank >> j) & 1) map->rank_or |= 1ull << map->pos[j];
}

int32_t plan_remote(const qca_rule_t& r, int world, int rank, std::vector<qca_remote_op_t>& out) {
    out.clear();
    int rank_bits = 0;
    while ((1 << rank_bits) < world) ++rank_bits;
    const int n = r.ncells - rank_bits;
    QCA_REQUIRE(n >= 1, QCA_ERR_ARG, "ncells %d too small for %d ranks", r.ncells, world);
    std::vector<qca_pass_t> passes;
    plan_passes(n, passes);
    ShardMap map{};
    plan_shard(r, world, &map, rank);
    const uint32_t imask = interval_mask_of(r.act_lo, r.act_hi);
    for (int j = 0; j < rank_bits; ++j) {
        qca_remote_op_t op{};
        op.qubit = map.pos[j];
        op.partner = rank ^ (1 << j);
        op.sign = ((rank >> j) & 1) ? -1 : 1;
        // the predicate of this qubit reads the local bits around the hole it leaves in the local index
        const int hole = map.pos[j] - j;                    // local position of the first bit above it
        const int below = std::min(r.distance, hole), above = std::min(r.distance, n - hole);
        const int wbits = below + above;
        op.shift = hole - below;
        op.window_bits = wbits;
        bool any = false;
        for (unsigned v = 0; v < (1u << wbits); ++v) {
            const unsigned long long xg = expand_index((unsigned long long)v << op.shift, map);
            const bool on = (activity_word<unsigned long long>(xg, r.distance, imask) >> op.qubit) & 1ull;
            any |= on;
            if (on && v < 16) op.mask |= (1u << v);
        }
        if (wbits > 4) op.mask = 0;   // too wide for the fast kernel's 16-entry mask: generic kernel only
        else  // the kernel indexes with (x >> shift) & 15: bits above the window must not matter
            for (unsigned v = (1u << wbits); v < 16; ++v)
                if ((op.mask >> (v & ((1u << wbits) - 1u))) & 1u) op.mask |= (1u << v);
        if (!any) continue;  // this rank never sees the term (e.g. not enough alive cells above)
        out.push_back(op);
    }
    // Placement: NVLink reads should overlap the HBM traffic of every pass, so the terms are
    // spread over the passes, heaviest first onto the least loaded pass.  Pass 0 already streams two
    // recurrence operands and takes at most one term; later passes at most three.
    const int npass = (int)passes.size();
    std::vector<double> load(npass, 0.0);
    std::vector<int> count(npass, 0);
    std::vector<int> order(out.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    auto weight = [&](int i) { return (double)__builtin_popcount(out[i].mask) + (out[i].mask ? 0.0 : 1.0); };
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return weight(x) > weight(y); });
    for (int i : order) {
        int best = -1;
        for (int p = 0; p < npass; ++p) {
            const int cap = (npass == 1) ? 8 : (p == 0 ? 1 : 3);
            if (count[p] >= cap) continue;
            // ties go to the later pass (its ring is deeper: one recurrence operand instead of two)
            if (best < 0 || load[p] < load[best] - 1e-9 || (fabs(load[p] - load[best]) <= 1e-9 && p > best)) best = p;
        }
        if (best < 0) best = npass - 1;
        out[i].pass = best;
        load[best] += weight(i);
        count[best] += 1;
    }
    return QCA_OK;
}

// Rotation of the remote terms over the passes (fast kernel).  Term j goes to pass (j + r) mod npasses,
// slot j / npasses, for the amplitudes of rotation r; r is read from four local index bits just above
// the first tile (they are constant over every contiguous piece the kernel copies, and no term's
// predicate depends on them for the registers that matter, so each pass gets ~1/npasses of every term).
int32_t plan_rotation(const qca_rule_t& r, int world, int rank, qca_remote_rotation_t* out) {
    *out = qca_remote_rotation_t{};
    for (auto& p : out->op_of) for (auto& s : p) for (int& v : s) v = -1;
    std::vector<qca_remote_op_t> ops;
    QCA_CHECK(plan_remote(r, world, rank, ops));
    int rank_bits = 0;
    while ((1 << rank_bits) < world) ++rank_bits;
    std::vector<qca_pass_t> passes;
    plan_passes(r.ncells - rank_bits, passes);
    const int np = (int)std::min<size_t>(passes.size(), 4);
    out->npasses = np;
    out->rot_shift = kTileBits;
    for (unsigned v = 0; v < 16; ++v) out->rot_word |= (v % (unsigned)np) << (2 * v);
    out->nslots = ((int)ops.size() + np - 1) / np;
    if (out->nslots > 2) {   // single-pass registers (< 13 local qubits) with three terms: generic kernel only
        out->nslots = 0;
        return QCA_OK;
    }
    // QCA_FORCE_SLOTS2 (test aid): run the two-slot kernels even when one slot would do
    const bool force2 = getenv("QCA_FORCE_SLOTS2") != nullptr && !ops.empty() && out->nslots == 1;
    if (force2) out->nslots = 2;
    for (int j = 0; j < (int)ops.size(); ++j)
        for (int rot = 0; rot < np; ++rot) out->op_of[(j + rot) % np][force2 ? 1 : j / np][rot] = j;
    return QCA_OK;
}

}  // namespace qca

extern "C" {

const char* qca_version(void) { return "qca_b200 0.1 (sm_100a)"; }
const char* qca_last_error(void) { return qca::g_error; }

int32_t qca_spectral_bound(const qca_rule_t* rule, double* bound) {
    QCA_CHECK(qca::validate_rule(rule, 1 << 20));  // chains of any TDVP length: the DP is O(ncells)
    QCA_REQUIRE(bound != nullptr, QCA_ERR_ARG, "bound is NULL");
    *bound = qca::spectral_bound(*rule);
    return QCA_OK;
}

int32_t qca_chebyshev_plan(double z, double tol, double* a, int32_t capacity, int32_t* nterms) {
    QCA_REQUIRE(nterms != nullptr, QCA_ERR_ARG, "nterms is NULL");
    std::vector<double> v;
    QCA_CHECK(qca::chebyshev_plan(z, tol, v));
    *nterms = (int32_t)v.size();
    if (a != nullptr) {
        QCA_REQUIRE(capacity >= (int32_t)v.size(), QCA_ERR_ARG, "coefficient buffer too small (%d < %zu)",
                    capacity, v.size());
        memcpy(a, v.data(), v.size() * sizeof(double));
    }
    return QCA_OK;
}

int32_t qca_plan_passes(int32_t local_bits, qca_pass_t* passes, int32_t capacity, int32_t* npasses) {
    QCA_REQUIRE(local_bits >= 1 && local_bits <= 40, QCA_ERR_ARG, "local_bits out of range (%d)", local_bits);
    QCA_REQUIRE(npasses != nullptr, QCA_ERR_ARG, "npasses is NULL");
    std::vector<qca_pass_t> v;
    qca::plan_passes(local_bits, v);
    *npasses = (int32_t)v.size();
    if (passes != nullptr) {
        QCA_REQUIRE(capacity >= (int32_t)v.size(), QCA_ERR_ARG, "pass buffer too small");
        memcpy(passes, v.data(), v.size() * sizeof(qca_pass_t));
    }
    return QCA_OK;
}

int32_t qca_plan_passes_v3(int32_t local_bits, int32_t max_cluster_bits, int32_t min_low, qca_pass_t* passes, int32_t capacity,
                           int32_t* npasses) {
    QCA_REQUIRE(local_bits >= 1 && local_bits <= 40, QCA_ERR_ARG, "local_bits out of range (%d)", local_bits);
    QCA_REQUIRE(npasses != nullptr, QCA_ERR_ARG, "npasses is NULL");
    std::vector<qca_pass_t> v;
    qca::plan_passes_v3(local_bits, max_cluster_bits, min_low, v);
    *npasses = (int32_t)v.size();
    if (passes != nullptr) {
        QCA_REQUIRE(capacity >= (int32_t)v.size(), QCA_ERR_ARG, "pass buffer too small");
        memcpy(passes, v.data(), v.size() * sizeof(qca_pass_t));
    }
    return QCA_OK;
}

int32_t qca_plan_shard(const qca_rule_t* rule, int32_t world_size, int32_t* positions) {
    QCA_CHECK(qca::validate_rule(rule));
    QCA_REQUIRE(world_size == 1 || world_size == 2 || world_size == 4 || world_size == 8, QCA_ERR_ARG,
                "world_size must be 1, 2, 4 or 8 (got %d)", world_size);
    QCA_REQUIRE(positions != nullptr, QCA_ERR_ARG, "positions is NULL");
    qca::ShardMap map{};
    qca::plan_shard(*rule, world_size, &map, 0);
    for (int j = 0; j < map.nins; ++j) positions[j] = map.pos[j];
    return QCA_OK;
}

int32_t qca_plan_rotation(const qca_rule_t* rule, int32_t world_size, int32_t rank, qca_remote_rotation_t* out) {
    QCA_REQUIRE(rule && out, QCA_ERR_ARG, "NULL argument");
    QCA_CHECK(qca::validate_rule(rule));
    QCA_REQUIRE(world_size == 1 || world_size == 2 || world_size == 4 || world_size == 8, QCA_ERR_ARG, "bad world size");
    QCA_REQUIRE(rank >= 0 && rank < world_size, QCA_ERR_ARG, "bad rank");
    return qca::plan_rotation(*rule, world_size, rank, out);
}

int32_t qca_plan_remote(const qca_rule_t* rule, int32_t world_size, int32_t rank, qca_remote_op_t* ops,
                        int32_t capacity, int32_t* nops) {
    QCA_CHECK(qca::validate_rule(rule));
    QCA_REQUIRE(world_size == 1 || world_size == 2 || world_size == 4 || world_size == 8, QCA_ERR_ARG,
                "world_size must be 1, 2, 4 or 8 (got %d)", world_size);
    QCA_REQUIRE(rank >= 0 && rank < world_size && nops != nullptr, QCA_ERR_ARG, "bad rank / NULL nops");
    std::vector<qca_remote_op_t> v;
    QCA_CHECK(qca::plan_remote(*rule, world_size, rank, v));
    *nops = (int32_t)v.size();
    if (ops != nullptr) {
        QCA_REQUIRE(capacity >= (int32_t)v.size(), QCA_ERR_ARG, "remote-op buffer too small");
        memcpy(ops, v.data(), v.size() * sizeof(qca_remote_op_t));
    }
    return QCA_OK;
}

}  // extern "C"
