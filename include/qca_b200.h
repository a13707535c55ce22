/*
 * qca_b200.h -- C ABI of the B200-native time-evolution path of the quantum
 * cellular automaton ("quantum game of life") simulator.
 *
 * Drop-in boundary: every entry point below replaces one piece of the
 * reference's (BenjaminDecker/quantum-cellular-automaton) Python hot path; the
 * reference file:line is cited beside each.  Plain pointers and sizes only, no
 * torch/numpy types.  All functions return QCA_OK (0) or a QCA_ERR_* code; the
 * message of the last failure on the calling thread is in qca_last_error().
 *
 * Conventions (identical to the reference): cell 0 is the MOST significant bit
 * of the basis-state index (tensor_networks/mps.py:194-208), one time step
 * applies exp(-i*pi/2*step_size*H) (lautils/lautils.py:45-55), host state
 * vectors are interleaved complex128 (numpy complex128, re,im,re,im...).
 *
 * There is no CPU fallback: every compute entry point needs an sm_100 device
 * and fails with QCA_ERR_CUDA otherwise.
 */
#ifndef QCA_B200_H
#define QCA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCA_OK 0
#define QCA_ERR_ARG 1         /* bad argument (reference: assert / ValueError) */
#define QCA_ERR_CUDA 2        /* CUDA runtime failure or no device */
#define QCA_ERR_NOMEM 3       /* state does not fit in device memory */
#define QCA_ERR_STATE 4       /* call order (e.g. step before set_state) */
#define QCA_ERR_UNSUPPORTED 5 /* rule outside the kernel's range (distance > 7, ncells > 40) */

/* parameters/rules.py:1-6 (periodic is always False in the reference, parser.py:182-187) */
typedef struct qca_rule {
    int32_t ncells;   /* --num-cells */
    int32_t distance; /* --distance */
    int32_t act_lo;   /* --activation-interval LOWER (inclusive) */
    int32_t act_hi;   /* --activation-interval UPPER (exclusive) */
} qca_rule_t;

const char* qca_version(void);
const char* qca_last_error(void);
/* number of CUDA devices visible, or 0 */
int32_t qca_device_count(void);

/* ------------------------------------------------------------------------
 * Host-only planning helpers (no GPU needed; used by the CPU test-suite).
 * ---------------------------------------------------------------------- */

/* Upper bound of the spectral radius of H = sum_i sx_i P_i: the largest number
 * of simultaneously active cells over all configurations (Gershgorin row sum of
 * MPO.as_matrix(), mpo.py:221-230), by a transfer-matrix DP over the chain. */
int32_t qca_spectral_bound(const qca_rule_t* rule, double* bound);

/* Chebyshev/Bessel plan for exp(-i z x), |x| <= 1:  a[k] = (2 - delta_k0) J_k(z),
 * k = 0..*nterms-1, truncated where the tail sum drops below tol.  a may be
 * NULL to query *nterms.  Replaces the eigendecomposition in calculate_U
 * (lautils.py:52-55). */
int32_t qca_chebyshev_plan(double z, double tol, double* a, int32_t capacity, int32_t* nterms);

/* One tile pass of the matrix-free operator (see DESIGN.md "pass plan"):
 * a CTA stages the amplitudes whose index bits [0,low_bits) and
 * [high_start, high_start+high_bits) vary and applies the rule terms whose
 * flipped qubit is in flip_mask. */
typedef struct qca_pass {
    int32_t low_bits;
    int32_t high_start;
    int32_t high_bits;
    int32_t reserved;
    uint64_t flip_mask;
} qca_pass_t;
/* Plan for a local register of local_bits qubits; returns the number of passes
 * written to passes[0..capacity). */
int32_t qca_plan_passes(int32_t local_bits, qca_pass_t* passes, int32_t capacity, int32_t* npasses);
/* Plan of the cluster kernels (one GPU, >= 14 qubits): CTA tiles of 14 index bits, joined by up to
 * max_cluster_bits (<= 3) more through distributed shared memory.  `high_bits` counts the CTA-local strided bits,
 * `reserved` the cluster bits directly above them; later passes keep at least min_low (>= 4) contiguous low bits.
 * Returns 0 passes for registers below 14 qubits. */
int32_t qca_plan_passes_v3(int32_t local_bits, int32_t max_cluster_bits, int32_t min_low, qca_pass_t* passes,
                           int32_t capacity, int32_t* npasses);

/* Sharded register: global index-bit positions of the log2(world_size) sharded qubits, ascending
 * (rank bit j <-> positions[j]).  Non-adjacent cells {0, d+1, 2(d+1)} when the register is large
 * enough (every rank then pulls the same amount over NVLink), else the top qubits.  The local index
 * of a rank is the global index with those bits removed. */
int32_t qca_plan_shard(const qca_rule_t* rule, int32_t world_size, int32_t* positions);

/* A rule term that flips a qubit held by another rank (sharded register): this rank adds
 * sign * [mask bit v set] * partner_vector[x], same local index x on the partner. */
typedef struct qca_remote_op {
    int32_t pass;    /* tile pass whose epilogue carries the term */
    int32_t partner; /* rank whose vector is read over NVLink */
    int32_t qubit;   /* index bit of the flipped qubit (>= local_bits) */
    int32_t sign;    /* +1 if this rank holds the qubit dead, -1 if alive */
    uint32_t mask;   /* activity of the term per value v = (local index >> shift) & 15 (the local bits
                        within `distance` of the sharded qubit) */
    int32_t shift;
    int32_t window_bits; /* number of local bits the predicate reads (mask is valid when <= 4) */
    int32_t reserved;
} qca_remote_op_t;
int32_t qca_plan_remote(const qca_rule_t* rule, int32_t world_size, int32_t rank, qca_remote_op_t* ops,
                        int32_t capacity, int32_t* nops);

/* How the fast tile-pass kernel spreads those terms over the passes of one operator application.
 * Every launch has `nslots` remote operand slots; slot s of pass p carries, for the amplitudes x whose
 * rotation is r, term op_of[p][s][r] of the qca_plan_remote list (-1: none), with
 *   r(x) = (rot_word >> 2 * ((x_local >> rot_shift) & 15)) & 3  <  npasses.
 * nslots == 0: the terms do not fit two slots per pass (three terms on a single-pass register): such
 * registers run on the generic kernel.
 * Each term is applied in exactly one pass for every x, and every pass pulls (nearly) the same share of
 * every term over NVLink, so the NVLink time overlaps the HBM time of every launch.  (The `pass` field of
 * qca_remote_op_t is the static placement the generic kernel uses instead.) */
typedef struct qca_remote_rotation {
    int32_t nslots;
    int32_t npasses;
    int32_t rot_shift;
    uint32_t rot_word;
    int32_t op_of[4][2][4];
} qca_remote_rotation_t;
int32_t qca_plan_rotation(const qca_rule_t* rule, int32_t world_size, int32_t rank, qca_remote_rotation_t* out);

/* ------------------------------------------------------------------------
 * Exact evolution engine == algorithms/exact.py:9-27 (class Exact).
 * ---------------------------------------------------------------------- */
typedef struct qca_exact* qca_exact_t;

#define QCA_FLAG_FORCE_COMPLEX 1u /* keep both real planes even if the rotated state is real */
#define QCA_FLAG_PROFILE 2u       /* record a CUDA-event pair around every kernel launch */
#define QCA_FLAG_LOOSE_BOUND 4u   /* scale H by the Gershgorin bound only (skip the block-Lanczos bound) */
#define QCA_FLAG_FUSED_MEASURE 8u   /* measure with one read of the state per tile pass (csrc/qca_measure.cu; single-plane
                                       states on >= 13 local qubits) instead of one read per cell.  The default (also when
                                       sharded) */
#define QCA_FLAG_TILE_PATH_ONLY 32u /* registers <= 13 qubits: use the tile-pass kernels instead of the one-kernel step (tests) */
#define QCA_FLAG_NO_GRAPH 64u      /* registers of 14..24 qubits: launch a step kernel by kernel instead of replaying its CUDA graph */
#define QCA_FLAG_V2_KERNELS 128u    /* one GPU, >= 14 qubits: use the 13-bit tile-pass kernels (pass_kernel_v2) instead of the
                                      cluster kernels (pass_kernel_v3); also: environment QCA_V2_KERNELS */
#define QCA_FLAG_NO_PERSISTENT 256u /* sharded engines: one CTA per tile (pass_kernel_v2) instead of persistent CTAs whose operand
                                      rings run across tile boundaries (pass_kernel_v2p); environment QCA_PERSISTENT_CTAS=0 likewise,
                                      QCA_PERSISTENT_CTAS=n > 0 limits the persistent grid to n CTAs (tests) */
#define QCA_FLAG_PERCELL_MEASURE 16u /* always use the per-cell measurement kernels (also: environment QCA_PERCELL_MEASURE) */

/* Exact.__init__ (exact.py:15-17).  Instead of MPO.as_matrix() + calculate_U
 * (dense 2^N x 2^N) the engine keeps only the rule.  world_size/rank shard the
 * state over the top log2(world_size) qubits (world_size in {1,2,4,8}); with
 * world_size > 1 the caller must exchange peer handles (below) before stepping.
 * stream: a cudaStream_t to launch on, or NULL for an engine-owned stream. */
int32_t qca_exact_create(qca_exact_t* out, const qca_rule_t* rule, int32_t device,
                         int32_t world_size, int32_t rank, uint32_t flags, void* stream);
int32_t qca_exact_destroy(qca_exact_t h);

/* number of amplitudes held by this rank: 2^(ncells - log2(world_size)) */
uint64_t qca_exact_local_amps(qca_exact_t h);

/* Exact.psi setter (exact.py:22-24): psi is this rank's contiguous slice of
 * MPS.as_vector() (mps.py:194-208), namps interleaved complex128. */
int32_t qca_exact_set_state(qca_exact_t h, const double* psi, uint64_t namps);
/* MPS.from_density_distribution (mps.py:35-52) + as_vector, evaluated on the
 * device: amplitude of cell i is (sqrt(1-p_i), sqrt(p_i)). */
int32_t qca_exact_set_product_state(qca_exact_t h, const double* p_alive, int32_t ncells);
/* Sharded engines only: whether the rotated state of this rank has non-zero real / imaginary
 * parts, and the collective decision (logical OR over ranks) of how many planes to keep.  A
 * sharded engine must be resolved after every set_state / set_product_state. */
int32_t qca_exact_plane_flags(qca_exact_t h, int32_t* has_re, int32_t* has_im);
int32_t qca_exact_resolve_planes(qca_exact_t h, int32_t has_re, int32_t has_im);
/* Exact.psi getter, vector part (exact.py:19-20). */
int32_t qca_exact_get_state(qca_exact_t h, double* psi, uint64_t namps);

/* Exact.do_time_step (exact.py:26-27) x nsteps with U = calculate_U(H, step_size)
 * (lautils.py:45-55): psi <- exp(-i*pi/2*step_size*H)^nsteps psi, device resident. */
int32_t qca_exact_step(qca_exact_t h, double step_size, int32_t nsteps);

/* MPS.measure (mps.py:100-140) on MPS.from_vector(psi) (mps.py:55-73): per-cell
 * population <P1>, np.round of it, single-site von-Neumann entropy in bits, and
 * the bond dimensions min(2^i, 2^(N-i)) of the exact MPS.  Any pointer may be
 * NULL.  With world_size > 1 the raw partial sums must be all-reduced by the
 * caller: use qca_exact_measure_partial + qca_measure_finish instead. */
int32_t qca_exact_measure(qca_exact_t h, double* population, double* d_population,
                          double* entropy, double* bond_dims);
/* sums[4*cell + {0,1,2,3}] = sum|psi_0|^2, sum|psi_1|^2, Re w, Im w with
 * w = sum_{rest} psi[cell=0,rest] conj(psi[cell=1,rest]) over this rank's slice. */
int32_t qca_exact_measure_partial(qca_exact_t h, double* sums /* 4*ncells */);
int32_t qca_measure_finish(const double* sums, int32_t ncells, double* population,
                           double* d_population, double* entropy, double* bond_dims);

/* Test hook: out = MPO.as_matrix() @ in (mpo.py:221-230) without changing the
 * engine's state; in/out are this rank's slices, interleaved complex128. */
int32_t qca_exact_apply_h(qca_exact_t h, const double* in, double* out, uint64_t namps);

/* Squared norm of the resident state (this rank's slice). */
int32_t qca_exact_norm2(qca_exact_t h, double* norm2);

/* The engine scales H by a bound R of its spectral radius: min(Gershgorin, sum of block norms found
 * by Lanczos on the device at creation; a block whose Lanczos run did not converge leaves the provable
 * Gershgorin bound in force).  The number of Chebyshev terms of a step -- and with it the number of
 * cross-rank barriers -- follows from R, so sharded engines MUST agree on it: qca_exact_step fails with
 * QCA_ERR_STATE on a sharded engine until this call has set the common value (the maximum of
 * qca_exact_get_stats().spectral_bound over all ranks). */
int32_t qca_exact_set_spectral_bound(qca_exact_t h, double bound);

typedef struct qca_exact_stats {
    double spectral_bound;       /* R used to scale H */
    int32_t planes;              /* 1: rotated state is real, 2: general complex */
    int32_t passes_per_apply;    /* tile passes per operator application */
    int32_t last_terms;          /* Chebyshev terms of the last step */
    int32_t local_bits;
    uint64_t kernel_launches;    /* launches of this library's kernels so far */
    uint64_t pass_launches;      /* ... of which tile-pass kernels */
    double pass_bytes;           /* algorithmic bytes moved by those pass launches */
    double profiled_pass_ms;     /* sum of event-timed pass durations (QCA_FLAG_PROFILE) */
    uint64_t profiled_pass_launches;
    double device_bytes;         /* bytes of device memory held */
    double remote_bytes;         /* bytes read from partner ranks over NVLink by pass launches */
    double profiled_ms_by_pass[4]; /* profiled_pass_ms split by tile pass (0: the contiguous first tile) */
} qca_exact_stats_t;
int32_t qca_exact_get_stats(qca_exact_t h, qca_exact_stats_t* out);
int32_t qca_exact_reset_stats(qca_exact_t h);

/* ------------------------------------------------------------------------
 * TDVP support: Householder QR with LAPACK's conventions (numpy.linalg.qr as used by
 * MPS.left_qr_tensors, tensor_networks/mps.py:84-88): beta = -sign(Re alpha)*norm (zlarfg), Q = H_1...H_k.
 * All pointers are DEVICE pointers to column-major complex128.  a (m x n) is overwritten by the
 * factorisation; q receives m x kq (kq = min(m,n) "reduced" or m "complete"), r receives kq x n.
 * The reference's results depend on this sign convention (see csrc/qca_linalg.cu).
 * The factorisation is a COOPERATIVE launch over up to one CTA per SM (one grid barrier per column): the device must
 * support cooperative launches, the call must not sit inside a CUDA-graph capture, and it starts once its whole grid
 * can be resident.  Environment (tests / A-B): QCA_QR_SINGLE_CTA = the one-CTA kernels of round 1,
 * QCA_QR_FORM_GLOBAL = form Q with the column in global memory also for m <= 512.
 * ---------------------------------------------------------------------- */
int32_t qca_qr_householder(void* a, int32_t m, int32_t n, void* tau, void* q, int32_t kq, void* r, void* stream);

/* Batched, segmented complex128 GEMM on the FP64 tensor cores (DMMA) for the effective-Hamiltonian
 * contractions (np.tensordot in tdvp.py:299-347):  C_g[m,n] = sum_{s<S} sum_{k<K} A_{g,s}[m,k] B_{g,s}[k,n].
 * DEVICE pointers, strides in complex128 elements: A at g*a_sg + s*a_ss + m*a_sm + k*a_sk (a_sm == 1 or
 * a_sk == 1), B at g*b_sg + s*b_ss + k*b_sk + n, C at g*c_sg + m*c_sm + n.  conj_a: use conj(A).
 * nsplit > 1 splits the (s, k) reduction over nsplit CTAs per tile; split i writes its partial sum to
 * c + i*c_ssplit and the caller adds the partials (fixed order: deterministic). */
int32_t qca_zgemm_batched(const void* a, const void* b, void* c, int32_t M, int32_t N, int32_t K, int32_t S, int32_t G,
                          int64_t a_sg, int64_t a_ss, int64_t a_sm, int64_t a_sk, int64_t b_sg, int64_t b_ss,
                          int64_t b_sk, int64_t c_sg, int64_t c_sm, int32_t conj_a, int32_t nsplit, int64_t c_ssplit,
                          void* stream);

/* Per-launch timing of qca_zgemm_batched (including the launches qca_heff_* make): returns the summed
 * CUDA-event duration (ms), the FP64 operations (8 M N K S G per launch) and the number of launches recorded
 * since the last call, then switches recording on (enable != 0) or off.  Synchronises; not thread safe. */
int32_t qca_zgemm_profile(int32_t enable, double* ms, double* flops, uint64_t* launches);

/* ------------------------------------------------------------------------
 * Matrix-free effective Hamiltonian of TDVP and its Krylov exponential (csrc/qca_heff.cu).  Replaces
 * TDVP._assemble_H_eff / _evolve_A (algorithms/tdvp.py:299-310, 350-365: a dense (g dl dr)^2 matrix per
 * site) and lautils.timestep (lautils/lautils.py:58-82: eigh of it).  All pointers are DEVICE pointers,
 * complex128 row-major in the reference's index order:
 *   psi[g][x][u]   g = 1 (bond matrix), 2 (one site), 4 (two sites: (a, c) -> 2a + c); x < dl, u < dr
 *   left[x][w][y]  (dl, wl, dl)        right[u][w][v]  (dr, wr, dr)
 *   mix: CSR matrix with g*wr rows and g*wl columns, Mx[(g', n), (g, w)]: the site operator(s)
 *        (one site: W[g, g', w, n]; two sites: sum_m W1[a,b,w,m] W2[c,d,m,n]; bond: identity)
 *   H_eff psi [g'][y][v] = sum  Mx[(g',n),(g,w)] left[x][w][y] right[u][n][v] psi[g][x][u]
 * Everything is enqueued on `stream`; nothing synchronises or allocates: the caller supplies
 * `workspace` of qca_heff_workspace_bytes(h, krylov_dim) bytes (krylov_dim = 0 for qca_heff_apply).
 * ---------------------------------------------------------------------- */
typedef struct qca_heff {
    const void* left;
    const void* right;
    const int32_t* mix_rowptr;   /* [g*wr + 1] */
    const int32_t* mix_col;      /* [nnz] */
    const void* mix_val;         /* [nnz] complex128 */
    int32_t dl, dr, wl, wr, g;
    /* structural zeros (HOST values; honoured when use_masks != 0 and wl, wr <= 32): col_mask[g] bit w set iff
     * column (g, w) of Mx has an entry, row_mask[g'] bit n set iff row (g', n) has one.  Contractions that
     * only feed unused columns or produce empty rows are skipped. */
    int32_t use_masks;
    uint32_t col_mask[4], row_mask[4];
} qca_heff_t;
int32_t qca_heff_workspace_bytes(const qca_heff_t* h, int32_t krylov_dim, uint64_t* bytes);
/* out = H_eff psi */
int32_t qca_heff_apply(const qca_heff_t* h, const void* psi, void* out, void* workspace, uint64_t workspace_bytes,
                       void* stream);
/* out = exp(-i t H_eff) psi by krylov_dim (<= 64) Lanczos steps with full re-orthogonalisation; the small
 * tridiagonal exponential is evaluated on the device too: a Chebyshev series when spectral_bound >= ||H_eff||
 * is given (e.g. qca_spectral_bound of the rule, valid for environments built from isometries), Jacobi
 * otherwise (spectral_bound = 0) or when the bound turns out too small.  out may alias psi. */
int32_t qca_heff_expm(const qca_heff_t* h, const void* psi, void* out, int32_t krylov_dim, double t, double spectral_bound,
                      void* workspace, uint64_t workspace_bytes, void* stream);

/* Environment update of TDVP (algorithms/tdvp.py:329-347, `_update_left_environment` / `_update_right_environment`:
 * three np.tensordot calls) on the same DMMA kernel:
 *   out[r][m][s] = sum  Mx[(b,m),(a,w)] left[x][w][y] site[a][x][r] conj(site[b][y][s])
 * h->left = the previous environment (dl, wl, dl), h->mix* = the ONE-site operator (g = 2), h->dr = the site
 * tensor's other bond dimension; h->right is not read (pass h->left).  out: (dr, wr, dr).  A right
 * environment is the same call on the mirrored tensors: site[a][u][l], W with its two bond indices swapped.
 * Workspace: qca_env_grow_workspace_bytes(h).  Stream-ordered, no synchronisation. */
int32_t qca_env_grow_workspace_bytes(const qca_heff_t* h, uint64_t* bytes);
int32_t qca_env_grow(const qca_heff_t* h, const void* site, void* out, void* workspace, uint64_t workspace_bytes,
                     void* stream);

/* ------------------------------------------------------------------------
 * Multi-GPU (one process per GPU).  The state is sharded over the top
 * log2(world_size) qubits; terms that flip a sharded qubit read the partner
 * rank's vector directly over NVLink (CUDA IPC peer mapping).  The host side
 * (torch.distributed) only moves these 64-byte handles once at start-up.
 * ---------------------------------------------------------------------- */
#define QCA_IPC_HANDLE_BYTES 64
/* number of device buffers this rank exports */
int32_t qca_exact_ipc_count(qca_exact_t h);
int32_t qca_exact_ipc_export(qca_exact_t h, int32_t index, uint8_t handle[QCA_IPC_HANDLE_BYTES]);
/* handles: [world_size][count][QCA_IPC_HANDLE_BYTES] gathered from all ranks */
int32_t qca_exact_ipc_import(qca_exact_t h, const uint8_t* handles, int32_t world_size, int32_t count);

/* Profiling aid (one GPU): a sharded engine whose "partner" vectors are its own planes and whose
 * cross-rank barrier is a no-op.  The launches, kernels and byte counts are those of rank `rank` of a
 * world_size job (remote operands come from local HBM instead of NVLink); the numbers it computes are
 * meaningless.  Used by scratch/loopback_prof.py to profile the sharded tile-pass kernel under ncu. */
int32_t qca_exact_loopback_peers(qca_exact_t h);

#ifdef __cplusplus
}
#endif
#endif /* QCA_B200_H */
