/*
 * crs_b200.h — C ABI of libcrs_b200.so: the B200 (sm_100a) implementation of contextual_rs's
 * per-iteration decision hot path.
 *
 * The reference (saitcakmak/contextual_rs) is pure Python and has no FFI layer; its boundary for
 * this path is the Python function signature plus a duck-typed BoTorch model.
 * Each entry point below names the reference interface it replaces; INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, enums; no C++/torch types cross the boundary;
 *   - every pointer is DEVICE memory owned by the caller unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every launch
 *     goes to that stream and no entry point synchronises unless documented;
 *   - return value: 0 (CRS_OK) or a crs_status code, never an exception;
 *   - tables are row-major `[n_c, ld]` with the ARM index fastest, i.e. exactly the layout of the
 *     reference's `posterior.mean` / `posterior.variance` (`n_c x K`).
 */
#ifndef CRS_B200_H
#define CRS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CRS_API __attribute__((visibility("default")))
#else
#define CRS_API
#endif

#define CRS_ABI_VERSION 1
#define CRS_MAX_DIM 16   /* context dimension d <= 16 */
#define CRS_MAX_OBS 160  /* observations per arm handled by the shared-memory Cholesky */

typedef enum crs_status {
  CRS_OK = 0,
  CRS_ERR_INVALID_ARG = 1,
  CRS_ERR_UNSUPPORTED = 2,   /* d > CRS_MAX_DIM, n_obs > CRS_MAX_OBS, ... */
  CRS_ERR_CUDA = 3,          /* a CUDA runtime call failed; see crs_last_cuda_error() */
  CRS_ERR_NOT_PSD = 4,       /* Cholesky pivot <= 0 for some arm (see state.status) */
  CRS_TIES_PENDING = 5       /* crs_gao_decide_host: >1 exact minimisers; caller must resolve */
} crs_status;

typedef enum crs_dtype { CRS_F64 = 0, CRS_F32 = 1 } crs_dtype;
typedef enum crs_kernel { CRS_MATERN52 = 0, CRS_RBF = 1 } crs_kernel;

/* flags for crs_gp_state.flags */
#define CRS_LEARN_NORMALIZE 1   /* learn Normalize bounds from each arm's train_x (else use norm_*) */
#define CRS_LEARN_STANDARDIZE 2 /* learn Standardize mean/std from each arm's train_y (else use y_*) */
#define CRS_TRAIN_INPUTS_RAW 4  /* `xn` (= model.train_inputs, read by OCBA step 2(b): exact-match counts / KDE,
                                   contextual_rs_strategies.py:168-193) holds the RAW inputs — what the reference's
                                   recorded runs show (tests/golden/reference_trace_ml_gao.json) — instead of the
                                   Normalize-transformed ones (BoTorch >= 0.4 eval mode) */

/* p-surrogate of OCBA step 2(b): contextual_rs/contextual_rs_strategies.py:158-193 */
typedef enum crs_p_mode {
  CRS_P_COUNT = 0,   /* exact-match training counts   (:187-193) */
  CRS_P_KDE = 1,     /* sum exp(-scale * ||x - c*||)  (:168-186; continuous_context.py:112-122) */
  CRS_P_INFER = 2    /* prior / posterior variance    (:158-167) */
} crs_p_mode;

/*
 * One ModelListGP of K independent SingleTaskGP arms
 * (contextual_rs/experiment_utils.py:29-40: Standardize(m=1) + Normalize(d) + ScaleKernel(ARD
 * Matern-5/2 | RBF) + ConstantMean + homoskedastic noise).  "Fitted description" arrays are inputs
 * (always fp64); "prepared" arrays are written by crs_gp_prepare in the compute dtype T.
 * `dim_pad` is d rounded up to {1,2,4,8,16}; padded coordinates are zero.
 */
typedef struct crs_gp_state {
  int32_t num_arms;   /* K */
  int32_t n_cap;      /* row capacity per arm, >= max n_obs, multiple of 8 */
  int32_t dim;        /* d */
  int32_t dim_pad;    /* padded d used by the prepared arrays */
  int32_t kernel;     /* crs_kernel */
  int32_t dtype;      /* crs_dtype of the prepared arrays and of all tables */
  int32_t flags;      /* CRS_LEARN_* | CRS_TRAIN_INPUTS_RAW */
  int32_t n_max;      /* caller's bound on max n_obs (sizes the kernels' tiles); 0 = use n_cap */
  /* fitted description */
  const double* train_x;      /* [K, n_cap, d]  raw inputs */
  const double* train_y;      /* [K, n_cap]     raw targets */
  const int32_t* n_obs;       /* [K] */
  const double* lengthscale;  /* [K, d] */
  const double* outputscale;  /* [K] */
  const double* noise;        /* [K] */
  const double* mean_const;   /* [K] */
  double* norm_mins;          /* [K, d]  in (frozen) or out (learnt) */
  double* norm_ranges;        /* [K, d] */
  double* y_mean;             /* [K] */
  double* y_std;              /* [K] */
  /* prepared */
  void* xn;        /* [K, n_cap, d]        T  model.train_inputs: raw (CRS_TRAIN_INPUTS_RAW) or normalised inputs */
  void* xs;        /* [K, n_cap, dim_pad]  T  normalised / lengthscale */
  void* alpha;     /* [K, n_cap]           T  o * K^-1 (y~ - m) */
  void* linv_t;    /* [K, crs_linv_stride(n_cap, dtype)] T  o * L^-T in 32-column panels */
  void* arm_head;  /* [K, 40] T  Normalize mins[16] | 1/(range*lengthscale)[16] | (outputscale,
                      mean_const, y_mean, y_std) | pad[4] */
  int32_t* status; /* [K]  0 = ok, j+1 = non-positive Cholesky pivot at column j */
} crs_gp_state;

/* Result of the allocation scan + step 2(b); lives in device memory, 64 bytes. */
typedef struct crs_decision {
  double min_zeta;     /* global min of zeta_{c,j}                 (strategies.py:145) */
  double psi_best;     /* y_s_1                                     (strategies.py:195) */
  double psi_rest;     /* y_s_2                                     (strategies.py:196) */
  int64_t tie_count;   /* # of (context, arm) entries equal to min_zeta (strategies.py:147-149) */
  int32_t context;     /* min_context                               (strategies.py:155) */
  int32_t zeta_arm;    /* arm id attaining min_zeta at that context (strategies.py:157,201) */
  int32_t best_arm;    /* predicted best arm at that context */
  int32_t next_arm;    /* the decision                              (strategies.py:199-201) */
  int32_t reserved[4]; /* [0] = (first arm whose Cholesky failed)+1 or 0; [1] = completion tag */
} crs_decision;

/*
 * Multi-GPU combine of a context-sharded decision without NCCL or the host (one process per GPU).
 * Every rank owns a MAILBOX of crs_mailbox_bytes(world) bytes of device memory that all ranks have
 * mapped (crs_ipc_export / crs_ipc_open).  `mailboxes` is a DEVICE array of the world base pointers as
 * seen from this rank.  The kernel that finishes a rank's local decision stores its 64-byte record
 * (context + ctx_begin) into every rank's mailbox, waits for the world records of exchange `seq` in its
 * own and folds them (smallest zeta, then lowest global context; tie counts added).  `seq` must be the
 * same non-zero number on every rank and change by one per exchange.
 */
typedef struct crs_peer_exchange {
  void* const* mailboxes;  /* device array [world] */
  int32_t rank, world;
  int32_t seq, reserved;
  int64_t ctx_begin;       /* global index of this rank's first context */
} crs_peer_exchange;

/* Result of crs_levi_scan; device memory, 24 bytes. */
typedef struct crs_levi_decision {
  double value;      /* max over contexts of (weight *) max over arms of EI  (levi.py:220-224) */
  int32_t context;   /* context_idx                                          (levi.py:224) */
  int32_t arm;       /* next_alternative = max_idcs[context_idx]             (levi.py:225) */
  int64_t reserved;
} crs_levi_decision;

CRS_API int crs_abi_version(void);
CRS_API const char* crs_status_string(int status);
CRS_API const char* crs_last_cuda_error(void);

/* Elements per arm of crs_gp_state.linv_t for a given row capacity: o * L^-T is stored panel-major,
 * panel p = rows j in [0, min(n_cap, 32 (p+1))) x columns [32 p, 32 p + 32), row-major with a row
 * stride of 40 (fp64) / 32 (fp32) elements, so that the grid kernel stages each panel with one TMA
 * bulk copy and reads it bank-conflict-free. */
CRS_API int64_t crs_linv_stride(int32_t n_cap, int32_t dtype);

/* Bytes of workspace crs_ocba_scan needs; zero it once, the kernel leaves it zeroed. */
CRS_API int64_t crs_scan_workspace_bytes(void);

/*
 * Replaces the one-off work of `ModelListGP.posterior` that GPyTorch caches after the first
 * eval-mode call: input/outcome transforms, K^ = K(X,X) + noise*I, its Cholesky factor and
 * alpha.  Arms [arm_begin, arm_end); one CTA per arm, fp64 throughout,
 * results cast to T.  Re-run it for one arm after appending an observation to get the
 * `condition_on_observations` update of experiments/wsc_experiments/main.py:356-362.
 */
CRS_API int crs_gp_prepare(const crs_gp_state* st, int32_t arm_begin, int32_t arm_end, void* stream);

/*
 * Replaces `model.posterior(context_set)` -> (.mean, .variance)
 * (contextual_rs/contextual_rs_strategies.py:133-136, generalized_pcs.py:443-444,
 * continuous_context.py:68,96, experiments/wsc_experiments/main.py:460).
 * ctx: [n_c, d] T.  Writes columns [col_offset + arm - arm_begin] of mean/var ([n_c, ld], T) for
 * arms [arm_begin, arm_end).  Variance is clamped at gpytorch's min_variance (1e-10 / 1e-6).
 */
CRS_API int crs_posterior_grid(const crs_gp_state* st, const void* ctx, int64_t n_c,
                       int32_t arm_begin, int32_t arm_end,
                       void* mean, void* var, int64_t ld, int32_t col_offset, void* stream);

/*
 * Replaces sort + gather + zeta + global argmin of gao_modellist
 * (contextual_rs/contextual_rs_strategies.py:137-157), `_get_hat_Z` / `MinZeta.forward`
 * (contextual_rs/continuous_context.py:21-36, 57-74; min_zeta[c] == -MinZeta(c)) and the best-arm
 * readout `posterior.mean.argmax` (experiments/wsc_experiments/main.py:459-468).
 * mean/var: [n_c, ld] T, K valid columns.  Per-context outputs (any may be NULL):
 *   best_arm[n_c] int32, min_zeta[n_c] T, min_arm[n_c] int32, tie_cnt[n_c] int32.
 * If ext_best_* are non-NULL the predicted best of every context is taken from them (global arm
 * id, mean, variance — the all-gathered best of an arm-sharded run) instead of the local row.
 * arm_offset is added to local column indices to form arm ids.  `decision` (may be NULL) receives
 * min_zeta/context/zeta_arm/best_arm/tie_count.  workspace: crs_scan_workspace_bytes() bytes, zeroed.
 * Tie rule (the reference's is implementation-defined): lowest arm id wins the
 * arg-max; among equal zeta the larger mean then the lower arm id; among contexts the lowest.
 */
CRS_API int crs_ocba_scan(const void* mean, const void* var, int64_t n_c, int32_t K, int64_t ld,
                  int32_t dtype, int32_t arm_offset,
                  const int32_t* ext_best_arm, const void* ext_best_mean, const void* ext_best_var,
                  int32_t* best_arm, void* min_zeta, int32_t* min_arm, int32_t* tie_cnt,
                  crs_decision* decision, void* workspace, void* stream);

/*
 * Replaces the tie randomisation of contextual_rs/contextual_rs_strategies.py:147-154: given the
 * caller's draw r = torch.randint(tie_count, (1,)) rewrites decision->{context, zeta_arm, best_arm}
 * to the r-th tied entry in flat (context-major) order.
 */
CRS_API int crs_ocba_resolve_tie(const void* mean, const void* var, int64_t n_c, int32_t K, int64_t ld,
                         int32_t dtype, const int32_t* best_arm, const void* min_zeta,
                         const int32_t* tie_cnt, int64_t r, crs_decision* decision, void* stream);

/*
 * Replaces OCBA step 2(b) (contextual_rs/contextual_rs_strategies.py:158-202,
 * contextual_rs/continuous_context.py:110-132): y_i = p_i / var_i at the selected context, psi of
 * the best arm vs the sum over the others, and the final arm.  Reads decision->context/zeta_arm/
 * best_arm from device memory (no host round-trip) and writes psi_best, psi_rest, next_arm.
 * ctx: [n_c, d] T; var: [n_c, ld] T with st->num_arms valid columns; arm_offset maps local columns
 * to arm ids (arm-sharded runs: psi_* are then partial sums the host all-reduces).  decision->
 * reserved[0] receives (first arm with a failed Cholesky)+1, or 0.
 */
CRS_API int crs_ocba_select(const crs_gp_state* st, const void* ctx, const void* mean, const void* var,
                    int64_t ld, int32_t arm_offset, int32_t p_mode, double kernel_scale,
                    crs_decision* decision, void* stream);

/*
 * crs_ocba_scan + crs_ocba_select in ONE launch (the last-arriving CTA of the scan runs step 2(b)):
 * the decision path of gao_modellist without a kernel boundary between argmin and arm choice.
 * Same outputs as the two calls; K = st->num_arms, dtype = st->dtype.
 */
CRS_API int crs_ocba_scan_select(const crs_gp_state* st, const void* ctx, const void* mean, const void* var,
                                 int64_t n_c, int64_t ld, int32_t arm_offset, int32_t p_mode,
                                 double kernel_scale, int32_t* best_arm, void* min_zeta,
                                 int32_t* min_arm, int32_t* tie_cnt, crs_decision* decision,
                                 void* workspace, void* stream);

/*
 * Replaces the sampling half of estimate_current_generalized_pcs for a ModelListGP
 * (contextual_rs/generalized_pcs.py:441-475): for every context, the number of the S posterior
 * draws in which the predicted-best arm beats every other arm (func_I = strict indicator).  Draws
 * are per-(context, arm) marginals from a Philox4x32-10 stream keyed by
 * (seed; sample, context + ctx_offset, (arm + arm_offset) >> 2, offset) — one call yields the draws
 * of a sample at 4 consecutive arms, so arm_offset must be a multiple of 4; nothing S-sized touches
 * HBM.  counts: [n_c] uint32.  stats (may be NULL): [2] uint64 = {draws executed, nominal S*K*n_c
 * minus executed}: samples stop at their first failure / when no remaining arm can win (exact).
 */
CRS_API int crs_pcs_counts(const void* mean, const void* var, int64_t n_c, int32_t K, int64_t ld,
                   int32_t dtype, int64_t num_samples, uint64_t seed, uint32_t offset,
                   int32_t arm_offset, int64_t ctx_offset, uint32_t* counts, uint64_t* stats,
                   void* stream);

/*
 * Same estimate as crs_pcs_counts — counts[c] = number of the S posterior draws in which the
 * predicted-best arm of context c beats every other arm (contextual_rs/generalized_pcs.py:441-475,
 * func_I = strict indicator) — under the "max-tree" Philox stream: the K - 1 competitors are the
 * five most threatening ones, drawn directly, plus the leaves of a 4-ary tree in threat order; a
 * node carries the largest standard normal among its leaves and the leaf that owns it, and children
 * are drawn conditionally (truncated maxima), so a sample is decided after ~1-3 Philox calls instead
 * of ~K/4 although the K - 1 competitor draws are iid N(0,1) exactly as before (stream definition +
 * brute-force restatement: oracle/pcs_tree_ref.py).
 * Counters (sample, context + ctx_offset, node, offset): counts do not depend on how contexts are
 * sharded.  The full row of K arms must be passed (2 <= K <= 4102, else CRS_ERR_UNSUPPORTED).
 * counts: [n_c] uint32.  stats (may be NULL): [3] uint64 = {root tests, node expansions, direct-arm
 * tests} (one Philox call each).
 */
CRS_API int crs_pcs_counts_tree(const void* mean, const void* var, int64_t n_c, int32_t K, int64_t ld,
                                int32_t dtype, int64_t num_samples, uint64_t seed, uint32_t offset,
                                int64_t ctx_offset, uint32_t* counts, uint64_t* stats, void* stream);

/*
 * crs_ocba_scan_select for ONE decision sharded over `px->world` GPUs by contexts — the multi-GPU form
 * of gao_modellist's scan (contextual_rs/contextual_rs_strategies.py:137-202): the local scan + step
 * 2(b) run as in crs_ocba_scan_select, then the ranks' records are combined on the device through the
 * peer mailboxes (crs_peer_exchange) and the GLOBAL decision (global context index) lands in `decision`
 * and, if decision_host is mapped pinned memory, in *decision_host with reserved[1] = px->seq written
 * last (poll it with crs_wait_decision).  n_c = 0 (a rank without contexts) takes part with an empty
 * record.  reserved[0] = -2 reports a peer that did not post within 2 s.
 */
CRS_API int crs_ocba_scan_select_sharded(const crs_gp_state* st, const void* ctx, const void* mean, const void* var,
                                         int64_t n_c, int64_t ld, int32_t p_mode, double kernel_scale,
                                         int32_t* best_arm, void* min_zeta, int32_t* min_arm, int32_t* tie_cnt,
                                         crs_decision* decision, void* workspace, const crs_peer_exchange* px,
                                         crs_decision* decision_host, void* stream);
/* Spin (GIL-free for ctypes callers) until decision_host->reserved[1] == seq; CRS_ERR_CUDA after timeout_ms. */
CRS_API int crs_wait_decision(const crs_decision* decision_host, int32_t seq, int32_t timeout_ms);
/*
 * ARM-sharded decision (the partition BASELINE.json's north_star names: rank g owns arms [a0, a1) of ALL
 * contexts).  These are the device-side steps between its collectives, so that
 * gao_modellist's sort head / zeta argmin / step 2(b) (contextual_rs/contextual_rs_strategies.py:137-157,
 * 195-201) stay on the stream:
 *   crs_arm_pack_best     best_arm[n_c] (GLOBAL ids, from crs_ocba_scan with arm_offset = a0) + the local
 *                         tables -> pack[3][n_c] doubles (mean, variance, arm id): the all-gather payload
 *   crs_arm_combine_best  pack_all[world][3][n_c] -> global best per context (largest mean, lowest arm id)
 *                         as ext_arm / ext_mean / ext_var, the ext_best_* inputs of crs_ocba_scan
 *   crs_arm_record_mean   stores the mean of the record's zeta arm in decision->psi_best (tie rule below)
 *   crs_arm_fold          records[world] (all-gathered crs_decision of the ranks' scans against the global
 *                         best) -> the global minimiser: smallest zeta, lowest context, larger mean of the
 *                         zeta arm, lower arm id; tie counts at the minimum added; best_arm = ext_arm[context]
 *   crs_arm_finish        psi[2] = all-reduced (psi_best, psi_rest) partials of crs_ocba_select ->
 *                         next_arm = psi_best < psi_rest ? best_arm : zeta_arm; *bad != 0 -> reserved[0]
 */
CRS_API int crs_arm_pack_best(const void* mean, const void* var, int64_t n_c, int64_t ld, int32_t dtype, int32_t arm_offset,
                              const int32_t* best_arm, double* pack, void* stream);
CRS_API int crs_arm_combine_best(const double* pack_all, int32_t world, int64_t n_c, int32_t dtype, int32_t* ext_arm,
                                 void* ext_mean, void* ext_var, void* stream);
CRS_API int crs_arm_record_mean(crs_decision* decision, const void* mean, int64_t ld, int32_t dtype, int32_t arm_offset,
                                void* stream);
CRS_API int crs_arm_fold(const crs_decision* records, int32_t world, const int32_t* ext_arm, crs_decision* out, void* stream);
CRS_API int crs_arm_finish(crs_decision* decision, const double* psi, const int32_t* bad, void* stream);

/*
 * Sum n <= 8 counters over the ranks of a mailbox group (the "one allreduce" of the PCS success counts of
 * a context-sharded estimate before rho, contextual_rs/generalized_pcs.py:470-479) in
 * one single-CTA launch: `values` (device) -> `out` (device, may be NULL) and, if out_host is mapped
 * pinned memory of 10 x uint64, out_host[0..7] = sums, [9] = 1 on a peer time-out, [8] = px->seq written
 * LAST (poll with crs_wait_u64).  world = 1 copies.
 */
CRS_API int crs_peer_sum_u64(const uint64_t* values, int32_t n, uint64_t* out, uint64_t* out_host,
                             const crs_peer_exchange* px, void* stream);
/* Spin until *tag_host == seq (mapped pinned memory written by a kernel); CRS_ERR_CUDA after timeout_ms. */
CRS_API int crs_wait_u64(const uint64_t* tag_host, uint64_t seq, int32_t timeout_ms);
/* Mailbox plumbing: plain cudaMalloc'd (zeroed) memory + CUDA IPC handles (64 bytes) to map it in peers. */
CRS_API int64_t crs_mailbox_bytes(int32_t world);
CRS_API int crs_device_alloc(int64_t bytes, void** dev_ptr);
CRS_API int crs_device_free(void* dev_ptr);
CRS_API int crs_ipc_export(void* dev_ptr, unsigned char handle[64]);
CRS_API int crs_ipc_open(const unsigned char handle[64], void** dev_ptr);
CRS_API int crs_ipc_close(void* dev_ptr);

/*
 * Same estimate again — counts[c] = number of the S posterior draws in which the predicted-best arm of
 * context c beats every other arm (contextual_rs/generalized_pcs.py:441-475, func_I = strict
 * indicator) — under the "cdf" Philox stream, the default sampler: arms are independent, so given the
 * best arm's draw y_b = mu_b + sd_b z the sample succeeds with probability
 * F_c(z) = prod_{j != b} Phi((y_b - mu_j) / sd_j); one sample is z ~ N(0,1), u ~ U(0,1), success iff
 * u < F_c(z) — the same Bernoulli law as drawing all K normals, for one normal and one uniform.  A
 * build kernel tabulates F_c per context over its live range (cubic on 29 / 61 / 93 / 125 cells, competitors too
 * sharp for the grid kept exact), a sampling kernel spends one Philox4x32-10 call per two samples.  Stream definition and CPU
 * restatement: oracle/pcs_cdf_ref.py.  Counters (call, context + ctx_offset, 0xCDF00001, offset):
 * counts do not depend on how contexts are sharded.  Any K >= 2.
 * workspace: crs_pcs_cdf_workspace_bytes(n_c) bytes of device memory, 16-byte aligned, no
 * initialisation needed.  counts: [n_c] uint32.  stats (may be NULL): [6] uint64 = {Philox calls,
 * ln Phi terms evaluated by the build, contexts with exactly-evaluated sharp competitors, contexts
 * with more than 8 of them (table accuracy not guaranteed), total successes, 0}.
 */
CRS_API int64_t crs_pcs_cdf_workspace_bytes(int64_t n_c);
CRS_API int crs_pcs_counts_cdf(const void* mean, const void* var, int64_t n_c, int32_t K, int64_t ld,
                               int32_t dtype, int64_t num_samples, uint64_t seed, uint32_t offset,
                               int64_t ctx_offset, uint32_t* counts, uint64_t* stats, void* workspace,
                               void* stream);

/*
 * The same counts from CALLER-SUPPLIED standard normals — the `base_samples` argument of
 * estimate_current_generalized_pcs (contextual_rs/generalized_pcs.py:325, 449-451; the tensor the
 * reference passes to `posterior.rsample`, shape num_samples x n_c x K).  base: [S, n_c, K] T,
 * contiguous; element (s, c, k) is used as the marginal draw of arm k at context c in sample s:
 * y = mean + sqrt(var) * z (IEEE sqrt, mul, add — no FMA contraction, so a torch restatement of the
 * same formula gives identical counts), counts[c] = #{s : y_best - max_rest > 0}
 * (generalized_pcs.py:464-473).  Reads S * n_c * K normals: for the reference's call shapes.
 */
CRS_API int crs_pcs_counts_base(const void* mean, const void* var, int64_t n_c, int32_t K, int64_t ld,
                                int32_t dtype, const void* base, int64_t num_samples, uint32_t* counts,
                                void* stream);

/* The tree sampler's fp32 inverse normal for tests: out[i] = z with Phi(z) = exp(-t[i]), t[i] >= 0
 * (odd i with t < 1/8 take the small-t variant used at the root — same values by construction). */
CRS_API int crs_debug_zfun(const float* t, int64_t n, float* out, void* stream);

/* Raw Philox4x32-10 words for known-answer tests: out[4*i..4*i+3] = philox(ctr[4*i..], key). */
CRS_API int crs_philox_words(const uint32_t* ctr, int64_t n, uint32_t key0, uint32_t key1, uint32_t* out,
                     void* stream);

/*
 * LEVI grid acquisition on a posterior grid — replaces `_compute_all_EI_reviewer`
 * (contextual_rs/levi.py:132-172), `PredictiveEI.forward` (:183-196) and `discrete_levi`
 * (:199-227):  Delta_a = -|mu_a - max_{j != a} mu_j|,  sigma~_a = var_a / sqrt(var_a + noise_a
 * stdv_a^2) with var clamped at 1e-9,  EI_a = sigma~_a (pdf(u) + u cdf(u)),  u = Delta_a / sigma~_a.
 * mean/var: [n_c, ld] T tables of the st->num_arms arms (T = st->dtype); noise and y_std are read
 * from `st` (levi.py:158-161: likelihood noise x outcome_transform._stdvs_sq).  Outputs:
 * max_ei[n_c] T = row maximum of EI (times weights[c] if weights != NULL, levi.py:221-223),
 * max_arm[n_c] = its arm (first index on ties), optional ei_out [n_c, ld_ei] T = the full EI table,
 * optional decision = first arg-max over contexts.  K >= 2.
 */
CRS_API int crs_levi_scan(const crs_gp_state* st, const void* mean, const void* var, int64_t n_c,
                          int64_t ld, const void* weights, void* ei_out, int64_t ld_ei, void* max_ei,
                          int32_t* max_arm, crs_levi_decision* decision, void* stream);

/*
 * Backward of `MinZeta.forward` (contextual_rs/continuous_context.py:57-74): grad[c, :] =
 * d(-min_j zeta_{c,j}) / d(ctx[c, :]), what `optimize_acqf` obtains by autograd in
 * experiments/continuous_experiments/main.py:252-263.  mean/var: the [n_c, ld] tables of
 * crs_posterior_grid for `ctx`; best_arm/min_arm: the per-context outputs of crs_ocba_scan (global
 * arm ids of `st`, -1 = none).  grad: [n_c, d] T.  Only the two arms zeta* involves are touched.
 */
CRS_API int crs_min_zeta_grad(const crs_gp_state* st, const void* ctx, int64_t n_c, const void* mean,
                              const void* var, int64_t ld, const int32_t* best_arm,
                              const int32_t* min_arm, void* grad, void* stream);

/*
 * Exact marginal log-likelihood of arms [arm_begin, arm_end) at the hyper-parameters currently in
 * `st` (lengthscale, outputscale, noise, mean_const) and its gradient — the objective of
 * `fit_gpytorch_mll(ExactMarginalLogLikelihood(m.likelihood, m))` in
 * contextual_rs/experiment_utils.py:41-47, 108-115.  Uses the frozen transforms of `st`
 * (norm_mins/norm_ranges, y_mean/y_std, as learnt by crs_gp_prepare) and only the fitted-description
 * arrays; the prepared arrays are not touched.
 *   value[k]          = log N(y~_k | m 1, o C(l) + noise I)        (-inf if not positive definite)
 *   grad[k, 0..d-1]   = d value / d lengthscale_j;  grad[k, d] = d/d outputscale;
 *   grad[k, d+1]      = d/d noise;  grad[k, d+2] = d/d mean_const
 * value: [K] f64, grad: [K, d+3] f64 (rows outside the arm range are not written).  n_obs <= 128.
 */
CRS_API int crs_gp_mll(const crs_gp_state* st, int32_t arm_begin, int32_t arm_end, double* value,
                       double* grad, void* stream);

/* Test hook for the grid kernel's fp64 math: out[i] = f(r2[i]) with variant 0 = the kernel's
 * correlation (table exp + Newton sqrt), 1 = the same formula on libdevice exp/sqrt, 2/3 = the
 * kernel's sqrt with 1/2 Newton steps, 4 = the kernel's exp(-x). */
CRS_API int crs_debug_corr(const double* r2, int64_t n, int32_t kernel, int32_t variant, double* out,
                           void* stream);

/* Pipe-peak probe for the roofline denominators that MEASURED_PEAKS.json lacks: `blocks` CTAs of
 * 256 threads each run `iters` x 64 dependent-chain FMAs on 16 independent accumulators
 * (kind 0 = FP64, 1 = FP32): flops = blocks * 256 * iters * 64 * 2; kind 2 = FP64 mma.sync m8n8k4;
 * kind 5 = the bare PCS draw sequence (Philox4x32-10 + 2 Box-Muller + 4 FMA + max per iteration,
 * registers only): draws = blocks * 256 * iters * 4; kinds 6 / 7 = the bare root test / node
 * expansion of the max-tree PCS kernel: steps = blocks * 256 * iters.  out: >= 8 bytes, device. */
CRS_API int crs_peak_probe(int32_t kind, int32_t iters, int32_t blocks, void* out, void* stream);

/*
 * Whole decision through one call with HOST context buffers — the reference-facing entry that
 * stands in for `gao_modellist(model, context_set, ...)` when context_set lives on the CPU, as it
 * does in the reference's driver (experiments/wsc_experiments/main.py:226-228).
 * Copies ctx_host (pinned or pageable, [n_c, d] T) to ws->ctx, optionally re-prepares arms
 * [prep_begin, prep_end) and runs grid -> fused scan + select on `stream`.  A mapped pinned ctx_host
 * of <= 1 MiB is pulled over PCIe by extra CTAs of the prepare launch (the copy overlaps prepare);
 * otherwise a cudaMemcpyAsync precedes the kernels.  The three kernels are chained by programmatic
 * dependent launch.  If decision_host is
 * pinned memory the last kernel writes the decision straight into it (mapped, zero-copy) and the
 * call spins on a completion tag (decision_host->reserved[1]); otherwise it copies the decision
 * back and synchronises the stream.  Either way the decision is complete on return.  Returns CRS_TIES_PENDING if randomize_ties and
 * tie_count > 1 (the caller then draws r, calls crs_ocba_resolve_tie + crs_ocba_select).
 * ws_* are caller-owned device buffers: ctx [n_c,d] T, mean/var [n_c,K] T, best_arm/min_arm/
 * tie_cnt [n_c] int32, min_zeta [n_c] T, decision (device), scan workspace.
 */
typedef struct crs_decide_ws {
  void* ctx; void* mean; void* var;
  int32_t* best_arm; void* min_zeta; int32_t* min_arm; int32_t* tie_cnt;
  crs_decision* decision; void* scan_ws;
} crs_decide_ws;

CRS_API int crs_gao_decide_host(const crs_gp_state* st, const void* ctx_host, int64_t n_c,
                        int32_t prep_begin, int32_t prep_end, int32_t p_mode, double kernel_scale,
                        int32_t randomize_ties, const crs_decide_ws* ws,
                        crs_decision* decision_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRS_B200_H */
