/*
 * gse.h -- C ABI of libgse_b200.so: the B200 (sm_100a) state-estimation hot path of gpu_se.
 *
 * The reference (AlgorithmicAmoeba/gpu_se) is pure Python and has no FFI of its own; the
 * interface this library sits behind is the reference's filter-class API
 *     filter/particle.py:43-114,151-327   ParticleFilter / ParallelParticleFilter
 *     filter/gs_ukf.py:45-183,223-449     (Parallel)GaussianSumUnscentedKalmanFilter
 *     gaussian_sum_dist/MultivariateGaussianSum.py:27-97
 * Each entry point below names the reference lines whose work it replaces.  The Python classes in
 * gpu_se_b200/ mirror the reference's constructors and methods and call these functions through
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *  - every function returns 0 on success, a negative gse_status otherwise; gse_last_error()
 *    returns a thread-local message for the last failure.
 *  - pointers named *_dev are raw DEVICE pointers owned by the caller (the Python side allocates
 *    them as torch tensors); the library allocates only the per-context workspace, the
 *    per-context step-parameter block, and what gse_peer_alloc is asked for.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *  - particle / component state is struct-of-arrays float32: column c of n rows starts at
 *    base + c*ld  (ld = leading dimension in elements, a multiple of 4 so that every column is
 *    16-byte aligned; base 16-byte aligned).
 *  - weights are kept as a float32 log-likelihood array `loglik` (sum of log pdf values since the
 *    last reset) times an optional float64 base weight array `base` (NULL = uniform):
 *        weight_k  proportional to  base_k * exp(loglik_k)
 *    A NULL `loglik` input means "all zero" (the state right after a resample: nothing is read).
 *    `stats_dev` is 4 doubles: [0] = M = max_k loglik_k, [1] = S = sum_k exp(loglik_k - M)
 *    (both written by the update kernels and gse_loglik_max, consumed by scan / moments; the
 *    sharded driver merges the shards' pairs in between: gse_peer_allgather_stats), [2..3] reserved.
 *  - resampling is LAZY: gse_resample_search produces the int32 ancestor index `idx` (the
 *    reference's sample_index, particle.py:100 / :314); the kernels that consume the population
 *    next (predict, moments) take `idx_dev` and read row idx[i] of the pre-resample state for
 *    output row i, so `particles[sample_index]` (particle.py:102 / :315) never makes its own trip
 *    through HBM.  idx_dev == NULL means "rows are in place".  gse_gather_rows materialises the
 *    gather when the rows themselves are wanted.  idx buffers are 16-byte aligned and hold at
 *    least round_up(n, 4) entries.
 *  - one context per GPU/filter; a context is not thread-safe; distinct contexts are independent.
 */
#ifndef GSE_H_
#define GSE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSE_ABI_VERSION 7

#define GSE_NX 5        /* states  (Cg, Cx, Cfa, Ce, Ch)   model/BioreactorModel.py:191 */
#define GSE_NU 2        /* inputs  (Fg_in, Fm_in)          model/BioreactorModel.py:195 */
#define GSE_NY 2        /* outputs (Cg*180, Cfa*116)       model/BioreactorModel.py:251-253 */
#define GSE_NSIGMA 11   /* 2*Nx+1 sigma points             filter/gs_ukf.py:61 */
#define GSE_NCOV 15     /* lower triangle of a 5x5 covariance, row-major (00,10,11,20,21,22,...) */
#define GSE_MAX_ND 8    /* mixture components supported */
#define GSE_MAX_SHARDS 8   /* GPUs of one node a population can be sharded over */
#define GSE_IPC_HANDLE_BYTES 64

#define GSE_MODEL_BIOREACTOR 1   /* Bioreactor.homeostatic_DEs / static_outputs */

typedef enum gse_status {
    GSE_OK = 0,
    GSE_EINVAL = -1,      /* bad argument (null pointer, misaligned, n out of range ...) */
    GSE_ECUDA = -2,       /* a CUDA runtime call or launch failed; see gse_last_error() */
    GSE_ENOMEM = -3,
    GSE_ELINALG = -4      /* mixture covariance not positive definite */
} gse_status;

/* Gaussian-sum parameters as the reference holds them
 * (MultivariateGaussianSum.__init__, MultivariateGaussianSum.py:27-37): row-major
 * means (nd, nx), covs (nd, nx, nx), weights (nd).  nx is 5 for state mixtures, 2 for the
 * measurement mixture. */
typedef struct gse_mixture {
    int32_t nd;
    int32_t nx;
    double weights[GSE_MAX_ND];
    double means[GSE_MAX_ND * GSE_NX];
    double covs[GSE_MAX_ND * GSE_NX * GSE_NX];
} gse_mixture;

typedef struct gse_ctx gse_ctx;

/* Conditions detected INSIDE kernels cannot fail the call that launched them (nothing synchronises).  The kernels OR
 * these bits into a per-context error word in host-mapped memory; gse_ctx_errors() returns it (and sets
 * gse_last_error() to a description) whenever the caller is at a synchronisation point anyway -- the Python classes
 * poll it after every moments read-back and raise. */
#define GSE_ERR_CHOLESKY 1u        /* GS-UKF: covariance not positive definite even with the +1e-10 I retry (gs_ukf.py:72-75 raises LinAlgError) */
#define GSE_ERR_SINGULAR_PYY 2u    /* GS-UKF update: innovation covariance P_yy singular / not positive (gs_ukf.py:132) */
#define GSE_ERR_PEER_TIMEOUT 4u    /* sharded run: a peer's mailbox flag did not arrive within the spin bound */
#define GSE_ERR_QUEUE_OVERFLOW 8u  /* fused resample: heavy-run queue full (workspace sized for fewer outputs) */
#define GSE_ERR_ZERO_WEIGHTS 16u   /* resample of weights that are all zero (the reference divides by zero, particle.py:90) */

int gse_abi_version(void);
const char* gse_last_error(void);

/* Context: mixture constants (inverse covariances, normalising constants as
 * MultivariateGaussianSum.py:33-37; Cholesky factors for sampling) + scan/reduction workspace
 * for up to n_max rows.  `state` is the process-noise mixture (nx = 5), `meas` the measurement
 * noise mixture (nx = 2).  Replaces the JIT of f/g in ParallelParticleFilter.__init__
 * (particle.py:151-208): f and g are compiled in, selected by model_id. */
int gse_ctx_create(int device, int model_id, int64_t n_max, const gse_mixture* state,
                   const gse_mixture* meas, gse_ctx** out);
int gse_ctx_destroy(gse_ctx* ctx);

/* The context's device-error word (GSE_ERR_* bits, 0 = none); clear != 0 resets it.  Does not synchronise: call it
 * after the stream the kernels ran on has been synchronised. */
unsigned int gse_ctx_errors(gse_ctx* ctx, int clear);

/* A 64-double result area in host-mapped memory owned by the context: `*dev_out` may be passed wherever a kernel takes
 * a moments pointer (gse_resample_fused's `moments_dev`), the block then lands in `*host_out` without a copy of its
 * own -- the read-back of `point_estimate()` (particle.py:318-320) in the filter loop.  Valid once the stream has been
 * synchronised (gse_ctx_wait). */
int gse_ctx_result_block(gse_ctx* ctx, double** host_out, double** dev_out);

/* cudaStreamSynchronize(stream) + gse_ctx_errors(ctx, clear = 1) in one call; `errors_out` (may be NULL) receives the
 * GSE_ERR_* bits. */
int gse_ctx_wait(gse_ctx* ctx, void* stream, unsigned int* errors_out);

/* Per-step scalars in DEVICE memory.  While a context has a parameter block attached
 * (gse_ctx_set_step_params), every launch made through it reads u, dt, z, r and the Philox step
 * counter from the block instead of from the by-value arguments of the call.  A captured CUDA
 * graph of predict -> update -> resample can then be replayed step after step: the host only
 * refreshes the 64-byte block (one small H2D copy) before each replay.  For the launch-bound sizes
 * (all of the reference's GS-UKF sweep, particle filters up to ~2^18). */
typedef struct gse_step_params {
    double u[GSE_NU];
    double dt;
    double z[GSE_NY];
    double r;
    uint64_t step;
    uint64_t reserved;
} gse_step_params;

/* The context owns one parameter block in device memory and a pinned staging ring.
 * gse_ctx_upload_step_params copies `values` (host) into the block on `stream` (asynchronous: the
 * values are staged in the ring first, so the caller's struct may be reused at once);
 * gse_ctx_use_step_params(ctx, 1) makes every later launch through the context read the block,
 * (ctx, 0) restores the by-value arguments. */
int gse_ctx_upload_step_params(gse_ctx* ctx, const gse_step_params* values, void* stream);
int gse_ctx_use_step_params(gse_ctx* ctx, int enable);

/* ---- sampling / density of a Gaussian sum (MultivariateGaussianSum.py:39-97) ---------------- */

/* x[c*ld + i] = component c of sample i of `mix` (nx <= 5 columns), i in [0, n): Philox4x32-10 stream keyed by
 * (seed; counter = (index0 + i, step, subsequence)).  Replaces x0.draw(N) in
 * ParticleFilter.__init__ (particle.py:49) and MultivariateGaussianSum.draw (:65-97); rows are
 * NOT grouped by component (quirk Q5 of SURVEY.md is not reproduced). */
int gse_mixture_draw(gse_ctx* ctx, const gse_mixture* mix, float* x_dev, int64_t ld, int64_t n,
                     uint64_t seed, uint64_t step, int64_t index0, void* stream);

/* out[i] = pdf of `mix` at row i of the SoA points x (nx columns), float64
 * (MultivariateGaussianSum.pdf, :39-63).  log_out != 0 writes log pdf instead. */
int gse_mixture_pdf(gse_ctx* ctx, const gse_mixture* mix, const float* x_dev, int64_t ld,
                    int64_t n, double* out_dev, int log_out, void* stream);

/* ---- particle filter ------------------------------------------------------------------------ */

/* predict (particle.py:54-67 / :265-277): x_i += f(x_i, u, dt) [n_sub explicit-Euler sub-steps of
 * dt/n_sub; the reference has n_sub = 1], then x_i += state noise.  Reads row idx[i] (or row i when
 * idx_dev is NULL) of x_src, writes row i of x_dst; x_src == x_dst is allowed only without idx.
 * noise_dev == NULL: noise drawn in-kernel (Philox, counter = (index0 + i, step)).
 * noise_dev != NULL: SoA (5, ld_noise) float32 host-supplied draws are added instead (the
 * DeterministicGaussianSum cross-check mode, DeterministicGaussianSum.py:32-65). */
int gse_pf_predict(gse_ctx* ctx, const float* x_src_dev, int64_t ld_src, const int32_t* idx_dev,
                   float* x_dst_dev, int64_t ld_dst, int64_t n, const double u[GSE_NU], double dt,
                   int n_sub, uint64_t seed, uint64_t step, int64_t index0, const float* noise_dev,
                   int64_t ld_noise, void* stream);

/* update (particle.py:69-83 / :279-294): loglik_i = loglik_in_i + log pdf_meas(z - g(x_i, u))
 * (loglik_in_dev NULL = zeros; may equal loglik_dev);
 * stats_dev[0] = max_i loglik_i, stats_dev[1] = sum_i exp(loglik_i - max). */
int gse_pf_update(gse_ctx* ctx, const float* x_dev, int64_t ld, int64_t n, const float* loglik_in_dev,
                  float* loglik_dev, const double u[GSE_NU], const double z[GSE_NY], double* stats_dev,
                  void* stream);

/* point_estimate + point_covariance in one pass (particle.py:105-114 / :318-327), over rows
 * idx[i] (or i) of x:
 * out_dev[0] = S = sum_i w_i, out_dev[1..5] = sum_i w_i (x_i - p), out_dev[6..20] = lower triangle of
 * sum_i w_i (x_i - p)(x_i - p)' with pivot p = out_dev[21..25] (written by the kernel: row 0),
 * where w_i = base_i * exp(loglik_i - stats_dev[0]).  26 doubles.  mean_only != 0 skips the second
 * moments (out_dev[6..20] untouched): point_estimate alone at a quarter of the float64 work. */
int gse_pf_moments(gse_ctx* ctx, const float* x_dev, int64_t ld, int64_t n, const int32_t* idx_dev,
                   const float* loglik_dev, const double* base_dev, const double* stats_dev,
                   int mean_only, double* out_dev, void* stream);

/* ---- weights and systematic resampling (shared by PF and GS-UKF) ------------------------------ */

/* stats_dev[0] = max_i loglik_i, stats_dev[1] = sum_i exp(loglik_i - max) (for callers that set
 * loglik themselves). */
int gse_loglik_max(gse_ctx* ctx, const float* loglik_dev, int64_t n, double* stats_dev,
                   void* stream);

/* out[i] = base_i * exp(loglik_i) * scale in float64: the reference's (un-normalised) `weights`
 * attribute (particle.py:50,83).  base_dev NULL = 1. */
int gse_weights_linear(gse_ctx* ctx, const float* loglik_dev, const double* base_dev, int64_t n,
                       double scale, double* out_dev, void* stream);

/* cumsum (particle.py:89 / :301-303, numpy.cumsum / torch.cumsum): inclusive scan of the
 * fixed-point weights  q_i = rint(base_i * exp(loglik_i - stats_dev[0]) * 2^s)  into uint64
 * cumsum_dev, with s = 52 - ceil(log2(stats_dev[1])) so that the total stays below 2^53 (exact in float64)
 * (stats_dev[1] >= sum_i base_i * exp(loglik_i - stats_dev[0]); the update kernels write it).
 * Integer addition is associative, so the result does not depend on the scan structure, the
 * launch geometry or the number of GPUs.  loglik_dev may be NULL (weights = base; the caller
 * then sets stats_dev[1] = sum base).  total_dev[0] receives cumsum[n-1] (NULL to skip). */
int gse_scan_weights(gse_ctx* ctx, const float* loglik_dev, const double* base_dev,
                     const double* stats_dev, int64_t n, uint64_t* cumsum_dev,
                     uint64_t* total_dev, void* stream);

/* Systematic resample of outputs [out0, out0 + n_out) of n_total from the local cumulative
 * weights (particle.py:92-100; the reference GPU kernel _parallel_resample, :223-263, differs
 * only on exact ties -- the CPU comparison `cumsum[k] < u` is the one followed):
 *     u_i   = (i + r) / n_total                                  (float64, as the reference)
 *     idx_i = min{ k : (offset + cumsum[k]) / total >= u_i }     (float64 divide, as `cumsum /= cumsum[-1]`)
 * evaluated exactly through integer thresholds; idx_out_dev[i - out0] = idx_i (int32, local source
 * row).  offset/total are read from offtot_dev[0..1] (uint64; the sharded driver writes the
 * exclusive prefix of the shard totals and the global total there). */
int gse_resample_search(gse_ctx* ctx, const uint64_t* cumsum_dev, int64_t n_src,
                        const uint64_t* offtot_dev, double r, int64_t n_total, int64_t out0,
                        int64_t n_out, int32_t* idx_out_dev, void* stream);

/* resample() in ONE launch (particle.py:85-100 / :296-314): scan of the fixed-point weights (as gse_scan_weights:
 * q_i = rint(base_i * exp(loglik_i - stats_dev[0]) * 2^s), either array may be NULL), global total, and for every
 * local source row k the outputs it owns,
 *     idx_j = src_row0 + k   for   e_{k-1} <= j < e_k,    e_k = #{ j : (j + r) / n_total <= cumsum_k / total }
 * -- the same indices as gse_scan_weights + gse_resample_search, bit for bit, without the cumulative weights ever
 * reaching HBM.  Outputs j in [out0, out0 + n_out) are written to idx_out_dev[j - out0] (int32 global ancestor row);
 * a single-GPU population passes out0 = 0, n_out = n_total = n_src, src_row0 = 0.  total_dev receives the integer
 * total (NULL to skip).  The kernel's CTAs wait on one another: the grid is one wave of co-resident CTAs.
 *   moments_dev != NULL additionally yields the estimate of the RESAMPLED population (point_estimate straight after
 * resample, particle.py:105-108, the pattern of every filter loop of the reference) without reading the ancestor index
 * back: sum_k c_k x_k over the source rows, c_k = e_k - e_{k-1} the offspring of row k, taken from state_dev (5 SoA
 * columns ld apart).  Written as a moment block like gse_pf_moments(mean_only): [0] = outputs sourced, [1..5] = the
 * column sums, [21..25] = pivot (0), [41..42] = (0, n_total).  Needs log-likelihood weights (base_dev NULL) and the
 * whole output range.
 *   reset_stats != 0: when the kernel is done it stores (M, S) = (0, n_total) -- the uniform weights a resample leaves
 * (particle.py:103 / :316) -- into stats_dev[0..1], which saves the caller a launch. */
int gse_resample_fused(gse_ctx* ctx, const float* loglik_dev, const double* base_dev, const double* stats_dev,
                       int64_t n_src, double r, int64_t n_total, int64_t out0, int64_t n_out, int64_t src_row0,
                       int32_t* idx_out_dev, uint64_t* total_dev, const float* state_dev, int64_t ld,
                       double* moments_dev, int reset_stats, void* stream);

/* `resample_from_cumsum` (contract (ii)): systematic resample fed the CALLER'S float64
 * cumulative sum -- e.g. the reference's own numpy.cumsum / torch.cumsum array (particle.py:89-90 / :301-304).
 *     normalise == 0: cumsum_dev is already normalised (`cumsum /= cumsum[-1]` done by the caller)
 *     normalise != 0: every element is divided by cumsum_dev[n_src - 1] on the fly (div.rn.f64, as numpy)
 *     ties_right == 0: idx_j = #{ k : cumsum[k] <  u_j }  the reference CPU loop (:96-100), searchsorted 'left'
 *     ties_right != 0: idx_j = min(#{ k : cumsum[k] <= u_j }, n_src - 1)  the reference GPU kernel _parallel_resample
 *                      (:223-263; searchsorted 'right' -- the two differ only where a sample position ties with a
 *                      DUPLICATED cumsum value, and the kernel indexes one past the end for u = 1.0)
 * with u_j = (j + r) / n_total in float64.  Outputs [out0, out0 + n_out) -> idx_out_dev[j - out0] (int32).
 * Bit-exact for any non-decreasing cumsum. */
int gse_resample_search_f64(gse_ctx* ctx, const double* cumsum_dev, int64_t n_src, int normalise, int ties_right,
                            double r, int64_t n_total, int64_t out0, int64_t n_out, int32_t* idx_out_dev,
                            void* stream);

/* dst[:, i] = src[:, idx[i]] for ncols SoA columns, i in [0, n_out) (particles[sample_index],
 * particle.py:102 / :315); loglik_out_dev (NULL to skip) is zeroed (weights reset, :103 / :316). */
int gse_gather_rows(gse_ctx* ctx, const int32_t* idx_dev, int64_t n_out, const float* src_dev,
                    int64_t ld_src, float* dst_dev, int64_t ld_dst, int ncols, float* loglik_out_dev,
                    void* stream);

/* ---- sharded populations: peer memory over NVLink / NVSwitch (no reference counterpart: the
 * reference is single-GPU) ---------------------------------------------------- */

/* Device memory other processes of the node can map (cudaMalloc + cudaIpcGetMemHandle).  The
 * sharded filter keeps its state and cumulative-weight buffers here; every rank opens every other
 * rank's buffers once and the resample kernels then read them directly. */
int gse_peer_alloc(int device, int64_t bytes, void** ptr_out, unsigned char handle_out[GSE_IPC_HANDLE_BYTES]);
int gse_peer_free(int device, void* ptr);
int gse_peer_open(int device, const unsigned char handle[GSE_IPC_HANDLE_BYTES], void** ptr_out);
int gse_peer_close(int device, void* ptr);

/* The shards of one population: shard s owns the global rows [rows[s], rows[s+1]); cumsum_dev[s] /
 * state_dev[s] are its local cumulative weights (gse_scan_weights) and SoA state (leading
 * dimension ld[s]) -- the caller's own buffers for its own shard, gse_peer_open mappings for the
 * others.  offsets_dev is a DEVICE array of nshards + 1 uint64: the exclusive prefix of the shard
 * totals and, last, the global total (all-gathered on the stream, never seen by the host). */
typedef struct gse_shards {
    int32_t nshards;
    int64_t rows[GSE_MAX_SHARDS + 1];
    const uint64_t* cumsum_dev[GSE_MAX_SHARDS];
    const float* state_dev[GSE_MAX_SHARDS];
    int64_t ld[GSE_MAX_SHARDS];
    const uint64_t* offsets_dev;
    int32_t* idx_dev[GSE_MAX_SHARDS];    /* shard s's ancestor-index buffer (rows[s+1] - rows[s] entries, rounded up to 4) */
    int32_t rank;                        /* the caller's own shard (its rows are looked up without a search) */
} gse_shards;

/* gse_resample_search over the rows of ALL shards for the outputs [out0, out0 + n_out) (this
 * shard's own slots): idx_out_dev[i - out0] = GLOBAL ancestor row (int32).  The kernel reads the
 * other shards' cumulative weights through peer memory -- search and communication are one
 * kernel, nothing is staged and the host never synchronises. */
int gse_resample_search_sharded(gse_ctx* ctx, const gse_shards* shards, double r, int64_t out0,
                                int64_t n_out, int32_t* idx_out_dev, void* stream);

/* gse_resample_fused for a sharded population: EVERY rank calls it with its own weights; rank `rank` scans its rows
 * [rows[rank], rows[rank+1]), the shards' totals are exchanged through the peer mailboxes INSIDE the kernel (sequence
 * number epoch_totals), every rank ranks its rows against the global total and writes the GLOBAL ancestor row of each
 * output it sources into shards->idx_dev[t] of the shard t that owns the output slot (NVLink stores for t != rank);
 * a second mailbox exchange (epoch_done) at the end of the kernel makes every rank's index buffer complete when its
 * kernel completes.  One launch per rank per resample; total_dev receives the global integer total (NULL to skip).
 * moments_dev / state_dev / ld as in gse_resample_fused: this rank's block covers the outputs it SOURCES (whichever
 * shard owns their slots); the blocks of all ranks add up to the moments of the resampled population
 * (gse_peer_allgather_moments). */
int gse_resample_fused_sharded(gse_ctx* ctx, const float* loglik_dev, const double* base_dev, const double* stats_dev,
                               double r, const gse_shards* shards, void* const mailboxes[GSE_MAX_SHARDS], int rank,
                               unsigned int epoch_totals, unsigned int epoch_done, uint64_t* total_dev,
                               const float* state_dev, int64_t ld, double* moments_dev, int reset_stats,
                               void* stream);

/* dst[:, i] = state row idx[i] (global) pulled from the owning shard's memory, ncols SoA columns. */
int gse_gather_rows_sharded(gse_ctx* ctx, const gse_shards* shards, const int32_t* idx_dev,
                            int64_t n_out, float* dst_dev, int64_t ld_dst, int ncols, void* stream);

/* All-gather of one small per-shard record over peer-mapped mailboxes, fused with the reduction
 * that consumes it: ONE single-warp kernel per exchange, no NCCL launch, nothing on the host.
 * mailboxes[t] is shard t's mailbox (GSE_MAILBOX_BYTES of gse_peer_alloc memory, zero at start;
 * the caller's own for t == rank, gse_peer_open mappings otherwise).  `epoch` must be the same on
 * every rank, non-zero, and increase by one with every exchange (of either kind) on a mailbox set.
 *   _stats : in/out stats_dev[0..1] = this shard's (M_s, S_s) -> (max_s M_s, sum_s S_s exp(M_s - M))
 *   _totals: total_dev[0] = this shard's total -> offsets_dev[0..nshards] (exclusive prefix, total last)
 * Every rank of the population must enqueue the same exchanges in the same order. */
#define GSE_MAILBOX_BYTES (2 * GSE_MAX_SHARDS * 512)
int gse_peer_allgather_stats(gse_ctx* ctx, void* const mailboxes[GSE_MAX_SHARDS], int rank, int nshards,
                             unsigned int epoch, double* stats_dev, void* stream);
int gse_peer_allgather_totals(gse_ctx* ctx, void* const mailboxes[GSE_MAX_SHARDS], int rank, int nshards,
                              unsigned int epoch, const uint64_t* total_dev, uint64_t* offsets_dev,
                              void* stream);

/*   _moments: mom_dev[0..47] = this shard's moment block (gse_pf_moments / gse_gsf_moments layout) -> out_dev[0..40]
 *             (NULL: mom_dev itself) = the moments of the whole population about shard 0's pivot, merged in shard
 *             order: identical on every rank.  stats_dev (may be NULL): the global (M, S) are copied to out_dev[41..42].
 *             out_dev may be the context's host-mapped result block (gse_ctx_result_block).
 * The waits are bounded (4 s): a peer that never arrives sets GSE_ERR_PEER_TIMEOUT in the context's error word. */
int gse_peer_allgather_moments(gse_ctx* ctx, void* const mailboxes[GSE_MAX_SHARDS], int rank, int nshards,
                               unsigned int epoch, double* mom_dev, const double* stats_dev, double* out_dev,
                               void* stream);

/* stats_dev[0..1] = (max_s M_s, sum_s S_s exp(M_s - M)) from the nshards all-gathered pairs
 * pairs_dev[2 s .. 2 s + 1] = (M_s, S_s) written by each shard's update kernel. */
int gse_merge_stats(gse_ctx* ctx, const double* pairs_dev, int nshards, double* stats_dev, void* stream);

/* predict() directly followed by update() -- the filter loop of the reference, particle.py:265-294 -- as ONE pass:
 * gse_pf_predict's arguments, then gse_pf_update's (z, loglik_in_dev or NULL when the accumulated log-likelihood is all
 * zero, loglik_dev, stats_dev).  The new rows are still in registers when their likelihood is evaluated; rows and
 * log-likelihoods are bit-identical to the two separate calls, stats_dev[1] differs in the last float32 bits (other
 * partition of the sum).  Built for the benchmark's specialisation only: gse_pf_can_fuse_update() says whether this
 * context / n_sub / index0 has it (diagonal two-component state noise, one Euler step, index0 a multiple of four,
 * two-component measurement mixture); callers fall back to the two calls otherwise. */
int gse_pf_can_fuse_update(const gse_ctx* ctx, int n_sub, int64_t index0);
int gse_pf_predict_update(gse_ctx* ctx, const float* x_src_dev, int64_t ld_src, const int32_t* idx_dev, float* x_dst_dev,
                          int64_t ld_dst, int64_t n, const double u[GSE_NU], double dt, int n_sub, uint64_t seed,
                          uint64_t step, int64_t index0, const double z[GSE_NY], const float* loglik_in_dev,
                          float* loglik_dev, double* stats_dev, void* stream);
/* ... reading its rows through the global ancestor index of a sharded population (gse_pf_predict_sharded) */
int gse_pf_predict_update_sharded(gse_ctx* ctx, const gse_shards* shards, const int32_t* idx_dev, float* x_dst_dev,
                                  int64_t ld_dst, int64_t n, const double u[GSE_NU], double dt, int n_sub, uint64_t seed,
                                  uint64_t step, int64_t index0, const double z[GSE_NY], const float* loglik_in_dev,
                                  float* loglik_dev, double* stats_dev, void* stream);

/* gse_pf_predict / gse_pf_moments reading row idx[i] (GLOBAL ancestor row from
 * gse_resample_search_sharded) straight out of the owning shard's memory: the lazy resample of a
 * sharded population -- the rows cross NVLink inside the kernel that consumes them. */
int gse_pf_predict_sharded(gse_ctx* ctx, const gse_shards* shards, const int32_t* idx_dev, float* x_dst_dev,
                           int64_t ld_dst, int64_t n, const double u[GSE_NU], double dt, int n_sub,
                           uint64_t seed, uint64_t step, int64_t index0, const float* noise_dev,
                           int64_t ld_noise, void* stream);
int gse_pf_moments_sharded(gse_ctx* ctx, const gse_shards* shards, const int32_t* idx_dev, int64_t n,
                           const float* loglik_dev, const double* base_dev, const double* stats_dev,
                           int mean_only, double* out_dev, void* stream);

/* Number of outputs i in [0, n_total) whose u_i maps at or below integer cumulative weight
 * `bound` of `total`, i.e. #{ i : fl(fl(bound)/fl(total)) >= u_i }: host helper used by the sharded
 * driver to split the output range between shards with the device's exact predicate. */
int64_t gse_count_outputs_below(uint64_t bound, uint64_t total, double r, int64_t n_total);

/* Host evaluation of the device's threshold q*(u) = min{ C : fl(fl(C)/fl(total)) >= u } (the same
 * inline function the kernels use); exported for the parity tests. */
uint64_t gse_threshold_u64(double u, uint64_t total);

/* ---- Gaussian-sum unscented Kalman filter ------------------------------------------------------ */

/* State: means SoA (5, ld), covariances SoA (15, ld) lower triangles, loglik (n).
 *
 * predict (gs_ukf.py:82-103 / :348-367): Cholesky (retry +1e-10 I, :72-75), 11 sigma points
 * mean +- L[:, j] (no scaling, quirk Q6), f on each, an independent state-noise draw per sigma
 * point (:99), weighted mean (numpy.average) and weighted scatter.  Reads component idx[i] (or i)
 * of the *_src arrays, writes component i.  noise_dev != NULL: host noise, SoA (55, ld_noise) with
 * row s*5 + j = component j of sigma point s. */
int gse_gsf_predict(gse_ctx* ctx, const float* mean_src_dev, const float* cov_src_dev, int64_t ld_src,
                    const int32_t* idx_dev, float* mean_dev, float* cov_dev, int64_t ld, int64_t n,
                    const double u[GSE_NU], double dt, uint64_t seed, uint64_t step,
                    int64_t index0, const float* noise_dev, int64_t ld_noise, void* stream);

/* gse_gsf_predict / gse_gsf_moments of a sharded population: component idx[i] (GLOBAL ancestor row) is read out of the
 * owning shard's (20, ld) state (5 mean rows + 15 covariance rows) through peer memory. */
int gse_gsf_predict_sharded(gse_ctx* ctx, const gse_shards* shards, const int32_t* idx_dev, float* mean_dev,
                            float* cov_dev, int64_t ld, int64_t n, const double u[GSE_NU], double dt, uint64_t seed,
                            uint64_t step, int64_t index0, const float* noise_dev, int64_t ld_noise, void* stream);
int gse_gsf_moments_sharded(gse_ctx* ctx, const gse_shards* shards, const int32_t* idx_dev, int64_t n,
                            const float* loglik_dev, const double* base_dev, const double* stats_dev, double* out_dev,
                            void* stream);

/* update (gs_ukf.py:105-149 / :369-407): sigma points, g, P_xy, P_yy, K, mean/cov update,
 * loglik_i = loglik_in_i + log pdf_meas(z - g(mean_i)) (loglik_in_dev NULL = zeros);
 * stats_dev[0] = max loglik, stats_dev[1] = sum exp(loglik - max). */
int gse_gsf_update(gse_ctx* ctx, float* mean_dev, float* cov_dev, int64_t ld, int64_t n,
                   const float* loglik_in_dev, float* loglik_dev, const double u[GSE_NU],
                   const double z[GSE_NY], double* stats_dev, void* stream);

/* sigma points (gs_ukf.py:69-80 / :332-346): out SoA (55, ld_out), row s*5 + j. */
int gse_gsf_sigma_points(gse_ctx* ctx, const float* mean_dev, const float* cov_dev, int64_t ld,
                         int64_t n, float* out_dev, int64_t ld_out, void* stream);

/* point_estimate / point_covariance (gs_ukf.py:173-183 / :438-449): as gse_pf_moments on the
 * means, plus out_dev[26..40] = sum_i w_i P_i (lower triangle).  41 doubles. */
int gse_gsf_moments(gse_ctx* ctx, const float* mean_dev, const float* cov_dev, int64_t ld,
                    int64_t n, const int32_t* idx_dev, const float* loglik_dev, const double* base_dev,
                    const double* stats_dev, double* out_dev, void* stream);

/* Debugging aid: with GSE_FUSED_TRACE set in the environment when the context is created, gse_resample_fused records
 * four %globaltimer stamps per CTA (start, after its sum, after the grid-wide aggregate exchange, after its fill).
 * Copies up to max_words uint64 words to host_out (synchronises); returns the number copied, 0 when tracing is off. */
int64_t gse_ctx_read_trace(gse_ctx* ctx, uint64_t* host_out, int64_t max_words);

/* ---- introspection ------------------------------------------------------------------------------ */

/* kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t gse_launch_count(const gse_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* GSE_H_ */
