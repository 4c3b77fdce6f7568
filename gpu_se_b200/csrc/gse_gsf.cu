// Gaussian-sum unscented Kalman filter kernels (G1 predict, G2 update, sigma points).
// One thread owns one Gaussian component; every SoA plane (5 mean rows + 15 covariance rows) is
// read and written coalesced across the threads of a warp.  The 5x5 Cholesky, the 11 sigma points
// and the weighted moments live in registers; nothing of size N x 11 x 5 is materialised
// (the reference materialises it three times per step, gs_ukf.py:342-346,363,386).
#include "gse_common.cuh"

#define GSF_THREADS 128


// sigma weights (gs_ukf.py:66-67), float32 as the reference stores them
#define W_SIGMA_0 ((float)(1.0 / (1.0 + 5.0 / 4.0 * 5.0)))
#define W_SIGMA_I ((float)(1.0 / (2.0 * 5.0 + 8.0 / 5.0)))

// packed lower-triangular index
__device__ __forceinline__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// Cholesky of the packed symmetric matrix P (+ jitter on the diagonal); returns false when a pivot
// is not positive (numpy.linalg.cholesky raising LinAlgError, gs_ukf.py:72-75).
__device__ __forceinline__ bool cholesky5(const float P[15], float jitter, float L[15]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        float d = P[tri(j, j)] + jitter;
#pragma unroll
        for (int k = 0; k < j; ++k) d = fmaf(-L[tri(j, k)], L[tri(j, k)], d);
        ok = ok && (d > 0.0f);
        const float s = sqrtf(d);
        L[tri(j, j)] = s;
        const float inv = 1.0f / s;
#pragma unroll
        for (int i = j + 1; i < 5; ++i) {
            float v = P[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) v = fmaf(-L[tri(i, k)], L[tri(j, k)], v);
            L[tri(i, j)] = v * inv;
        }
    }
    return ok;
}

// float64 Cholesky of P + jitter * I (the reference's retry: `covariances + 1e-10 * numpy.eye(Nx)` is a float64 array,
// gs_ukf.py:75), result rounded to float32 as `sigmas[:, 1:Nx+1, :] += stds` rounds it (:77)
__device__ __noinline__ bool cholesky5_f64(const float P[15], double jitter, float L[15]) {
    double Ld[15];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        double d = (double)P[tri(j, j)] + jitter;
#pragma unroll
        for (int k = 0; k < j; ++k) d -= Ld[tri(j, k)] * Ld[tri(j, k)];
        ok = ok && (d > 0.0);
        const double sd = sqrt(d);
        Ld[tri(j, j)] = sd;
#pragma unroll
        for (int i = j + 1; i < 5; ++i) {
            double v = (double)P[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= Ld[tri(i, k)] * Ld[tri(j, k)];
            Ld[tri(i, j)] = v / sd;
        }
    }
#pragma unroll
    for (int t = 0; t < 15; ++t) L[t] = (float)Ld[t];
    return ok;
}

// numpy.linalg.cholesky(covariances), and on LinAlgError the retry with + 1e-10 I (gs_ukf.py:72-75).  The reference
// retries the WHOLE batch when any component fails; here every component decides for itself (a batch-wide retry
// cannot be reproduced by a population that is streamed and sharded).  A component that still fails -- the reference
// raises LinAlgError -- sets GSE_ERR_CHOLESKY in the context's error word.
__device__ __forceinline__ void cholesky5_retry(const float P[15], float L[15], unsigned int* err) {
    if (!cholesky5(P, 0.0f, L)) {
        if (!cholesky5_f64(P, 1e-10, L)) atomicOr(err, GSE_ERR_CHOLESKY);
    }
}

// sigma point s of (m, L): m, m + L[:, j], m - L[:, j]   (gs_ukf.py:76-78), float32 adds
__device__ __forceinline__ void sigma_point(const float m[5], const float L[15], int s, float out[5]) {
#pragma unroll
    for (int i = 0; i < 5; ++i) out[i] = m[i];
    if (s == 0) return;
    const int j = (s - 1) % 5;
    const float sign = (s <= 5) ? 1.0f : -1.0f;
#pragma unroll
    for (int i = 0; i < 5; ++i)
        if (i >= j) out[i] = __fadd_rn(m[i], sign * L[tri(i, j)]);
}

// ------------------------------------------------------------------------------------------------
// sigma points read-back (gs_ukf.py:69-80)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GSF_THREADS)
k_gsf_sigma_points(const float* __restrict__ mean, const float* __restrict__ cov, int64_t ld, int64_t n,
                   float* __restrict__ out, int64_t ldo, unsigned int* err) {
    const int64_t i = (int64_t)blockIdx.x * GSF_THREADS + threadIdx.x;
    if (i >= n) return;
    float m[5], P[15], L[15];
#pragma unroll
    for (int j = 0; j < 5; ++j) m[j] = mean[j * ld + i];
#pragma unroll
    for (int j = 0; j < 15; ++j) P[j] = cov[j * ld + i];
    cholesky5_retry(P, L, err);
#pragma unroll
    for (int s = 0; s < GSE_NSIGMA; ++s) {
        float x[5];
        sigma_point(m, L, s, x);
#pragma unroll
        for (int j = 0; j < 5; ++j) out[(s * 5 + j) * ldo + i] = x[j];
    }
}

extern "C" int gse_gsf_sigma_points(gse_ctx* ctx, const float* mean_dev, const float* cov_dev, int64_t ld,
                                    int64_t n, float* out_dev, int64_t ld_out, void* stream) {
    GSE_REQUIRE(ctx != NULL && mean_dev != NULL && cov_dev != NULL && out_dev != NULL, "NULL argument");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 1 && ld >= n && ld_out >= n, "n / ld out of range");
    k_gsf_sigma_points<<<(unsigned)gse_div_up(n, GSF_THREADS), GSF_THREADS, 0, (cudaStream_t)stream>>>(
        mean_dev, cov_dev, ld, n, out_dev, ld_out, ctx->err_dev);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

// ------------------------------------------------------------------------------------------------
// G1: predict (gs_ukf.py:82-103)
// ------------------------------------------------------------------------------------------------
// SHARDED: idx holds GLOBAL ancestor rows of a multi-GPU population; the (20, ld) state of the owning shard is read
// through peer memory (rows 0-4 the mean, 5-19 the covariance triangle)
template <bool DIAG, bool HOST_NOISE, bool SHARDED>
__global__ void __launch_bounds__(GSF_THREADS)
k_gsf_predict(const float* mean_src, const float* cov_src, int64_t lds, const int32_t* __restrict__ idx,
              const __grid_constant__ GatherShards shards_arg, float* mean, float* cov, int64_t ld, int64_t n, ModelInputs in_arg,
              const __grid_constant__ MixSampler5 sp, uint32_t k0, uint32_t k1, uint32_t step, int64_t index0,
              const float* __restrict__ noise, int64_t ldn, const gse_step_params* __restrict__ params,
              unsigned int* err) {
    __shared__ GatherShards s_shards;
    if (SHARDED) stage_shards(&s_shards, shards_arg);
    const GatherShards& shards = SHARDED ? s_shards : shards_arg;
    const int64_t i = (int64_t)blockIdx.x * GSF_THREADS + threadIdx.x;
    if (i >= n) return;
    const ModelInputs in = model_inputs(in_arg, params, 1);
    if (params) step = (uint32_t)params->step;
    float m[5], P[15], L[15];
    if (SHARDED) {
        int64_t lsrc;
        const float* src = shard_row(shards, idx[i], lsrc);
#pragma unroll
        for (int j = 0; j < 5; ++j) m[j] = src[j * lsrc];
#pragma unroll
        for (int j = 0; j < 15; ++j) P[j] = src[(5 + j) * lsrc];
    } else {
        const int64_t is = idx ? (int64_t)idx[i] : i;      // pending resample: read through the ancestor index
#pragma unroll
        for (int j = 0; j < 5; ++j) m[j] = mean_src[j * lds + is];
#pragma unroll
        for (int j = 0; j < 15; ++j) P[j] = cov_src[j * lds + is];
    }
    cholesky5_retry(P, L, err);

    float sg[GSE_NSIGMA][5];
    double msum[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int s = 0; s < GSE_NSIGMA; ++s) {
        float x[5], d[5], e[5];
        sigma_point(m, L, s, x);
        bioreactor_increment(x, in, d);                              // :95-97
        if (HOST_NOISE) {
#pragma unroll
            for (int j = 0; j < 5; ++j) e[j] = noise[(s * 5 + j) * ldn + i];
        } else {
            draw_mixture5<DIAG, 0>(sp, (uint64_t)(index0 + i), step, (uint32_t)s, k0, k1, e);   // :99
        }
        const double w = (double)(s == 0 ? W_SIGMA_0 : W_SIGMA_I);
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            sg[s][j] = __fadd_rn(__fadd_rn(x[j], d[j]), e[j]);
            msum[j] = fma(w, (double)sg[s][j], msum[j]);
        }
    }
    // numpy.average divides by the sum of the weights (:101)
    // (div.rn.f64 issues at 1 lane/clk/SM on B200 -- tools/ubench_xu.cu -- so the constant divisor is inverted
    // at compile time; the quotient differs from numpy's by at most 1 ulp of float64 before the float32 store)
    constexpr double inv_wsum = 1.0 / ((double)W_SIGMA_0 + 10.0 * (double)W_SIGMA_I);
    float mn[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) mn[j] = (float)(msum[j] * inv_wsum);
    float C[15];
#pragma unroll
    for (int j = 0; j < 15; ++j) C[j] = 0.0f;
#pragma unroll
    for (int s = 0; s < GSE_NSIGMA; ++s) {
        const float w = (s == 0) ? W_SIGMA_0 : W_SIGMA_I;
        float dv[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) dv[j] = __fsub_rn(sg[s][j], mn[j]);   // sigmas -= means  (:102)
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            const float wa = w * dv[a];
#pragma unroll
            for (int b = 0; b <= a; ++b) C[tri(a, b)] = fmaf(wa, dv[b], C[tri(a, b)]);   // :103
        }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) mean[j * ld + i] = mn[j];
#pragma unroll
    for (int j = 0; j < 15; ++j) cov[j * ld + i] = C[j];
}

static int launch_gsf_predict(gse_ctx* ctx, const float* mean_src_dev, const float* cov_src_dev, int64_t ld_src,
                              const int32_t* idx_dev, const GatherShards* shards, float* mean_dev, float* cov_dev,
                              int64_t ld, int64_t n, const double u[GSE_NU], double dt, uint64_t seed, uint64_t step,
                              int64_t index0, const float* noise_dev, int64_t ld_noise, void* stream) {
    GSE_REQUIRE(ctx != NULL && u != NULL && mean_dev != NULL && cov_dev != NULL, "NULL argument");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 1 && n <= ctx->n_max && ld >= n, "n / ld out of range");
    GSE_REQUIRE(noise_dev == NULL || ld_noise >= n, "ld_noise too small");
    ModelInputs in;
    in.feed = (float)(u[0] * (5000.0 / 180.0));
    in.f_out = (float)(u[0] + u[1]);
    in.dt = (float)dt;
    const unsigned blocks = (unsigned)gse_div_up(n, GSF_THREADS);
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    GatherShards none;
    memset(&none, 0, sizeof(none));
    const GatherShards& sh = shards ? *shards : none;
#define LAUNCH_G1(DIAG, HOST, SH)                                                                                   \
    k_gsf_predict<DIAG, HOST, SH><<<blocks, GSF_THREADS, 0, s>>>(mean_src_dev, cov_src_dev, ld_src, idx_dev, sh, mean_dev, \
                                                                 cov_dev, ld, n, in, ctx->state_sampler, k0, k1,     \
                                                                 (uint32_t)step, index0, noise_dev, ld_noise,       \
                                                                 ctx->step_params, ctx->err_dev)
    if (shards) {
        if (noise_dev) LAUNCH_G1(true, true, true);
        else if (ctx->state_sampler.diag) LAUNCH_G1(true, false, true);
        else LAUNCH_G1(false, false, true);
    } else {
        if (noise_dev) LAUNCH_G1(true, true, false);
        else if (ctx->state_sampler.diag) LAUNCH_G1(true, false, false);
        else LAUNCH_G1(false, false, false);
    }
#undef LAUNCH_G1
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

extern "C" int gse_gsf_predict(gse_ctx* ctx, const float* mean_src_dev, const float* cov_src_dev, int64_t ld_src,
                               const int32_t* idx_dev, float* mean_dev, float* cov_dev, int64_t ld, int64_t n,
                               const double u[GSE_NU], double dt, uint64_t seed, uint64_t step, int64_t index0,
                               const float* noise_dev, int64_t ld_noise, void* stream) {
    GSE_REQUIRE(mean_src_dev != NULL && cov_src_dev != NULL, "NULL argument");
    GSE_REQUIRE(idx_dev == NULL || mean_src_dev != mean_dev, "a gathering predict cannot run in place");
    GSE_REQUIRE(idx_dev != NULL || ld_src >= n, "ld_src too small");
    return launch_gsf_predict(ctx, mean_src_dev, cov_src_dev, ld_src, idx_dev, NULL, mean_dev, cov_dev, ld, n, u, dt, seed,
                              step, index0, noise_dev, ld_noise, stream);
}

extern "C" int gse_gsf_predict_sharded(gse_ctx* ctx, const gse_shards* shards, const int32_t* idx_dev, float* mean_dev,
                                       float* cov_dev, int64_t ld, int64_t n, const double u[GSE_NU], double dt,
                                       uint64_t seed, uint64_t step, int64_t index0, const float* noise_dev,
                                       int64_t ld_noise, void* stream) {
    GSE_REQUIRE(idx_dev != NULL, "idx is NULL");
    GatherShards g;
    int rc = gse_build_gather_shards(shards, mean_dev, &g);
    if (rc) return rc;
    return launch_gsf_predict(ctx, NULL, NULL, 0, idx_dev, &g, mean_dev, cov_dev, ld, n, u, dt, seed, step, index0,
                              noise_dev, ld_noise, stream);
}

// ------------------------------------------------------------------------------------------------
// G2: update (gs_ukf.py:105-149).  The reference evaluates this stage in float64 (etas is a
// float64 array, :118), so the Kalman algebra here is float64 too; storage stays float32.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GSF_THREADS)
k_gsf_update(float* __restrict__ mean, float* __restrict__ cov, int64_t ld, int64_t n, const float* loglik_in,
             float* loglik, double z0, double z1, const __grid_constant__ MixDensity2 md, float* block_max, float* block_sum,
             unsigned int* ticket, double* stats, const gse_step_params* __restrict__ params, unsigned int* err) {
    if (params) { z0 = params->z[0]; z1 = params->z[1]; }
    const int64_t i = (int64_t)blockIdx.x * GSF_THREADS + threadIdx.x;
    float vals[1] = {0.0f};
    bool valid[1] = {false};
    if (i < n) {
        float m[5], P[15], L[15];
#pragma unroll
        for (int j = 0; j < 5; ++j) m[j] = mean[j * ld + i];
#pragma unroll
        for (int j = 0; j < 15; ++j) P[j] = cov[j * ld + i];
        cholesky5_retry(P, L, err);
        constexpr double inv_wsum = 1.0 / ((double)W_SIGMA_0 + 10.0 * (double)W_SIGMA_I);
        // pass 1: eta mean (:126)
        double eta[GSE_NSIGMA][2];
        double em0 = 0.0, em1 = 0.0;
#pragma unroll
        for (int s = 0; s < GSE_NSIGMA; ++s) {
            float x[5];
            sigma_point(m, L, s, x);
            eta[s][0] = (double)output_glucose(x[0]);                 // :118-123
            eta[s][1] = (double)output_fa(x[2]);
            const double w = (double)(s == 0 ? W_SIGMA_0 : W_SIGMA_I);
            em0 = fma(w, eta[s][0], em0);
            em1 = fma(w, eta[s][1], em1);
        }
        em0 *= inv_wsum;
        em1 *= inv_wsum;
        // pass 2: P_xy (5x2), P_yy (2x2)  (:127-131)
        double pxy[5][2], pyy00 = 0.0, pyy01 = 0.0, pyy11 = 0.0;
#pragma unroll
        for (int a = 0; a < 5; ++a) { pxy[a][0] = 0.0; pxy[a][1] = 0.0; }
#pragma unroll
        for (int s = 0; s < GSE_NSIGMA; ++s) {
            float x[5];
            sigma_point(m, L, s, x);
            const double w = (double)(s == 0 ? W_SIGMA_0 : W_SIGMA_I);
            const double d0 = eta[s][0] - em0, d1 = eta[s][1] - em1;
            pyy00 = fma(w * d0, d0, pyy00);
            pyy01 = fma(w * d0, d1, pyy01);
            pyy11 = fma(w * d1, d1, pyy11);
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                const double ds = (double)__fsub_rn(x[a], m[a]);        // sigmas -= means  (:127)
                pxy[a][0] = fma(w * ds, d0, pxy[a][0]);
                pxy[a][1] = fma(w * ds, d1, pxy[a][1]);
            }
        }
        // K = P_xy pinv(P_yy) (:132-133).  P_yy is symmetric positive semi-definite 2x2: closed-form inverse; when its
        // smaller eigenvalue is below numpy.linalg.pinv's cut-off (1e-15 of the larger) the pseudo-inverse of the
        // rank-one matrix, P_yy / trace^2 (zero for the zero matrix).  Anything else -- NaN, negative determinant: not
        // a covariance -- is flagged in the error word.
        const double det = pyy00 * pyy11 - pyy01 * pyy01;
        const double tr = pyy00 + pyy11;
        double i00, i01, i11;
        if (det > 1e-15 * tr * tr) {                                    // lambda_min / lambda_max ~ det / tr^2
            const double inv_det = 1.0 / det;                           // one division, three products
            i00 = pyy11 * inv_det; i01 = -pyy01 * inv_det; i11 = pyy00 * inv_det;
        } else {
            if (!(det >= -1e-12 * tr * tr) || !(tr >= 0.0)) atomicOr(err, GSE_ERR_SINGULAR_PYY);
            const double s2 = tr > 0.0 ? 1.0 / (tr * tr) : 0.0;
            i00 = pyy00 * s2; i01 = pyy01 * s2; i11 = pyy11 * s2;
        }
        double K[5][2];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            K[a][0] = pxy[a][0] * i00 + pxy[a][1] * i01;
            K[a][1] = pxy[a][0] * i01 + pxy[a][1] * i11;
        }
        const double e0 = z0 - em0, e1 = z1 - em1;                      // :136
        float mn[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) mn[a] = (float)((double)m[a] + (K[a][0] * e0 + K[a][1] * e1));   // :137
        // P -= K P_yy K'  (:139)
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            const double t0 = K[a][0] * pyy00 + K[a][1] * pyy01;
            const double t1 = K[a][0] * pyy01 + K[a][1] * pyy11;
#pragma unroll
            for (int b = 0; b <= a; ++b)
                P[tri(a, b)] = (float)((double)P[tri(a, b)] - (t0 * K[b][0] + t1 * K[b][1]));
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) mean[j * ld + i] = mn[j];
#pragma unroll
        for (int j = 0; j < 15; ++j) cov[j * ld + i] = P[j];
        // global update: weights *= pdf(z - g(mean))  (:141-149)
        const double g0 = z0 - (double)output_glucose(mn[0]);
        const double g1 = z1 - (double)output_fa(mn[2]);
        vals[0] = (float)((loglik_in ? (double)loglik_in[i] : 0.0) + meas_logpdf(md, g0, g1));
        valid[0] = true;
        loglik[i] = vals[0];
    }
    block_max_sumexp_finalize<GSF_THREADS, 1>(vals, valid, block_max, block_sum, ticket, stats);
}

extern "C" int gse_gsf_update(gse_ctx* ctx, float* mean_dev, float* cov_dev, int64_t ld, int64_t n,
                              const float* loglik_in_dev, float* loglik_dev, const double u[GSE_NU], const double z[GSE_NY],
                              double* stats_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && z != NULL && mean_dev != NULL && cov_dev != NULL && loglik_dev != NULL && stats_dev != NULL, "NULL argument");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 1 && n <= ctx->n_max && ld >= n, "n / ld out of range");
    (void)u;
    const unsigned blocks = (unsigned)gse_div_up(n, GSF_THREADS);
    GSE_REQUIRE((int64_t)blocks <= ctx->max_blocks, "workspace too small");
    k_gsf_update<<<blocks, GSF_THREADS, 0, (cudaStream_t)stream>>>(mean_dev, cov_dev, ld, n, loglik_in_dev, loglik_dev, z[0], z[1],
                                                                   ctx->meas_density, ctx->block_max, ctx->block_sum,
                                                                   ctx->ticket, stats_dev, ctx->step_params, ctx->err_dev);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}
