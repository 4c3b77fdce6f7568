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
template <bool DIAG, bool HOST_NOISE, bool SHARDED, int ND, int MINB>
__global__ void __launch_bounds__(GSF_THREADS, MINB)
k_gsf_predict(const float* mean_src, const float* cov_src, int64_t lds, const int32_t* __restrict__ idx,
              const __grid_constant__ GatherShards shards_arg, float* mean, float* cov, int64_t ld, int64_t n, ModelInputs in_arg,
              const __grid_constant__ MixSampler5 sp, uint32_t k0, uint32_t k1, uint32_t step, int64_t index0,
              const float* __restrict__ noise, int64_t ldn, const gse_step_params* __restrict__ params,
              unsigned int* err) {
    const GatherShards& shards = shards_arg;
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
    SigmaNoise stream((uint64_t)(index0 + i), step, k0, k1);
#pragma unroll
    for (int s = 0; s < GSE_NSIGMA; ++s) {
        float x[5], d[5], e[5];
        sigma_point(m, L, s, x);
        bioreactor_increment(x, in, d);                              // :95-97
        if (HOST_NOISE) {
#pragma unroll
            for (int j = 0; j < 5; ++j) e[j] = noise[(s * 5 + j) * ldn + i];
        } else {
            stream.template draw<DIAG, ND>(s, sp, e);                // an independent draw per sigma point  (:99)
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
#define LAUNCH_G1M(DIAG, HOST, SH, ND, MB)                                                                          \
    k_gsf_predict<DIAG, HOST, SH, ND, MB><<<blocks, GSF_THREADS, 0, s>>>(mean_src_dev, cov_src_dev, ld_src, idx_dev, sh, \
                                                                         mean_dev, cov_dev, ld, n, in, ctx->state_sampler, \
                                                                         k0, k1, (uint32_t)step, index0, noise_dev,      \
                                                                         ld_noise, ctx->step_params, ctx->err_dev)
#define LAUNCH_G1(DIAG, HOST, SH) LAUNCH_G1M(DIAG, HOST, SH, 0, 4)
    if (shards) {
        if (noise_dev) LAUNCH_G1(true, true, true);
        else if (ctx->state_sampler.diag) LAUNCH_G1(true, false, true);
        else LAUNCH_G1(false, false, true);
    } else {
        if (noise_dev) LAUNCH_G1(true, true, false);
        else if (ctx->state_sampler.diag && ctx->state_sampler.nd == 2) {       // the benchmark's noise; GSE_GSF_MINB tunes
            // measured at 2^20 components (us): 3 CTAs/SM 113, 4: 110, 5: 109 (96 registers), 6: 117 (80, spills)
            if (ctx->gsf_minb == 4) LAUNCH_G1M(true, false, false, 2, 4);
            else if (ctx->gsf_minb == 6) LAUNCH_G1M(true, false, false, 2, 6);
            else if (ctx->gsf_minb == 3) LAUNCH_G1M(true, false, false, 2, 3);
            else LAUNCH_G1M(true, false, false, 2, 5);
        } else if (ctx->state_sampler.diag) LAUNCH_G1(true, false, false);
        else LAUNCH_G1(false, false, false);
    }
#undef LAUNCH_G1M
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
template <int MINB>
__global__ void __launch_bounds__(GSF_THREADS, MINB)
k_gsf_update(float* __restrict__ mean, float* __restrict__ cov, int64_t ld, int64_t n, const float* loglik_in,
             float* loglik, double z0, double z1, const __grid_constant__ MixDensity2 md, float* block_max, float* block_sum,
             unsigned int* ticket, double* stats, const gse_step_params* __restrict__ params, unsigned int* err) {
    if (params) { z0 = params->z[0]; z1 = params->z[1]; }
    // persistent grid: a CTA walks over tiles of GSF_THREADS components and merges its (max, sum exp) ONCE at the end.
    // (One CTA per tile spent a fifth of its few microseconds in the five barriers, the fence and the ticket's round
    // trip of block_merge_max_sumexp: 20 % of the stall samples at 2^20 components.)
    MaxSumExp acc;
    for (int64_t i = (int64_t)blockIdx.x * GSF_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * GSF_THREADS) {
        float m[5], P[15], L[15];
#pragma unroll
        for (int j = 0; j < 5; ++j) m[j] = mean[j * ld + i];
#pragma unroll
        for (int j = 0; j < 15; ++j) P[j] = cov[j * ld + i];
        cholesky5_retry(P, L, err);
        constexpr double inv_wsum = 1.0 / ((double)W_SIGMA_0 + 10.0 * (double)W_SIGMA_I);
        constexpr double w0 = (double)W_SIGMA_0, wi = (double)W_SIGMA_I;
        // The sigma points are m and m +- L[:, j] (float32 adds, :76-78) and g reads rows 0 and 2 only (:251-253), so
        // most of the 11 x (5 + 2) values the reference forms coincide or vanish:
        //   eta[0] differs from g(m)[0] only for column 0 (L is lower triangular), eta[1] only for columns 0, 1, 2;
        //   sigma - m is zero above the diagonal: 15 non-zero deviations per sign instead of 55.
        // Every value that IS formed is the reference's (float32 sigma point, float32 output, float64 from there on).
        const double g0 = (double)output_glucose(m[0]), g1 = (double)output_fa(m[2]);                 // eta of sigma point 0
        const double g0p = (double)output_glucose(__fadd_rn(m[0], L[tri(0, 0)]));
        const double g0m = (double)output_glucose(__fadd_rn(m[0], -L[tri(0, 0)]));
        double g1p[3], g1m[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            g1p[j] = (double)output_fa(__fadd_rn(m[2], L[tri(2, j)]));
            g1m[j] = (double)output_fa(__fadd_rn(m[2], -L[tri(2, j)]));
        }
        // eta mean (:126): numpy.average over the 11 points
        double em0 = fma(wi, (g0p + g0m) + 8.0 * g0, w0 * g0);
        double em1 = fma(wi, ((g1p[0] + g1m[0]) + (g1p[1] + g1m[1])) + ((g1p[2] + g1m[2]) + 4.0 * g1), w0 * g1);
        em0 *= inv_wsum;
        em1 *= inv_wsum;
        // deviations of the outputs (:128)
        const double d0 = g0 - em0, d0p = g0p - em0, d0m = g0m - em0;
        const double d1 = g1 - em1;
        double d1p[3], d1m[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { d1p[j] = g1p[j] - em1; d1m[j] = g1m[j] - em1; }
        // P_yy (:130-131)
        double pyy00 = fma(wi, fma(d0p, d0p, d0m * d0m), (w0 + 8.0 * wi) * (d0 * d0));
        double pyy11 = (w0 + 4.0 * wi) * (d1 * d1);
        double pyy01 = fma(wi, fma(d0p, d1p[0], d0m * d1m[0]), (w0 + 4.0 * wi) * (d0 * d1));
#pragma unroll
        for (int j = 0; j < 3; ++j) pyy11 = fma(wi, fma(d1p[j], d1p[j], d1m[j] * d1m[j]), pyy11);
        pyy01 = fma(wi * d0, (d1p[1] + d1m[1]) + (d1p[2] + d1m[2]), pyy01);
        // P_xy (:127-129): row a collects the columns j <= a
        double pxy[5][2];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int j = 0; j <= a; ++j) {
                const double dsp = (double)__fsub_rn(__fadd_rn(m[a], L[tri(a, j)]), m[a]);      // sigmas -= means  (:127)
                const double dsm = (double)__fsub_rn(__fadd_rn(m[a], -L[tri(a, j)]), m[a]);
                const double e0p = j == 0 ? d0p : d0, e0m = j == 0 ? d0m : d0;
                const double e1p = j < 3 ? d1p[j < 3 ? j : 0] : d1, e1m = j < 3 ? d1m[j < 3 ? j : 0] : d1;
                s0 = fma(dsp, e0p, fma(dsm, e0m, s0));
                s1 = fma(dsp, e1p, fma(dsm, e1m, s1));
            }
            pxy[a][0] = wi * s0;
            pxy[a][1] = wi * s1;
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
        const double ge0 = z0 - (double)output_glucose(mn[0]);
        const double ge1 = z1 - (double)output_fa(mn[2]);
        const float vals[1] = {(float)((loglik_in ? (double)loglik_in[i] : 0.0) + meas_logpdf(md, ge0, ge1))};
        const bool valid[1] = {true};
        loglik[i] = vals[0];
        acc.add<1>(vals, valid);
    }
    block_merge_max_sumexp<GSF_THREADS>(acc.m, acc.s, block_max, block_sum, ticket, stats);
}

extern "C" int gse_gsf_update(gse_ctx* ctx, float* mean_dev, float* cov_dev, int64_t ld, int64_t n,
                              const float* loglik_in_dev, float* loglik_dev, const double u[GSE_NU], const double z[GSE_NY],
                              double* stats_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && z != NULL && mean_dev != NULL && cov_dev != NULL && loglik_dev != NULL && stats_dev != NULL, "NULL argument");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 1 && n <= ctx->n_max && ld >= n, "n / ld out of range");
    (void)u;
    const int64_t tiles = gse_div_up(n, GSF_THREADS);
    GSE_REQUIRE(tiles <= ctx->max_blocks, "workspace too small");
    const int minb = (ctx->gsf_minb >= 3 && ctx->gsf_minb <= 6) ? ctx->gsf_minb : 4;
    const int64_t resident = (int64_t)ctx->num_sms * minb * ctx->gsf_update_waves;
    const unsigned blocks = (unsigned)(tiles < resident ? tiles : resident);
#define LAUNCH_G2(MB)                                                                                              \
    k_gsf_update<MB><<<blocks, GSF_THREADS, 0, (cudaStream_t)stream>>>(mean_dev, cov_dev, ld, n, loglik_in_dev, loglik_dev, \
                                                                       z[0], z[1], ctx->meas_density, ctx->block_max,      \
                                                                       ctx->block_sum, ctx->ticket, stats_dev,             \
                                                                       ctx->step_params, ctx->err_dev)
    // measured at 2^20 components (us), persistent grid of one wave: 3 CTAs/SM 69, 4: 64 (128 registers, no spills), 5: 68,
    // 6: 74 (80 registers, spills); two waves: +1.5.  (One CTA per tile, round 1: 95 / 82 / 76 / 71.)
    if (minb == 5) LAUNCH_G2(5);
    else if (minb == 6) LAUNCH_G2(6);
    else if (minb == 3) LAUNCH_G2(3);
    else LAUNCH_G2(4);
#undef LAUNCH_G2
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}
