// Particle-filter kernels: initial draw (K0), predict (K1), update (K2), weighted moments (K6),
// and the stand-alone mixture pdf / weight read-back helpers.  One thread owns 4 consecutive rows
// so that every SoA column moves as 128-bit coalesced vectors.
#include "gse_common.cuh"
#include "gse_mailbox.cuh"

#define PF_THREADS 256
#ifndef PREDICT_MINB
#define PREDICT_MINB 5       // CTAs per SM the predict kernel is compiled for (48 registers)
#endif
#define ROWS_PER_THREAD 4

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

#define CHECK_SOA(ptr, ld, n)                                                                   \
    GSE_REQUIRE((ptr) != NULL && aligned16(ptr), "SoA base pointer must be non-NULL and 16-byte aligned"); \
    GSE_REQUIRE((ld) % 4 == 0 && (ld) >= ((n) + 3) / 4 * 4, "ld must be a multiple of 4 and >= round_up(n, 4)")

// ------------------------------------------------------------------------------------------------
// K0: mixture draw into SoA columns
// ------------------------------------------------------------------------------------------------
template <bool DIAG>
__global__ void __launch_bounds__(PF_THREADS)
k_mixture_draw(float* __restrict__ x, int64_t ld, int64_t n, int ncols, const __grid_constant__ MixSampler5 sp,
               uint32_t k0, uint32_t k1, uint32_t step, int64_t index0) {
    const int64_t g = (int64_t)blockIdx.x * PF_THREADS + threadIdx.x;
    const int64_t row0 = g * ROWS_PER_THREAD;
    if (row0 >= n) return;
    float v[5][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float o[5];
        draw_mixture5<DIAG, 0>(sp, (uint64_t)(index0 + row0 + r), step, 0u, k0, k1, o);
#pragma unroll
        for (int j = 0; j < 5; ++j) v[j][r] = o[j];
    }
#pragma unroll
    for (int j = 0; j < 5; ++j)
        if (j < ncols) st_stream4(x + j * ld + row0, make_float4(v[j][0], v[j][1], v[j][2], v[j][3]));
}

extern "C" int gse_mixture_draw(gse_ctx* ctx, const gse_mixture* mix, float* x_dev, int64_t ld, int64_t n,
                                uint64_t seed, uint64_t step, int64_t index0, void* stream) {
    GSE_REQUIRE(ctx != NULL, "ctx is NULL");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return GSE_OK;
    CHECK_SOA(x_dev, ld, n);
    MixSampler5 sp;
    int nx = 0;
    int rc = gse_build_sampler(mix, &sp, &nx);
    if (rc) return rc;
    const int64_t groups = gse_div_up(n, ROWS_PER_THREAD);
    const unsigned blocks = (unsigned)gse_div_up(groups, PF_THREADS);
    cudaStream_t s = (cudaStream_t)stream;
    if (sp.diag)
        k_mixture_draw<true><<<blocks, PF_THREADS, 0, s>>>(x_dev, ld, n, nx, sp, (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)step, index0);
    else
        k_mixture_draw<false><<<blocks, PF_THREADS, 0, s>>>(x_dev, ld, n, nx, sp, (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)step, index0);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

// ------------------------------------------------------------------------------------------------
// K1: predict.  x += f(x, u, dt) (n_sub Euler sub-steps), then x += noise (particle.py:65-67).
// ------------------------------------------------------------------------------------------------
// GMODE: 0 rows in place, 1 rows through a local ancestor index, 2 rows through a GLOBAL ancestor
// index into the shards of a multi-GPU population (peer memory)
// ALIGNED: global row index0 + row0 of every thread is a multiple of four (one grouped draw serves the thread's rows);
// always the case except for a shard that starts off a multiple of four
// FUSE_ND >= 0: predict() immediately followed by update() (the filter loop of the reference, particle.py:265-294) as ONE
// pass: the new rows are still in registers when their likelihood is evaluated, so the update neither re-reads the
// two measured columns nor costs a launch.  The arithmetic is k_pf_update's, on the same float32 values it would read
// back.  FUSE_ND = number of components of the measurement mixture (0: run-time loop).
struct FusedUpdate {
    float z0h, z0l, z1h, z1l;       // z split into float32 head and tail (gse_pf_update)
    const float* loglik_in;         // accumulated log-likelihood, NULL: all zero (fresh resample)
    float* loglik;
    float* block_max;
    float* block_sum;
    unsigned int* ticket;
    double* stats;
};

template <bool DIAG, bool HOST_NOISE, bool ONE_STEP, int ND, int GMODE, bool ALIGNED, int MINB, int FUSE_ND = -1>
__global__ void __launch_bounds__(PF_THREADS, MINB)
k_pf_predict(const float* xs, int64_t lds, const int32_t* __restrict__ idx, const __grid_constant__ GatherShards shards_arg,
             float* xd, int64_t ldd, int64_t n,
             ModelInputs in_arg, int n_sub, const __grid_constant__ MixSampler5 sp, uint32_t k0, uint32_t k1,
             uint32_t step, int64_t index0, const float* __restrict__ noise, int64_t ldn,
             const gse_step_params* __restrict__ params, const __grid_constant__ FusedUpdate fu,
             const __grid_constant__ MixDensity2f md) {
    constexpr bool FUSE = FUSE_ND >= 0;
    const GatherShards& shards = shards_arg;
    // Sharded: the rows whose ancestors live on another GPU sit at the two ENDS of the shard (the ancestor index is
    // non-decreasing).  In launch order the high end runs last, and the kernel's tail then waits out NVLink round
    // trips; with ends_first the blocks alternate between the two ends and finish in the all-local middle.
    unsigned int vblock = blockIdx.x;
    if (GMODE == 2 && shards.ends_first) vblock = (vblock & 1u) ? gridDim.x - 1u - (vblock >> 1) : (vblock >> 1);
    const int64_t g = (int64_t)vblock * PF_THREADS + threadIdx.x;
    // (fused: every thread reaches the block reduction; a thread past the end recomputes the last group and stores nothing)
    const bool active = g * ROWS_PER_THREAD < n;
    if (!FUSE && !active) return;
    const int64_t row0 = active ? g * ROWS_PER_THREAD : ((n - 1) >> 2) * ROWS_PER_THREAD;
    const ModelInputs in = model_inputs(in_arg, params, ONE_STEP ? 1 : n_sub);
    if (params) step = (uint32_t)params->step;
    // process noise of the four rows: one grouped draw (six Philox calls) when the rows are a whole group of
    // the global index space -- always, unless a shard starts off a multiple of four
    float e4[4][5];
    if (!HOST_NOISE) {
        const uint64_t r0g = (uint64_t)(index0 + row0);
        if (ALIGNED) {
            draw_mixture5_x4<DIAG, ND>(sp, r0g >> 2, step, k0, k1, e4);
        } else {
            // (fully unrolled: a loop over r would index e4 dynamically and push it into local memory)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float g4[4][5];
                draw_mixture5_x4<DIAG, ND>(sp, (r0g + r) >> 2, step, k0, k1, g4);
                const int l = (int)((r0g + r) & 3ull);
#pragma unroll
                for (int j = 0; j < 5; ++j) e4[r][j] = l == 0 ? g4[0][j] : (l == 1 ? g4[1][j] : (l == 2 ? g4[2][j] : g4[3][j]));
            }
        }
    }
    float v[5][4];
    if (GMODE == 2) {
        const int4 id4 = *reinterpret_cast<const int4*>(idx + row0);
        const int id[4] = {id4.x, (row0 + 1 < n) ? id4.y : id4.x, (row0 + 2 < n) ? id4.z : id4.x,
                           (row0 + 3 < n) ? id4.w : id4.x};
        // idx is non-decreasing, and nearly every ancestor is one of this rank's own rows (two int32 compares).  The home
        // shard is read WITHOUT a branch -- the very loads of GMODE 1, home_base indexed by the GLOBAL row, clamped to a
        // valid row for the few threads whose ancestors live elsewhere -- so that the compiler schedules them, like GMODE
        // 1's, after the Philox rounds that hide the latency of the index load.  (With the loads behind a branch on the
        // index every warp sat out that latency: +18 us at 2^24 rows.)  The other threads then fetch their rows through
        // the shard table, out of line.
        const bool home = id[0] >= shards.home_lo && id[3] < shards.home_hi;
        {
            const float* base = shards.home_base;
            const int64_t ld0 = shards.home_ld;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
#pragma unroll
                for (int r = 0; r < 4; ++r) v[j][r] = __ldg(base + j * ld0 + (home ? id[r] : shards.home_lo));
            }
        }
        if (!home) {
            // (fully unrolled: a loop over r would index id[] dynamically and park it in local memory for EVERY thread)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const ShardRef own = shard_ref_search(&shards, id[r]);
                const float* p = own.state + (id[r] - own.row0);
#pragma unroll
                for (int j = 0; j < 5; ++j) v[j][r] = p[j * own.ld];
            }
        }
    } else if (GMODE == 1) {
        // rows of the resampled population are read through the ancestor index (the pending
        // particles[sample_index] of the last resample, particle.py:102): idx is non-decreasing, so a
        // warp's reads fall into one short window of each column
        const int4 id4 = *reinterpret_cast<const int4*>(idx + row0);
        const int id[4] = {id4.x, (row0 + 1 < n) ? id4.y : id4.x, (row0 + 2 < n) ? id4.z : id4.x,
                           (row0 + 3 < n) ? id4.w : id4.x};
#pragma unroll
        for (int j = 0; j < 5; ++j) {
#pragma unroll
            for (int r = 0; r < 4; ++r) v[j][r] = __ldg(xs + j * lds + id[r]);
        }
    } else {
        float4 c[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) c[j] = ld_stream4(xs + j * lds + row0);
#pragma unroll
        for (int j = 0; j < 5; ++j) { v[j][0] = c[j].x; v[j][1] = c[j].y; v[j][2] = c[j].z; v[j][3] = c[j].w; }
    }
    float4 nz[5];
    if (HOST_NOISE) {
#pragma unroll
        for (int j = 0; j < 5; ++j) nz[j] = ld_stream4(noise + j * ldn + row0);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float xv[5] = {v[0][r], v[1][r], v[2][r], v[3][r], v[4][r]};
#pragma unroll 1
        for (int s = 0; s < (ONE_STEP ? 1 : n_sub); ++s) {
            float d[5];
            bioreactor_increment(xv, in, d);
#pragma unroll
            for (int j = 0; j < 5; ++j) xv[j] = __fadd_rn(xv[j], d[j]);    // particles[i] += f(...)  (:66)
        }
        float e[5];
        if (HOST_NOISE) {
#pragma unroll
            for (int j = 0; j < 5; ++j) e[j] = reinterpret_cast<const float*>(&nz[j])[r];
        } else {
#pragma unroll
            for (int j = 0; j < 5; ++j) e[j] = e4[r][j];
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) v[j][r] = __fadd_rn(xv[j], e[j]);      // particles += draw(N)   (:67)
    }
    if (!FUSE || active) {
#pragma unroll
        for (int j = 0; j < 5; ++j)
            st_stream4(xd + j * ldd + row0, make_float4(v[j][0], v[j][1], v[j][2], v[j][3]));
    }
    if (FUSE) {
        float z0h = fu.z0h, z0l = fu.z0l, z1h = fu.z1h, z1l = fu.z1l;
        if (params) {
            z0h = (float)params->z[0];
            z1h = (float)params->z[1];
            z0l = (float)(params->z[0] - (double)z0h);
            z1l = (float)(params->z[1] - (double)z1h);
        }
        float4 lw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fu.loglik_in) lw = ld_stream4(fu.loglik_in + row0);
        const float l[4] = {lw.x, lw.y, lw.z, lw.w};
        float vals[4];
        bool valid[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float e0 = __fadd_rn(__fsub_rn(z0h, output_glucose(v[0][r])), z0l);   // e = z - y   (:82)
            const float e1 = __fadd_rn(__fsub_rn(z1h, output_fa(v[2][r])), z1l);
            vals[r] = l[r] + meas_logpdf32<(FUSE ? FUSE_ND : 0)>(md, e0, e1);           // weights[i] *= pdf(e)  (:83)
            valid[r] = active && row0 + r < n;
        }
        if (active) st_stream4(fu.loglik + row0, make_float4(vals[0], vals[1], vals[2], vals[3]));
        // (max, sum exp) of the CTA's 1024 rows -> block_max / block_sum; k_merge_block_stats folds them into stats.
        // No ticket here: a CTA lives for a few microseconds, and waiting out an atomic's round trip at the end of
        // each (the way the persistent update kernel finds its last block) cost 25 % of the kernel.
        MaxSumExp acc;
        acc.add<4>(vals, valid);
        __shared__ float s_m[PF_THREADS / 32], s_s[PF_THREADS / 32];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const float wm = warp_max(acc.m);
        const float ws = warp_sum(acc.m > -INFINITY ? acc.s * fast_exp(acc.m - wm) : 0.0f);
        if (lane == 0) { s_m[wid] = wm; s_s[wid] = ws; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float bm = s_m[0];
#pragma unroll
            for (int w = 1; w < PF_THREADS / 32; ++w) bm = fmaxf(bm, s_m[w]);
            float bs = 0.0f;
#pragma unroll
            for (int w = 0; w < PF_THREADS / 32; ++w) bs += s_m[w] > -INFINITY ? s_s[w] * fast_exp(s_m[w] - bm) : 0.0f;
            fu.block_max[blockIdx.x] = bm;
            fu.block_sum[blockIdx.x] = bs;
        }
    }
}

// stats[0..1] = (M, S) from nblocks (max, sum exp) pairs: fixed order per thread, then a fixed tree, in float64
__global__ void __launch_bounds__(1024)
k_merge_block_stats(const float* __restrict__ block_max, const float* __restrict__ block_sum, unsigned int nblocks,
                    double* stats) {
    __shared__ float s_red[32];
    __shared__ double s_dred[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float gm = -INFINITY;
    for (unsigned int b = threadIdx.x; b < nblocks; b += 1024) gm = fmaxf(gm, block_max[b]);
    gm = warp_max(gm);
    if (lane == 0) s_red[wid] = gm;
    __syncthreads();
    float M = s_red[lane];
    M = warp_max(M);
    double acc = 0.0;
    for (unsigned int b = threadIdx.x; b < nblocks; b += 1024) {
        const float bmx = block_max[b];
        if (bmx > -INFINITY) acc += (double)block_sum[b] * (double)fast_exp(bmx - M);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_dred[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        double t = s_dred[lane];
        t = warp_sum(t);
        if (lane == 0) { stats[0] = (double)M; stats[1] = t; }
    }
}

// Can predict + update run as one pass?  Only the benchmark's specialisation is built fused: diagonal two-component state
// noise from the in-kernel Philox stream, one Euler step, a shard that starts at a multiple of four rows, two-component
// measurement mixture.
static bool can_fuse_update(const gse_ctx* ctx, int n_sub, int64_t index0) {
    return ctx->state_sampler.diag && ctx->state_sampler.nd == 2 && n_sub == 1 && (index0 & 3) == 0 &&
           ctx->meas_density32.nd == 2 && ctx->predict_minb == 4;
}

extern "C" int gse_pf_can_fuse_update(const gse_ctx* ctx, int n_sub, int64_t index0) {
    return (ctx != NULL && can_fuse_update(ctx, n_sub, index0)) ? 1 : 0;
}

static int launch_predict(gse_ctx* ctx, const float* x_src_dev, int64_t ld_src, const int32_t* idx_dev,
                          const GatherShards* shards, bool sharded, float* x_dst_dev, int64_t ld_dst, int64_t n,
                          const double u[GSE_NU], double dt, int n_sub, uint64_t seed, uint64_t step, int64_t index0,
                          const float* noise_dev, int64_t ld_noise, void* stream, const FusedUpdate* fuse = NULL) {
    GSE_REQUIRE(ctx != NULL && u != NULL, "ctx / u is NULL");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 0 && n <= ctx->n_max, "n out of range for this context");
    GSE_REQUIRE(n_sub >= 1, "n_sub must be >= 1");
    if (n == 0) return GSE_OK;
    CHECK_SOA(x_dst_dev, ld_dst, n);
    if (sharded) {
        GSE_REQUIRE(idx_dev != NULL && aligned16(idx_dev), "idx must be non-NULL and 16-byte aligned");
    } else if (idx_dev) {
        GSE_REQUIRE(x_src_dev != NULL && x_src_dev != x_dst_dev, "a gathering predict cannot run in place");
        GSE_REQUIRE(aligned16(idx_dev), "idx must be 16-byte aligned");
    } else {
        CHECK_SOA(x_src_dev, ld_src, n);
    }
    if (noise_dev) { CHECK_SOA(noise_dev, ld_noise, n); }
    ModelInputs in;
    in.feed = (float)(u[0] * (5000.0 / 180.0));
    in.f_out = (float)(u[0] + u[1]);
    in.dt = (float)(dt / n_sub);
    const int64_t groups = gse_div_up(n, ROWS_PER_THREAD);
    const unsigned blocks = (unsigned)gse_div_up(groups, PF_THREADS);
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    FusedUpdate fu_none;
    memset(&fu_none, 0, sizeof(fu_none));
    if (fuse) {
        GSE_REQUIRE(noise_dev == NULL && can_fuse_update(ctx, n_sub, index0), "this configuration has no fused predict + update");
        GSE_REQUIRE((int64_t)blocks <= ctx->max_blocks, "workspace too small");
#define LAUNCH_FUSED(GMODE)                                                                                      \
    k_pf_predict<true, false, true, 2, GMODE, true, 4, 2><<<blocks, PF_THREADS, 0, s>>>(                         \
        x_src_dev, ld_src, idx_dev, *shards, x_dst_dev, ld_dst, n, in, n_sub, ctx->state_sampler, k0, k1,        \
        (uint32_t)step, index0, noise_dev, ld_noise, ctx->step_params, *fuse, ctx->meas_density32)
        if (sharded) LAUNCH_FUSED(2);
        else if (idx_dev) LAUNCH_FUSED(1);
        else LAUNCH_FUSED(0);
#undef LAUNCH_FUSED
        GSE_CHECK_LAUNCH(ctx);
        k_merge_block_stats<<<1, 1024, 0, s>>>(fuse->block_max, fuse->block_sum, blocks, fuse->stats);
        GSE_CHECK_LAUNCH(ctx);
        return GSE_OK;
    }
#define LAUNCH_PREDICT_GAM(DIAG, HOST, ONE, ND, GMODE, AL, MB)                                                   \
    k_pf_predict<DIAG, HOST, ONE, ND, GMODE, AL, MB><<<blocks, PF_THREADS, 0, s>>>(                              \
        x_src_dev, ld_src, idx_dev, *shards, x_dst_dev, ld_dst, n, in, n_sub, ctx->state_sampler, k0, k1,        \
        (uint32_t)step, index0, noise_dev, ld_noise, ctx->step_params, fu_none, ctx->meas_density32)
// The benchmark's specialisation (diagonal two-component noise, one Euler step, aligned rows) runs at 4 CTAs per SM
// (64 registers): measured at 2^24 rows 132 us against 139 us at 5 CTAs (48 registers) and 143 us at 6 (40, spills);
// GSE_PREDICT_MINB=5 selects the 48-register build for comparison.  (Tying the gather addresses to the first Philox call
// so that the scheduler runs it under the ancestor-index load changed nothing: 132.1 us.)
#define LAUNCH_PREDICT_GA(DIAG, HOST, ONE, ND, GMODE, AL)                                                        \
    do {                                                                                                         \
        constexpr bool tuned = (ND) == 2 && (AL);                                                                \
        if (tuned && ctx->predict_minb == 4) LAUNCH_PREDICT_GAM(DIAG, HOST, ONE, ND, GMODE, AL, (tuned ? 4 : PREDICT_MINB)); \
        else LAUNCH_PREDICT_GAM(DIAG, HOST, ONE, ND, GMODE, AL, PREDICT_MINB);                                   \
    } while (0)
#define LAUNCH_PREDICT_G(DIAG, HOST, ONE, ND, GMODE)                                                             \
    do {                                                                                                         \
        if ((index0 & 3) == 0 || (HOST)) LAUNCH_PREDICT_GA(DIAG, HOST, ONE, ND, GMODE, true);                    \
        else LAUNCH_PREDICT_GA(DIAG, HOST, ONE, ND, GMODE, false);                                               \
    } while (0)
#define LAUNCH_PREDICT(DIAG, HOST, ONE, ND)                                                                      \
    do {                                                                                                         \
        if (sharded) LAUNCH_PREDICT_G(DIAG, HOST, ONE, ND, 2);                                                   \
        else if (idx_dev) LAUNCH_PREDICT_G(DIAG, HOST, ONE, ND, 1);                                              \
        else LAUNCH_PREDICT_G(DIAG, HOST, ONE, ND, 0);                                                           \
    } while (0)
    const bool one = (n_sub == 1);
    const int nd = ctx->state_sampler.nd;
    if (noise_dev) { if (one) LAUNCH_PREDICT(true, true, true, 0); else LAUNCH_PREDICT(true, true, false, 0); }
    else if (ctx->state_sampler.diag && one && nd == 2) LAUNCH_PREDICT(true, false, true, 2);     // the benchmark's noise
    else if (ctx->state_sampler.diag && one && nd == 1) LAUNCH_PREDICT(true, false, true, 1);
    else if (ctx->state_sampler.diag) { if (one) LAUNCH_PREDICT(true, false, true, 0); else LAUNCH_PREDICT(true, false, false, 0); }
    else { if (one) LAUNCH_PREDICT(false, false, true, 0); else LAUNCH_PREDICT(false, false, false, 0); }
#undef LAUNCH_PREDICT
#undef LAUNCH_PREDICT_G
#undef LAUNCH_PREDICT_GA
#undef LAUNCH_PREDICT_GAM
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

extern "C" int gse_pf_predict(gse_ctx* ctx, const float* x_src_dev, int64_t ld_src, const int32_t* idx_dev,
                              float* x_dst_dev, int64_t ld_dst, int64_t n, const double u[GSE_NU], double dt,
                              int n_sub, uint64_t seed, uint64_t step, int64_t index0, const float* noise_dev,
                              int64_t ld_noise, void* stream) {
    GatherShards none;
    memset(&none, 0, sizeof(none));
    static const bool force_table = getenv("GSE_DEBUG_PREDICT_TABLE") != NULL;     // A/B of the sharded read path on one GPU
    if (force_table && idx_dev != NULL) {
        none.nseg = 1;
        none.seg_row[0] = 0; none.seg_row[1] = n;
        none.state[0] = x_src_dev; none.ld[0] = ld_src;
        none.home_state = x_src_dev; none.home_ld = ld_src; none.home_row0 = 0; none.home_row1 = n;
        none.home_base = x_src_dev; none.home_lo = 0; none.home_hi = (int)n;
        return launch_predict(ctx, NULL, 0, idx_dev, &none, true, x_dst_dev, ld_dst, n, u, dt, n_sub, seed, step, index0,
                              noise_dev, ld_noise, stream);
    }
    return launch_predict(ctx, x_src_dev, ld_src, idx_dev, &none, false, x_dst_dev, ld_dst, n, u, dt, n_sub, seed, step,
                          index0, noise_dev, ld_noise, stream);
}

extern "C" int gse_pf_predict_sharded(gse_ctx* ctx, const gse_shards* shards, const int32_t* idx_dev, float* x_dst_dev,
                                      int64_t ld_dst, int64_t n, const double u[GSE_NU], double dt, int n_sub,
                                      uint64_t seed, uint64_t step, int64_t index0, const float* noise_dev,
                                      int64_t ld_noise, void* stream) {
    GatherShards g;
    int rc = gse_build_gather_shards(shards, x_dst_dev, &g);
    if (rc) return rc;
    return launch_predict(ctx, NULL, 0, idx_dev, &g, true, x_dst_dev, ld_dst, n, u, dt, n_sub, seed, step, index0,
                          noise_dev, ld_noise, stream);
}

static int fill_fused_update(gse_ctx* ctx, int64_t n, const double z[GSE_NY], const float* loglik_in_dev, float* loglik_dev,
                             double* stats_dev, FusedUpdate* fu) {
    GSE_REQUIRE(ctx != NULL && z != NULL && stats_dev != NULL, "ctx / z / stats is NULL");
    GSE_REQUIRE(n >= 1, "n out of range for this context");
    GSE_REQUIRE(loglik_dev != NULL && aligned16(loglik_dev), "loglik must be 16-byte aligned");
    GSE_REQUIRE(loglik_in_dev == NULL || aligned16(loglik_in_dev), "loglik_in must be 16-byte aligned");
    fu->z0h = (float)z[0];
    fu->z1h = (float)z[1];
    fu->z0l = (float)(z[0] - (double)fu->z0h);
    fu->z1l = (float)(z[1] - (double)fu->z1h);
    fu->loglik_in = loglik_in_dev;
    fu->loglik = loglik_dev;
    fu->block_max = ctx->block_max;
    fu->block_sum = ctx->block_sum;
    fu->ticket = ctx->ticket;
    fu->stats = stats_dev;
    return GSE_OK;
}

extern "C" int gse_pf_predict_update(gse_ctx* ctx, const float* x_src_dev, int64_t ld_src, const int32_t* idx_dev,
                                     float* x_dst_dev, int64_t ld_dst, int64_t n, const double u[GSE_NU], double dt,
                                     int n_sub, uint64_t seed, uint64_t step, int64_t index0, const double z[GSE_NY],
                                     const float* loglik_in_dev, float* loglik_dev, double* stats_dev, void* stream) {
    FusedUpdate fu;
    int rc = fill_fused_update(ctx, n, z, loglik_in_dev, loglik_dev, stats_dev, &fu);
    if (rc) return rc;
    GatherShards none;
    memset(&none, 0, sizeof(none));
    return launch_predict(ctx, x_src_dev, ld_src, idx_dev, &none, false, x_dst_dev, ld_dst, n, u, dt, n_sub, seed, step,
                          index0, NULL, 0, stream, &fu);
}

extern "C" int gse_pf_predict_update_sharded(gse_ctx* ctx, const gse_shards* shards, const int32_t* idx_dev,
                                             float* x_dst_dev, int64_t ld_dst, int64_t n, const double u[GSE_NU],
                                             double dt, int n_sub, uint64_t seed, uint64_t step, int64_t index0,
                                             const double z[GSE_NY], const float* loglik_in_dev, float* loglik_dev,
                                             double* stats_dev, void* stream) {
    FusedUpdate fu;
    int rc = fill_fused_update(ctx, n, z, loglik_in_dev, loglik_dev, stats_dev, &fu);
    if (rc) return rc;
    GatherShards g;
    rc = gse_build_gather_shards(shards, x_dst_dev, &g);
    if (rc) return rc;
    return launch_predict(ctx, NULL, 0, idx_dev, &g, true, x_dst_dev, ld_dst, n, u, dt, n_sub, seed, step, index0, NULL, 0,
                          stream, &fu);
}

// ------------------------------------------------------------------------------------------------
// K2: update.  loglik_i += log pdf_meas(z - g(x_i, u))   (particle.py:80-83), fused max / sum-exp.
// ------------------------------------------------------------------------------------------------
// Persistent grid (a few CTAs per SM): every thread walks groups of 4 rows with a grid stride and
// keeps a running (max, sum exp) pair, so the block-level reduction and its barriers run once per
// CTA instead of once per 1024 rows.
// PIPE: the loads of iteration k + 1 are issued before iteration k is evaluated (the rows of two iterations in
// registers: 64 instead of 48, four CTAs per SM instead of five).  Without it 45 % of the kernel's stall samples sit on
// the first use of the freshly loaded rows: a warp has no load in flight while it evaluates its eight log-pdfs.
template <int ND, bool LL_ZERO, bool PIPE>      // LL_ZERO: the accumulated log-likelihood is all zero (fresh resample)
__global__ void __launch_bounds__(PF_THREADS, PIPE ? 4 : 5)
k_pf_update(const float* __restrict__ xg, const float* __restrict__ xfa, const float* loglik_in,
            float* loglik, int64_t n, float z0h, float z0l, float z1h, float z1l,
            const __grid_constant__ MixDensity2f md, float* block_max, float* block_sum, unsigned int* ticket,
            double* stats, const gse_step_params* __restrict__ params) {
    if (params) {                                           // the host's hi/lo split of z (gse_pf_update), bit for bit
        z0h = (float)params->z[0];
        z1h = (float)params->z[1];
        z0l = (float)(params->z[0] - (double)z0h);
        z1l = (float)(params->z[1] - (double)z1h);
    }
    const int64_t groups = n >> 2;                          // whole groups of four rows; a ragged tail is handled below
    const int64_t stride = (int64_t)gridDim.x * PF_THREADS;
    MaxSumExp acc;
    // two groups per iteration, all six loads issued before the first use
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 cg[2], cf[2], lw[2];
    int64_t g = (int64_t)blockIdx.x * PF_THREADS + threadIdx.x;
#define UPDATE_LOAD(G, CG, CF, LW)                                                              \
    do {                                                                                        \
        const int64_t rA_ = (G) * ROWS_PER_THREAD, rB_ = ((G) + stride) * ROWS_PER_THREAD;      \
        const bool hB_ = (G) + stride < groups;                                                 \
        CG[0] = ld_stream4(xg + rA_);                                                           \
        CF[0] = ld_stream4(xfa + rA_);                                                          \
        LW[0] = LL_ZERO ? zero4 : ld_stream4(loglik_in + rA_);                                  \
        CG[1] = hB_ ? ld_stream4(xg + rB_) : zero4;                                             \
        CF[1] = hB_ ? ld_stream4(xfa + rB_) : zero4;                                            \
        LW[1] = (hB_ && !LL_ZERO) ? ld_stream4(loglik_in + rB_) : zero4;                        \
    } while (0)
    if (PIPE && g < groups) UPDATE_LOAD(g, cg, cf, lw);
    for (; g < groups; g += 2 * stride) {
        const int64_t rowA = g * ROWS_PER_THREAD, rowB = (g + stride) * ROWS_PER_THREAD;
        const bool hasB = g + stride < groups;
        float4 ncg[2], ncf[2], nlw[2];
        if (PIPE) {
            if (g + 2 * stride < groups) {
                UPDATE_LOAD(g + 2 * stride, ncg, ncf, nlw);
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) { ncg[h] = zero4; ncf[h] = zero4; nlw[h] = zero4; }      // last iteration: never used
            }
        } else {
            UPDATE_LOAD(g, cg, cf, lw);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && !hasB) break;
            const float a[4] = {cg[h].x, cg[h].y, cg[h].z, cg[h].w};
            const float b[4] = {cf[h].x, cf[h].y, cf[h].z, cf[h].w};
            const float l[4] = {lw[h].x, lw[h].y, lw[h].z, lw[h].w};
            float vals[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float e0 = __fadd_rn(__fsub_rn(z0h, output_glucose(a[r])), z0l);   // e = z - y   (:82)
                const float e1 = __fadd_rn(__fsub_rn(z1h, output_fa(b[r])), z1l);
                vals[r] = l[r] + meas_logpdf32<ND>(md, e0, e1);                 // weights[i] *= pdf(e)  (:83)
            }
            st_stream4(loglik + (h ? rowB : rowA), make_float4(vals[0], vals[1], vals[2], vals[3]));
            acc.add_all<4>(vals);
        }
        if (PIPE) {
#pragma unroll
            for (int h = 0; h < 2; ++h) { cg[h] = ncg[h]; cf[h] = ncf[h]; lw[h] = nlw[h]; }
        }
    }
#undef UPDATE_LOAD
    if ((n & 3) && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {      // the last one to three rows
        for (int64_t i = groups * ROWS_PER_THREAD; i < n; ++i) {
            const float e0 = __fadd_rn(__fsub_rn(z0h, output_glucose(xg[i])), z0l);
            const float e1 = __fadd_rn(__fsub_rn(z1h, output_fa(xfa[i])), z1l);
            const float v[1] = {(LL_ZERO ? 0.0f : loglik_in[i]) + meas_logpdf32<ND>(md, e0, e1)};
            loglik[i] = v[0];
            acc.add_all<1>(v);
        }
    }
    block_merge_max_sumexp<PF_THREADS>(acc.m, acc.s, block_max, block_sum, ticket, stats);
}

extern "C" int gse_pf_update(gse_ctx* ctx, const float* x_dev, int64_t ld, int64_t n, const float* loglik_in_dev,
                             float* loglik_dev, const double u[GSE_NU], const double z[GSE_NY], double* stats_dev,
                             void* stream) {
    GSE_REQUIRE(ctx != NULL && z != NULL && stats_dev != NULL, "ctx / z / stats is NULL");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 1 && n <= ctx->n_max, "n out of range for this context");
    CHECK_SOA(x_dev, ld, n);
    GSE_REQUIRE(loglik_dev != NULL && aligned16(loglik_dev), "loglik must be 16-byte aligned");
    GSE_REQUIRE(loglik_in_dev == NULL || aligned16(loglik_in_dev), "loglik_in must be 16-byte aligned");
    (void)u;   // static_outputs ignores u (BioreactorModel.py:250)
    const int64_t groups = gse_div_up(n, 2 * ROWS_PER_THREAD);          // a thread takes two groups of four rows per iteration
    int64_t nblk = gse_div_up(groups, PF_THREADS);
    const bool pipe = ctx->update_pipe != 0;
    const int per_sm = pipe && ctx->update_ctas_per_sm > 4 ? 4 : ctx->update_ctas_per_sm;
    if (nblk > (int64_t)ctx->num_sms * per_sm) nblk = (int64_t)ctx->num_sms * per_sm;   // persistent grid
    const unsigned blocks = (unsigned)nblk;
    GSE_REQUIRE((int64_t)blocks <= ctx->max_blocks, "workspace too small");
    const float z0h = (float)z[0], z1h = (float)z[1];
    const float z0l = (float)(z[0] - (double)z0h), z1l = (float)(z[1] - (double)z1h);
#define LAUNCH_UPDATE_ZP(ND, ZERO, PIPE)                                                                    \
    k_pf_update<ND, ZERO, PIPE><<<blocks, PF_THREADS, 0, (cudaStream_t)stream>>>(                              \
        x_dev + 0 * ld, x_dev + 2 * ld, loglik_in_dev, loglik_dev, n, z0h, z0l, z1h, z1l, ctx->meas_density32, \
        ctx->block_max, ctx->block_sum, ctx->ticket, stats_dev, ctx->step_params)
#define LAUNCH_UPDATE_Z(ND, ZERO) do { if (pipe) LAUNCH_UPDATE_ZP(ND, ZERO, true); else LAUNCH_UPDATE_ZP(ND, ZERO, false); } while (0)
#define LAUNCH_UPDATE(ND) do { if (loglik_in_dev) LAUNCH_UPDATE_Z(ND, false); else LAUNCH_UPDATE_Z(ND, true); } while (0)
    switch (ctx->meas_density32.nd) {
        case 1: LAUNCH_UPDATE(1); break;
        case 2: LAUNCH_UPDATE(2); break;
        case 3: LAUNCH_UPDATE(3); break;
        case 4: LAUNCH_UPDATE(4); break;
        default: LAUNCH_UPDATE(0); break;
    }
#undef LAUNCH_UPDATE
#undef LAUNCH_UPDATE_Z
#undef LAUNCH_UPDATE_ZP
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

__global__ void __launch_bounds__(PF_THREADS)
k_loglik_max(const float* __restrict__ loglik, int64_t n, float* block_max, float* block_sum,
             unsigned int* ticket, double* stats) {
    const int64_t groups = (n + 3) >> 2;
    const int64_t stride = (int64_t)gridDim.x * PF_THREADS;
    MaxSumExp acc;
    for (int64_t g = (int64_t)blockIdx.x * PF_THREADS + threadIdx.x; g < groups; g += stride) {
        const int64_t row0 = g * ROWS_PER_THREAD;
        const float4 lw = ld_stream4(loglik + row0);
        const float vals[4] = {lw.x, lw.y, lw.z, lw.w};
        bool valid[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) valid[r] = (row0 + r) < n;
        acc.add<4>(vals, valid);
    }
    block_merge_max_sumexp<PF_THREADS>(acc.m, acc.s, block_max, block_sum, ticket, stats);
}

extern "C" int gse_loglik_max(gse_ctx* ctx, const float* loglik_dev, int64_t n, double* stats_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && stats_dev != NULL, "ctx / stats is NULL");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 1 && n <= ctx->n_max, "n out of range for this context");
    GSE_REQUIRE(loglik_dev != NULL && aligned16(loglik_dev), "loglik must be 16-byte aligned");
    int64_t nblk = gse_div_up(gse_div_up(n, ROWS_PER_THREAD), PF_THREADS);
    if (nblk > (int64_t)ctx->num_sms * 8) nblk = (int64_t)ctx->num_sms * 8;
    const unsigned blocks = (unsigned)nblk;
    k_loglik_max<<<blocks, PF_THREADS, 0, (cudaStream_t)stream>>>(loglik_dev, n, ctx->block_max, ctx->block_sum,
                                                                   ctx->ticket, stats_dev);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

// ------------------------------------------------------------------------------------------------
// All-gathers over the peer mailboxes (gse_mailbox.cuh) fused with the reduction that consumes them: ONE single-warp
// kernel per exchange, no NCCL launch, no host involvement.
// ------------------------------------------------------------------------------------------------
// records (M_s, S_s) -> stats[0..1] = (max M_s, sum S_s exp(M_s - M)), merged in shard order
__global__ void __launch_bounds__(32)
k_mbox_stats(const __grid_constant__ MailboxTable mb, int rank, int nshards, unsigned int epoch, double* stats,
             unsigned int* err) {
    const int lane = threadIdx.x;
    unsigned long long r0, r1;
    mbox_exchange(mb, rank, nshards, epoch, (unsigned long long)__double_as_longlong(stats[0]),
                  (unsigned long long)__double_as_longlong(stats[1]), lane, r0, r1, err);
    const double m_s = lane < nshards ? __longlong_as_double((long long)r0) : -INFINITY;
    const double s_s = lane < nshards ? __longlong_as_double((long long)r1) : 0.0;
    double M = m_s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmax(M, __shfl_xor_sync(0xffffffffu, M, o));
    const double term = (m_s > -INFINITY) ? s_s * exp(m_s - M) : 0.0;
    double S = 0.0;
    for (int t = 0; t < nshards; ++t) S += __shfl_sync(0xffffffffu, term, t);      // fixed (shard) order
    if (lane == 0) { stats[0] = M; stats[1] = S; }
}

// records T_s (uint64) -> offsets[0..nshards] = exclusive prefix of the shard totals, total last
__global__ void __launch_bounds__(32)
k_mbox_offsets(const __grid_constant__ MailboxTable mb, int rank, int nshards, unsigned int epoch,
               const uint64_t* __restrict__ total, uint64_t* __restrict__ offsets, unsigned int* err) {
    const int lane = threadIdx.x;
    unsigned long long r0, r1;
    mbox_exchange(mb, rank, nshards, epoch, (unsigned long long)total[0], 0ull, lane, r0, r1, err);
    unsigned long long incl = r0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane < nshards) offsets[lane + 1] = incl;
    if (lane == 0) offsets[0] = 0;
}

// records of 48 doubles (the output block of the moments kernels: S0, S1[5], S2[15], pivot[5], extra[15], M, S, ...):
// merged in shard order into the moments of the whole population about shard 0's pivot,
//   S0 = sum s0,  S1 = sum (s1 + s0 d),  S2 = sum (s2 + s1 d' + d s1' + s0 d d'),  d = pivot_s - pivot_0,
// the extra block (GS-UKF: sum w P) is a plain sum.  Every rank ends up with the same 48 doubles.
#define MOM_WORDS 48
__global__ void __launch_bounds__(32)
k_mbox_moments(const __grid_constant__ MailboxTable mb, int rank, int nshards, unsigned int epoch, const double* mom,
               const double* __restrict__ stats, double* out, unsigned int* err) {
    const int lane = threadIdx.x;
    mbox_exchange_block(mb, rank, nshards, epoch, reinterpret_cast<const unsigned long long*>(mom), MOM_WORDS, lane, err);
    // lane t reads shard t's block and moves it to shard 0's pivot (one round trip to the mailbox for all shards, not
    // one per shard); the sums then run over the lanes in shard order, the same on every rank
    const int t = lane < nshards ? lane : 0;
    volatile unsigned long long* src = mbox_slot(mb, rank, t, epoch);
    double m[41];
#pragma unroll
    for (int k = 0; k < 41; ++k) m[k] = __longlong_as_double((long long)src[k]);
    double c[36];                                            // S0, S1[5], S2[15], X[15] of this shard about the common pivot
    double d[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) d[j] = m[21 + j] - __shfl_sync(0xffffffffu, m[21 + j], 0);
    c[0] = m[0];
#pragma unroll
    for (int j = 0; j < 5; ++j) c[1 + j] = m[1 + j] + m[0] * d[j];
    {
        int q = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j, ++q)
                c[6 + q] = m[6 + q] + m[1 + i] * d[j] + d[i] * m[1 + j] + m[0] * d[i] * d[j];
    }
#pragma unroll
    for (int k = 0; k < 15; ++k) c[21 + k] = m[26 + k];
#pragma unroll
    for (int k = 0; k < 36; ++k) {
        double tot = 0.0;
        for (int u = 0; u < nshards; ++u) tot += __shfl_sync(0xffffffffu, c[k], u);      // fixed (shard) order
        c[k] = tot;
    }
    if (lane != 0) return;
    out[0] = c[0];
#pragma unroll
    for (int j = 0; j < 5; ++j) { out[1 + j] = c[1 + j]; out[21 + j] = m[21 + j]; }     // lane 0 holds shard 0: its pivot
#pragma unroll
    for (int k = 0; k < 15; ++k) { out[6 + k] = c[6 + k]; out[26 + k] = c[21 + k]; }
    if (stats) { out[41] = stats[0]; out[42] = stats[1]; }          // the global (M, S) ride along in the same read-back
}

extern "C" int gse_peer_allgather_stats(gse_ctx* ctx, void* const mailboxes[GSE_MAX_SHARDS], int rank, int nshards,
                                        unsigned int epoch, double* stats_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && stats_dev != NULL && epoch != 0, "bad arguments");
    gse_device_guard guard(ctx->device);
    MailboxTable mb;
    int rc = gse_build_mailboxes(mailboxes, rank, nshards, &mb);
    if (rc) return rc;
    k_mbox_stats<<<1, 32, 0, (cudaStream_t)stream>>>(mb, rank, nshards, epoch, stats_dev, ctx->err_dev);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

extern "C" int gse_peer_allgather_totals(gse_ctx* ctx, void* const mailboxes[GSE_MAX_SHARDS], int rank, int nshards,
                                         unsigned int epoch, const uint64_t* total_dev, uint64_t* offsets_dev,
                                         void* stream) {
    GSE_REQUIRE(ctx != NULL && total_dev != NULL && offsets_dev != NULL && epoch != 0, "bad arguments");
    gse_device_guard guard(ctx->device);
    MailboxTable mb;
    int rc = gse_build_mailboxes(mailboxes, rank, nshards, &mb);
    if (rc) return rc;
    k_mbox_offsets<<<1, 32, 0, (cudaStream_t)stream>>>(mb, rank, nshards, epoch, total_dev, offsets_dev, ctx->err_dev);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

extern "C" int gse_peer_allgather_moments(gse_ctx* ctx, void* const mailboxes[GSE_MAX_SHARDS], int rank, int nshards,
                                          unsigned int epoch, double* mom_dev, const double* stats_dev, double* out_dev,
                                          void* stream) {
    GSE_REQUIRE(ctx != NULL && mom_dev != NULL && epoch != 0, "bad arguments");
    if (!out_dev) out_dev = mom_dev;
    gse_device_guard guard(ctx->device);
    MailboxTable mb;
    int rc = gse_build_mailboxes(mailboxes, rank, nshards, &mb);
    if (rc) return rc;
    k_mbox_moments<<<1, 32, 0, (cudaStream_t)stream>>>(mb, rank, nshards, epoch, mom_dev, stats_dev, out_dev, ctx->err_dev);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

// Sharded run: combine the per-shard (M_s, S_s) pairs (all-gathered on the stream) into the global
// stats[0] = M = max_s M_s, stats[1] = S = sum_s S_s exp(M_s - M), in shard order.
__global__ void k_merge_stats(const double* __restrict__ pairs, int nshards, double* __restrict__ stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double M = -INFINITY;
    for (int s = 0; s < nshards; ++s) M = fmax(M, pairs[2 * s]);
    double S = 0.0;
    for (int s = 0; s < nshards; ++s)
        if (pairs[2 * s] > -INFINITY) S += pairs[2 * s + 1] * exp(pairs[2 * s] - M);
    stats[0] = M;
    stats[1] = S;
}

extern "C" int gse_merge_stats(gse_ctx* ctx, const double* pairs_dev, int nshards, double* stats_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && pairs_dev != NULL && stats_dev != NULL && nshards >= 1, "bad arguments");
    gse_device_guard guard(ctx->device);
    k_merge_stats<<<1, 32, 0, (cudaStream_t)stream>>>(pairs_dev, nshards, stats_dev);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

// ------------------------------------------------------------------------------------------------
// weights read-back: base * exp(loglik) * scale in float64 (the reference's `weights` attribute)
// ------------------------------------------------------------------------------------------------
__global__ void k_weights_linear(const float* __restrict__ loglik, const double* __restrict__ base,
                                 int64_t n, double scale, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double w = scale;
    if (loglik) w *= exp((double)loglik[i]);
    if (base) w *= base[i];
    out[i] = w;
}

extern "C" int gse_weights_linear(gse_ctx* ctx, const float* loglik_dev, const double* base_dev, int64_t n,
                                  double scale, double* out_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && out_dev != NULL, "ctx / out is NULL");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return GSE_OK;
    k_weights_linear<<<(unsigned)gse_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(loglik_dev, base_dev, n, scale, out_dev);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

// ------------------------------------------------------------------------------------------------
// K6: weighted moments about a pivot row, float64 accumulation, fixed-order two-level reduction.
// out: [0] S0, [1..5] S1, [6..20] S2 lower triangle, [21..25] pivot; (GSF: [26..40] sum w P)
// ------------------------------------------------------------------------------------------------
#define MOM_THREADS 256
// NEXTRA: 0 particle filter, 15 GS-UKF (+ sum w P), -1 means only (S0, S1: point_estimate alone)
template <int NEXTRA, int GMODE>      // GMODE as in k_pf_predict
__global__ void __launch_bounds__(MOM_THREADS, NEXTRA < 0 ? 4 : (NEXTRA == 0 ? 2 : 1))
k_moments(const float* __restrict__ x, const float* __restrict__ extra, int64_t ld, int64_t n,
          const int32_t* __restrict__ idx, const __grid_constant__ GatherShards shards_arg,
          const float* __restrict__ loglik, const double* __restrict__ base,
          const double* __restrict__ stats, double* partials, unsigned int* ticket, double* out) {
    const GatherShards& shards = shards_arg;
    constexpr bool MEAN = NEXTRA < 0;
    constexpr int NV = MEAN ? 6 : 21 + NEXTRA;
    const float M = (float)stats[0];
    float p[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        if (GMODE == 2) {
            int64_t l0;
            const float* q = shard_row(shards, idx[0], l0);
            p[j] = q[j * l0];
        } else {
            p[j] = __ldg(x + j * ld + (GMODE == 1 ? idx[0] : 0));
        }
    }
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    const int64_t groups = (n + 3) / 4;
    for (int64_t g = (int64_t)blockIdx.x * MOM_THREADS + threadIdx.x; g < groups; g += (int64_t)gridDim.x * MOM_THREADS) {
        const int64_t row0 = g * 4;
        float4 c[5];
        int id[4] = {0, 0, 0, 0};
        const float* qs[4] = {NULL, NULL, NULL, NULL};     // GMODE 2: column 0 of each row in its owner's buffer
        int64_t ls[4] = {0, 0, 0, 0};
        if (GMODE != 0) {
            const int4 id4 = *reinterpret_cast<const int4*>(idx + row0);
            id[0] = id4.x; id[1] = (row0 + 1 < n) ? id4.y : id4.x; id[2] = (row0 + 2 < n) ? id4.z : id4.x;
            id[3] = (row0 + 3 < n) ? id4.w : id4.x;
            if (GMODE == 2) {
                shard_rows4(shards, id, qs, ls);
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    c[j] = make_float4(qs[0][j * ls[0]], qs[1][j * ls[1]], qs[2][j * ls[2]], qs[3][j * ls[3]]);
            } else {
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    c[j] = make_float4(__ldg(x + j * ld + id[0]), __ldg(x + j * ld + id[1]), __ldg(x + j * ld + id[2]),
                                       __ldg(x + j * ld + id[3]));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 5; ++j) c[j] = ld_stream4(x + j * ld + row0);
        }
        float4 lw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (loglik) lw = ld_stream4(loglik + row0);
        const float l[4] = {lw.x, lw.y, lw.z, lw.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (row0 + r < n) {
                double w = loglik ? (double)fast_exp(l[r] - M) : 1.0;
                if (base) w *= base[row0 + r];
                double d[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) d[j] = (double)reinterpret_cast<const float*>(&c[j])[r] - (double)p[j];   // exact
                acc[0] += w;
                int t = 6;
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const double wd = w * d[j];
                    acc[1 + j] += wd;
                    if (!MEAN) {
#pragma unroll
                        for (int k = 0; k <= j; ++k) { acc[t] = fma(wd, d[k], acc[t]); ++t; }
                    }
                }
                if (NEXTRA > 0) {
#pragma unroll
                    for (int k = 0; k < NEXTRA; ++k) {
                        const float ex = GMODE == 2 ? qs[r][(5 + k) * ls[r]]
                                                    : extra[k * ld + (GMODE == 1 ? (int64_t)id[r] : row0 + r)];
                        acc[21 + k] = fma(w, (double)ex, acc[21 + k]);
                    }
                }
            }
        }
    }
    __shared__ double s_part[MOM_THREADS / 32][NV];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) s_part[wid][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < MOM_THREADS / 32; ++w) t += s_part[w][threadIdx.x];
        partials[(size_t)blockIdx.x * NV + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < NV) {
        double t = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) t += __ldcg(partials + (size_t)b * NV + threadIdx.x);
        out[threadIdx.x < 21 ? threadIdx.x : threadIdx.x + 5] = t;
    }
    if (threadIdx.x < 5) out[21 + threadIdx.x] = (double)p[threadIdx.x];
    if (threadIdx.x == 0) *ticket = 0u;
}

// ------------------------------------------------------------------------------------------------
// point_estimate alone (weights without a float64 base): S0 = sum w, S1 = sum w x about pivot 0.
// Products and the sum over a thread's 4 rows in float32 (4 terms: relative error < 3e-7 per group,
// zero-mean over the groups), accumulation over groups in float64, fixed-order reduction.  HBM-bound
// where the all-float64 k_moments is not; the groups are the same whatever the launch geometry or
// the sharding (shards start at multiples of four), so single-GPU and sharded sums agree to the
// float64 summation order.
// ------------------------------------------------------------------------------------------------
template <int GMODE>
__global__ void __launch_bounds__(MOM_THREADS, GMODE == 2 ? 3 : 4)
k_means(const float* __restrict__ x, int64_t ld, int64_t n, const int32_t* __restrict__ idx,
        const __grid_constant__ GatherShards shards_arg, const float* __restrict__ loglik,
        const double* __restrict__ stats, double* partials, unsigned int* ticket, double* out) {
    const GatherShards& shards = shards_arg;
    constexpr int NV = 6;
    const float M = (float)stats[0];
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    const int64_t groups = (n + 3) / 4;
    for (int64_t g = (int64_t)blockIdx.x * MOM_THREADS + threadIdx.x; g < groups; g += (int64_t)gridDim.x * MOM_THREADS) {
        const int64_t row0 = g * 4;
        float4 c[5];
        if (GMODE != 0) {
            const int4 id4 = *reinterpret_cast<const int4*>(idx + row0);
            const int id[4] = {id4.x, (row0 + 1 < n) ? id4.y : id4.x, (row0 + 2 < n) ? id4.z : id4.x,
                               (row0 + 3 < n) ? id4.w : id4.x};
            if (GMODE == 2) {
                const float* q[4];
                int64_t l[4];
                shard_rows4(shards, id, q, l);
#pragma unroll
                for (int j = 0; j < 5; ++j) c[j] = make_float4(q[0][j * l[0]], q[1][j * l[1]], q[2][j * l[2]], q[3][j * l[3]]);
            } else {
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const float* col = x + j * ld;
                    c[j] = make_float4(__ldg(col + id[0]), __ldg(col + id[1]), __ldg(col + id[2]), __ldg(col + id[3]));
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 5; ++j) c[j] = ld_stream4(x + j * ld + row0);
        }
        float w[4] = {1.f, 1.f, 1.f, 1.f};
        if (loglik) {
            const float4 lw = ld_stream4(loglik + row0);
            w[0] = fast_exp(lw.x - M); w[1] = fast_exp(lw.y - M); w[2] = fast_exp(lw.z - M); w[3] = fast_exp(lw.w - M);
        }
#pragma unroll
        for (int r = 1; r < 4; ++r) if (row0 + r >= n) w[r] = 0.f;
        acc[0] += (double)((w[0] + w[1]) + (w[2] + w[3]));
#pragma unroll
        for (int j = 0; j < 5; ++j)
            acc[1 + j] += (double)(fmaf(w[0], c[j].x, w[1] * c[j].y) + fmaf(w[2], c[j].z, w[3] * c[j].w));
    }
    __shared__ double s_part[MOM_THREADS / 32][NV];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) s_part[wid][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < MOM_THREADS / 32; ++w) t += s_part[w][threadIdx.x];
        partials[(size_t)blockIdx.x * NV + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < NV) {
        double t = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) t += __ldcg(partials + (size_t)b * NV + threadIdx.x);
        out[threadIdx.x] = t;
    }
    if (threadIdx.x < 5) out[21 + threadIdx.x] = 0.0;      // pivot 0
    if (threadIdx.x == 0) *ticket = 0u;
}

static int launch_moments(gse_ctx* ctx, const float* x, const float* extra, bool gsf, int64_t ld, int64_t n,
                          const int32_t* idx, const GatherShards* shards, bool mean_only, const float* loglik, const double* base, const double* stats, double* out,
                          void* stream) {
    GSE_REQUIRE(ctx != NULL && stats != NULL && out != NULL, "ctx / stats / out is NULL");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 1 && n <= ctx->n_max, "n out of range for this context");
    if (idx == NULL) { CHECK_SOA(x, ld, n); }
    GatherShards none;
    memset(&none, 0, sizeof(none));
    const GatherShards& sh = shards ? *shards : none;
    GSE_REQUIRE(shards == NULL || idx != NULL, "sharded moments need idx");
    const int64_t groups = gse_div_up(n, 4);
    int64_t blocks = gse_div_up(groups, MOM_THREADS);
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks > 2048) blocks = 2048;
    GSE_REQUIRE(idx == NULL || aligned16(idx), "idx must be 16-byte aligned");
#define LAUNCH_MOM(NE, G, EX)                                                                                \
    k_moments<NE, G><<<(unsigned)blocks, MOM_THREADS, 0, (cudaStream_t)stream>>>(x, EX, ld, n, idx, sh, loglik,  \
                                                                                 base, stats, ctx->red_partials, \
                                                                                 ctx->ticket + 2, out)
#define LAUNCH_MEANS(G)                                                                                      \
    k_means<G><<<(unsigned)blocks, MOM_THREADS, 0, (cudaStream_t)stream>>>(x, ld, n, idx, sh, loglik, stats,     \
                                                                           ctx->red_partials, ctx->ticket + 2, out)
    if (gsf) { if (shards) LAUNCH_MOM(15, 2, extra); else if (idx) LAUNCH_MOM(15, 1, extra); else LAUNCH_MOM(15, 0, extra); }
    else if (mean_only && base == NULL) {                  // the common point_estimate: float32 group sums
        if (shards) LAUNCH_MEANS(2);
        else if (idx) LAUNCH_MEANS(1);
        else LAUNCH_MEANS(0);
    } else if (mean_only) {
        if (shards) LAUNCH_MOM(-1, 2, NULL);
        else if (idx) LAUNCH_MOM(-1, 1, NULL);
        else LAUNCH_MOM(-1, 0, NULL);
    } else if (shards) LAUNCH_MOM(0, 2, NULL);
    else { if (idx) LAUNCH_MOM(0, 1, NULL); else LAUNCH_MOM(0, 0, NULL); }
#undef LAUNCH_MOM
#undef LAUNCH_MEANS
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

extern "C" int gse_pf_moments(gse_ctx* ctx, const float* x_dev, int64_t ld, int64_t n, const int32_t* idx_dev,
                              const float* loglik_dev, const double* base_dev, const double* stats_dev,
                              int mean_only, double* out_dev, void* stream) {
    return launch_moments(ctx, x_dev, NULL, false, ld, n, idx_dev, NULL, mean_only != 0, loglik_dev, base_dev, stats_dev,
                          out_dev, stream);
}

extern "C" int gse_pf_moments_sharded(gse_ctx* ctx, const gse_shards* shards, const int32_t* idx_dev, int64_t n,
                                      const float* loglik_dev, const double* base_dev, const double* stats_dev,
                                      int mean_only, double* out_dev, void* stream) {
    GatherShards g;
    int rc = gse_build_gather_shards(shards, NULL, &g);
    if (rc) return rc;
    return launch_moments(ctx, NULL, NULL, false, 0, n, idx_dev, &g, mean_only != 0, loglik_dev, base_dev, stats_dev, out_dev,
                          stream);
}

extern "C" int gse_gsf_moments(gse_ctx* ctx, const float* mean_dev, const float* cov_dev, int64_t ld, int64_t n,
                               const int32_t* idx_dev, const float* loglik_dev, const double* base_dev,
                               const double* stats_dev, double* out_dev, void* stream) {
    GSE_REQUIRE(cov_dev != NULL, "cov is NULL");
    return launch_moments(ctx, mean_dev, cov_dev, true, ld, n, idx_dev, NULL, false, loglik_dev, base_dev, stats_dev, out_dev,
                          stream);
}

extern "C" int gse_gsf_moments_sharded(gse_ctx* ctx, const gse_shards* shards, const int32_t* idx_dev, int64_t n,
                                       const float* loglik_dev, const double* base_dev, const double* stats_dev,
                                       double* out_dev, void* stream) {
    GatherShards g;
    int rc = gse_build_gather_shards(shards, NULL, &g);
    if (rc) return rc;
    return launch_moments(ctx, NULL, NULL, true, 0, n, idx_dev, &g, false, loglik_dev, base_dev, stats_dev, out_dev, stream);
}

// ------------------------------------------------------------------------------------------------
// stand-alone mixture pdf (MultivariateGaussianSum.pdf, :39-63): float64 throughout, not a hot path
// ------------------------------------------------------------------------------------------------
__global__ void k_mixture_pdf(const float* __restrict__ x, int64_t ld, int64_t n,
                              const __grid_constant__ MixDensityN md, double* __restrict__ out, int log_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double xv[GSE_NX];
    for (int j = 0; j < md.nx; ++j) xv[j] = (double)x[j * ld + i];
    double a[GSE_MAX_ND];
    double m = -1.0e300;
    for (int d = 0; d < md.nd; ++d) {
        double q = 0.0;
        for (int r = 0; r < md.nx; ++r) {
            double t = 0.0;
            for (int c = 0; c < md.nx; ++c) t += md.P[d][r * md.nx + c] * (xv[c] - md.mean[d][c]);
            q += (xv[r] - md.mean[d][r]) * t;
        }
        a[d] = md.logc[d] - 0.5 * q;
        m = fmax(m, a[d]);
    }
    double s = 0.0;
    for (int d = 0; d < md.nd; ++d) s += exp(a[d] - m);
    const double lp = m + log(s);
    out[i] = log_out ? lp : exp(lp);
}

extern "C" int gse_mixture_pdf(gse_ctx* ctx, const gse_mixture* mix, const float* x_dev, int64_t ld, int64_t n,
                               double* out_dev, int log_out, void* stream) {
    GSE_REQUIRE(ctx != NULL && x_dev != NULL && out_dev != NULL, "ctx / x / out is NULL");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 0 && ld >= n, "n / ld out of range");
    if (n == 0) return GSE_OK;
    MixDensityN md;
    int rc = gse_build_densityN(mix, &md);
    if (rc) return rc;
    k_mixture_pdf<<<(unsigned)gse_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(x_dev, ld, n, md, out_dev, log_out);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}
