// Shared definitions for libgse_b200.so (sm_100a).  See include/gse.h for the ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/gse.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
void gse_set_error(const char* fmt, ...);

#define GSE_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            gse_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,            \
                          cudaGetErrorString(_e));                                        \
            return GSE_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

#define GSE_CHECK_LAUNCH(ctx)                                                             \
    do {                                                                                  \
        (ctx)->launches++;                                                                \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            gse_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,        \
                          cudaGetErrorString(_e));                                        \
            return GSE_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

#define GSE_REQUIRE(cond, msg)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            gse_set_error("invalid argument at %s:%d: %s", __FILE__, __LINE__, msg);      \
            return GSE_EINVAL;                                                            \
        }                                                                                 \
    } while (0)

// ------------------------------------------------------------------------------------------------
// kernel-parameter structs (passed by value: no __constant__ globals, so that several contexts
// with different mixtures can live in one process)
// ------------------------------------------------------------------------------------------------

// Sampler for a 5-d Gaussian sum: component chosen by inverse cdf, then mean + L z.
struct MixSampler5 {
    int nd;
    int diag;                          // every L is diagonal (the benchmark's noise, sim_base.py:141-160)
    float cdf[GSE_MAX_ND];             // inclusive cumulative weights, cdf[nd-1] = 1
    float mean[GSE_MAX_ND][GSE_NX];
    float L[GSE_MAX_ND][GSE_NCOV];     // lower-triangular Cholesky factor, row-major packed
};

// log-density of a 2-d Gaussian sum (measurement noise), float64 quadratic forms.
struct MixDensity2 {
    int nd;
    double logc[GSE_MAX_ND];           // log(w_d * const_d)          (MultivariateGaussianSum.py:36-37,60)
    double mean[GSE_MAX_ND][2];
    double p00[GSE_MAX_ND], p01[GSE_MAX_ND], p11[GSE_MAX_ND];   // inverse covariance; p01 = P01 + P10 (:33)
};

// float32 copy used by the particle-filter update kernel (see meas_logpdf32)
struct MixDensity2f {
    int nd;
    float logc[GSE_MAX_ND];
    float mean[GSE_MAX_ND][2];
    float p00[GSE_MAX_ND], p01[GSE_MAX_ND], p11[GSE_MAX_ND];
};

// Generic density (nx <= 5) used by gse_mixture_pdf.
struct MixDensityN {
    int nd, nx;
    double logc[GSE_MAX_ND];
    double mean[GSE_MAX_ND][GSE_NX];
    double P[GSE_MAX_ND][GSE_NX * GSE_NX];
};

struct gse_ctx {
    int device;
    int model_id;
    int64_t n_max;
    int num_sms;
    int64_t launches;
    MixSampler5 state_sampler;
    MixDensity2 meas_density;
    MixDensity2f meas_density32;
    // workspace
    void* ws;                 // one allocation, carved below
    size_t ws_bytes;
    float* block_max;         // per-block partial maxima (update / loglik_max)
    float* block_sum;         // per-block partial sums of exp(loglik - block max)
    unsigned int* ticket;     // [0] update reduction, [1] tile sums, [2] moments, [4..6] look-back scan (start ticket, finished,
                              // epoch), [8..11] fused resample (start ticket, past phase 3, queue length, past the drain)
    double* red_partials;     // per-block partial moments
    uint64_t* tile_agg;       // scan: sum of every warp's run of tiles
    uint64_t* tile_inc;       // scan: exclusive offset of every warp's run
    uint64_t* tile_status;    // single-pass kernels: one status word per CTA (look-back scan, fused resample)
    int64_t scan_tiles_prev;  // tiles the previous look-back launch rewrote (0: none yet)
    int scan_resident_blocks; // co-resident CTAs of the look-back kernel on this device (0: not queried yet)
    int update_ctas_per_sm;   // persistent grid of the update kernel (5; GSE_UPDATE_CTAS overrides, for tuning)
    int update_pipe;          // update kernel with the next iteration's loads in flight (GSE_UPDATE_PIPE=0 disables)
    int gsf_update_waves;     // persistent grid of the GS-UKF update kernel, in waves of resident CTAs (GSE_GSF_UPDATE_WAVES)
    int scan_single_pass;     // use the look-back scan for loglik-only weights (GSE_SCAN=twopass disables)
    uint64_t* fused_status;   // fused resample: one aggregate word per CTA, zero between launches
    int4* heavy_queue;        // fused resample: runs of one heavy source handed to the whole grid (start, end, ancestor)
    int heavy_queue_cap;
    int fused_resident[24];   // co-resident CTAs of each k_resample_fused instantiation (0: not queried yet)
    int gsf_minb;             // GSE_GSF_MINB = 3..6: CTAs (of 128 threads) per SM for both GS-UKF kernels; 0: predict 5, update 6
    int predict_minb;         // CTAs per SM of the benchmark's predict specialisation (4; GSE_PREDICT_MINB=5: the 48-register build)
    unsigned long long* fused_trace;   // GSE_FUSED_TRACE=1: per-CTA phase time stamps of the last fused resample (debugging)
    int fused_minb;           // CTAs per SM the fused kernel is compiled for (3; GSE_FUSED_MINB=4 to compare)
    unsigned int* err_host;   // device-error word: pinned, mapped host memory the kernels OR their GSE_ERR_* bits into
    unsigned int* err_dev;    // its device alias
    double* result_host;      // 64 doubles in the same mapped block: kernels may write a moment block straight to the host
    double* result_dev;       // its device alias
    int64_t* part;            // merge-path split points
    int64_t* range;           // [k_lo, k_hi): sources that interleave with a shard's outputs
    const gse_step_params* step_params;   // device block overriding the per-step scalars (CUDA-graph replay), or NULL
    gse_step_params* params_block;        // the context's device block
    gse_step_params* params_ring;         // pinned staging ring (GSE_PARAM_RING slots)
    cudaEvent_t params_event[4];          // one per quarter of the ring: guards against overrunning queued copies
    int params_pos;
    int64_t max_blocks;
    int64_t max_tiles;
};

int gse_build_sampler5(const gse_mixture* m, MixSampler5* out);
int gse_build_sampler(const gse_mixture* m, MixSampler5* out, int* nx_out);
int gse_build_density2(const gse_mixture* m, MixDensity2* out);
int gse_build_densityN(const gse_mixture* m, MixDensityN* out);

#define GSE_PARAM_RING 1024

// Every entry point that launches, allocates or frees selects the context's device for the duration of the call and
// restores the caller's current device on the way out (a filter may live on a device other than the current one).
struct gse_device_guard {
    int prev;
    bool switched;
    explicit gse_device_guard(int device) : prev(-1), switched(false) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = (cudaSetDevice(device) == cudaSuccess);
    }
    ~gse_device_guard() {
        if (switched) cudaSetDevice(prev);
    }
};

static inline int64_t gse_div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// State rows of a sharded population (kernel parameter): shard s holds the global rows
// [seg_row[s], seg_row[s+1]) in its own SoA buffer -- local memory for this rank's shard, peer
// memory (NVLink) for the others.
struct GatherShards {
    int nseg;
    int64_t seg_row[GSE_MAX_SHARDS + 1];
    const float* state[GSE_MAX_SHARDS];
    int64_t ld[GSE_MAX_SHARDS];
    // the calling rank's own shard once more, as scalars: nearly every ancestor is local (the shards' weight totals
    // are balanced), and these fields are read without any look-up
    const float* home_state;
    int64_t home_ld, home_row0, home_row1;
    const float* home_base;        // home_state - home_row0: row k of the home shard is home_base[k] (k a GLOBAL index)
    int home_lo, home_hi;          // home_row0 / home_row1 as int32 (ancestor indices are int32)
    int ends_first;                // block order of the gathering kernels: both ends of the shard first (see k_pf_predict)
};

static inline int gse_build_gather_shards(const gse_shards* sh, const void* dst, GatherShards* g) {
    GSE_REQUIRE(sh != NULL && sh->nshards >= 1 && sh->nshards <= GSE_MAX_SHARDS, "bad shard table");
    memset(g, 0, sizeof(*g));
    g->nseg = sh->nshards;
    for (int t = 0; t <= sh->nshards; ++t) g->seg_row[t] = sh->rows[t];
    for (int t = 0; t < sh->nshards; ++t) {
        GSE_REQUIRE(sh->state_dev[t] != NULL && (const void*)sh->state_dev[t] != dst, "bad shard state pointer");
        g->state[t] = sh->state_dev[t];
        g->ld[t] = sh->ld[t];
    }
    const int home = (sh->rank >= 0 && sh->rank < sh->nshards) ? sh->rank : 0;
    g->home_state = g->state[home];
    g->home_ld = g->ld[home];
    g->home_row0 = g->seg_row[home];
    g->home_row1 = g->seg_row[home + 1];
    g->home_base = (const float*)((uintptr_t)g->home_state - (uintptr_t)g->home_row0 * sizeof(float));
    g->home_lo = (int)g->home_row0;
    g->home_hi = (int)g->home_row1;
    {
        const char* v = getenv("GSE_PREDICT_ENDS_FIRST");
        g->ends_first = (sh->nshards > 1 && !(v && atoi(v) == 0)) ? 1 : 0;     // measured at 2 x 2^24 rows: 140.4 -> 136.7 us
    }
    return GSE_OK;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Counter-based: the stream of row i at step t
// is a pure function of (seed, i, t, subsequence) -- independent of launch geometry and of how
// rows are sharded over GPUs.
// ------------------------------------------------------------------------------------------------
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

struct Philox4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    Philox4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// uniform in (0, 1]: (x + 0.5) * 2^-32 evaluated in float32 (the curand convention)
__device__ __forceinline__ float u32_to_unit(uint32_t x) {
    return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

// Box-Muller with the MUFU approximations (lg2.approx, sin/cos.approx); the angle is folded to
// (-pi, pi] where sin.approx / cos.approx are most accurate.
__device__ __forceinline__ float mufu_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mufu_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// exp(x) for x <= ~0 as one multiply and one MUFU: ex2.approx.ftz (results below 2^-126 flush to zero; __expf spends
// three more instructions per call on that range)
__device__ __forceinline__ float fast_exp(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}
__device__ __forceinline__ float box_muller_radius(uint32_t a) {
    // sqrt(-2 ln u1) = sqrt(lg2(u1) * (-2 ln 2))
    return mufu_sqrt(mufu_lg2(u32_to_unit(a)) * -1.3862943611198906f);
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float r = box_muller_radius(a);
    const float th = fmaf(u32_to_unit(b), 6.2831853071795865f, -3.1415926535897932f);
    z0 = r * __cosf(th);
    z1 = r * __sinf(th);
}
__device__ __forceinline__ float box_muller_cos(uint32_t a, uint32_t b) {
    const float th = fmaf(u32_to_unit(b), 6.2831853071795865f, -3.1415926535897932f);
    return box_muller_radius(a) * __cosf(th);
}

// mixture sample from five standard normals z and the component selector uc: mean[c] + L[c] z
template <bool DIAG, int ND>      // ND = 0: component count read from the sampler at run time
__device__ __forceinline__ void mix_apply(const MixSampler5& sp, const float z[5], float uc, float out[5]) {
    const int dg[5] = {0, 2, 5, 9, 14};
    if (ND == 2 && DIAG) {
        const bool second = uc > sp.cdf[0];
#pragma unroll
        for (int j = 0; j < 5; ++j)
            out[j] = fmaf(second ? sp.L[1][dg[j]] : sp.L[0][dg[j]], z[j], second ? sp.mean[1][j] : sp.mean[0][j]);
        return;
    }
    int comp = 0;
    if (ND != 1) {
#pragma unroll
        for (int d = 0; d < GSE_MAX_ND - 1; ++d)
            comp += ((ND == 0 ? d < sp.nd - 1 : d < ND - 1) && uc > sp.cdf[d]) ? 1 : 0;
    }
    if (DIAG) {
#pragma unroll
        for (int j = 0; j < 5; ++j) out[j] = fmaf(sp.L[comp][dg[j]], z[j], sp.mean[comp][j]);
    } else {
        int t = 0;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            float acc = sp.mean[comp][j];
#pragma unroll
            for (int m = 0; m <= j; ++m) acc = fmaf(sp.L[comp][t++], z[m], acc);
            out[j] = acc;
        }
    }
}

// Five state-noise values for row `index` at `step`, subsequence pair (sub, sub+1).
// Draw layout (restated in oracle/philox.py):
//   A = philox(index_lo, index_hi, step, 2*sub)    -> (z0, z1) = BM(A.x, A.y), (z2, z3) = BM(A.z, A.w)
//   B = philox(index_lo, index_hi, step, 2*sub+1)  -> (z4, _ ) = BM(B.x, B.y), component from B.z
template <bool DIAG, int ND>      // ND = 0: component count read from the sampler at run time
__device__ __forceinline__ void draw_mixture5(const MixSampler5& sp, uint64_t index, uint32_t step,
                                              uint32_t sub, uint32_t k0, uint32_t k1, float out[5]) {
    const Philox4 A = philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), step, 2u * sub, k0, k1);
    const Philox4 B = philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), step, 2u * sub + 1u, k0, k1);
    float z[5];
    box_muller(A.x, A.y, z[0], z[1]);
    box_muller(A.z, A.w, z[2], z[3]);
    z[4] = box_muller_cos(B.x, B.y);
    mix_apply<DIAG, ND>(sp, z, u32_to_unit(B.z), out);
}

// State noise of the ELEVEN sigma points of one Gaussian component (GS-UKF predict, gs_ukf.py:99) from 17 Philox calls
// instead of 22: 55 normals + 11 selectors = 66 of 68 words.
//   P_j = philox(ctr = (i_lo, i_hi, step, 0x40000000 + j), key),  j = 0..16;  words w[4 j + {0,1,2,3}] = P_j.{x,y,z,w}
//   (n[2 p], n[2 p + 1]) = BM(w[2 p], w[2 p + 1]),  p = 0..27;  sigma point s takes n[5 s .. 5 s + 4] and selector w[56 + s]
// Consumed as a stream (four normals and four selector words live at a time); restated in oracle/philox.py
// (sigma_grouped_normals).
struct SigmaNoise {
    float q[4];
    uint32_t sel[4];
    uint32_t lo, hi, step, k0, k1;
    __device__ __forceinline__ SigmaNoise(uint64_t index, uint32_t step_, uint32_t k0_, uint32_t k1_)
        : lo((uint32_t)index), hi((uint32_t)(index >> 32)), step(step_), k0(k0_), k1(k1_) {}
    // S = sigma point (compile-time after unrolling)
    template <bool DIAG, int ND>
    __device__ __forceinline__ void draw(int S, const MixSampler5& sp, float out[5]) {
        float z[5];
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int k = 5 * S + t;                              // index of the normal in the component's stream
            if ((k & 3) == 0) {
                const Philox4 P = philox4x32_10(lo, hi, step, 0x40000000u + (uint32_t)(k >> 2), k0, k1);
                box_muller(P.x, P.y, q[0], q[1]);
                box_muller(P.z, P.w, q[2], q[3]);
            }
            z[t] = q[k & 3];
        }
        if ((S & 3) == 0) {
            const Philox4 P = philox4x32_10(lo, hi, step, 0x40000000u + 14u + (uint32_t)(S >> 2), k0, k1);
            sel[0] = P.x; sel[1] = P.y; sel[2] = P.z; sel[3] = P.w;
        }
        mix_apply<DIAG, ND>(sp, z, u32_to_unit(sel[S & 3]), out);
    }
};

// State noise of FOUR consecutive rows (a group: global rows 4 g .. 4 g + 3) from six Philox calls instead
// of eight: 20 normals = 10 Box-Muller pairs (words 0..19) and 4 component selectors (words 20..23).
//   P_j = philox(ctr = (g_lo, g_hi, step, 0x80000000 + j), key),  j = 0..5;  words w[4 j + {0,1,2,3}] = P_j.{x,y,z,w}
//   (n[2 p], n[2 p + 1]) = BM(w[2 p], w[2 p + 1]),  p = 0..9;  row r of the group takes n[5 r .. 5 r + 4] and
//   selector w[20 + r]
// Still a pure function of the GLOBAL row index (restated in oracle/philox.py: grouped_normals5).
// the four component selectors of a group (Philox call 5 of the grouped layout)
__device__ __forceinline__ Philox4 draw_selectors_x4(uint64_t group, uint32_t step, uint32_t k0, uint32_t k1) {
    return philox4x32_10((uint32_t)group, (uint32_t)(group >> 32), step, 0x80000005u, k0, k1);
}

template <bool DIAG, int ND>
__device__ __forceinline__ void draw_mixture5_x4(const MixSampler5& sp, uint64_t group, uint32_t step, uint32_t k0,
                                                 uint32_t k1, const Philox4& S, float out[4][5]) {
    // every Philox call is turned into its four normals at once (its words die there): 20 normals + 4 selectors live
    float z[20];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const Philox4 P = philox4x32_10((uint32_t)group, (uint32_t)(group >> 32), step, 0x80000000u + j, k0, k1);
        box_muller(P.x, P.y, z[4 * j], z[4 * j + 1]);
        box_muller(P.z, P.w, z[4 * j + 2], z[4 * j + 3]);
    }
    const uint32_t sel[4] = {S.x, S.y, S.z, S.w};
    const int dg[5] = {0, 2, 5, 9, 14};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float uc = u32_to_unit(sel[r]);
        int comp = 0;
        if (ND != 1) {
#pragma unroll
            for (int d = 0; d < GSE_MAX_ND - 1; ++d)
                comp += ((ND == 0 ? d < sp.nd - 1 : d < ND - 1) && uc > sp.cdf[d]) ? 1 : 0;
        }
        if (ND == 2 && DIAG) {
            const bool second = comp != 0;
#pragma unroll
            for (int j = 0; j < 5; ++j)
                out[r][j] = fmaf(second ? sp.L[1][dg[j]] : sp.L[0][dg[j]], z[5 * r + j],
                                 second ? sp.mean[1][j] : sp.mean[0][j]);
        } else if (DIAG) {
#pragma unroll
            for (int j = 0; j < 5; ++j) out[r][j] = fmaf(sp.L[comp][dg[j]], z[5 * r + j], sp.mean[comp][j]);
        } else {
            int t = 0;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                float acc = sp.mean[comp][j];
#pragma unroll
                for (int m = 0; m <= j; ++m) acc = fmaf(sp.L[comp][t++], z[5 * r + m], acc);
                out[r][j] = acc;
            }
        }
    }
}

template <bool DIAG, int ND>
__device__ __forceinline__ void draw_mixture5_x4(const MixSampler5& sp, uint64_t group, uint32_t step, uint32_t k0,
                                                 uint32_t k1, float out[4][5]) {
    draw_mixture5_x4<DIAG, ND>(sp, group, step, k0, k1, draw_selectors_x4(group, step, k0, k1), out);
}

// ------------------------------------------------------------------------------------------------
// Bioreactor model (model/BioreactorModel.py:170-253), float32, hard-coded.
// ------------------------------------------------------------------------------------------------
struct ModelInputs {
    float feed;    // Fg_in * Cg_in           (:196,225)
    float f_out;   // Fg_in + Fm_in           (:197)
    float dt;
};

// per-step scalars: by value, or from the device block when one is attached to the context
__device__ __forceinline__ ModelInputs model_inputs(ModelInputs in, const gse_step_params* sp, int n_sub) {
    if (sp) {
        in.feed = (float)(sp->u[0] * (5000.0 / 180.0));       // the host's expressions (gse_pf_predict), bit for bit
        in.f_out = (float)(sp->u[0] + sp->u[1]);
        in.dt = (float)(sp->dt / n_sub);
    }
    return in;
}

__device__ __forceinline__ void bioreactor_increment(const float x[5], const ModelInputs& in, float d[5]) {
    // :192-193 -- the rate expressions see clamped Cg, Cx, Cfa, Ce; Ch is not clamped
    const float Cg = fmaxf(x[0], 0.0f), Cx = fmaxf(x[1], 0.0f), Cfa = fmaxf(x[2], 0.0f),
                Ce = fmaxf(x[3], 0.0f), Ch = x[4];
    const float K_FA = (float)(0.25 / 116 * 24.6);            // :205
    const float K_T1 = (float)((0.4 - 0.25) / 180 * 24.6);    // :209
    const float K_E = (float)(0.025 / 46 * 24.6);             // :214
    const float K_T2 = (float)((0.1 - 0.025) / 180 * 24.6);   // :219
    const float K_H = (float)(1.0 / 2000 / (0.28 / 180));     // :210
    const float rH = (float)(280.0 / 180) - Cg;               // :202
    const float den = 1e-2f + Cg;
    float rc = mufu_rcp(den);
    rc = fmaf(rc, fmaf(-den, rc, 1.0f), rc);                  // one Newton step: <= 1 ulp
    const float sat = Cg * rc;                                // Cg / (1e-2 + Cg)   :206,211
    const float rFA = K_FA * Cx * sat;                        // :206
    const float t1max = K_T1 * Cx;                            // :209
    const float t1req = t1max - fmaf(t1max * K_H, rH, 0.01f * Ch);          // :210
    const float t1 = fminf(t1max, fmaxf(0.0f, t1req)) * sat;                // :211
    const float over = t1req - t1max;                                      // :215
    const float rE = fminf(K_E * Cx, fmaxf(0.0f, over));                    // :216
    const float t2 = fminf(K_T2 * Cx, fmaxf(0.0f, over - rE));              // :219-221
    const float rG = -rFA * (float)(116.0 / 180) - t1 - rE * (float)(46.0 / 180) - t2;   // :223
    d[0] = (in.feed - in.f_out * Cg + rG) * in.dt;            // :225
    d[1] = 0.0f;                                              // :226 (rX = 0 * Cx)
    d[2] = (rFA - in.f_out * Cfa) * in.dt;                    // :227
    d[3] = (rE - in.f_out * Ce) * in.dt;                      // :228
    d[4] = rH * in.dt;                                        // :229
}

// static_outputs (:233-253): float32 products, as the reference produces them (see oracle/bioreactor.py)
__device__ __forceinline__ float output_glucose(float Cg) { return __fmul_rn(Cg, 180.0f); }
__device__ __forceinline__ float output_fa(float Cfa) { return __fmul_rn(Cfa, 116.0f); }

// log pdf of the measurement mixture at e = z - y  (MultivariateGaussianSum.py:39-63), quadratic
// forms in float64, the log-sum-exp tail in float32 (terms <= 1, so the absolute error of the
// result is ~1e-7).
__device__ __forceinline__ double meas_logpdf(const MixDensity2& md, double e0, double e1) {
    double a[GSE_MAX_ND];
    double m = -1.0e300;
#pragma unroll
    for (int d = 0; d < GSE_MAX_ND; ++d) {
        if (d < md.nd) {
            const double v0 = e0 - md.mean[d][0], v1 = e1 - md.mean[d][1];
            const double q = v0 * (md.p00[d] * v0 + md.p01[d] * v1) + md.p11[d] * v1 * v1;
            a[d] = md.logc[d] - 0.5 * q;
            m = fmax(m, a[d]);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int d = 0; d < GSE_MAX_ND; ++d)
        if (d < md.nd) s += fast_exp((float)(a[d] - m));
    return m + (double)(mufu_lg2(s) * 0.69314718055994531f);
}

// float32 variant for the particle-filter update (K2).  e = z - y is formed from a float32 hi/lo
// split of z: (z_hi - y) is exact (Sterbenz) whenever the innovation is small against the output,
// so e carries a relative error of ~6e-8 and the log-density an error of ~3e-7 * max(1, |log pdf|)
// -- inside the stated 1e-6 * max(1, |log w|) tolerance -- at a third of the issue slots of the
// float64 form (K2 is instruction-issue bound, not HBM bound, in float64).
template <int ND>                 // ND = 0: component count read from the density at run time
__device__ __forceinline__ float meas_logpdf32(const MixDensity2f& md, float e0, float e1) {
    if (ND > 0) {
        float a[ND > 0 ? ND : 1];
        float m = -3.0e38f;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const float v0 = e0 - md.mean[d][0], v1 = e1 - md.mean[d][1];
            const float q = fmaf(v0, fmaf(md.p00[d], v0, md.p01[d] * v1), md.p11[d] * v1 * v1);
            a[d] = fmaf(-0.5f, q, md.logc[d]);
            m = fmaxf(m, a[d]);
        }
        if (ND == 1) return a[0];
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d < ND; ++d) s += fast_exp(a[d] - m);
        return fmaf(mufu_lg2(s), 0.69314718055994531f, m);
    }
    float a[GSE_MAX_ND];
    float m = -3.0e38f;
#pragma unroll
    for (int d = 0; d < GSE_MAX_ND; ++d) {
        if (d < md.nd) {
            const float v0 = e0 - md.mean[d][0], v1 = e1 - md.mean[d][1];
            const float q = fmaf(v0, fmaf(md.p00[d], v0, md.p01[d] * v1), md.p11[d] * v1 * v1);
            a[d] = fmaf(-0.5f, q, md.logc[d]);
            m = fmaxf(m, a[d]);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int d = 0; d < GSE_MAX_ND; ++d)
        if (d < md.nd) s += fast_exp(a[d] - m);
    return fmaf(mufu_lg2(s), 0.69314718055994531f, m);
}

// ------------------------------------------------------------------------------------------------
// warp / block reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// block-level (max, sum exp) reduction shared by K2 and gse_loglik_max.
// Each block publishes (m_b, s_b = sum exp(v - m_b)); the last block to finish merges them in a
// fixed order into stats[0] = M = max, stats[1] = S = sum exp(v - M).
// ------------------------------------------------------------------------------------------------
// Online (max, sum exp) accumulator: after add(v...) the pair (m, s) satisfies
// s = sum_k exp(v_k - m), m = max_k v_k.  One rescale per group of values, not per value.
struct MaxSumExp {
    float m, s;
    __device__ __forceinline__ MaxSumExp() : m(-INFINITY), s(0.0f) {}
    template <int NV>
    __device__ __forceinline__ void add_all(const float vals[NV]) {
        float g = vals[0];
#pragma unroll
        for (int r = 1; r < NV; ++r) g = fmaxf(g, vals[r]);
        const float mn = fmaxf(m, g);
        if (mn > -INFINITY) {
            float t = s * fast_exp(m - mn);          // m = -inf: exp(-inf) = 0, s = 0
#pragma unroll
            for (int r = 0; r < NV; ++r) t += fast_exp(vals[r] - mn);
            s = t;
            m = mn;
        }
    }
    template <int NV>
    __device__ __forceinline__ void add(const float vals[NV], const bool valid[NV]) {
        float g = -INFINITY;
#pragma unroll
        for (int r = 0; r < NV; ++r) if (valid[r]) g = fmaxf(g, vals[r]);
        const float mn = fmaxf(m, g);
        if (mn > -INFINITY) {
            float t = s * fast_exp(m - mn);          // m = -inf: exp(-inf) = 0, s = 0
#pragma unroll
            for (int r = 0; r < NV; ++r) if (valid[r]) t += fast_exp(vals[r] - mn);
            s = t;
            m = mn;
        }
    }
};

template <int THREADS>
__device__ __forceinline__ void block_merge_max_sumexp(float m_t, float s_t, float* block_max, float* block_sum,
                                                       unsigned int* ticket, double* stats) {
    __shared__ float s_red[THREADS / 32];
    __shared__ float s_bcast;
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float m = warp_max(m_t);
    if (lane == 0) s_red[wid] = m;
    __syncthreads();
    if (wid == 0) {
        float t = (lane < THREADS / 32) ? s_red[lane] : -INFINITY;
        t = warp_max(t);
        if (lane == 0) s_bcast = t;
    }
    __syncthreads();
    const float bm = s_bcast;
    float sum = (m_t > -INFINITY) ? s_t * fast_exp(m_t - bm) : 0.0f;
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) s_red[wid] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) t += s_red[w];
        block_max[blockIdx.x] = bm;
        block_sum[blockIdx.x] = t;
        __threadfence();
        const unsigned int done = atomicAdd(ticket, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: merge all partials (fixed order per thread, then a fixed tree)
    __shared__ double s_dred[THREADS / 32];
    float gm = -INFINITY;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += THREADS) gm = fmaxf(gm, __ldcg(block_max + b));
    gm = warp_max(gm);
    __syncthreads();
    if (lane == 0) s_red[wid] = gm;
    __syncthreads();
    if (wid == 0) {
        float t = (lane < THREADS / 32) ? s_red[lane] : -INFINITY;
        t = warp_max(t);
        if (lane == 0) s_bcast = t;
    }
    __syncthreads();
    const float M = s_bcast;
    double acc = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += THREADS) {
        const float bmx = __ldcg(block_max + b);
        if (bmx > -INFINITY) acc += (double)__ldcg(block_sum + b) * (double)fast_exp(bmx - M);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_dred[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) t += s_dred[w];
        stats[0] = (double)M;
        stats[1] = t;
        *ticket = 0u;
    }
}

template <int THREADS, int NV>
__device__ __forceinline__ void block_max_sumexp_finalize(const float vals[NV], const bool valid[NV],
                                                          float* block_max, float* block_sum,
                                                          unsigned int* ticket, double* stats) {
    MaxSumExp acc;
    acc.add<NV>(vals, valid);
    block_merge_max_sumexp<THREADS>(acc.m, acc.s, block_max, block_sum, ticket, stats);
}

__device__ __forceinline__ unsigned int ld_status32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The shard table arrives as a kernel parameter (__grid_constant__: its address can be taken).  Nearly every ancestor a
// rank reads is one of its own rows -- the shards' weight totals are balanced -- so the table's "home" fields are tested
// first (two compares); the search through the other shards sits in a function of its own so that it costs the hot path
// no instructions and no registers.  (Inlined as a chain of selects it cost every thread ~60 instructions: +18 us on
// the 2^24-row predict.  Indexing the parameter array in line makes the compiler copy the struct to local memory;
// staging it in shared memory puts a block barrier in front of every short-lived CTA: +30 us.)
struct ShardRef {
    const float* state;    // column 0, row 0 of the owning shard's buffer
    int64_t ld;
    int64_t row0, row1;    // the shard holds global rows [row0, row1)
};
static __device__ __noinline__ ShardRef shard_ref_search(const GatherShards* g, int64_t k) {
    int t = 0;
    for (int u = 1; u < g->nseg; ++u) t += (k >= g->seg_row[u]) ? 1 : 0;
    ShardRef r;
    r.state = g->state[t]; r.ld = g->ld[t]; r.row0 = g->seg_row[t]; r.row1 = g->seg_row[t + 1];
    return r;
}
__device__ __forceinline__ ShardRef shard_ref(const GatherShards& g, int64_t k) {
    ShardRef r;
    r.state = g.home_state; r.ld = g.home_ld; r.row0 = g.home_row0; r.row1 = g.home_row1;
    if (k < r.row0 || k >= r.row1) r = shard_ref_search(&g, k);
    return r;
}
// column 0 of global row k and the leading dimension of the shard that owns it
__device__ __forceinline__ const float* shard_row(const GatherShards& g, int64_t k, int64_t& ld) {
    const ShardRef r = shard_ref(g, k);
    ld = r.ld;
    return r.state + (k - r.row0);
}

// four rows with non-decreasing global indices: one look-up serves all four unless they straddle a shard boundary
__device__ __forceinline__ void shard_rows4(const GatherShards& g, const int id[4], const float* q[4], int64_t l[4]) {
    if (id[0] >= g.home_lo && id[3] < g.home_hi) {
#pragma unroll
        for (int r = 0; r < 4; ++r) { q[r] = g.home_base + id[r]; l[r] = g.home_ld; }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) q[r] = shard_row(g, id[r], l[r]);
    }
}

// streaming 128-bit accesses for touch-once columns
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------
// Exact systematic-resampling thresholds (host + device).
//
// The reference compares  cumsum[k] / cumsum[-1] < (i + r) / N  in float64 (particle.py:89-98).
// Here cumulative weights are exact integers C <= T < 2^53 (so fl(C), fl(T) are exact) and the
// comparison value is g(C) = fl(C / T), a monotone function of C.  Hence
//     g(C) >= u   <=>   C >= q*(u),     q*(u) = min{ C : fl(C / T) >= u }.
// fl(x) >= u  <=>  x >= m, m = (pred(u) + u) / 2 the rounding boundary below u (ties to even: a
// quotient exactly on the boundary rounds to u iff u's significand is even).  With u = Mu * 2^eu:
// m = A * 2^ea, A = 2 Mu - 1 (4 Mu - 1 when u is a power of two), so
//     q* = ceil(A * T * 2^ea)     (+1 on an exact tie with Mu odd)
// evaluated with one 64 x 64 -> 128-bit product and a shift: no floating-point division at all.
// tests/test_thresholds.py checks g(q*-1) < u <= g(q*) against Python integers / numpy float64.
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#define GSE_HD __host__ __device__ __forceinline__
#else
#define GSE_HD static inline
#endif

#define GSE_TOTAL_BITS 52      /* the quantised weights sum to T <= 2^52 + n/2 < 2^53 */

GSE_HD uint64_t gse_double_to_bits(double d) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b; memcpy(&b, &d, 8); return b;
#endif
}
GSE_HD double gse_u64_to_double(uint64_t c) {
#ifdef __CUDA_ARCH__
    return __ull2double_rn(c);
#else
    return (double)c;   // round-to-nearest-even on x86-64
#endif
}
GSE_HD void gse_mul128(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    lo = (uint64_t)p;
    hi = (uint64_t)(p >> 64);
#endif
}

// u_i = (i + r) / N exactly as the reference evaluates it (particle.py:97); di = (double)i
GSE_HD double gse_sample_position(double di, double r, double n_total, double inv_n, bool n_pow2) {
#ifdef __CUDA_ARCH__
    const double s = __dadd_rn(di, r);
    return n_pow2 ? __dmul_rn(s, inv_n) : __ddiv_rn(s, n_total);
#else
    const double s = di + r;
    return n_pow2 ? s * inv_n : s / n_total;
#endif
}

// q*(u): smallest integer C with fl(C / T) >= u, for 0 <= u <= 1 and 1 <= T < 2^53.
GSE_HD uint64_t gse_threshold(double u, uint64_t T) {
    if (!(u > 0.0)) return 0;
    const uint64_t bits = gse_double_to_bits(u);
    const uint64_t frac = bits & 0xfffffffffffffull;
    const int ebits = (int)(bits >> 52);                 // u > 0: sign bit clear
    if (ebits == 0) return 1;                            // subnormal u: any C >= 1 qualifies (1/T >> u)
    const uint64_t Mu = frac | (1ull << 52);
    const bool pow2 = (frac == 0);
    const uint64_t A = pow2 ? (Mu << 2) - 1 : (Mu << 1) - 1;      // m = A * 2^ea
    const int s = (pow2 ? 1077 : 1076) - ebits;                    // s = -ea = -(eu - 1 [-1]), eu = ebits - 1075
    uint64_t hi, lo;
    gse_mul128(A, T, hi, lo);
    uint64_t c0;
    bool rem;
    if (s >= 128) {
        c0 = 0;
        rem = true;                                      // A * T > 0
    } else if (s >= 64) {
        const int t = s - 64;
        c0 = hi >> t;
        rem = (lo != 0) || ((t > 0) && ((hi & ((1ull << t) - 1)) != 0));
    } else {                                             // 53 <= s < 64 (u in (1/2, 1]... down to 2^-10)
        c0 = (hi << (64 - s)) | (lo >> s);
        rem = (lo & ((1ull << s) - 1)) != 0;
    }
    return c0 + ((rem || (Mu & 1ull)) ? 1 : 0);
}
