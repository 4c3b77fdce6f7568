// Context management, error reporting and host-side mixture preprocessing for libgse_b200.so.
#include <stdarg.h>
#include <stdlib.h>

#include "gse_common.cuh"

static thread_local char g_err[512] = "";

void gse_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* gse_last_error(void) { return g_err; }
extern "C" int gse_abi_version(void) { return GSE_ABI_VERSION; }

// ---- small dense helpers (double, row-major n x n, n <= 5) -------------------------------------
static int chol_lower(const double* A, int n, double* L) {
    for (int i = 0; i < n * n; ++i) L[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        double d = A[j * n + j];
        for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0.0)) return -1;
        L[j * n + j] = sqrt(d);
        for (int i = j + 1; i < n; ++i) {
            double s = A[i * n + j];
            for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = s / L[j * n + j];
        }
    }
    return 0;
}

// inverse and determinant by Gauss-Jordan with partial pivoting
static int inv_det(const double* A, int n, double* inv, double* det) {
    double a[25], b[25];
    for (int i = 0; i < n * n; ++i) { a[i] = A[i]; b[i] = 0.0; }
    for (int i = 0; i < n; ++i) b[i * n + i] = 1.0;
    double dt = 1.0;
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r) if (fabs(a[r * n + c]) > fabs(a[p * n + c])) p = r;
        if (a[p * n + c] == 0.0) return -1;
        if (p != c) {
            for (int k = 0; k < n; ++k) {
                double t = a[c * n + k]; a[c * n + k] = a[p * n + k]; a[p * n + k] = t;
                t = b[c * n + k]; b[c * n + k] = b[p * n + k]; b[p * n + k] = t;
            }
            dt = -dt;
        }
        const double piv = a[c * n + c];
        dt *= piv;
        for (int k = 0; k < n; ++k) { a[c * n + k] /= piv; b[c * n + k] /= piv; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = a[r * n + c];
            if (f == 0.0) continue;
            for (int k = 0; k < n; ++k) { a[r * n + k] -= f * a[c * n + k]; b[r * n + k] -= f * b[c * n + k]; }
        }
    }
    for (int i = 0; i < n * n; ++i) inv[i] = b[i];
    *det = dt;
    return 0;
}

static int check_mixture(const gse_mixture* m, int nx) {
    GSE_REQUIRE(m != NULL, "mixture is NULL");
    GSE_REQUIRE(m->nd >= 1 && m->nd <= GSE_MAX_ND, "mixture nd out of range");
    GSE_REQUIRE(m->nx == nx, "mixture nx mismatch");
    return GSE_OK;
}

// parameters are stored float32 by the reference (MultivariateGaussianSum.py:29-31)
static double f32(double v) { return (double)(float)v; }

// Sampler of an nx-dimensional mixture, nx <= 5: mean and Cholesky factor are padded to five dimensions (zero mean, zero
// factor rows), the kernels write the first nx columns.
int gse_build_sampler(const gse_mixture* m, MixSampler5* out, int* nx_out) {
    GSE_REQUIRE(m != NULL, "mixture is NULL");
    GSE_REQUIRE(m->nd >= 1 && m->nd <= GSE_MAX_ND, "mixture nd out of range");
    GSE_REQUIRE(m->nx >= 1 && m->nx <= GSE_NX, "mixture nx out of range");
    const int nx = m->nx;
    memset(out, 0, sizeof(*out));
    out->nd = m->nd;
    double wsum = 0.0;
    for (int d = 0; d < m->nd; ++d) wsum += f32(m->weights[d]);
    GSE_REQUIRE(wsum > 0.0, "mixture weights sum to zero");
    double acc = 0.0;
    int diag = 1;
    for (int d = 0; d < m->nd; ++d) {
        acc += f32(m->weights[d]) / wsum;
        out->cdf[d] = (float)acc;
        double C[25], L[25];
        for (int i = 0; i < nx * nx; ++i) C[i] = f32(m->covs[d * nx * nx + i]);
        if (chol_lower(C, nx, L) != 0) {
            gse_set_error("mixture component %d covariance is not positive definite", d);
            return GSE_ELINALG;
        }
        int t = 0;
        for (int i = 0; i < 5; ++i) {
            out->mean[d][i] = i < nx ? (float)m->means[d * nx + i] : 0.0f;
            for (int j = 0; j <= i; ++j) {
                const double v = (i < nx) ? L[i * nx + j] : 0.0;
                out->L[d][t++] = (float)v;
                if (i != j && v != 0.0) diag = 0;
            }
        }
    }
    out->cdf[m->nd - 1] = 1.0f;
    for (int d = m->nd; d < GSE_MAX_ND; ++d) out->cdf[d] = 2.0f;
    out->diag = diag;
    if (nx_out) *nx_out = nx;
    return GSE_OK;
}

int gse_build_sampler5(const gse_mixture* m, MixSampler5* out) {
    int rc = check_mixture(m, GSE_NX);
    if (rc) return rc;
    return gse_build_sampler(m, out, NULL);
}

int gse_build_densityN(const gse_mixture* m, MixDensityN* out) {
    GSE_REQUIRE(m != NULL, "mixture is NULL");
    GSE_REQUIRE(m->nd >= 1 && m->nd <= GSE_MAX_ND, "mixture nd out of range");
    GSE_REQUIRE(m->nx >= 1 && m->nx <= GSE_NX, "mixture nx out of range");
    memset(out, 0, sizeof(*out));
    const int n = m->nx;
    out->nd = m->nd;
    out->nx = n;
    for (int d = 0; d < m->nd; ++d) {
        double C[25], C32[25], P[25], P32[25], det, det32;
        for (int i = 0; i < n * n; ++i) { C[i] = m->covs[d * n * n + i]; C32[i] = f32(C[i]); }
        // inverse of the float64 input (MultivariateGaussianSum.py:33); constant from the float32
        // copy (:36-37)
        if (inv_det(C, n, P, &det) != 0 || inv_det(C32, n, P32, &det32) != 0 || !(det32 > 0.0)) {
            gse_set_error("mixture component %d covariance is singular", d);
            return GSE_ELINALG;
        }
        const double cst = pow(2.0 * M_PI, -0.5 * n) / sqrt(det32);
        const double w = f32(m->weights[d]);
        out->logc[d] = (w > 0.0) ? log(w * cst) : -1.0e300;
        for (int i = 0; i < n; ++i) out->mean[d][i] = f32(m->means[d * n + i]);
        for (int i = 0; i < n * n; ++i) out->P[d][i] = P[i];
    }
    return GSE_OK;
}

int gse_build_density2(const gse_mixture* m, MixDensity2* out) {
    int rc = check_mixture(m, GSE_NY);
    if (rc) return rc;
    MixDensityN g;
    rc = gse_build_densityN(m, &g);
    if (rc) return rc;
    memset(out, 0, sizeof(*out));
    out->nd = g.nd;
    for (int d = 0; d < g.nd; ++d) {
        out->logc[d] = g.logc[d];
        out->mean[d][0] = g.mean[d][0];
        out->mean[d][1] = g.mean[d][1];
        out->p00[d] = g.P[d][0];
        out->p01[d] = g.P[d][1] + g.P[d][2];
        out->p11[d] = g.P[d][3];
    }
    return GSE_OK;
}

// ---- context ------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int gse_ctx_create(int device, int model_id, int64_t n_max, const gse_mixture* state,
                              const gse_mixture* meas, gse_ctx** out) {
    GSE_REQUIRE(out != NULL, "out is NULL");
    GSE_REQUIRE(model_id == GSE_MODEL_BIOREACTOR, "unknown model_id (only GSE_MODEL_BIOREACTOR is compiled in)");
    GSE_REQUIRE(n_max >= 1 && n_max <= ((int64_t)1 << 40), "n_max out of range");
    int ndev = 0;
    GSE_CHECK_CUDA(cudaGetDeviceCount(&ndev));
    GSE_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    gse_device_guard guard(device);
    cudaDeviceProp prop;
    GSE_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        gse_set_error("libgse_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        return GSE_ECUDA;
    }
    gse_ctx* c = (gse_ctx*)calloc(1, sizeof(gse_ctx));
    if (!c) { gse_set_error("out of host memory"); return GSE_ENOMEM; }
    c->device = device;
    c->model_id = model_id;
    c->n_max = n_max;
    c->num_sms = prop.multiProcessorCount;
    int rc = gse_build_sampler5(state, &c->state_sampler);
    if (rc == GSE_OK) rc = gse_build_density2(meas, &c->meas_density);
    if (rc != GSE_OK) { free(c); return rc; }
    c->meas_density32.nd = c->meas_density.nd;
    for (int d = 0; d < c->meas_density.nd; ++d) {
        c->meas_density32.logc[d] = (float)c->meas_density.logc[d];
        c->meas_density32.mean[d][0] = (float)c->meas_density.mean[d][0];
        c->meas_density32.mean[d][1] = (float)c->meas_density.mean[d][1];
        c->meas_density32.p00[d] = (float)c->meas_density.p00[d];
        c->meas_density32.p01[d] = (float)c->meas_density.p01[d];
        c->meas_density32.p11[d] = (float)c->meas_density.p11[d];
    }

    // workspace: sized for the largest grid any kernel uses on n_max rows
    c->max_blocks = gse_div_up(n_max, 128) + 8;        // >= blocks of any kernel (GS-UKF: 128 components per block)
    c->max_tiles = gse_div_up(n_max, 512) + 16;        // scan warp runs (<= n/512 + 1) / merge partitions (2n/4096 + 1)
    size_t off = 0;
    const size_t o_bmax = off; off = align_up(off + sizeof(float) * c->max_blocks, 256);
    const size_t o_bsum = off; off = align_up(off + sizeof(float) * c->max_blocks, 256);
    const size_t o_tick = off; off = align_up(off + sizeof(unsigned int) * 16, 256);
    const size_t o_red = off; off = align_up(off + sizeof(double) * 48 * 2048, 256);
    const size_t o_agg = off; off = align_up(off + sizeof(uint64_t) * c->max_tiles, 256);
    const size_t o_inc = off; off = align_up(off + sizeof(uint64_t) * c->max_tiles, 256);
    const size_t o_status = off; off = align_up(off + sizeof(uint64_t) * c->max_tiles, 256);
    const size_t o_part = off; off = align_up(off + sizeof(int64_t) * (c->max_tiles + 2), 256);
    const size_t o_range = off; off = align_up(off + sizeof(int64_t) * 8, 256);    // [0..1] range, [4] 1/T
    // fused resample: one status word per co-resident CTA (<= 32 per SM); queue of heavy runs -- every entry covers
    // more than 4096 outputs of its own, cut into pieces of <= 65536 (gse_resample_fused.cu)
    const size_t n_fused_status = 2 * 4096 + 8;       // aggregates + prefixes (gse_resample_fused.cu: RF_MAX_BLOCKS)
    c->heavy_queue_cap = (int)(n_max / 4096 + n_max / 65536 + 64);
    const size_t o_fstatus = off; off = align_up(off + sizeof(uint64_t) * n_fused_status, 256);
    const size_t o_queue = off; off = align_up(off + sizeof(int4) * (size_t)c->heavy_queue_cap, 256);
    c->ws_bytes = off;
    cudaError_t e = cudaMalloc(&c->ws, c->ws_bytes);
    if (e != cudaSuccess) {
        gse_set_error("cudaMalloc(%zu) for the workspace failed: %s", c->ws_bytes, cudaGetErrorString(e));
        free(c);
        return GSE_ENOMEM;
    }
    e = cudaMemset(c->ws, 0, c->ws_bytes);
    if (e != cudaSuccess) {
        gse_set_error("cudaMemset failed: %s", cudaGetErrorString(e));
        cudaFree(c->ws);
        free(c);
        return GSE_ECUDA;
    }
    char* base = (char*)c->ws;
    c->block_max = (float*)(base + o_bmax);
    c->block_sum = (float*)(base + o_bsum);
    c->ticket = (unsigned int*)(base + o_tick);
    c->red_partials = (double*)(base + o_red);
    c->tile_agg = (uint64_t*)(base + o_agg);
    c->tile_inc = (uint64_t*)(base + o_inc);
    c->tile_status = (uint64_t*)(base + o_status);
    c->scan_tiles_prev = 0;
    {
        const char* v = getenv("GSE_UPDATE_CTAS");
        c->update_ctas_per_sm = (v && atoi(v) >= 1 && atoi(v) <= 32) ? atoi(v) : 5;
    }
    {
        const char* v = getenv("GSE_UPDATE_PIPE");
        c->update_pipe = (v && atoi(v) == 0) ? 0 : 1;
    }
    c->scan_resident_blocks = 0;
    {
        const char* mode = getenv("GSE_SCAN");
        c->scan_single_pass = !(mode && strcmp(mode, "twopass") == 0);
    }
    c->part = (int64_t*)(base + o_part);
    c->range = (int64_t*)(base + o_range);
    c->fused_status = (uint64_t*)(base + o_fstatus);
    c->heavy_queue = (int4*)(base + o_queue);
    {
        const char* v = getenv("GSE_GSF_MINB");
        c->gsf_minb = (v && (atoi(v) == 3 || atoi(v) == 4 || atoi(v) == 5 || atoi(v) == 6)) ? atoi(v) : 0;    // 0: per-kernel defaults
    }
    {
        const char* v = getenv("GSE_GSF_UPDATE_WAVES");           // tuning knob: grid of the GS-UKF update kernel = waves x resident CTAs
        c->gsf_update_waves = (v && atoi(v) >= 1 && atoi(v) <= 64) ? atoi(v) : 1;
    }
    {
        const char* v = getenv("GSE_PREDICT_MINB");               // tuning knob: CTAs per SM of the predict kernel
        c->predict_minb = (v && atoi(v) == 5) ? 5 : 4;
    }
    {
        const char* v = getenv("GSE_FUSED_MINB");                 // 3 CTAs per SM (80 registers, no spills) measured 70 us against 72 us at 4 (64 registers); GSE_FUSED_MINB=4 to compare
        c->fused_minb = (v && atoi(v) == 4) ? 4 : 3;
    }
    if (getenv("GSE_FUSED_TRACE")) {
        if (cudaMalloc((void**)&c->fused_trace, sizeof(unsigned long long) * 8 * 4096) != cudaSuccess) c->fused_trace = NULL;
        else cudaMemset(c->fused_trace, 0, sizeof(unsigned long long) * 8 * 4096);
    }
    // device-error word: host-mapped so that reading it never needs a copy or a synchronisation of its own
    // (the same block holds the result area of gse_ctx_result_block: error word at byte 0, 64 doubles at byte 64)
    e = cudaHostAlloc((void**)&c->err_host, 64 + 64 * sizeof(double), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        memset(c->err_host, 0, 64 + 64 * sizeof(double));
        e = cudaHostGetDevicePointer((void**)&c->err_dev, c->err_host, 0);
    }
    if (e == cudaSuccess) {
        c->result_host = (double*)((char*)c->err_host + 64);
        c->result_dev = (double*)((char*)c->err_dev + 64);
    }
    if (e != cudaSuccess) {
        gse_set_error("allocating the device-error word failed: %s", cudaGetErrorString(e));
        if (c->err_host) cudaFreeHost(c->err_host);
        cudaFree(c->ws);
        free(c);
        return GSE_ECUDA;
    }
    *out = c;
    return GSE_OK;
}

extern "C" int gse_ctx_destroy(gse_ctx* ctx) {
    if (!ctx) return GSE_OK;
    gse_device_guard guard(ctx->device);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->err_host) cudaFreeHost(ctx->err_host);
    if (ctx->fused_trace) cudaFree(ctx->fused_trace);
    if (ctx->params_block) {
        cudaFree(ctx->params_block);
        cudaFreeHost(ctx->params_ring);
        for (int q = 0; q < 4; ++q) cudaEventDestroy(ctx->params_event[q]);
    }
    free(ctx);
    return GSE_OK;
}

// ---- peer memory ----------------------------------------------------------------------------------
extern "C" int gse_peer_alloc(int device, int64_t bytes, void** ptr_out, unsigned char handle_out[GSE_IPC_HANDLE_BYTES]) {
    GSE_REQUIRE(ptr_out != NULL && handle_out != NULL && bytes > 0, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == GSE_IPC_HANDLE_BYTES, "IPC handle size");
    gse_device_guard guard(device);
    void* p = NULL;
    GSE_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        gse_set_error("peer allocation of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
        cudaFree(p);
        return GSE_ECUDA;
    }
    memcpy(handle_out, &h, GSE_IPC_HANDLE_BYTES);
    *ptr_out = p;
    return GSE_OK;
}

extern "C" int gse_peer_free(int device, void* ptr) {
    if (!ptr) return GSE_OK;
    gse_device_guard guard(device);
    GSE_CHECK_CUDA(cudaFree(ptr));
    return GSE_OK;
}

extern "C" int gse_peer_open(int device, const unsigned char handle[GSE_IPC_HANDLE_BYTES], void** ptr_out) {
    GSE_REQUIRE(ptr_out != NULL && handle != NULL, "bad arguments");
    gse_device_guard guard(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, GSE_IPC_HANDLE_BYTES);
    GSE_CHECK_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return GSE_OK;
}

extern "C" int gse_peer_close(int device, void* ptr) {
    if (!ptr) return GSE_OK;
    gse_device_guard guard(device);
    GSE_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return GSE_OK;
}

extern "C" int gse_ctx_upload_step_params(gse_ctx* ctx, const gse_step_params* values, void* stream) {
    GSE_REQUIRE(ctx != NULL && values != NULL, "ctx / values is NULL");
    gse_device_guard guard(ctx->device);
    if (!ctx->params_block) {
        GSE_CHECK_CUDA(cudaMalloc((void**)&ctx->params_block, sizeof(gse_step_params)));
        GSE_CHECK_CUDA(cudaHostAlloc((void**)&ctx->params_ring, sizeof(gse_step_params) * GSE_PARAM_RING, cudaHostAllocDefault));
        for (int q = 0; q < 4; ++q) GSE_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->params_event[q], cudaEventDisableTiming));
        ctx->params_pos = 0;
    }
    const int pos = ctx->params_pos;
    const int quarter = GSE_PARAM_RING / 4;
    if (pos % quarter == 0) {
        // entering a quarter: every copy queued from it a lap ago must have run (a host far ahead of the GPU)
        GSE_CHECK_CUDA(cudaEventSynchronize(ctx->params_event[pos / quarter]));
    }
    ctx->params_ring[pos] = *values;
    GSE_CHECK_CUDA(cudaMemcpyAsync(ctx->params_block, ctx->params_ring + pos, sizeof(gse_step_params),
                                   cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if ((pos + 1) % quarter == 0) GSE_CHECK_CUDA(cudaEventRecord(ctx->params_event[pos / quarter], (cudaStream_t)stream));
    ctx->params_pos = (pos + 1) % GSE_PARAM_RING;
    return GSE_OK;
}

extern "C" int gse_ctx_use_step_params(gse_ctx* ctx, int enable) {
    GSE_REQUIRE(ctx != NULL, "ctx is NULL");
    GSE_REQUIRE(!enable || ctx->params_block != NULL, "upload a parameter block first");
    ctx->step_params = enable ? ctx->params_block : NULL;
    return GSE_OK;
}

extern "C" int64_t gse_launch_count(const gse_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int64_t gse_ctx_read_trace(gse_ctx* ctx, uint64_t* host_out, int64_t max_words) {
    if (!ctx || !ctx->fused_trace || !host_out || max_words <= 0) return 0;
    gse_device_guard guard(ctx->device);
    const int64_t n = max_words < 8 * 4096 ? max_words : 8 * 4096;
    if (cudaMemcpy(host_out, ctx->fused_trace, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return n;
}

extern "C" unsigned int gse_ctx_errors(gse_ctx* ctx, int clear) {
    if (!ctx || !ctx->err_host) return 0u;
    const unsigned int bits = *(volatile unsigned int*)ctx->err_host;
    if (bits) {
        gse_set_error("device-side error(s):%s%s%s%s%s",
                      (bits & GSE_ERR_CHOLESKY) ? " [covariance not positive definite (Cholesky failed after the +1e-10 I retry)]" : "",
                      (bits & GSE_ERR_SINGULAR_PYY) ? " [innovation covariance P_yy singular]" : "",
                      (bits & GSE_ERR_PEER_TIMEOUT) ? " [peer mailbox exchange timed out: a rank died or issued its calls in another order]" : "",
                      (bits & GSE_ERR_QUEUE_OVERFLOW) ? " [resample heavy-run queue overflow]" : "",
                      (bits & GSE_ERR_ZERO_WEIGHTS) ? " [resample of all-zero weights]" : "");
        if (clear) *(volatile unsigned int*)ctx->err_host = 0u;
    }
    return bits;
}

extern "C" int gse_ctx_result_block(gse_ctx* ctx, double** host_out, double** dev_out) {
    GSE_REQUIRE(ctx != NULL && host_out != NULL && dev_out != NULL, "ctx / out is NULL");
    *host_out = ctx->result_host;
    *dev_out = ctx->result_dev;
    return GSE_OK;
}

extern "C" int gse_ctx_wait(gse_ctx* ctx, void* stream, unsigned int* errors_out) {
    GSE_REQUIRE(ctx != NULL, "ctx is NULL");
    gse_device_guard guard(ctx->device);
    GSE_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    const unsigned int bits = gse_ctx_errors(ctx, 1);
    if (errors_out) *errors_out = bits;
    return GSE_OK;
}

// Host evaluation of the device's output-count predicate (see gse_common.cuh): number of outputs
// i in [0, n_total) with q*(u_i) <= bound, i.e. those sourced at or below cumulative weight `bound`.
extern "C" int64_t gse_count_outputs_below(uint64_t bound, uint64_t total, double r, int64_t n_total) {
    if (n_total <= 0 || total == 0) return 0;
    const double nd = (double)n_total;
    const bool pow2 = (n_total & (n_total - 1)) == 0;
    const double inv = 1.0 / nd;
    // q*(u_i) is non-decreasing in i: binary search for the first i with q*(u_i) > bound
    int64_t lo = 0, hi = n_total;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        const uint64_t q = gse_threshold(gse_sample_position((double)mid, r, nd, inv, pow2), total);
        if (q <= bound) lo = mid + 1; else hi = mid;
    }
    return lo;
}

extern "C" uint64_t gse_threshold_u64(double u, uint64_t total) {
    return gse_threshold(u, total);
}
