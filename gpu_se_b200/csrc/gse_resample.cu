// Systematic resampling: fixed-point weight scan (K3), merge-path partition and the fused
// search + gather (K4+K5).  Shared by the particle filter and the GS-UKF.
//
// Weights are quantised to integers q_i = rint(w_i * 2^s) and scanned with integer adds.  Integer
// addition is associative, so the cumulative weights are independent of the scan structure, the
// launch geometry and the number of GPUs, and every comparison below is exact.
#include <cuda/atomic>

#include "gse_common.cuh"

#define SCAN_THREADS 1024
#define SCAN_ITEMS 4
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

#define FLAG_INVALID 0u
#define FLAG_AGGREGATE 1u
#define FLAG_PREFIX 2u

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }
static inline bool aligned32(const void* p) { return ((uintptr_t)p & 31u) == 0; }

__device__ __forceinline__ void st_stream_u64x4(uint64_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// scale = 2^(61 - e) with S <= 2^e: the total of the quantised weights stays below 2^62
__device__ __forceinline__ double quantisation_scale(double S) {
    if (!(S > 0.0) || !isfinite(S)) return 1.0;
    int e;
    frexp(S, &e);
    return ldexp(1.0, 61 - e);
}

// ------------------------------------------------------------------------------------------------
// K3: single-pass inclusive scan with decoupled look-back over 4096-row tiles.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_weights(const float* __restrict__ loglik, const double* __restrict__ base,
               const double* __restrict__ stats, int64_t n, uint64_t* __restrict__ cumsum,
               uint64_t* __restrict__ total_out, uint64_t* tile_agg, uint64_t* tile_inc,
               unsigned int* tile_flag, unsigned int* ticket, unsigned int epoch, unsigned int num_tiles) {
    __shared__ unsigned int s_tile;
    __shared__ uint64_t s_warp[SCAN_THREADS / 32];
    __shared__ uint64_t s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    if (tid == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == num_tiles - 1) *ticket = 0u;   // nobody else draws a ticket in this launch
        s_tile = t;
    }
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t row0 = (int64_t)tile * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;

    const float M = loglik ? (float)stats[0] : 0.0f;
    const double scale = quantisation_scale(stats[1]);

    // quantised weights of this thread's 4 rows
    uint64_t q[SCAN_ITEMS] = {0, 0, 0, 0};
    if (row0 < n) {
        float l[4] = {0.f, 0.f, 0.f, 0.f};
        if (loglik) {
            const float4 lw = ld_stream4(loglik + row0);
            l[0] = lw.x; l[1] = lw.y; l[2] = lw.z; l[3] = lw.w;
        }
#pragma unroll
        for (int r = 0; r < SCAN_ITEMS; ++r) {
            if (row0 + r < n) {
                double w = loglik ? (double)__expf(l[r] - M) : 1.0;
                if (base) w *= base[row0 + r];
                q[r] = __double2ull_rn(w * scale);
            }
        }
    }
    // thread-local inclusive scan, warp scan of thread totals, block scan of warp totals
    q[1] += q[0]; q[2] += q[1]; q[3] += q[2];
    uint64_t incl = q[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = s_warp[lane];
        uint64_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        s_warp[lane] = wi - w;                  // exclusive prefix of warp totals
        const uint64_t tile_total = __shfl_sync(0xffffffffu, wi, 31);

        // ---- decoupled look-back (warp 0) ----
        cuda::atomic_ref<unsigned int, cuda::thread_scope_device> my_flag(tile_flag[tile]);
        uint64_t exclusive = 0;
        if (tile == 0) {
            if (lane == 0) {
                cuda::atomic_ref<uint64_t, cuda::thread_scope_device>(tile_inc[0]).store(tile_total, cuda::memory_order_relaxed);
                my_flag.store((epoch << 2) | FLAG_PREFIX, cuda::memory_order_release);
            }
        } else {
            if (lane == 0) {
                cuda::atomic_ref<uint64_t, cuda::thread_scope_device>(tile_agg[tile]).store(tile_total, cuda::memory_order_relaxed);
                my_flag.store((epoch << 2) | FLAG_AGGREGATE, cuda::memory_order_release);
            }
            int64_t look = (int64_t)tile - 1 - lane;     // lane 0 looks at the nearest predecessor
            while (true) {
                unsigned int st = FLAG_PREFIX;           // lanes before tile 0 count as a zero prefix
                uint64_t val = 0;
                if (look >= 0) {
                    cuda::atomic_ref<unsigned int, cuda::thread_scope_device> f(tile_flag[look]);
                    unsigned int fv;
                    do {
                        fv = f.load(cuda::memory_order_acquire);
                    } while ((fv >> 2) != epoch || (fv & 3u) == FLAG_INVALID);
                    st = fv & 3u;
                    uint64_t* src = (st == FLAG_PREFIX) ? &tile_inc[look] : &tile_agg[look];
                    val = cuda::atomic_ref<uint64_t, cuda::thread_scope_device>(*src).load(cuda::memory_order_relaxed);
                }
                const unsigned int has_prefix = __ballot_sync(0xffffffffu, st == FLAG_PREFIX);
                const int first = __ffs(has_prefix) - 1;             // nearest tile with a full prefix
                uint64_t contrib = (has_prefix == 0u || lane <= first) ? val : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                exclusive += contrib;
                if (has_prefix != 0u) break;
                look -= 32;
            }
            if (lane == 0) {
                cuda::atomic_ref<uint64_t, cuda::thread_scope_device>(tile_inc[tile]).store(exclusive + tile_total, cuda::memory_order_relaxed);
                my_flag.store((epoch << 2) | FLAG_PREFIX, cuda::memory_order_release);
            }
        }
        if (lane == 0) {
            s_prefix = exclusive;
            if (tile == num_tiles - 1 && total_out) *total_out = exclusive + tile_total;
        }
    }
    __syncthreads();
    const uint64_t off = s_prefix + s_warp[wid] + (incl - q[3]);
    if (row0 + SCAN_ITEMS <= n) {
        st_stream_u64x4(cumsum + row0, off + q[0], off + q[1], off + q[2], off + q[3]);
    } else {
#pragma unroll
        for (int r = 0; r < SCAN_ITEMS; ++r)
            if (row0 + r < n) cumsum[row0 + r] = off + q[r];
    }
}

extern "C" int gse_scan_weights(gse_ctx* ctx, const float* loglik_dev, const double* base_dev,
                                const double* stats_dev, int64_t n, uint64_t* cumsum_dev,
                                uint64_t* total_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && stats_dev != NULL && cumsum_dev != NULL, "ctx / stats / cumsum is NULL");
    GSE_REQUIRE(n >= 1 && n <= ctx->n_max, "n out of range for this context");
    GSE_REQUIRE(loglik_dev != NULL || base_dev != NULL, "need loglik or base weights");
    GSE_REQUIRE(loglik_dev == NULL || aligned16(loglik_dev), "loglik must be 16-byte aligned");
    GSE_REQUIRE(aligned32(cumsum_dev), "cumsum must be 32-byte aligned");
    const int64_t tiles = gse_div_up(n, SCAN_TILE);
    GSE_REQUIRE(tiles <= ctx->max_tiles, "workspace too small");
    ctx->scan_epoch++;
    if ((ctx->scan_epoch >> 30) != 0) {     // epoch wrapped: clear the flags once
        GSE_CHECK_CUDA(cudaMemsetAsync(ctx->tile_flag, 0, sizeof(unsigned int) * ctx->max_tiles, (cudaStream_t)stream));
        ctx->scan_epoch = 1;
    }
    k_scan_weights<<<(unsigned)tiles, SCAN_THREADS, 0, (cudaStream_t)stream>>>(
        loglik_dev, base_dev, stats_dev, n, cumsum_dev, total_dev, ctx->tile_agg, ctx->tile_inc,
        ctx->tile_flag, ctx->ticket + 1, ctx->scan_epoch, (unsigned)tiles);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

// ------------------------------------------------------------------------------------------------
// Merge-path partition.  Sources k (cumulative weights C_k) and outputs i (thresholds q*_i) are two
// sorted sequences; source k precedes output i in the merged order iff C_k < q*_i, so that
// idx_i = #{k : C_k < q*_i} = number of sources merged before output i.  Block b owns merged
// elements [b*W, (b+1)*W): at most W sources staged in shared memory and at most W outputs.
// ------------------------------------------------------------------------------------------------
#define RG_THREADS 512
#define RG_WORK 4096

struct ResampleArgs {
    const uint64_t* cumsum;
    const uint64_t* offtot;    // [0] offset of this shard's cumulative weights, [1] global total
    int64_t n_src;
    int64_t n_out;
    int64_t out0;              // global index of local output 0
    double r;
    double n_total;
};

// threshold of local output j relative to the local cumulative weights
__device__ __forceinline__ uint64_t local_threshold(const ResampleArgs& a, int64_t j, uint64_t off, double Td) {
    const uint64_t q = gse_threshold(gse_sample_position(a.out0 + j, a.r, a.n_total), Td);
    return q > off ? q - off : 0ull;
}

__global__ void k_resample_partition(const ResampleArgs a, int64_t* __restrict__ part, int64_t nparts) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nparts) return;
    const uint64_t off = a.offtot[0];
    const double Td = gse_u64_to_double(a.offtot[1]);
    int64_t diag = b * RG_WORK;
    const int64_t total = a.n_src + a.n_out;
    if (diag > total) diag = total;
    int64_t lo = diag > a.n_out ? diag - a.n_out : 0;
    int64_t hi = diag < a.n_src ? diag : a.n_src;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        const uint64_t c = a.cumsum[mid];
        const uint64_t q = local_threshold(a, diag - 1 - mid, off, Td);
        if (c < q) lo = mid + 1; else hi = mid;
    }
    part[b] = lo;
}

// ------------------------------------------------------------------------------------------------
// K4+K5: per block, stage the source window of cumulative weights in shared memory, binary-search
// every output's threshold in it, gather the SoA columns, reset the log-likelihoods.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RG_THREADS)
k_resample_gather(const ResampleArgs a, const int64_t* __restrict__ part, const float* __restrict__ src,
                  int64_t ld_src, float* __restrict__ dst, int64_t ld_dst, int ncols,
                  float* __restrict__ loglik_out, int64_t* __restrict__ idx_out) {
    __shared__ uint64_t s_c[RG_WORK];
    const int64_t b = blockIdx.x;
    const int64_t total = a.n_src + a.n_out;
    int64_t d0 = b * RG_WORK, d1 = d0 + RG_WORK;
    if (d1 > total) d1 = total;
    const int64_t a0 = part[b], a1 = part[b + 1];
    const int64_t o0 = d0 - a0, o1 = d1 - a1;
    if (o1 <= o0) return;                                  // a stretch of sources with no offspring
    const int ns = (int)(a1 - a0);
    for (int k = threadIdx.x; k < ns; k += RG_THREADS) s_c[k] = a.cumsum[a0 + k];
    __syncthreads();
    const uint64_t off = a.offtot[0];
    const double Td = gse_u64_to_double(a.offtot[1]);
    for (int64_t j = o0 + threadIdx.x; j < o1; j += RG_THREADS) {
        const uint64_t q = local_threshold(a, j, off, Td);
        int lo = 0, hi = ns;                               // first k in the window with C_k >= q
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_c[mid] < q) lo = mid + 1; else hi = mid;
        }
        int64_t idx = a0 + lo;
        if (idx >= a.n_src) idx = a.n_src - 1;             // only reachable through rounding of the total
        if (dst) {
            for (int c = 0; c < ncols; ++c) dst[c * ld_dst + j] = __ldg(src + c * ld_src + idx);
        }
        if (loglik_out) loglik_out[j] = 0.0f;
        if (idx_out) idx_out[j] = idx;
    }
}

extern "C" int gse_resample_gather(gse_ctx* ctx, const uint64_t* cumsum_dev, int64_t n_src,
                                   const uint64_t* offtot_dev, double r, int64_t n_total, int64_t out0,
                                   int64_t n_out, const float* src_dev, int64_t ld_src, float* dst_dev,
                                   int64_t ld_dst, int ncols, float* loglik_out_dev, int64_t* idx_out_dev,
                                   void* stream) {
    GSE_REQUIRE(ctx != NULL && cumsum_dev != NULL && offtot_dev != NULL, "ctx / cumsum / offtot is NULL");
    GSE_REQUIRE(n_src >= 1 && n_src <= ctx->n_max, "n_src out of range for this context");
    GSE_REQUIRE(n_out >= 0 && n_out <= ctx->n_max, "n_out out of range for this context");
    GSE_REQUIRE(n_total >= 1 && out0 >= 0 && out0 + n_out <= n_total, "output range outside [0, n_total)");
    GSE_REQUIRE(r >= 0.0 && r < 1.0, "r must be in [0, 1)");
    if (n_out == 0) return GSE_OK;
    if (dst_dev) {
        GSE_REQUIRE(src_dev != NULL && ncols >= 1, "src is NULL / ncols < 1");
        GSE_REQUIRE(ld_src >= n_src && ld_dst >= n_out, "ld too small");
        GSE_REQUIRE(src_dev != dst_dev, "gather cannot be done in place");
    }
    ResampleArgs a;
    a.cumsum = cumsum_dev;
    a.offtot = offtot_dev;
    a.n_src = n_src;
    a.n_out = n_out;
    a.out0 = out0;
    a.r = r;
    a.n_total = (double)n_total;
    const int64_t nparts = gse_div_up(n_src + n_out, RG_WORK);
    GSE_REQUIRE(nparts + 1 <= ctx->max_tiles + 2, "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    k_resample_partition<<<(unsigned)gse_div_up(nparts + 1, 128), 128, 0, s>>>(a, ctx->part, nparts);
    GSE_CHECK_LAUNCH(ctx);
    k_resample_gather<<<(unsigned)nparts, RG_THREADS, 0, s>>>(a, ctx->part, src_dev, ld_src, dst_dev, ld_dst, ncols,
                                                              loglik_out_dev, idx_out_dev);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}
