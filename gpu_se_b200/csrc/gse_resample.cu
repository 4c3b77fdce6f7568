// Systematic resampling: fixed-point weight scan (K3a tile sums, K3b scan), merge-path partition
// and the fused search + gather (K4+K5).  Shared by the particle filter and the GS-UKF.
//
// Weights are quantised to integers q_i = rint(w_i * 2^s) and summed with integer adds.  Integer
// addition is associative, so the cumulative weights are independent of the scan structure, the
// launch geometry and the number of GPUs, and every comparison below is exact.
//
// The scan is reduce-then-scan (tile sums -> offsets -> in-tile scan) rather than a single pass
// with look-back: both kernels are pure streaming kernels with no inter-block waiting.
#include "gse_common.cuh"

#define TILE_THREADS 256
#define TILE_ITEMS 16
#define TILE_ROWS (TILE_THREADS * TILE_ITEMS)      // 4096 rows per tile

static inline bool aligned32(const void* p) { return ((uintptr_t)p & 31u) == 0; }

__device__ __forceinline__ void ld_f32x8(const float* p, float v[8]) {
    asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void ld_f64x4(const double* p, double v[4]) {
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st_u64x4(uint64_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// scale = 2^(52 - e) with S <= 2^e: the quantised weights sum to at most 2^52 + n/2 < 2^53
__device__ __forceinline__ int quantisation_exponent(double S) {
    if (!(S > 0.0) || !isfinite(S)) return 0;
    int e;
    frexp(S, &e);
    return GSE_TOTAL_BITS - e;
}

// Quantised weights of the 16 consecutive rows owned by this thread.
template <bool HAS_LL, bool HAS_BASE>
__device__ __forceinline__ void quantise16(const float* __restrict__ loglik, const double* __restrict__ base,
                                           float M, int sexp, int64_t row0, int64_t n, uint64_t q[TILE_ITEMS]) {
    const bool full = row0 + TILE_ITEMS <= n;
    float e[TILE_ITEMS];
    if (HAS_LL) {
        float l[TILE_ITEMS];
        if (full) {
            ld_f32x8(loglik + row0, l);
            ld_f32x8(loglik + row0 + 8, l + 8);
        } else {
#pragma unroll
            for (int r = 0; r < TILE_ITEMS; ++r) l[r] = (row0 + r < n) ? loglik[row0 + r] : -INFINITY;
        }
#pragma unroll
        for (int r = 0; r < TILE_ITEMS; ++r) e[r] = __expf(l[r] - M);
    }
    if (!HAS_BASE) {
        const float scale = __int_as_float((127 + sexp) << 23);       // 2^sexp, 0 <= sexp <= 52
#pragma unroll
        for (int r = 0; r < TILE_ITEMS; ++r) q[r] = __float2ull_rn(e[r] * scale);
        if (!full) {
#pragma unroll
            for (int r = 0; r < TILE_ITEMS; ++r) if (row0 + r >= n) q[r] = 0;
        }
    } else {
        const double scale = ldexp(1.0, sexp);
        double b[TILE_ITEMS];
        if (full) {
#pragma unroll
            for (int v = 0; v < TILE_ITEMS / 4; ++v) ld_f64x4(base + row0 + 4 * v, b + 4 * v);
        } else {
#pragma unroll
            for (int r = 0; r < TILE_ITEMS; ++r) b[r] = (row0 + r < n) ? base[row0 + r] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < TILE_ITEMS; ++r) {
            double w = b[r];
            if (HAS_LL) w *= (double)e[r];
            q[r] = __double2ull_rn(w * scale);
        }
    }
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint64_t warp_inclusive_scan_u64(uint64_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// K3a: per-tile sums of the quantised weights; the last block to finish turns them into exclusive
// tile offsets (fixed order) and writes the grand total.
// ------------------------------------------------------------------------------------------------
template <bool HAS_LL, bool HAS_BASE>
__global__ void __launch_bounds__(TILE_THREADS)
k_weight_tile_sums(const float* __restrict__ loglik, const double* __restrict__ base,
                   const double* __restrict__ stats, int64_t n, uint64_t* tile_sum, uint64_t* tile_off,
                   uint64_t* total_out, unsigned int* ticket) {
    __shared__ uint64_t s_w[TILE_THREADS / 32];
    __shared__ uint64_t s_scan[TILE_THREADS];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float M = HAS_LL ? (float)stats[0] : 0.0f;
    const int sexp = quantisation_exponent(stats[1]);
    const int64_t row0 = (int64_t)blockIdx.x * TILE_ROWS + (int64_t)tid * TILE_ITEMS;
    uint64_t q[TILE_ITEMS];
    uint64_t sum = 0;
    if (row0 < n) {
        quantise16<HAS_LL, HAS_BASE>(loglik, base, M, sexp, row0, n, q);
#pragma unroll
        for (int r = 0; r < TILE_ITEMS; ++r) sum += q[r];
    }
    sum = warp_sum_u64(sum);
    if (lane == 0) s_w[wid] = sum;
    __syncthreads();
    if (tid == 0) {
        uint64_t t = 0;
#pragma unroll
        for (int w = 0; w < TILE_THREADS / 32; ++w) t += s_w[w];
        tile_sum[blockIdx.x] = t;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // exclusive scan of gridDim.x tile sums by one block: contiguous chunk per thread
    const unsigned int nt = gridDim.x;
    const unsigned int chunk = (nt + TILE_THREADS - 1) / TILE_THREADS;
    const unsigned int b0 = tid * chunk;
    uint64_t part = 0;
    for (unsigned int k = 0; k < chunk; ++k)
        if (b0 + k < nt) part += __ldcg(tile_sum + b0 + k);
    s_scan[tid] = part;
    __syncthreads();
    if (wid == 0) {                                     // 256 partials: 8 per lane
        uint64_t loc[TILE_THREADS / 32];
        uint64_t run = 0;
#pragma unroll
        for (int k = 0; k < TILE_THREADS / 32; ++k) { loc[k] = run; run += s_scan[lane * (TILE_THREADS / 32) + k]; }
        const uint64_t incl = warp_inclusive_scan_u64(run, lane);
        const uint64_t excl = incl - run;
#pragma unroll
        for (int k = 0; k < TILE_THREADS / 32; ++k) s_scan[lane * (TILE_THREADS / 32) + k] = excl + loc[k];
        if (lane == 31 && total_out) *total_out = incl;
    }
    __syncthreads();
    uint64_t run = s_scan[tid];
    for (unsigned int k = 0; k < chunk; ++k) {
        if (b0 + k < nt) {
            tile_off[b0 + k] = run;
            run += __ldcg(tile_sum + b0 + k);
        }
    }
    if (tid == 0) *ticket = 0u;
}

// ------------------------------------------------------------------------------------------------
// K3b: in-tile inclusive scan + tile offset -> cumulative weights.
// ------------------------------------------------------------------------------------------------
template <bool HAS_LL, bool HAS_BASE>
__global__ void __launch_bounds__(TILE_THREADS)
k_weight_scan(const float* __restrict__ loglik, const double* __restrict__ base,
              const double* __restrict__ stats, int64_t n, const uint64_t* __restrict__ tile_off,
              uint64_t* __restrict__ cumsum) {
    __shared__ uint64_t s_w[TILE_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float M = HAS_LL ? (float)stats[0] : 0.0f;
    const int sexp = quantisation_exponent(stats[1]);
    const int64_t row0 = (int64_t)blockIdx.x * TILE_ROWS + (int64_t)tid * TILE_ITEMS;
    uint64_t q[TILE_ITEMS];
#pragma unroll
    for (int r = 0; r < TILE_ITEMS; ++r) q[r] = 0;
    if (row0 < n) quantise16<HAS_LL, HAS_BASE>(loglik, base, M, sexp, row0, n, q);
#pragma unroll
    for (int r = 1; r < TILE_ITEMS; ++r) q[r] += q[r - 1];
    const uint64_t incl = warp_inclusive_scan_u64(q[TILE_ITEMS - 1], lane);
    if (lane == 31) s_w[wid] = incl;
    __syncthreads();
    uint64_t off = tile_off[blockIdx.x] + (incl - q[TILE_ITEMS - 1]);
#pragma unroll
    for (int w = 0; w < TILE_THREADS / 32; ++w) off += (w < wid) ? s_w[w] : 0;
    if (row0 + TILE_ITEMS <= n) {
#pragma unroll
        for (int v = 0; v < TILE_ITEMS / 4; ++v)
            st_u64x4(cumsum + row0 + 4 * v, off + q[4 * v], off + q[4 * v + 1], off + q[4 * v + 2], off + q[4 * v + 3]);
    } else {
#pragma unroll
        for (int r = 0; r < TILE_ITEMS; ++r)
            if (row0 + r < n) cumsum[row0 + r] = off + q[r];
    }
}

extern "C" int gse_scan_weights(gse_ctx* ctx, const float* loglik_dev, const double* base_dev,
                                const double* stats_dev, int64_t n, uint64_t* cumsum_dev,
                                uint64_t* total_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && stats_dev != NULL && cumsum_dev != NULL, "ctx / stats / cumsum is NULL");
    GSE_REQUIRE(n >= 1 && n <= ctx->n_max, "n out of range for this context");
    GSE_REQUIRE(loglik_dev != NULL || base_dev != NULL, "need loglik or base weights");
    GSE_REQUIRE(loglik_dev == NULL || aligned32(loglik_dev), "loglik must be 32-byte aligned");
    GSE_REQUIRE(base_dev == NULL || aligned32(base_dev), "base must be 32-byte aligned");
    GSE_REQUIRE(aligned32(cumsum_dev), "cumsum must be 32-byte aligned");
    const int64_t tiles = gse_div_up(n, TILE_ROWS);
    GSE_REQUIRE(tiles <= ctx->max_tiles, "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = (unsigned)tiles;
#define LAUNCH_SCAN(LL, BASE)                                                                                   \
    do {                                                                                                        \
        k_weight_tile_sums<LL, BASE><<<g, TILE_THREADS, 0, s>>>(loglik_dev, base_dev, stats_dev, n, ctx->tile_agg, \
                                                                ctx->tile_inc, total_dev, ctx->ticket + 1);     \
        GSE_CHECK_LAUNCH(ctx);                                                                                  \
        k_weight_scan<LL, BASE><<<g, TILE_THREADS, 0, s>>>(loglik_dev, base_dev, stats_dev, n, ctx->tile_inc,   \
                                                           cumsum_dev);                                         \
        GSE_CHECK_LAUNCH(ctx);                                                                                  \
    } while (0)
    if (loglik_dev && base_dev) LAUNCH_SCAN(true, true);
    else if (loglik_dev) LAUNCH_SCAN(true, false);
    else LAUNCH_SCAN(false, true);
#undef LAUNCH_SCAN
    return GSE_OK;
}

// ------------------------------------------------------------------------------------------------
// Merge-path partition.  Sources k (cumulative weights C_k) and outputs i (thresholds q*_i) are two
// sorted sequences; source k precedes output i in the merged order iff C_k < q*_i, so that
// idx_i = #{k : C_k < q*_i} = number of sources merged before output i.  Block b owns merged
// elements [b*W, (b+1)*W): at most W sources staged in shared memory and at most W outputs.
// One warp finds each split point with a 32-ary search (5 rounds of dependent loads at 2^24).
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
    double inv_n;
    int n_pow2;
};

// The exact predicate is the reference's own comparison (particle.py:98): source k precedes output
// i iff fl(C_k / T) < u_i.  Its integer threshold q*(u) lies in [qa - 2, qa + 2] with
// qa = floor(fl(u * T)) (T < 2^53: the product is within 1/2 of the real value and the rounding
// boundary below u is less than 2 away from u * T), so the comparison is decided by two integer
// compares except inside that 5-wide window, where the float64 division is evaluated for real.
// The spacing of consecutive cumulative weights is ~T / n, so the window is hit with probability
// ~5 n / 2^52 per probe.
struct OutputKey {
    double u;          // sample position (i + r) / N
    uint64_t q_lo;     // local lower bound  (qa - 2) - offset, saturating at 0
    uint64_t q_hi;     // local upper bound  (qa + 2) - offset, saturating at 0
};

__device__ __forceinline__ OutputKey output_key(const ResampleArgs& a, double di, uint64_t off, double Td) {
    OutputKey k;
    k.u = gse_sample_position(di, a.r, a.n_total, a.inv_n, a.n_pow2 != 0);
    const uint64_t qa = __double2ull_rd(__dmul_rn(k.u, Td));
    const uint64_t lo = qa > 2 ? qa - 2 : 0;
    const uint64_t hi = qa + 2;
    k.q_lo = lo > off ? lo - off : 0ull;
    k.q_hi = hi > off ? hi - off : 0ull;
    return k;
}

// does the source with LOCAL cumulative weight c precede the output?  (exact)
__device__ __forceinline__ bool precedes(uint64_t c, const OutputKey& k, uint64_t off, double Td) {
    if (c < k.q_lo) return true;
    if (c >= k.q_hi) return false;
    return __ddiv_rn(__ull2double_rn(c + off), Td) < k.u;
}

__global__ void __launch_bounds__(128)
k_resample_partition(const ResampleArgs a, int64_t* __restrict__ part, int64_t nparts) {
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b > nparts) return;
    const uint64_t off = a.offtot[0];
    const double Td = __ull2double_rn(a.offtot[1]);
    int64_t diag = b * RG_WORK;
    const int64_t total = a.n_src + a.n_out;
    if (diag > total) diag = total;
    int64_t lo = diag > a.n_out ? diag - a.n_out : 0;
    int64_t hi = diag < a.n_src ? diag : a.n_src;
    // smallest s in [lo, hi] with NOT precedes(C[s], output diag - 1 - s); true on a prefix
    while (lo < hi) {
        const int64_t span = hi - lo;
        int64_t mid;
        bool active;
        if (span <= 32) {
            mid = lo + lane;
            active = mid < hi;
        } else {
            const int64_t step = span / 33;
            mid = lo + (lane + 1) * step;
            active = true;
        }
        bool pred = false;
        if (active) {
            const uint64_t c = a.cumsum[mid];
            const OutputKey k = output_key(a, (double)(a.out0 + (diag - 1 - mid)), off, Td);
            pred = precedes(c, k, off, Td);
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, pred);
        const int ntrue = __popc(bal);                              // trues form a prefix of the lanes
        const int64_t mid_last_true = __shfl_sync(0xffffffffu, mid, ntrue > 0 ? ntrue - 1 : 0);
        const int64_t mid_first_false = __shfl_sync(0xffffffffu, mid, ntrue < 32 ? ntrue : 31);
        const unsigned int act = __ballot_sync(0xffffffffu, active);
        const int nact = __popc(act);
        if (ntrue > 0) lo = mid_last_true + 1;
        if (ntrue < nact) hi = mid_first_false;
        else if (span <= 32) hi = lo;                               // every candidate was true
    }
    if (lane == 0) part[b] = lo;
}

// ------------------------------------------------------------------------------------------------
// K4+K5: per block, stage the source window of cumulative weights in shared memory.  Each warp
// owns a run of 256 consecutive outputs and walks it 32 at a time: lane l holds output base + l, so
// the 32 thresholds of a round are adjacent and non-decreasing -- the search range starts at the
// previous round's last position and is capped 64 sources ahead when that bound holds (it does
// unless the round crosses a stretch of sources without offspring), and the gathered rows are
// written by consecutive lanes to consecutive addresses.
// ------------------------------------------------------------------------------------------------
#define RG_WARPS (RG_THREADS / 32)
#define RG_CHUNK (RG_WORK / RG_WARPS)      // 256 outputs per warp

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
    const int no = (int)(o1 - o0);
    for (int k = threadIdx.x; k < ns; k += RG_THREADS) s_c[k] = a.cumsum[a0 + k];
    __syncthreads();
    const uint64_t off = a.offtot[0];
    const double Td = __ull2double_rn(a.offtot[1]);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int jbeg = wid * RG_CHUNK;
    const int jend = min(jbeg + RG_CHUNK, no);
    if (jbeg >= jend) return;
    double di = (double)(a.out0 + o0 + jbeg + lane);
    int lo = 0;                                            // warp-uniform: every earlier source precedes
    for (int base = jbeg; base < jend; base += 32, di += 32.0) {
        const int j = base + lane;
        const bool valid = j < jend;
        const int last = min(31, jend - 1 - base);         // last valid lane of this round
        const OutputKey key = output_key(a, di, off, Td);
        // upper end of the search range, warp-uniform
        const uint64_t qmax = __shfl_sync(0xffffffffu, key.q_lo, last);
        int hi = min(lo + 64, ns);
        if (hi < ns && s_c[hi - 1] < qmax) hi = ns;
        int l = lo, h = hi;                                // first k in [lo, hi) with C_k >= q_lo, else hi
        while (l < h) {
            const int mid = (l + h) >> 1;
            if (s_c[mid] < key.q_lo) l = mid + 1; else h = mid;
        }
        lo = __shfl_sync(0xffffffffu, l, last);
        // resolve the (rare) sources inside the rounding window with the reference's own division
        while (l < ns && s_c[l] < key.q_hi && precedes(s_c[l], key, off, Td)) ++l;
        if (valid) {
            int64_t idx = a0 + l;
            if (idx >= a.n_src) idx = a.n_src - 1;         // only reachable through a degenerate total
            const int64_t jo = o0 + j;
            if (dst) {
                for (int c = 0; c < ncols; ++c) dst[c * ld_dst + jo] = __ldg(src + c * ld_src + idx);
            }
            if (loglik_out) loglik_out[jo] = 0.0f;
            if (idx_out) idx_out[jo] = idx;
        }
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
    a.inv_n = 1.0 / (double)n_total;
    a.n_pow2 = ((n_total & (n_total - 1)) == 0) ? 1 : 0;
    const int64_t nparts = gse_div_up(n_src + n_out, RG_WORK);
    GSE_REQUIRE(nparts + 1 <= ctx->max_tiles + 2, "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    k_resample_partition<<<(unsigned)gse_div_up((nparts + 1) * 32, 128), 128, 0, s>>>(a, ctx->part, nparts);
    GSE_CHECK_LAUNCH(ctx);
    k_resample_gather<<<(unsigned)nparts, RG_THREADS, 0, s>>>(a, ctx->part, src_dev, ld_src, dst_dev, ld_dst, ncols,
                                                              loglik_out_dev, idx_out_dev);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}
