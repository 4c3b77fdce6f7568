// Systematic resampling: fixed-point weight scan (K3a tile sums, K3b scan), merge-path partition
// the search (K4) and the materialising gather (K5).  Shared by the particle filter and the GS-UKF.
//
// Weights are quantised to integers q_i = rint(w_i * 2^s) and summed with integer adds.  Integer
// addition is associative, so the cumulative weights are independent of the scan structure, the
// launch geometry and the number of GPUs, and every comparison below is exact.
//
// Two scans: the single-pass look-back kernel (k_weight_scan_lookback, the default for weights without a float64
// base) and reduce-then-scan (k_weight_tile_sums + k_weight_scan: per-warp run sums -> offsets -> in-run scan).
// The filters' resample() itself runs the fused kernel of gse_resample_fused.cu; the kernels here serve the
// stand-alone ABI entry points (gse_scan_weights, gse_resample_search), the slab exchange and the sharded search.
#include "gse_resample_common.cuh"

#define TILE_THREADS 256
#define TILE_ITEMS 16
#define TILE_ROWS (TILE_THREADS * TILE_ITEMS)      // 4096 rows per tile

// q = rint(exp(l - M) * 2^sexp): the one expression both passes (K3a, K3b) and the fused kernel evaluate (weight_exp();
// the argument named M is -M log2(e)), with explicit round-to-nearest operations so that no contraction can make them differ
// (SMALL = convert through cvt.rni.u32.f32 when the scale is <= 2^31: measured slower than the plain
// 64-bit conversion, kept only as the documented experiment)
template <bool SMALL>
__device__ __forceinline__ uint64_t quantise1(float l, float M, float scale) {
    const float v = __fmul_rn(weight_exp(l, M), scale);
    return SMALL ? (uint64_t)__float2uint_rn(v) : __float2ull_rn(v);
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
        for (int r = 0; r < TILE_ITEMS; ++r) e[r] = weight_exp(l[r], M);
    }
    if (!HAS_BASE) {
        const float scale = __int_as_float((127 + sexp) << 23);       // 2^sexp, 0 <= sexp <= 52
#pragma unroll
        for (int r = 0; r < TILE_ITEMS; ++r) q[r] = __float2ull_rn(__fmul_rn(e[r], scale));
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

// ------------------------------------------------------------------------------------------------
// K3a / K3b.  Persistent grid; every WARP owns a contiguous run of 512-row tiles and walks it with
// a running carry in registers, so neither pass has a block-level barrier on its data path.
//   K3a: sum of the warp's run -> warp_sum[w]; the last block to finish turns the (at most a few
//        thousand) warp sums into exclusive offsets warp_off[w] in a fixed order and writes the total.
//   K3b: re-quantise, scan inside the warp (shuffles), add the carry, store the cumulative weights.
// ------------------------------------------------------------------------------------------------
#define WTILE_ROWS (32 * TILE_ITEMS)               // 512 rows per warp tile
#define SCAN_WARPS (TILE_THREADS / 32)

struct ScanGeom {
    int64_t ntiles;        // ceil(n / 512)
    int64_t tiles_per_warp;
    int nwarps;            // warps that own at least one tile
};

static ScanGeom scan_geometry(int64_t n, int num_sms, unsigned* blocks) {
    ScanGeom g;
    g.ntiles = gse_div_up(n, WTILE_ROWS);
    const int64_t max_warps = (int64_t)num_sms * 4 * SCAN_WARPS;     // 4 resident CTAs per SM at 64 registers
    g.tiles_per_warp = gse_div_up(g.ntiles, max_warps);
    g.nwarps = (int)gse_div_up(g.ntiles, g.tiles_per_warp);
    *blocks = (unsigned)gse_div_up(g.nwarps, SCAN_WARPS);
    return g;
}

template <bool HAS_LL, bool HAS_BASE>
__global__ void __launch_bounds__(TILE_THREADS)
k_weight_tile_sums(const float* __restrict__ loglik, const double* __restrict__ base,
                   const double* __restrict__ stats, int64_t n, const ScanGeom geo, uint64_t* warp_sum,
                   uint64_t* warp_off, uint64_t* total_out, unsigned int* ticket) {
    __shared__ uint64_t s_scan[TILE_THREADS];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float M = HAS_LL ? weight_exp_offset((float)stats[0]) : 0.0f;    // -M log2(e), see weight_exp()
    const int sexp = quantisation_exponent(stats[1]);
    const int w = blockIdx.x * SCAN_WARPS + wid;
    if (w < geo.nwarps) {
        const int64_t t0 = (int64_t)w * geo.tiles_per_warp;
        const int64_t t1 = min(t0 + geo.tiles_per_warp, geo.ntiles);
        uint64_t sum = 0;
        if (HAS_LL && !HAS_BASE) {
            // the sum does not care about the order: 128-bit loads, 4 groups in flight per lane
            const float scale = __int_as_float((127 + sexp) << 23);
            const int64_t r0 = t0 * WTILE_ROWS, r1 = min(t1 * WTILE_ROWS, n);
            const int64_t r1_full = r0 + ((r1 - r0) & ~(int64_t)127);
            int64_t row = r0 + 4 * lane;
#pragma unroll 4
            for (; row < r1_full; row += 128) {
                const float4 l = ld_stream4(loglik + row);
                sum += quantise1<false>(l.x, M, scale) + quantise1<false>(l.y, M, scale) +
                       quantise1<false>(l.z, M, scale) + quantise1<false>(l.w, M, scale);
            }
            for (int r = 0; r < 4; ++r)
                if (row + r < r1) sum += quantise1<false>(loglik[row + r], M, scale);
        } else {
            for (int64_t t = t0; t < t1; ++t) {
                const int64_t row0 = t * WTILE_ROWS + (int64_t)lane * TILE_ITEMS;
                if (row0 < n) {
                    uint64_t q[TILE_ITEMS];
                    quantise16<HAS_LL, HAS_BASE>(loglik, base, M, sexp, row0, n, q);
#pragma unroll
                    for (int r = 0; r < TILE_ITEMS; ++r) sum += q[r];
                }
            }
        }
        sum = warp_sum_u64(sum);
        if (lane == 0) warp_sum[w] = sum;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // exclusive scan of nwarps sums by one block: contiguous chunk per thread
    const int nt = geo.nwarps;
    const int chunk = (nt + TILE_THREADS - 1) / TILE_THREADS;
    const int b0 = tid * chunk;
    uint64_t part = 0;
    for (int k = 0; k < chunk; ++k)
        if (b0 + k < nt) part += __ldcg(warp_sum + b0 + k);
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
    for (int k = 0; k < chunk; ++k) {
        if (b0 + k < nt) {
            warp_off[b0 + k] = run;
            run += __ldcg(warp_sum + b0 + k);
        }
    }
    if (tid == 0) *ticket = 0u;
}

template <bool HAS_LL, bool HAS_BASE>
__global__ void __launch_bounds__(TILE_THREADS)
k_weight_scan(const float* __restrict__ loglik, const double* __restrict__ base,
              const double* __restrict__ stats, int64_t n, const ScanGeom geo,
              const uint64_t* __restrict__ warp_off, uint64_t* __restrict__ cumsum) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int w = blockIdx.x * SCAN_WARPS + wid;
    if (w >= geo.nwarps) return;
    const float M = HAS_LL ? weight_exp_offset((float)stats[0]) : 0.0f;    // -M log2(e), see weight_exp()
    const int sexp = quantisation_exponent(stats[1]);
    const int64_t t0 = (int64_t)w * geo.tiles_per_warp;
    const int64_t t1 = min(t0 + geo.tiles_per_warp, geo.ntiles);
    uint64_t carry = warp_off[w];
    for (int64_t t = t0; t < t1; ++t) {
        const int64_t row0 = t * WTILE_ROWS + (int64_t)lane * TILE_ITEMS;
        uint64_t q[TILE_ITEMS];
#pragma unroll
        for (int r = 0; r < TILE_ITEMS; ++r) q[r] = 0;
        if (row0 < n) quantise16<HAS_LL, HAS_BASE>(loglik, base, M, sexp, row0, n, q);
#pragma unroll
        for (int r = 1; r < TILE_ITEMS; ++r) q[r] += q[r - 1];
        const uint64_t incl = warp_inclusive_scan_u64(q[TILE_ITEMS - 1], lane);
        const uint64_t off = carry + (incl - q[TILE_ITEMS - 1]);
        if (row0 + TILE_ITEMS <= n) {
#pragma unroll
            for (int v = 0; v < TILE_ITEMS / 4; ++v)
                st_u64x4(cumsum + row0 + 4 * v, off + q[4 * v], off + q[4 * v + 1], off + q[4 * v + 2], off + q[4 * v + 3]);
        } else {
#pragma unroll
            for (int r = 0; r < TILE_ITEMS; ++r)
                if (row0 + r < n) cumsum[row0 + r] = off + q[r];
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// ------------------------------------------------------------------------------------------------
// K3 in one kernel (weights = exp(loglik - M), no float64 base): every block owns one contiguous run of
// rows; it sums the run (HBM read), chains to the blocks before it with a decoupled look-back, then
// re-reads the run (L2: a run is ~100 KB) and writes the cumulative weights -- 4 R + 8 W of DRAM
// traffic, no second launch and no serial last-block pass.  A block publishes ONE 64-bit status word
// (2 bits state | 8 bits epoch | 54 bits value: the values are < 2^53), so publishing and reading need
// no fences; warp 0 looks back 32 predecessors at a time (at most gridDim/32 rounds).  Blocks number
// themselves in the order they start (an atomic ticket), so every block a warp waits for is running.  The epoch
// lives in device memory and is advanced by the last block to finish, so that the kernel can be
// replayed from a captured CUDA graph.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS)
k_weight_scan_lookback(const float* __restrict__ loglik, const double* __restrict__ stats, int64_t n,
                       int64_t rows_per_block, uint64_t* status, unsigned int* counters,
                       uint64_t* __restrict__ cumsum, uint64_t* total_out) {
    __shared__ unsigned int s_vb, s_epoch;
    __shared__ uint64_t s_wsum[SCAN_WARPS];
    __shared__ uint64_t s_excl;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        s_epoch = *(volatile unsigned int*)(counters + 2);      // advanced only after every block has finished
        s_vb = atomicAdd(counters, 1u);                         // blocks are numbered in the order they start
    }
    __syncthreads();
    const unsigned int epoch = s_epoch;
    const int64_t vb = s_vb;
    const float M = weight_exp_offset((float)stats[0]);          // -M log2(e), see weight_exp()
    const int sexp = quantisation_exponent(stats[1]);
    const float scale = __int_as_float((127 + sexp) << 23);
    // this block owns rows [b0, b1), warp w the contiguous sub-run [w0, w1) of it (whole 512-row tiles)
    const int64_t b0 = vb * rows_per_block, b1 = min(b0 + rows_per_block, n);
    const int64_t rows_per_warp = rows_per_block / SCAN_WARPS;
    const int64_t w0 = min(b0 + wid * rows_per_warp, b1), w1 = min(w0 + rows_per_warp, b1);
    // pass 1: sum of the warp's sub-run (order-free: 128-bit loads)
    uint64_t sum = 0;
    {
        const int64_t full = w0 + ((w1 - w0) & ~(int64_t)127);
        int64_t row = w0 + 4 * lane;
#pragma unroll 4
        for (; row < full; row += 128) {
            const float4 l = ld_stream4(loglik + row);
            sum += quantise1<false>(l.x, M, scale) + quantise1<false>(l.y, M, scale) +
                   quantise1<false>(l.z, M, scale) + quantise1<false>(l.w, M, scale);
        }
        for (int r = 0; r < 4; ++r)
            if (row + r < w1) sum += quantise1<false>(loglik[row + r], M, scale);
    }
    sum = warp_sum_u64(sum);
    if (lane == 0) s_wsum[wid] = sum;
    __syncthreads();
    // warp 0: publish the block aggregate, look back over the blocks that started earlier
    if (wid == 0) {
        uint64_t agg = 0;
#pragma unroll
        for (int w = 0; w < SCAN_WARPS; ++w) agg += s_wsum[w];
        uint64_t excl = 0;
        if (vb > 0) {
            if (lane == 0) st_status(status + vb, status_pack(ST_AGGREGATE, epoch, agg));
            int64_t pos = vb - 1;                                // lane l looks at block pos - l
            for (;;) {
                const int64_t mine = pos - lane;
                uint64_t word = status_pack(ST_PREFIX, epoch, 0);               // before block 0: prefix 0
                if (mine >= 0) word = ld_status(status + mine);
                const bool valid = ((word >> 54) & 0xffull) == (uint64_t)(epoch & 0xffu) && (word >> 62) != 0ull;
                const unsigned int v_mask = __ballot_sync(0xffffffffu, valid);
                const unsigned int p_mask = __ballot_sync(0xffffffffu, valid && (word >> 62) == ST_PREFIX);
                const int first = p_mask ? __ffs(p_mask) - 1 : 32;              // nearest block with a full prefix
                const unsigned int need = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);
                if ((v_mask & need) != need) continue;                          // a needed block is not published yet
                const uint64_t v = (lane <= first) ? (word & ((1ull << 54) - 1ull)) : 0ull;
                excl += warp_sum_u64(v);
                if (first < 32) break;
                pos -= 32;
            }
        }
        if (lane == 0) {
            st_status(status + vb, status_pack(ST_PREFIX, epoch, excl + agg));
            s_excl = excl;
            if (b1 == n && total_out) *total_out = excl + agg;
        }
    }
    __syncthreads();
    // pass 2: re-read the sub-run (L2), scan inside the warp with a register carry, store
    uint64_t carry = s_excl;
#pragma unroll
    for (int w = 0; w < SCAN_WARPS; ++w) carry += (w < wid) ? s_wsum[w] : 0ull;
    for (int64_t t0 = w0; t0 < w1; t0 += WTILE_ROWS) {
        const int64_t row0 = t0 + (int64_t)lane * TILE_ITEMS;
        uint64_t q[TILE_ITEMS];
#pragma unroll
        for (int r = 0; r < TILE_ITEMS; ++r) q[r] = 0;
        if (row0 < w1) quantise16<true, false>(loglik, NULL, M, sexp, row0, w1, q);
#pragma unroll
        for (int r = 1; r < TILE_ITEMS; ++r) q[r] += q[r - 1];
        const uint64_t incl = warp_inclusive_scan_u64(q[TILE_ITEMS - 1], lane);
        const uint64_t off = carry + (incl - q[TILE_ITEMS - 1]);
        if (row0 + TILE_ITEMS <= w1) {
#pragma unroll
            for (int v = 0; v < TILE_ITEMS / 4; ++v)
                st_u64x4(cumsum + row0 + 4 * v, off + q[4 * v], off + q[4 * v + 1], off + q[4 * v + 2], off + q[4 * v + 3]);
        } else {
#pragma unroll
            for (int r = 0; r < TILE_ITEMS; ++r)
                if (row0 + r < w1) cumsum[row0 + r] = off + q[r];
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(counters + 1, 1u) == gridDim.x - 1) {      // last block out: reset for the next launch
            counters[0] = 0u;
            counters[1] = 0u;
            counters[2] = epoch + 1u;
            __threadfence();
        }
    }
}

extern "C" int gse_scan_weights(gse_ctx* ctx, const float* loglik_dev, const double* base_dev,
                                const double* stats_dev, int64_t n, uint64_t* cumsum_dev,
                                uint64_t* total_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && stats_dev != NULL && cumsum_dev != NULL, "ctx / stats / cumsum is NULL");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n >= 1 && n <= ctx->n_max, "n out of range for this context");
    GSE_REQUIRE(loglik_dev != NULL || base_dev != NULL, "need loglik or base weights");
    GSE_REQUIRE(loglik_dev == NULL || aligned32(loglik_dev), "loglik must be 32-byte aligned");
    GSE_REQUIRE(base_dev == NULL || aligned32(base_dev), "base must be 32-byte aligned");
    GSE_REQUIRE(aligned32(cumsum_dev), "cumsum must be 32-byte aligned");
    unsigned g = 0;
    const ScanGeom geo = scan_geometry(n, ctx->num_sms, &g);
    GSE_REQUIRE(geo.nwarps <= ctx->max_tiles && geo.ntiles <= ctx->max_tiles, "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    if (loglik_dev && !base_dev && ctx->scan_single_pass) {
        // status words this context did not rewrite in its previous launch may carry an epoch that has wrapped
        // around (8 bits) since: clear the ones the tile count grows into
        if (ctx->scan_resident_blocks == 0) {
            // every block must be resident at once: warp 0 of a block waits for blocks that have started, and
            // a block that cannot start while the others spin would never publish its aggregate
            int per_sm = 0;
            GSE_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_weight_scan_lookback, TILE_THREADS, 0));
            GSE_REQUIRE(per_sm >= 1, "look-back scan kernel does not fit an SM");
            ctx->scan_resident_blocks = per_sm * ctx->num_sms;
        }
        // one contiguous run of whole 4096-row groups per block, one wave of blocks
        const int64_t groups = gse_div_up(n, (int64_t)WTILE_ROWS * SCAN_WARPS);
        int64_t blocks = groups < ctx->scan_resident_blocks ? groups : ctx->scan_resident_blocks;
        const int64_t rows_per_block = gse_div_up(groups, blocks) * WTILE_ROWS * SCAN_WARPS;
        blocks = gse_div_up(n, rows_per_block);
        // status words this context did not rewrite in its previous launch may carry an epoch that has wrapped
        // around (8 bits) since: clear the ones the block count grows into
        if (blocks > ctx->scan_tiles_prev)
            GSE_CHECK_CUDA(cudaMemsetAsync(ctx->tile_status + ctx->scan_tiles_prev, 0,
                                           sizeof(uint64_t) * (size_t)(blocks - ctx->scan_tiles_prev), s));
        ctx->scan_tiles_prev = blocks;
        k_weight_scan_lookback<<<(unsigned)blocks, TILE_THREADS, 0, s>>>(loglik_dev, stats_dev, n, rows_per_block,
                                                                          ctx->tile_status, ctx->ticket + 4, cumsum_dev,
                                                                          total_dev);
        GSE_CHECK_LAUNCH(ctx);
        return GSE_OK;
    }
#define LAUNCH_SCAN(LL, BASE)                                                                                   \
    do {                                                                                                        \
        k_weight_tile_sums<LL, BASE><<<g, TILE_THREADS, 0, s>>>(loglik_dev, base_dev, stats_dev, n, geo,        \
                                                                ctx->tile_agg, ctx->tile_inc, total_dev,        \
                                                                ctx->ticket + 1);                               \
        GSE_CHECK_LAUNCH(ctx);                                                                                  \
        k_weight_scan<LL, BASE><<<g, TILE_THREADS, 0, s>>>(loglik_dev, base_dev, stats_dev, n, geo,             \
                                                           ctx->tile_inc, cumsum_dev);                          \
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
// elements [b*W, (b+1)*W): at most W sources and at most W outputs, both staged in shared memory.
// One warp finds each split point with a 32-ary search (5 rounds of dependent loads at 2^24).
// ------------------------------------------------------------------------------------------------
#define RS_THREADS 256
#define RS_VT 16
#define RS_WORK (RS_THREADS * RS_VT)      // 4096 merged elements per block

struct ResampleArgs {
    const uint64_t* cumsum;    // single segment: the local cumulative weights
    const uint64_t* offtot;    // [s] offset of segment s's cumulative weights, [nseg] the global total
    // segmented sources (sharded run, peer memory): segment s holds global rows [seg_row[s], seg_row[s+1])
    const int64_t* range;      // device [k_lo, k_hi): the sources that interleave with the outputs (NULL = all)
    double* consts;            // device scratch: [0] = 1 / T (written by the partition kernel)
    int nseg;
    int64_t seg_row[GSE_MAX_SHARDS + 1];
    const uint64_t* seg_cumsum[GSE_MAX_SHARDS];
    int64_t n_src;
    int64_t n_out;
    int64_t out0;              // global index of local output 0
    double r;
    const double* r_dev;       // device override of r (parameter block of a captured graph), or NULL
    double n_total;
    double inv_n;
    int n_pow2;
};

// The exact predicate is the reference's own comparison (particle.py:98): source k precedes output
// i iff fl(C_k / T) < u_i.  Its integer threshold q*(u) lies in [qa - 2, qa + 2] with
// qa = floor(fl(u * T)) (T < 2^53: the product is within 1/2 of the real value and the rounding
// boundary below u is less than 2 away from u * T), so the comparison is decided by two integer
// compares against q_lo = qa - 2 (local: minus the shard offset, saturating at 0) and q_lo + 4,
// except inside that window, where the float64 division is evaluated for real.  The spacing of
// consecutive cumulative weights is ~T / n, so the window is hit with probability ~5 n / 2^52.
// global cumulative weight of global source row k: the owning shard's local value (read from that
// GPU's memory over NVLink when it is a peer) plus the shard's offset
template <bool SEG>
__device__ __forceinline__ uint64_t source_weight(const ResampleArgs& a, int64_t k, uint64_t off0) {
    if (!SEG) return __ldg(a.cumsum + k) + off0;
    int s = 0;
#pragma unroll
    for (int t = 1; t < GSE_MAX_SHARDS; ++t) s += (t < a.nseg && k >= a.seg_row[t]) ? 1 : 0;
    return a.seg_cumsum[s][k - a.seg_row[s]] + a.offtot[s];
}

__device__ __forceinline__ double offset_r(const ResampleArgs& a) { return a.r_dev ? __ldg(a.r_dev) : a.r; }
__device__ __forceinline__ double sample_u(const ResampleArgs& a, double di) {
    return gse_sample_position(di, offset_r(a), a.n_total, a.inv_n, a.n_pow2 != 0);
}
__device__ __forceinline__ uint64_t output_qlo(const ResampleArgs& a, double di, uint64_t off, double Td) {
    const uint64_t qa = __double2ull_rd(__dmul_rn(sample_u(a, di), Td));
    const uint64_t lo = qa > 2 ? qa - 2 : 0;
    return lo > off ? lo - off : 0ull;
}
// does the source with LOCAL cumulative weight c precede the output with global index di?  (exact)
__device__ __forceinline__ bool precedes(const ResampleArgs& a, uint64_t c, uint64_t q_lo, double di,
                                         uint64_t off, double Td) {
    if (c < q_lo) return true;
    if (c >= q_lo + 4) return false;
    return __ddiv_rn(__ull2double_rn(c + off), Td) < sample_u(a, di);
}

// Warp-cooperative search for the smallest s in [lo, hi] with NOT precedes(C[s], output j(s)), where the
// predicate holds on a prefix; j(s) = fixed_out when that is >= 0 (split for one output), otherwise
// diag - 1 - (s - s_base) (merge-path diagonal).  32-ary: ~5 rounds of dependent loads at 2^24.
template <bool SEG>
__device__ __forceinline__ int64_t warp_split(const ResampleArgs& a, int64_t lo, int64_t hi, int64_t diag,
                                              int64_t s_base, int64_t fixed_out, uint64_t off0, double Td, int lane) {
    const uint64_t off = 0;                                     // source_weight() returns global weights
#define GSE_OUT_OF(mid) (double)(a.out0 + (fixed_out >= 0 ? fixed_out : diag - 1 - ((mid) - s_base)))
    if (SEG) {
        // the last row of shard t - 1 carries exactly the offset of shard t: probe the shard boundaries
        // with the offsets alone, so that the search below stays inside one shard's array
        for (int t = 1; t < a.nseg; ++t) {
            const int64_t mid = a.seg_row[t] - 1;
            if (mid < lo || mid >= hi) continue;
            const double di = GSE_OUT_OF(mid);
            if (precedes(a, a.offtot[t], output_qlo(a, di, off, Td), di, off, Td)) lo = mid + 1; else hi = mid;
        }
    }
    // end points first (one round of loads): ranges that lie entirely among sources without offspring
    // -- the other shards' rows in a sharded run, the tail of a degenerate weight vector -- finish here
    if (lo < hi) {
        const int64_t mid = lane == 0 ? lo : hi - 1;
        bool pred = false;
        if (lane < 2) {
            const uint64_t c = source_weight<SEG>(a, mid, off0);
            const double di = GSE_OUT_OF(mid);
            pred = precedes(a, c, output_qlo(a, di, off, Td), di, off, Td);
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, pred);
        if (!(bal & 1u)) hi = lo;                      // the first candidate already does not precede
        else if (bal & 2u) lo = hi;                    // every candidate precedes
        else { lo = lo + 1; hi = hi - 1; }
    }
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
            const uint64_t c = source_weight<SEG>(a, mid, off0);
            const double di = GSE_OUT_OF(mid);
            pred = precedes(a, c, output_qlo(a, di, off, Td), di, off, Td);
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
#undef GSE_OUT_OF
    return lo;
}

// Sharded run: only the sources between the ancestor of this shard's first output and the ancestor of
// its last output interleave with its outputs.  Two warps find that range [k_lo, k_hi) so that the
// partition and the search below cost what they cost on one GPU, whatever the number of shards.
template <bool SEG>
__global__ void __launch_bounds__(64)
k_resample_source_range(const ResampleArgs a, int64_t* __restrict__ range) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t off0 = SEG ? 0ull : a.offtot[0];
    const double Td = __ull2double_rn(a.offtot[SEG ? a.nseg : 1]);
    const int64_t s = warp_split<SEG>(a, 0, a.n_src, 0, 0, w == 0 ? 0 : a.n_out - 1, off0, Td, lane);
    if (lane == 0) range[w] = (w == 0) ? s : min(s + 1, a.n_src);
}

template <bool SEG>
__global__ void __launch_bounds__(128)
k_resample_partition(const ResampleArgs a, int64_t* __restrict__ part, int64_t nparts) {
    const int lane = threadIdx.x & 31;
    const int64_t k_lo = a.range ? a.range[0] : 0;
    const int64_t span = a.range ? a.range[1] - k_lo : a.n_src;
    const uint64_t off0 = SEG ? 0ull : a.offtot[0];
    const double Td = __ull2double_rn(a.offtot[SEG ? a.nseg : 1]);
    const int64_t total = span + a.n_out;
    // split points 0 .. ceil(total / W) are needed; the grid covers them with a stride (it is sized for the
    // expected span of a sharded run, not for all G shards' rows)
    const int64_t last = min(nparts, (total + RS_WORK - 1) / RS_WORK);
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b <= last; b += warps) {
        const int64_t diag = b * RS_WORK;
        if (diag >= total) {                                     // the end of the merged sequence
            if (lane == 0) part[b] = k_lo + span;
            continue;
        }
        const int64_t lo = k_lo + (diag > a.n_out ? diag - a.n_out : 0);
        const int64_t hi = k_lo + (diag < span ? diag : span);
        const int64_t s = warp_split<SEG>(a, lo, hi, diag, k_lo, -1, off0, Td, lane);
        if (lane == 0) part[b] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// K4: search by rank-and-fill.  Block b owns the merged elements [b*W, (b+1)*W): a window of ns
// sources and no outputs (ns + no <= 4096).  With g_k = fl(C_k / T) and u_j = (j + r) / N evaluated
// exactly as the reference does (particle.py:90,97), source k precedes output j iff g_k < u_j, and
// u_j is non-decreasing in j.  So every SOURCE computes its rank among the block's outputs
//     e_k = #{ j : u_j <= g_k } = first j with u_j > g_k
// directly: one float64 division, a float64 guess  floor(g_k N - r) + 1  and one or two exact
// evaluations of u around it -- no search at all.  Output j then has idx_j = a0 + #{k : e_k <= j}:
// the last source of every distinct rank drops the marker k + 1 at position e_k and a block-wide
// prefix maximum fills the runs.  Blocks without outputs (sources without offspring) return at
// once; blocks with few sources (heavy ancestors) cost a fill only.
// ------------------------------------------------------------------------------------------------
template <bool POW2>
__device__ __forceinline__ double sample_pos(const ResampleArgs& a, double di) {
    return gse_sample_position(di, offset_r(a), a.n_total, a.inv_n, POW2);
}

template <bool POW2, bool SEG>
__global__ void __launch_bounds__(RS_THREADS, 6)
k_resample_search(const ResampleArgs a, const int64_t* __restrict__ part, int32_t* __restrict__ idx_out) {
    __shared__ __align__(16) int s_mark[RS_WORK];          // markers, then their prefix maximum
    __shared__ int s_warp[RS_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t k_lo = a.range ? a.range[0] : 0;
    const int64_t total = (a.range ? a.range[1] - k_lo : a.n_src) + a.n_out;
    // windows with a grid stride: one per CTA on one GPU; in a sharded run the grid is sized for the expected
    // span of sources and the stride absorbs a skewed one
    int64_t b = blockIdx.x;
    if (b * RS_WORK >= total) return;
    do {                                                   // (one pass on one GPU: the grid covers every window)
    int64_t d0 = b * RS_WORK, d1 = d0 + RS_WORK;
    if (d1 > total) d1 = total;
    const int64_t a0 = part[b], a1 = part[b + 1];
    const int64_t o0 = d0 - (a0 - k_lo), o1 = d1 - (a1 - k_lo);
    if (o1 <= o0) continue;                                // a stretch of sources with no offspring
    const int ns = (int)(a1 - a0);
    const int no = (int)(o1 - o0);
    // sharded: a window inside one shard (all but G - 1 of them) is read like a local one
    int seg = 0;
    bool direct = true;
    if (SEG && ns > 0) {
        int seg_last = 0;
#pragma unroll
        for (int t = 1; t < GSE_MAX_SHARDS; ++t) {
            seg += (t < a.nseg && a0 >= a.seg_row[t]) ? 1 : 0;
            seg_last += (t < a.nseg && a1 - 1 >= a.seg_row[t]) ? 1 : 0;
        }
        direct = seg == seg_last;
    }
    const uint64_t off = SEG ? (direct ? a.offtot[seg] : 0ull) : a.offtot[0];
    const double Td = __ull2double_rn(a.offtot[SEG ? a.nseg : 1]);
    const double dbase = (double)(a.out0 + o0);            // global index of the block's first output
    const double dend = dbase + (double)no;
#pragma unroll
    for (int m = 0; m < RS_VT / 4; ++m)
        *reinterpret_cast<int4*>(s_mark + 4 * tid + m * (4 * RS_THREADS)) = make_int4(0, 0, 0, 0);
    // Fast path: with t* = (C/T) N - r in real arithmetic, u_i <= g_k  <=>  i <= t* up to the rounding
    // of u_i and g_k, which moves the boundary by less than 3 N 2^-53; t below is within another
    // 3 N 2^-53 of t*.  So when t is further than eps = N 2^-48 from an integer the rank is
    // floor(t) + 1, no division needed; otherwise (ties, ~2 eps of all sources) settle it exactly.
    const double rr = offset_r(a);
    const double inv_T = 1.0 / Td;     // (hoisting this division into the partition kernel measured 8 us SLOWER:
                                       //  it overlaps the window's load latency here, a dependent load does not)
    const double eps = a.n_total * 3.5527136788005009e-15;         // N * 2^-48
    const double one_m_eps = 1.0 - eps;
    const double one_m_base = 1.0 - dbase;
    // warp w owns the sources [512 w, 512 w + 512) of the window; lane l holds k = 512 w + 32 m + l
    const int kw = wid * (RS_VT * 32) + lane;
    int e[RS_VT];
    if (wid * (RS_VT * 32) < ns) {
        const uint64_t* src = (SEG ? a.seg_cumsum[seg] + (a0 - a.seg_row[seg]) : a.cumsum + a0) + kw;
#pragma unroll
        for (int h = 0; h < RS_VT; h += 8) {                // 8 loads in flight per thread
            uint64_t c[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                c[m] = 0ull;
                if (kw + 32 * (h + m) < ns)
                    c[m] = direct ? __ldg(src + 32 * (h + m)) : source_weight<true>(a, a0 + kw + 32 * (h + m), 0ull);
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                int r = -2;                                 // no source: equals no rank
                if (kw + 32 * (h + m) < ns) {
                    // (the cvt / floor instructions run on the XU pipe, 7-14 lanes/clk/SM, but exponent-trick
                    // replacements on the ALU pipe measured slower: tools/ubench_xu.cu, profiles/)
                    const double cd = __ull2double_rn(c[m] + off);
                    const double t = __fma_rn(__dmul_rn(cd, inv_T), a.n_total, -rr);
                    const double fl = floor(t);
                    const double fr = t - fl;
                    if (fr > eps && fr < one_m_eps)
                        r = __double2int_rz(fl + one_m_base);      // saturating; the rank is floor(t) + 1 - dbase
                    else
                        r = __double2int_rz(rank_exact<POW2>(rr, a.n_total, a.inv_n, cd, Td, fl + 1.0, dbase, dend) - dbase);
                    r = min(max(r, 0), no);
                }
                e[h + m] = r;
            }
        }
    }
    __syncthreads();                                        // markers are zeroed
    if (wid * (RS_VT * 32) < ns) {
        // the last source of every distinct rank drops its marker: the successor's rank comes from the
        // next lane (or lane 0 of the next round); the warp's very last source cannot see its successor
        // and uses a maximum instead (a later writer of the same rank always carries a larger k)
#pragma unroll
        for (int m = 0; m < RS_VT; ++m) {
            int nxt = __shfl_down_sync(0xffffffffu, e[m], 1);
            const int wrap = __shfl_sync(0xffffffffu, e[m + 1 < RS_VT ? m + 1 : m], 0);
            if (lane == 31) nxt = wrap;
            const int r = e[m];
            if (r >= 0 && r < no) {
                if (m == RS_VT - 1 && lane == 31) atomicMax(&s_mark[r], kw + 32 * m + 1);
                else if (nxt != r) s_mark[r] = kw + 32 * m + 1;
            }
        }
    }
    __syncthreads();
    // prefix maximum of the markers: lane l of warp w owns positions [512 w + 16 l, + 16)
    int v[RS_VT];
    const bool active = wid * (RS_VT * 32) < no;           // warps beyond the outputs have nothing to fill
    if (active) {
        const int4* p = reinterpret_cast<const int4*>(s_mark + wid * (RS_VT * 32) + RS_VT * lane);
#pragma unroll
        for (int m = 0; m < RS_VT / 4; ++m) {
            const int4 x = p[m];
            v[4 * m + 0] = x.x; v[4 * m + 1] = x.y; v[4 * m + 2] = x.z; v[4 * m + 3] = x.w;
        }
#pragma unroll
        for (int r = 1; r < RS_VT; ++r) v[r] = max(v[r], v[r - 1]);
        int incl = v[RS_VT - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl = max(incl, t);
        }
        int excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0;
#pragma unroll
        for (int r = 0; r < RS_VT; ++r) v[r] = max(v[r], excl);
        if (lane == 31) s_warp[wid] = incl;
    }
    __syncthreads();
    if (active) {
        int wbase = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32 - 1; ++w) wbase = max(wbase, w < wid ? s_warp[w] : 0);
        int4* p = reinterpret_cast<int4*>(s_mark + wid * (RS_VT * 32) + RS_VT * lane);
#pragma unroll
        for (int m = 0; m < RS_VT / 4; ++m)
            p[m] = make_int4(max(v[4 * m], wbase), max(v[4 * m + 1], wbase), max(v[4 * m + 2], wbase), max(v[4 * m + 3], wbase));
    }
    __syncthreads();
    // idx = a0 + (sources of this block merged before the output), clamped to n_src - 1 (only reachable
    // through a degenerate total)
    const int64_t room = a.n_src - 1 - a0;
    const int cap = room < 0 ? 0 : (room > RS_WORK ? RS_WORK : (int)room);
    const int base = (int)(a0 + (room < 0 ? room : 0));
    int32_t* out = idx_out + o0;
    for (int j = tid; j < no; j += RS_THREADS) out[j] = base + min(s_mark[j], cap);
    if (SEG) __syncthreads();                              // s_mark is reused by the next window
    } while (SEG && (b += gridDim.x) * RS_WORK < total);
}

static int launch_search(gse_ctx* ctx, const ResampleArgs& a, int64_t nparts_all, int32_t* idx_out_dev, bool seg,
                         cudaStream_t s) {
    int64_t nparts = nparts_all;
    // sharded: the sources that interleave with this shard's outputs are ~n_out of them in steady state
    // (balanced weights); both kernels stride over the windows, so a grid sized for 2.5 n_out merged
    // elements is correct for any skew and does not grow with the number of shards
    if (seg) {
        const int64_t expected = gse_div_up(a.n_out * 5 / 2 + RS_WORK, RS_WORK);
        if (nparts > expected) nparts = expected;
    }
    const unsigned pgrid = (unsigned)gse_div_up((nparts + 1) * 32, 128);
    if (seg) {
        k_resample_source_range<true><<<1, 64, 0, s>>>(a, ctx->range);
        GSE_CHECK_LAUNCH(ctx);
        k_resample_partition<true><<<pgrid, 128, 0, s>>>(a, ctx->part, nparts_all);
    } else {
        k_resample_partition<false><<<pgrid, 128, 0, s>>>(a, ctx->part, nparts);
    }
    GSE_CHECK_LAUNCH(ctx);
    const unsigned g = (unsigned)nparts;
    if (seg) {
        if (a.n_pow2) k_resample_search<true, true><<<g, RS_THREADS, 0, s>>>(a, ctx->part, idx_out_dev);
        else k_resample_search<false, true><<<g, RS_THREADS, 0, s>>>(a, ctx->part, idx_out_dev);
    } else {
        if (a.n_pow2) k_resample_search<true, false><<<g, RS_THREADS, 0, s>>>(a, ctx->part, idx_out_dev);
        else k_resample_search<false, false><<<g, RS_THREADS, 0, s>>>(a, ctx->part, idx_out_dev);
    }
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

static void fill_common(gse_ctx* ctx, ResampleArgs& a, double r, int64_t n_total, int64_t out0, int64_t n_out) {
    a.consts = (double*)(ctx->range + 4);
    a.r_dev = ctx->step_params ? &ctx->step_params->r : NULL;
    a.n_out = n_out;
    a.out0 = out0;
    a.r = r;
    a.n_total = (double)n_total;
    a.inv_n = 1.0 / (double)n_total;
    a.n_pow2 = ((n_total & (n_total - 1)) == 0) ? 1 : 0;
}

// Sharded run: the sources are the rows of ALL shards (their cumulative-weight arrays are peer
// memory mapped into this process), the outputs are this shard's own slots.  One kernel does the
// search and, through source_weight(), the communication: nothing is staged or sent.
extern "C" int gse_resample_search_sharded(gse_ctx* ctx, const gse_shards* sh, double r, int64_t out0, int64_t n_out,
                                           int32_t* idx_out_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && sh != NULL && idx_out_dev != NULL, "ctx / shards / idx is NULL");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(sh->nshards >= 1 && sh->nshards <= GSE_MAX_SHARDS && sh->offsets_dev != NULL, "bad shard table");
    const int64_t n_total = sh->rows[sh->nshards];
    GSE_REQUIRE(sh->rows[0] == 0 && n_total >= 1 && n_total <= 0x7fffffff, "global row count out of range (int32 index)");
    GSE_REQUIRE(n_out >= 0 && n_out <= ctx->n_max && out0 >= 0 && out0 + n_out <= n_total, "output range out of range");
    GSE_REQUIRE(r >= 0.0 && r < 1.0, "r must be in [0, 1)");
    if (n_out == 0) return GSE_OK;
    ResampleArgs a;
    memset(&a, 0, sizeof(a));
    a.offtot = sh->offsets_dev;
    a.n_src = n_total;
    a.nseg = sh->nshards;
    for (int t = 0; t <= sh->nshards; ++t) a.seg_row[t] = sh->rows[t];
    for (int t = 0; t < sh->nshards; ++t) {
        GSE_REQUIRE(sh->cumsum_dev[t] != NULL && sh->rows[t + 1] > sh->rows[t], "bad shard entry");
        a.seg_cumsum[t] = sh->cumsum_dev[t];
    }
    fill_common(ctx, a, r, n_total, out0, n_out);
    a.range = ctx->range;
    const int64_t nparts = gse_div_up(a.n_src + n_out, RS_WORK);
    GSE_REQUIRE(nparts + 1 <= ctx->max_tiles + 2, "workspace too small (create the context with n_max >= global rows)");
    return launch_search(ctx, a, nparts, idx_out_dev, true, (cudaStream_t)stream);
}

extern "C" int gse_resample_search(gse_ctx* ctx, const uint64_t* cumsum_dev, int64_t n_src,
                                   const uint64_t* offtot_dev, double r, int64_t n_total, int64_t out0,
                                   int64_t n_out, int32_t* idx_out_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && cumsum_dev != NULL && offtot_dev != NULL && idx_out_dev != NULL,
                "ctx / cumsum / offtot / idx is NULL");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n_src >= 1 && n_src <= ctx->n_max && n_src <= 0x7fffffff, "n_src out of range for this context");
    GSE_REQUIRE(n_out >= 0 && n_out <= ctx->n_max, "n_out out of range for this context");
    GSE_REQUIRE(n_total >= 1 && out0 >= 0 && out0 + n_out <= n_total, "output range outside [0, n_total)");
    GSE_REQUIRE(r >= 0.0 && r < 1.0, "r must be in [0, 1)");
    if (n_out == 0) return GSE_OK;
    ResampleArgs a;
    memset(&a, 0, sizeof(a));
    a.cumsum = cumsum_dev;
    a.offtot = offtot_dev;
    a.n_src = n_src;
    a.nseg = 1;
    fill_common(ctx, a, r, n_total, out0, n_out);
    const int64_t nparts = gse_div_up(n_src + n_out, RS_WORK);
    GSE_REQUIRE(nparts + 1 <= ctx->max_tiles + 2, "workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    return launch_search(ctx, a, nparts, idx_out_dev, false, s);
}

// ------------------------------------------------------------------------------------------------
// K5: materialising gather  dst[:, i] = src[:, idx_i]  (particles[sample_index], particle.py:102 /
// :315).  Only used when the resampled rows themselves are needed (the `particles` attribute, the
// sharded slab exchange): the filters normally hand idx to the next predict / moments kernel, which
// reads its rows through it.  idx is non-decreasing, so the reads of a warp fall in one short
// window of each column (or on one address when a heavy source has many offspring).
// ------------------------------------------------------------------------------------------------
#define GR_THREADS 256
template <bool VEC>
__global__ void __launch_bounds__(GR_THREADS)
k_gather_rows(const int32_t* __restrict__ idx, int64_t n_out, const float* __restrict__ src, int64_t ld_src,
              float* __restrict__ dst, int64_t ld_dst, int ncols, float* __restrict__ loglik_out) {
    if (VEC) {
        const int64_t row0 = ((int64_t)blockIdx.x * GR_THREADS + threadIdx.x) * 4;
        if (row0 >= n_out) return;
        if (row0 + 4 <= n_out) {
            const int4 id = *reinterpret_cast<const int4*>(idx + row0);
            for (int c = 0; c < ncols; ++c) {
                const float* col = src + c * ld_src;
                st_stream4(dst + c * ld_dst + row0,
                           make_float4(__ldg(col + id.x), __ldg(col + id.y), __ldg(col + id.z), __ldg(col + id.w)));
            }
            if (loglik_out) st_stream4(loglik_out + row0, make_float4(0.f, 0.f, 0.f, 0.f));
        } else {                                           // ragged tail: never touch rows >= n_out
            for (int64_t i = row0; i < n_out; ++i) {
                const int k = idx[i];
                for (int c = 0; c < ncols; ++c) dst[c * ld_dst + i] = __ldg(src + c * ld_src + k);
                if (loglik_out) loglik_out[i] = 0.0f;
            }
        }
    } else {
        const int64_t i = (int64_t)blockIdx.x * GR_THREADS + threadIdx.x;
        if (i >= n_out) return;
        const int k = idx[i];
        for (int c = 0; c < ncols; ++c) dst[c * ld_dst + i] = __ldg(src + c * ld_src + k);
        if (loglik_out) loglik_out[i] = 0.0f;
    }
}

// dst[:, i] = state of global row idx[i], pulled from whichever shard owns it (peer memory over NVLink)
template <bool VEC>
__global__ void __launch_bounds__(GR_THREADS)
k_gather_rows_sharded(const __grid_constant__ GatherShards g_arg, const int32_t* __restrict__ idx, int64_t n_out,
                      float* __restrict__ dst, int64_t ld_dst, int ncols) {
    const GatherShards& g = g_arg;
    if (VEC) {
        const int64_t row0 = ((int64_t)blockIdx.x * GR_THREADS + threadIdx.x) * 4;
        if (row0 >= n_out) return;
        if (row0 + 4 <= n_out) {
            const int4 id4 = *reinterpret_cast<const int4*>(idx + row0);
            const int id[4] = {id4.x, id4.y, id4.z, id4.w};
            const float* p[4];
            int64_t l[4];
            shard_rows4(g, id, p, l);
            for (int c = 0; c < ncols; ++c)
                st_stream4(dst + c * ld_dst + row0, make_float4(p[0][c * l[0]], p[1][c * l[1]], p[2][c * l[2]], p[3][c * l[3]]));
            return;
        }
        for (int64_t i = row0; i < n_out; ++i) {
            int64_t ld;
            const float* src = shard_row(g, idx[i], ld);
            for (int c = 0; c < ncols; ++c) dst[c * ld_dst + i] = src[c * ld];
        }
    } else {
        const int64_t i = (int64_t)blockIdx.x * GR_THREADS + threadIdx.x;
        if (i >= n_out) return;
        int64_t ld;
        const float* src = shard_row(g, idx[i], ld);
        for (int c = 0; c < ncols; ++c) dst[c * ld_dst + i] = src[c * ld];
    }
}

extern "C" int gse_gather_rows_sharded(gse_ctx* ctx, const gse_shards* sh, const int32_t* idx_dev, int64_t n_out,
                                       float* dst_dev, int64_t ld_dst, int ncols, void* stream) {
    GSE_REQUIRE(ctx != NULL && sh != NULL && idx_dev != NULL && dst_dev != NULL, "NULL argument");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(sh->nshards >= 1 && sh->nshards <= GSE_MAX_SHARDS, "bad shard table");
    GSE_REQUIRE(n_out >= 0 && ncols >= 1 && ld_dst >= n_out, "n_out / ncols / ld out of range");
    if (n_out == 0) return GSE_OK;
    GatherShards g;
    int rc = gse_build_gather_shards(sh, dst_dev, &g);
    if (rc) return rc;
    const bool vec = (((uintptr_t)idx_dev | (uintptr_t)dst_dev) & 15u) == 0 && ld_dst % 4 == 0;
    if (vec)
        k_gather_rows_sharded<true><<<(unsigned)gse_div_up(gse_div_up(n_out, 4), GR_THREADS), GR_THREADS, 0,
                                      (cudaStream_t)stream>>>(g, idx_dev, n_out, dst_dev, ld_dst, ncols);
    else
        k_gather_rows_sharded<false><<<(unsigned)gse_div_up(n_out, GR_THREADS), GR_THREADS, 0, (cudaStream_t)stream>>>(
            g, idx_dev, n_out, dst_dev, ld_dst, ncols);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

extern "C" int gse_gather_rows(gse_ctx* ctx, const int32_t* idx_dev, int64_t n_out, const float* src_dev,
                               int64_t ld_src, float* dst_dev, int64_t ld_dst, int ncols,
                               float* loglik_out_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && idx_dev != NULL && src_dev != NULL && dst_dev != NULL, "NULL argument");
    gse_device_guard guard(ctx->device);
    GSE_REQUIRE(n_out >= 0 && ncols >= 1 && ld_dst >= n_out, "n_out / ncols / ld out of range");
    GSE_REQUIRE(src_dev != dst_dev, "gather cannot be done in place");
    if (n_out == 0) return GSE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    // vector path: idx, dst columns and loglik 16-byte aligned
    const bool vec = (((uintptr_t)idx_dev | (uintptr_t)dst_dev | (uintptr_t)loglik_out_dev) & 15u) == 0 &&
                     ld_dst % 4 == 0;
    if (vec)
        k_gather_rows<true><<<(unsigned)gse_div_up(gse_div_up(n_out, 4), GR_THREADS), GR_THREADS, 0, s>>>(
            idx_dev, n_out, src_dev, ld_src, dst_dev, ld_dst, ncols, loglik_out_dev);
    else
        k_gather_rows<false><<<(unsigned)gse_div_up(n_out, GR_THREADS), GR_THREADS, 0, s>>>(
            idx_dev, n_out, src_dev, ld_src, dst_dev, ld_dst, ncols, loglik_out_dev);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}
