// Systematic resampling in ONE launch: weight scan + rank + fill (K3 + K4 fused; particle.py:85-100 / :296-314).
//
// The cumulative weights never reach HBM.  Every CTA owns one contiguous run of source rows and the grid is one wave
// of co-resident CTAs, so the kernel may wait on itself:
//   phase 1  each warp sums the fixed-point weights of its sub-run (streaming read of loglik, 4 B/row);
//   phase 2  every CTA publishes its aggregate in one status word and reads ALL aggregates: the sum of the CTAs before
//            it is its exclusive prefix, the sum of all is the total T the ranks need;
//   phase 3  warp-autonomous, no block barrier: per tile of 32*ITEMS rows the warp re-reads its rows (L2), scans them
//            with shuffles on top of a register carry, turns every cumulative weight C_k into the RANK
//                e_k = #{ outputs j : u_j <= fl(C_k / T) },   u_j = (j + r) / N        (particle.py:90,97-98)
//            -- so that source k owns exactly the outputs [e_{k-1}, e_k) -- and fills them: the first output of every
//            source that has offspring gets a marker in a per-warp shared-memory window, a prefix maximum spreads it
//            over the run, and the window is stored coalesced.  4 B/row read + 4 B/row written.
//   tail     runs of more than HEAVY_MIN outputs of one source (degenerate weights) are not filled by the warp that
//            found them but queued; once every CTA has finished phase 3 all warps of the grid drain the queue.
//
// The cumulative weights are exact integers (< 2^53) carried in float64 -- the same values the uint64 scan of
// gse_resample.cu produces -- so the ancestor indices are bit-identical to searchsorted(cumsum / cumsum[-1], u, 'left')
// on those integers.  Everything on the per-row path avoids the XU pipe except the one ex2: integer <-> float64
// conversions, floor and cvt.u64.f32 (7-14 lanes/clk/SM, profiles/r1_ubench_xu_pipe.txt) are replaced by float64
// adds of 2^52-type constants.
//
// gse_resample_search_f64 feeds the same rank + fill code with a caller's own float64 cumulative sum
// (`resample_from_cumsum`, SURVEY.md section 7 contract (ii)).
#include "gse_resample_common.cuh"

#define RF_THREADS 256
#define RF_WARPS (RF_THREADS / 32)
#define RF_HEAVY_MIN 4096        // a warp fills at most this many outputs of one source beyond the current window itself
#define RF_PIECE 65536           // queued runs are cut into pieces of at most this many outputs (one warp each)

#define RF_MAGIC 6755399441055744.0      // 1.5 * 2^52: (t + MAGIC) - MAGIC = rint(t), low word of (t + MAGIC) = (int)rint(t)
#define RF_TWO52 4503599627370496.0

struct FusedArgs {
    const float* loglik;       // NULL: weights = base
    const double* base;        // NULL: weights = exp(loglik - M)
    const double* stats;       // [0] M, [1] S
    const double* cumsum;      // f64 entry: the caller's cumulative sum (normalised unless `normalise`)
    int64_t n_src;
    int64_t rows_per_block;    // multiple of RF_WARPS * 32 * ITEMS
    uint64_t* status;          // one word per CTA, zero between launches
    unsigned int* counters;    // [0] start ticket, [1] CTAs past phase 3, [2] queue length, [3] CTAs past the drain
    int4* queue;
    int queue_cap;
    unsigned int* err;         // device error word of the context (host-mapped)
    double r;
    const double* r_dev;       // device override of r (parameter block of a captured graph), or NULL
    double n_total;
    double inv_n;
    int n_total_i;
    int out_lo, out_hi;        // global outputs this launch writes: idx_out[j - out_lo], j in [out_lo, out_hi)
    int32_t* idx_out;
    int src_row0;              // global row of local source row 0 (added to the stored ancestor index)
    int first_shard;           // local row 0 is the first row of the whole population (e_{-1} = 0)
    int normalise;             // f64 entry: divide by cumsum[n_src - 1] on the fly (`cumsum /= cumsum[-1]`, :90)
    uint64_t* total_out;       // receives the integer total (NULL to skip)
};

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// rint for 0 <= x (already an integer from 2^52 on), without the conversion pipe
__device__ __forceinline__ double rint_nonneg(double x) {
    const double y = __dadd_rn(__dadd_rn(x, RF_TWO52), -RF_TWO52);
    return x >= RF_TWO52 ? x : y;
}

// Fixed-point weights of ITEMS consecutive rows as float64 integers: the values of quantise16() in gse_resample.cu
// (rint(fl32(exp(l - M)) * 2^s), or rint(base * fl32(exp(l - M)) * 2^s) in float64 with a base), rows >= n_end are 0.
template <int ITEMS, bool HAS_LL, bool HAS_BASE>
__device__ __forceinline__ void quantise_rows(const float* __restrict__ loglik, const double* __restrict__ base, float M,
                                              float scale_f, double scale_d, int64_t row0, int64_t n_end,
                                              double q[ITEMS]) {
    static_assert(ITEMS == 8 || ITEMS == 16, "ITEMS");
    const bool full = row0 + ITEMS <= n_end;
    float e[ITEMS];
    if (HAS_LL) {
        float l[ITEMS];
        if (full) {
#pragma unroll
            for (int v = 0; v < ITEMS / 8; ++v) ld_f32x8(loglik + row0 + 8 * v, l + 8 * v);
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) l[k] = (row0 + k < n_end) ? loglik[row0 + k] : -INFINITY;
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) e[k] = __expf(__fsub_rn(l[k], M));
    }
    if (!HAS_BASE) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) q[k] = rint_nonneg((double)__fmul_rn(e[k], scale_f));
    } else {
        double b[ITEMS];
        if (full) {
#pragma unroll
            for (int v = 0; v < ITEMS / 4; ++v) ld_f64x4(base + row0 + 4 * v, b + 4 * v);
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) b[k] = (row0 + k < n_end) ? base[row0 + k] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            double w = b[k];
            if (HAS_LL) w = __dmul_rn(w, (double)e[k]);
            q[k] = rint_nonneg(fmax(__dmul_rn(w, scale_d), 0.0));
        }
    }
    if (HAS_LL && !HAS_BASE && !full) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) if (row0 + k >= n_end) q[k] = 0.0;
    }
}

struct RankConsts {
    double inv_T, Td, n_total, inv_n, r, eps;
    int n_total_i;
};

// e = #{ outputs j in [0, N) : u_j <= g }  (TIES_RIGHT: u_j < g), g = fl(C / T) for the integer weights, g = C for a
// caller's normalised cumulative sum (Td = 1).  Fast path: with t* = g N - r in real arithmetic, u_j <= g  <=>
// j <= t* up to the rounding of u_j and g, which moves the boundary by less than 3 N 2^-53; t below is within another
// 3 N 2^-53 of t*.  So when t is further than eps = N 2^-48 from an integer the rank is floor(t) + 1; otherwise
// (exact ties -- dyadic weights, r = 0 -- and ~2 eps of random sources) the comparison is evaluated as the reference does.
template <bool POW2, bool TIES_RIGHT, bool NORMALISED>
__device__ __forceinline__ int rank_of(double C, const RankConsts& k) {
    const double g_fast = NORMALISED ? C : __dmul_rn(C, k.inv_T);
    const double t = __fma_rn(g_fast, k.n_total, -k.r);
    const double y = __dadd_rn(t, RF_MAGIC);
    const double rn = __dadd_rn(y, -RF_MAGIC);
    const double diff = __dadd_rn(t, -rn);                        // t - rint(t), exact
    int e = __double2loint(y) + (diff >= 0.0 ? 1 : 0);            // floor(t) + 1
    if (!(fabs(diff) > k.eps)) {
        const double g = NORMALISED ? C : __ddiv_rn(C, k.Td);
        e = __double2int_rz(rank_exact_g<POW2, TIES_RIGHT>(k.r, k.n_total, k.inv_n, g, rn + 1.0, 0.0, k.n_total));
    }
    if (t >= k.n_total) e = k.n_total_i;
    return min(max(e, 0), k.n_total_i);
}

// Outputs [max(E0, out_lo), min(E1, out_hi)) of the warp's tile: source (lane, k) of the tile owns the outputs
// [start, e[k]) with start = e[k - 1] (ep for the lane's first row).  s_mark / s_end: 32 * ITEMS ints each, this warp's.
template <int ITEMS>
__device__ __forceinline__ void warp_fill(const int (&e)[ITEMS], int ep, int E0, int E1, int src_base, const FusedArgs& a,
                                          int* s_mark, int* s_end, int lane) {
    constexpr int TILE = 32 * ITEMS;
    const int lo = max(E0, a.out_lo), hi = min(E1, a.out_hi);
    if (lo >= hi) return;                                         // warp-uniform
    int* my_end = s_end + lane * ITEMS;
    int* my_mark = s_mark + lane * ITEMS;
#pragma unroll
    for (int v = 0; v < ITEMS / 4; ++v)
        *reinterpret_cast<int4*>(my_end + 4 * v) = make_int4(e[4 * v], e[4 * v + 1], e[4 * v + 2], e[4 * v + 3]);
    // the source that covers output lo: 1 + #{ sources of the tile with e <= lo }  (1-based index into the tile)
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) cnt += (e[k] <= lo) ? 1 : 0;
    int carry_m = warp_sum_i32(cnt) + 1;
    int wb = lo;
    while (wb < hi) {
#pragma unroll
        for (int v = 0; v < ITEMS / 4; ++v) *reinterpret_cast<int4*>(my_mark + 4 * v) = make_int4(0, 0, 0, 0);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int start = k ? e[k - 1] : ep;
            const unsigned int p = (unsigned int)(start - wb);   // start < wb wraps to a huge value
            if (e[k] > start && p < (unsigned int)TILE) s_mark[p] = lane * ITEMS + k + 1;
        }
        __syncwarp();
        int v[ITEMS];
#pragma unroll
        for (int m = 0; m < ITEMS / 4; ++m) {
            const int4 x = *reinterpret_cast<const int4*>(my_mark + 4 * m);
            v[4 * m] = x.x; v[4 * m + 1] = x.y; v[4 * m + 2] = x.z; v[4 * m + 3] = x.w;
        }
#pragma unroll
        for (int k = 1; k < ITEMS; ++k) v[k] = max(v[k], v[k - 1]);
        int incl = v[ITEMS - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl = max(incl, t);
        }
        int excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0;
        excl = max(excl, carry_m);
#pragma unroll
        for (int m = 0; m < ITEMS / 4; ++m)
            *reinterpret_cast<int4*>(my_mark + 4 * m) = make_int4(max(v[4 * m], excl), max(v[4 * m + 1], excl),
                                                                  max(v[4 * m + 2], excl), max(v[4 * m + 3], excl));
        __syncwarp();
        const int cntw = min(hi - wb, TILE);
        int32_t* out = a.idx_out + (wb - a.out_lo);
        const int vbase = src_base - 1;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int pos = i * 32 + lane;
            if (pos < cntw) out[pos] = vbase + s_mark[pos];
        }
        carry_m = max(__shfl_sync(0xffffffffu, incl, 31), carry_m);
        wb += TILE;
        if (wb < hi) {
            const int end_c = s_end[carry_m - 1];                // where the covering source's run ends
            if (end_c - wb > RF_HEAVY_MIN) {                     // a heavy source: queue the rest of its run
                const int target = min(end_c, hi);
                if (lane == 0) {
                    const int pieces = (target - wb + RF_PIECE - 1) / RF_PIECE;
                    const int slot = (int)atomicAdd(a.counters + 2, (unsigned int)pieces);
                    for (int p = 0; p < pieces; ++p) {
                        if (slot + p < a.queue_cap)
                            a.queue[slot + p] = make_int4(wb + p * RF_PIECE, min(wb + (p + 1) * RF_PIECE, target),
                                                          vbase + carry_m, 0);
                        else
                            atomicOr(a.err, GSE_ERR_QUEUE_OVERFLOW);
                    }
                }
                wb = target;
            }
        }
        __syncwarp();
    }
}

// drain of the heavy-run queue by every warp of the grid (after all CTAs have finished phase 3)
__device__ __forceinline__ void drain_queue(const FusedArgs& a, int vb, int nblocks, int lane, int wid) {
    const int qn = min((int)ld_status32(a.counters + 2), a.queue_cap);
    for (int i = vb * RF_WARPS + wid; i < qn; i += nblocks * RF_WARPS) {
        const int4 ent = __ldcg(a.queue + i);
        int32_t* out = a.idx_out - a.out_lo;
        int p = ent.x;
        const int end = ent.y, val = ent.z;
        // head up to 16-byte alignment, 128-bit body, tail
        const int head = min(end, p + (int)((4 - (((uintptr_t)(out + p) >> 2) & 3)) & 3));
        if (p + lane < head) out[p + lane] = val;
        p = head;
        const int4 v4 = make_int4(val, val, val, val);
        for (int q = p + 4 * lane; q + 4 <= end; q += 128) *reinterpret_cast<int4*>(out + q) = v4;
        const int body_end = p + ((end - p) & ~3);
        if (body_end + lane < end) out[body_end + lane] = val;
    }
}

template <int ITEMS, bool HAS_LL, bool HAS_BASE, bool POW2>
__global__ void __launch_bounds__(RF_THREADS, ITEMS == 8 ? 4 : 3)
k_resample_fused(const __grid_constant__ FusedArgs a) {
    constexpr int TILE = 32 * ITEMS;
    __shared__ __align__(16) int s_mark[RF_WARPS][TILE];
    __shared__ __align__(16) int s_end[RF_WARPS][TILE];
    __shared__ double s_wsum[RF_WARPS];
    __shared__ double s_excl, s_total;
    __shared__ unsigned int s_vb;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nblocks = gridDim.x;
    if (tid == 0) s_vb = atomicAdd(a.counters, 1u);               // CTAs are numbered in the order they start
    __syncthreads();
    const int vb = (int)s_vb;
    const float M = HAS_LL ? (float)a.stats[0] : 0.0f;
    const int sexp = quantisation_exponent(a.stats[1]);
    const float scale_f = __int_as_float((127 + sexp) << 23);     // 2^sexp, 0 <= sexp <= 52
    const double scale_d = ldexp(1.0, sexp);
    // this CTA owns rows [b0, b1), warp w the contiguous sub-run [w0, w1) of it (whole tiles)
    const int64_t b0 = (int64_t)vb * a.rows_per_block, b1 = min(b0 + a.rows_per_block, a.n_src);
    const int64_t rows_per_warp = a.rows_per_block / RF_WARPS;
    const int64_t w0 = min(b0 + wid * rows_per_warp, b1), w1 = min(w0 + rows_per_warp, b1);

    // ---- phase 1: sum of the warp's sub-run -------------------------------------------------------------------
    double sum = 0.0;
    for (int64_t t0 = w0; t0 < w1; t0 += TILE) {
        const int64_t row0 = t0 + (int64_t)lane * ITEMS;
        if (row0 < w1) {
            double q[ITEMS];
            quantise_rows<ITEMS, HAS_LL, HAS_BASE>(a.loglik, a.base, M, scale_f, scale_d, row0, w1, q);
#pragma unroll
            for (int k = 0; k < ITEMS; k += 2) sum += q[k] + q[k + 1];
        }
    }
    sum = warp_sum_f64(sum);
    if (lane == 0) s_wsum[wid] = sum;
    __syncthreads();

    // ---- phase 2: publish the CTA aggregate, read all of them -------------------------------------------------
    if (wid == 0) {
        double agg = 0.0;
#pragma unroll
        for (int w = 0; w < RF_WARPS; ++w) agg += s_wsum[w];
        if (lane == 0) st_status(a.status + vb, status_pack(ST_AGGREGATE, 0u, (uint64_t)__double2ll_rn(agg)));
        uint64_t excl = 0, tot = 0;
        for (int base = 0; base < nblocks; base += 32) {
            const int i = base + lane;
            if (i < nblocks) {
                uint64_t word = ld_status(a.status + i);
                while ((word >> 62) == 0ull) { __nanosleep(64); word = ld_status(a.status + i); }
                const uint64_t v = word & ((1ull << 54) - 1ull);
                tot += v;
                if (i < vb) excl += v;
            }
        }
        excl = warp_sum_u64(excl);
        tot = warp_sum_u64(tot);
        if (lane == 0) {
            s_excl = __ull2double_rn(excl);
            s_total = __ull2double_rn(tot);
            if (vb == 0 && a.total_out) *a.total_out = tot;
        }
    }
    __syncthreads();

    // ---- phase 3: scan + rank + fill, one warp per tile, no block barrier -------------------------------------
    RankConsts rc;
    rc.Td = s_total;
    rc.inv_T = 1.0 / rc.Td;
    rc.n_total = a.n_total;
    rc.inv_n = a.inv_n;
    rc.r = a.r_dev ? __ldg(a.r_dev) : a.r;
    rc.eps = a.n_total * 3.5527136788005009e-15;                  // N * 2^-48
    rc.n_total_i = a.n_total_i;
    const bool degenerate = !(rc.Td > 0.0);                       // all weights zero: everything descends from the last row
    double carry = s_excl;
#pragma unroll
    for (int w = 0; w < RF_WARPS; ++w) carry += (w < wid) ? s_wsum[w] : 0.0;
    int carry_rank = 0;
    if (w0 < w1) {
        if (!(a.first_shard && w0 == 0) && !degenerate) carry_rank = rank_of<POW2, false, false>(carry, rc);
        if (degenerate) {
            if (lane == 0 && vb == 0 && wid == 0) atomicOr(a.err, GSE_ERR_ZERO_WEIGHTS);
        }
    }
    for (int64_t t0 = w0; t0 < w1; t0 += TILE) {
        const int64_t row0 = t0 + (int64_t)lane * ITEMS;
        double q[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) q[k] = 0.0;
        if (row0 < w1) quantise_rows<ITEMS, HAS_LL, HAS_BASE>(a.loglik, a.base, M, scale_f, scale_d, row0, w1, q);
#pragma unroll
        for (int k = 1; k < ITEMS; ++k) q[k] += q[k - 1];
        double incl = q[ITEMS - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const double off = carry + (incl - q[ITEMS - 1]);
        int e[ITEMS];
        if (!degenerate) {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) e[k] = rank_of<POW2, false, false>(off + q[k], rc);
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k)          // only the last row of the whole population has offspring
                e[k] = (row0 + k >= a.n_src - 1 && a.src_row0 + a.n_src >= (int64_t)a.n_total_i) ? a.n_total_i : 0;
        }
        int ep = __shfl_up_sync(0xffffffffu, e[ITEMS - 1], 1);
        if (lane == 0) ep = carry_rank;
        const int E1 = __shfl_sync(0xffffffffu, e[ITEMS - 1], 31);
        warp_fill<ITEMS>(e, ep, carry_rank, E1, a.src_row0 + (int)t0, a, s_mark[wid], s_end[wid], lane);
        carry += __shfl_sync(0xffffffffu, incl, 31);
        carry_rank = E1;
    }

    // ---- tail: wait for every CTA, drain the heavy-run queue, reset the launch state --------------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicAdd(a.counters + 1, 1u);
        while (ld_status32(a.counters + 1) < (unsigned int)nblocks) __nanosleep(100);
        __threadfence();
    }
    __syncthreads();
    drain_queue(a, vb, nblocks, lane, wid);
    __syncthreads();
    if (tid == 0) s_vb = atomicAdd(a.counters + 3, 1u);
    __syncthreads();
    if (s_vb == (unsigned int)nblocks - 1u) {                     // last CTA out: leave everything zero for the next launch
        for (int i = tid; i < nblocks; i += RF_THREADS) a.status[i] = 0ull;
        if (tid < 4) a.counters[tid] = 0u;
    }
}

// ------------------------------------------------------------------------------------------------
// rank + fill on a caller's float64 cumulative sum: no scan, no inter-CTA dependency (every tile reads the element
// before it for e_{k-1}), any grid.  TIES_RIGHT = the reference's GPU kernel (`cumsum[k] > u` walk,
// particle.py:223-263, searchsorted side='right'); otherwise the CPU loop (`cumsum[k] < u`, :96-100, side='left').
// ------------------------------------------------------------------------------------------------
template <int ITEMS, bool POW2, bool TIES_RIGHT>
__global__ void __launch_bounds__(RF_THREADS, 4)
k_resample_search_f64(const __grid_constant__ FusedArgs a, int64_t ntiles) {
    constexpr int TILE = 32 * ITEMS;
    __shared__ __align__(16) int s_mark[RF_WARPS][TILE];
    __shared__ __align__(16) int s_end[RF_WARPS][TILE];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    RankConsts rc;
    rc.Td = a.normalise ? a.cumsum[a.n_src - 1] : 1.0;
    rc.inv_T = 1.0;
    rc.n_total = a.n_total;
    rc.inv_n = a.inv_n;
    rc.r = a.r_dev ? __ldg(a.r_dev) : a.r;
    rc.eps = a.n_total * 3.5527136788005009e-15;
    rc.n_total_i = a.n_total_i;
    const int64_t warps = (int64_t)gridDim.x * RF_WARPS;
    for (int64_t tile = (int64_t)blockIdx.x * RF_WARPS + wid; tile < ntiles; tile += warps) {
        const int64_t t0 = tile * TILE, row0 = t0 + (int64_t)lane * ITEMS;
        double c[ITEMS];
        if (row0 + ITEMS <= a.n_src) {
#pragma unroll
            for (int v = 0; v < ITEMS / 4; ++v) ld_f64x4(a.cumsum + row0 + 4 * v, c + 4 * v);
        } else {
            const double last = a.cumsum[a.n_src - 1];           // padding rows repeat the last value: no outputs
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) c[k] = (row0 + k < a.n_src) ? a.cumsum[row0 + k] : last;
        }
        if (a.normalise) {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) c[k] = __ddiv_rn(c[k], rc.Td);        // cumsum /= cumsum[-1]  (:90)
        }
        int e[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) e[k] = rank_of<POW2, TIES_RIGHT, true>(c[k], rc);
        if (row0 + ITEMS >= a.n_src) {
            // the last row takes whatever is left (TIES_RIGHT with u = 1.0: the reference kernel would index one past
            // the end, particle.py:259-263; a cumsum that does not end at 1.0)
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) if (row0 + k >= a.n_src - 1) e[k] = a.n_total_i;
        }
        int E0 = 0;
        if (t0 > 0) {
            double cp = a.cumsum[t0 - 1];
            if (a.normalise) cp = __ddiv_rn(cp, rc.Td);
            E0 = rank_of<POW2, TIES_RIGHT, true>(cp, rc);
        }
        int ep = __shfl_up_sync(0xffffffffu, e[ITEMS - 1], 1);
        if (lane == 0) ep = E0;
        const int E1 = __shfl_sync(0xffffffffu, e[ITEMS - 1], 31);
        warp_fill<ITEMS>(e, ep, E0, E1, a.src_row0 + (int)t0, a, s_mark[wid], s_end[wid], lane);
    }
}

// queued heavy runs of k_resample_search_f64 (its CTAs do not wait for one another): a second, tiny launch
__global__ void __launch_bounds__(RF_THREADS)
k_resample_drain(const __grid_constant__ FusedArgs a) {
    drain_queue(a, blockIdx.x, gridDim.x, threadIdx.x & 31, threadIdx.x >> 5);
    __syncthreads();
    __shared__ unsigned int s_done;
    if (threadIdx.x == 0) s_done = atomicAdd(a.counters + 3, 1u);
    __syncthreads();
    if (s_done == gridDim.x - 1u && threadIdx.x < 4) a.counters[threadIdx.x] = 0u;
}

static void fill_common(gse_ctx* ctx, FusedArgs& a, double r, int64_t n_total, int64_t out0, int64_t n_out,
                        int32_t* idx_out_dev, int64_t src_row0) {
    a.status = ctx->fused_status;
    a.counters = ctx->ticket + 8;
    a.queue = ctx->heavy_queue;
    a.queue_cap = ctx->heavy_queue_cap;
    a.err = ctx->err_dev;
    a.r = r;
    a.r_dev = ctx->step_params ? &ctx->step_params->r : NULL;
    a.n_total = (double)n_total;
    a.inv_n = 1.0 / (double)n_total;
    a.n_total_i = (int)n_total;
    a.out_lo = (int)out0;
    a.out_hi = (int)(out0 + n_out);
    a.idx_out = idx_out_dev;
    a.src_row0 = (int)src_row0;
}

#define GSE_FUSED_VARIANTS 12
template <int ITEMS, bool LL, bool BASE, bool POW2>
static int launch_fused(gse_ctx* ctx, FusedArgs& a, int variant, cudaStream_t s) {
    // every CTA must be resident at once (phase 2 and the tail wait on the other CTAs)
    if (ctx->fused_resident[variant] == 0) {
        int per_sm = 0;
        GSE_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_resample_fused<ITEMS, LL, BASE, POW2>,
                                                                     RF_THREADS, 0));
        GSE_REQUIRE(per_sm >= 1, "fused resample kernel does not fit an SM");
        ctx->fused_resident[variant] = per_sm * ctx->num_sms;
    }
    const int64_t group = (int64_t)RF_WARPS * 32 * ITEMS;           // rows of one tile per warp
    const int64_t groups = gse_div_up(a.n_src, group);
    int64_t blocks = groups < ctx->fused_resident[variant] ? groups : ctx->fused_resident[variant];
    a.rows_per_block = gse_div_up(groups, blocks) * group;
    blocks = gse_div_up(a.n_src, a.rows_per_block);
    GSE_REQUIRE(blocks <= ctx->max_tiles, "workspace too small");
    k_resample_fused<ITEMS, LL, BASE, POW2><<<(unsigned)blocks, RF_THREADS, 0, s>>>(a);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

extern "C" int gse_resample_fused(gse_ctx* ctx, const float* loglik_dev, const double* base_dev, const double* stats_dev,
                                  int64_t n_src, double r, int64_t n_total, int64_t out0, int64_t n_out,
                                  int64_t src_row0, int32_t* idx_out_dev, uint64_t* total_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && stats_dev != NULL && idx_out_dev != NULL, "ctx / stats / idx is NULL");
    GSE_REQUIRE(loglik_dev != NULL || base_dev != NULL, "need loglik or base weights");
    GSE_REQUIRE(loglik_dev == NULL || aligned32(loglik_dev), "loglik must be 32-byte aligned");
    GSE_REQUIRE(base_dev == NULL || aligned32(base_dev), "base must be 32-byte aligned");
    GSE_REQUIRE(n_src >= 1 && n_src <= ctx->n_max, "n_src out of range for this context");
    GSE_REQUIRE(n_total >= 1 && n_total <= 0x7ffffff0ll, "n_total out of range (int32 ancestor index)");
    GSE_REQUIRE(n_out >= 0 && n_out <= ctx->n_max && out0 >= 0 && out0 + n_out <= n_total, "output range out of range");
    GSE_REQUIRE(src_row0 >= 0 && src_row0 + n_src <= 0x7ffffff0ll, "source rows out of range (int32 ancestor index)");
    GSE_REQUIRE(r >= 0.0 && r < 1.0, "r must be in [0, 1)");
    if (n_out == 0) return GSE_OK;
    gse_device_guard guard(ctx->device);
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.loglik = loglik_dev;
    a.base = base_dev;
    a.stats = stats_dev;
    a.n_src = n_src;
    a.first_shard = (src_row0 == 0) ? 1 : 0;
    a.total_out = total_dev;
    fill_common(ctx, a, r, n_total, out0, n_out, idx_out_dev, src_row0);
    cudaStream_t s = (cudaStream_t)stream;
    const bool pow2 = (n_total & (n_total - 1)) == 0;
    const int items = ctx->fused_items;
#define FUSED_CASE(IT, LL, BASE, V)                                                       \
    do {                                                                                  \
        if (pow2) return launch_fused<IT, LL, BASE, true>(ctx, a, 2 * (V), s);            \
        return launch_fused<IT, LL, BASE, false>(ctx, a, 2 * (V) + 1, s);                 \
    } while (0)
    if (items == 16) {
        if (loglik_dev && base_dev) FUSED_CASE(16, true, true, 0);
        else if (loglik_dev) FUSED_CASE(16, true, false, 1);
        else FUSED_CASE(16, false, true, 2);
    } else {
        if (loglik_dev && base_dev) FUSED_CASE(8, true, true, 3);
        else if (loglik_dev) FUSED_CASE(8, true, false, 4);
        else FUSED_CASE(8, false, true, 5);
    }
#undef FUSED_CASE
    return GSE_OK;
}

extern "C" int gse_resample_search_f64(gse_ctx* ctx, const double* cumsum_dev, int64_t n_src, int normalise,
                                       int ties_right, double r, int64_t n_total, int64_t out0, int64_t n_out,
                                       int32_t* idx_out_dev, void* stream) {
    GSE_REQUIRE(ctx != NULL && cumsum_dev != NULL && idx_out_dev != NULL, "ctx / cumsum / idx is NULL");
    GSE_REQUIRE(aligned32(cumsum_dev), "cumsum must be 32-byte aligned");
    GSE_REQUIRE(n_src >= 1 && n_src <= 0x7ffffff0ll, "n_src out of range (int32 ancestor index)");
    GSE_REQUIRE(n_total >= 1 && n_total <= 0x7ffffff0ll, "n_total out of range");
    GSE_REQUIRE(n_out >= 0 && n_out <= ctx->n_max && out0 >= 0 && out0 + n_out <= n_total, "output range out of range");
    GSE_REQUIRE(r >= 0.0 && r < 1.0, "r must be in [0, 1)");
    if (n_out == 0) return GSE_OK;
    gse_device_guard guard(ctx->device);
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.cumsum = cumsum_dev;
    a.n_src = n_src;
    a.normalise = normalise ? 1 : 0;
    a.first_shard = 1;
    fill_common(ctx, a, r, n_total, out0, n_out, idx_out_dev, 0);
    cudaStream_t s = (cudaStream_t)stream;
    constexpr int ITEMS = 8;
    const int64_t ntiles = gse_div_up(n_src, 32 * ITEMS);
    int64_t blocks = gse_div_up(ntiles, RF_WARPS);
    const int64_t cap = (int64_t)ctx->num_sms * 8;
    if (blocks > cap) blocks = cap;
    const bool pow2 = (n_total & (n_total - 1)) == 0;
#define F64_CASE(P, T) k_resample_search_f64<ITEMS, P, T><<<(unsigned)blocks, RF_THREADS, 0, s>>>(a, ntiles)
    if (pow2) { if (ties_right) F64_CASE(true, true); else F64_CASE(true, false); }
    else { if (ties_right) F64_CASE(false, true); else F64_CASE(false, false); }
#undef F64_CASE
    GSE_CHECK_LAUNCH(ctx);
    k_resample_drain<<<(unsigned)ctx->num_sms, RF_THREADS, 0, s>>>(a);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}
