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
// conversions, floor and cvt.u64.f32 (7-14 lanes/clk/SM, profiles/r2_ubench_xu_pipe.txt) are replaced by adding
// 2^23 / 2^52-type constants and reading the mantissa.
//
// gse_resample_search_f64 feeds the same rank + fill code with a caller's own float64 cumulative sum
// (`resample_from_cumsum`, SURVEY.md section 7 contract (ii)).
#include "gse_resample_common.cuh"
#include "gse_mailbox.cuh"

#define RF_THREADS 256
#define RF_WARPS (RF_THREADS / 32)
#define RF_ITEMS 8               // consecutive source rows per lane and tile
#define RF_TILE (32 * RF_ITEMS)  // source rows per warp tile
#define RF_WIN 256               // outputs per window (one flush: 8 per lane)
#define RF_RING (2 * RF_WIN)     // per-warp marker ring: the window being completed and the one after it
#define RF_HEAVY_MIN 4096        // whole windows inside a longer run of one source are queued, not filled by the warp
#define RF_PIECE 65536           // queued runs are cut into pieces of at most this many outputs (one warp each)

#define RF_TWO52 4503599627370496.0
#define RF_NMEAN 6               // MEAN kernels: sum of c_k x_k (5 columns) and of c_k
#define RF_MAX_BLOCKS 4096       // status words: [0, 4096) CTA aggregates, [4096, 8192] exclusive prefixes + total

struct FusedArgs {
    const float* loglik;       // NULL: weights = base
    const double* base;        // NULL: weights = exp(loglik - M)
    const double* stats;       // [0] M, [1] S
    const double* cumsum;      // f64 entry: the caller's cumulative sum (normalised unless `normalise`)
    int64_t n_src;
    int64_t rows_per_block;    // multiple of RF_WARPS * RF_TILE
    uint64_t* status;          // one word per CTA, zero between launches
    unsigned int* counters;    // [0] start ticket, [1] CTAs past phase 3, [2] queue length, [3] CTAs past the drain, [4] aggregates published
    int4* queue;
    int queue_cap;
    unsigned int* err;         // device error word of the context (host-mapped)
    double r;
    double magic_minus_r;      // RF_MAGIC16 - r (see rank_of)
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
    double* stats_reset;       // weights are uniform after a resample: (M, S) <- (0, n_total) when the kernel is done (NULL to skip)
    // sharded population (SHARDED kernels): every rank ranks its OWN rows against the global total and writes the
    // ancestor of every output it sources straight into the index buffer of the shard that owns the output slot
    int nshards, rank;
    int out_home_lo, out_home_hi;              // this rank's own output slots (idx_out is its buffer)
    unsigned int epoch_totals, epoch_done;     // mailbox sequence numbers of the two exchanges inside the kernel
    int shard_lo[GSE_MAX_SHARDS + 1];          // shard t owns the output slots [shard_lo[t], shard_lo[t + 1])
    int32_t* shard_idx[GSE_MAX_SHARDS];        // shard t's index buffer (its local slot 0), peer memory for t != rank
    MailboxTable mb;
    // MEAN kernels: the post-resample estimate sum_k c_k x_k (c_k = offspring of row k) falls out of the ranks -- only the
    // few rows that have offspring are read
    const float* mean_state;   // SoA state of the source rows, 5 columns `mean_ld` apart
    int64_t mean_ld;
    double* mean_partials;     // 6 doubles per CTA
    double* mean_out;          // moment block: [0] S0 = outputs sourced here, [1..5] S1, [21..25] pivot (0)
    unsigned long long* trace; // GSE_FUSED_TRACE: 8 words per CTA (globaltimer at start / phase 1 / 2 / 3 done, SM id, ...), or NULL
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

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

// ring[slot] = val where `take` (a predicated store: the compiler otherwise branches around every one of them)
__device__ __forceinline__ void sts_if(unsigned int saddr, int val, bool take) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.shared.b32 [%0], %1;\n\t}"
                 :: "r"(saddr), "r"(val), "r"((int)take) : "memory");
}

// rint for 0 <= x <= 2^52 without the conversion pipe
__device__ __forceinline__ double rint_to_2p52(double x) { return __dadd_rn(__dadd_rn(x, RF_TWO52), -RF_TWO52); }
__device__ __forceinline__ double rint_nonneg(double x) { return x >= RF_TWO52 ? x : rint_to_2p52(x); }

// raw log-likelihoods of the lane's RF_ITEMS rows (rows >= n_end: -inf, weight 0)
__device__ __forceinline__ void load_loglik(const float* __restrict__ loglik, int64_t row0, int64_t n_end, float l[RF_ITEMS]) {
    if (row0 + RF_ITEMS <= n_end) {
        ld_f32x8(loglik + row0, l);
    } else {
#pragma unroll
        for (int k = 0; k < RF_ITEMS; ++k) l[k] = (row0 + k < n_end) ? loglik[row0 + k] : -INFINITY;
    }
}

// Fixed-point weights as float64 integers: the values of quantise16() in gse_resample.cu, rint(fl32(exp(l - M)) 2^s)
// (the float32 product is at most 2^52, so adding and subtracting 2^52 rounds it to an integer exactly as cvt.rni does)
__device__ __forceinline__ void quantise_loglik(const float l[RF_ITEMS], float nm, float scale_f, double q[RF_ITEMS]) {
#pragma unroll
    for (int k = 0; k < RF_ITEMS; ++k) q[k] = rint_to_2p52((double)__fmul_rn(weight_exp(l[k], nm), scale_f));
}

// ... and with float64 base weights: rint(base * fl32(exp(l - M)) * 2^s) in float64; rows >= n_end are 0
template <bool HAS_LL>
__device__ __forceinline__ void quantise_base(const float* __restrict__ loglik, const double* __restrict__ base, float nm,
                                              double scale_d, int64_t row0, int64_t n_end, double q[RF_ITEMS]) {
    const bool full = row0 + RF_ITEMS <= n_end;
    float l[RF_ITEMS];
    if (HAS_LL) load_loglik(loglik, row0, n_end, l);
    double b[RF_ITEMS];
    if (full) {
#pragma unroll
        for (int v = 0; v < RF_ITEMS / 4; ++v) ld_f64x4(base + row0 + 4 * v, b + 4 * v);
    } else {
#pragma unroll
        for (int k = 0; k < RF_ITEMS; ++k) b[k] = (row0 + k < n_end) ? base[row0 + k] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < RF_ITEMS; ++k) {
        double w = b[k];
        if (HAS_LL) w = __dmul_rn(w, (double)weight_exp(l[k], nm));
        q[k] = rint_nonneg(fmax(__dmul_rn(w, scale_d), 0.0));
    }
}

struct RankConsts {
    double inv_T, Td, n_total, inv_n, r, magic_minus_r;
    int n_total_i;
};

// 1.5 * 2^36: y = t + RF_MAGIC16 has ulp 2^-16, so the mantissa of y holds rint(t * 2^16) + 2^51: its low 16 bits are the
// fraction of t in units of 2^-16 and bits 16..47 are floor(t) (two's complement, modulo 2^32)
#define RF_MAGIC16 103079215104.0
__device__ __forceinline__ void rank_consts(RankConsts& k, double Td, const FusedArgs& a) {
    k.Td = Td;
    k.inv_T = 1.0 / Td;
    k.n_total = a.n_total;
    k.inv_n = a.inv_n;
    k.r = a.r_dev ? __ldg(a.r_dev) : a.r;
    k.magic_minus_r = a.r_dev ? RF_MAGIC16 - k.r : a.magic_minus_r;   // rounds r to a multiple of 2^-16: error <= 2^-17
    k.n_total_i = a.n_total_i;
}

// e = #{ outputs j in [0, N) : u_j <= g }  (TIES_RIGHT: u_j < g), g = fl(C / T) for the integer weights, g = C for a
// caller's normalised cumulative sum.  Fast path: with t* = g N - r in real arithmetic, u_j <= g  <=>  j <= t* up to
// the rounding of u_j and g, which moves the boundary by less than 6 N 2^-53 <= 0.19 * 2^-17 (N < 2^31).  One FMA
// gives y = fl(g N + (MAGIC - r)), i.e. t* in 2^-16 fixed point with an error below 2^-17 (rounding of r) + 2^-17
// (rounding of y) + 0.19 * 2^-17: when the 16-bit fraction is in [2, 65533] t* is further than that from an integer
// and the rank is floor(t) + 1 (in [0, N]: t* >= -r > -1 and t* <= N - r).  Otherwise (exact ties -- dyadic weights,
// r = 0 -- and 2^-14 of random sources) the comparison is evaluated exactly as the reference does.
// branch-free part: floor(t) + 1 and whether the fraction is too close to an integer to trust it
template <bool NORMALISED>
__device__ __forceinline__ int rank_fast(double C, const RankConsts& k, bool& near_integer) {
    const double g_fast = NORMALISED ? C : __dmul_rn(C, k.inv_T);
    const double y = __fma_rn(g_fast, k.n_total, k.magic_minus_r);
    const unsigned int lo = (unsigned int)__double2loint(y), hi = (unsigned int)__double2hiint(y);
    near_integer = ((lo & 0xffffu) - 2u) > 65531u;                // fraction in {0, 1, 65534, 65535}
    return (int)__funnelshift_r(lo, hi, 16) + 1;
}
template <bool POW2, bool TIES_RIGHT, bool NORMALISED>
__device__ __forceinline__ int rank_slow(double C, int guess, const RankConsts& k) {
    const double g = NORMALISED ? C : __ddiv_rn(C, k.Td);
    const int e = __double2int_rz(rank_exact_g<POW2, TIES_RIGHT>(k.r, k.n_total, k.inv_n, g, (double)guess, 0.0, k.n_total));
    return min(max(e, 0), k.n_total_i);
}
template <bool POW2, bool TIES_RIGHT, bool NORMALISED>
__device__ __forceinline__ int rank_of(double C, const RankConsts& k) {
    bool near_integer;
    int e = rank_fast<NORMALISED>(C, k, near_integer);
    if (near_integer) e = rank_slow<POW2, TIES_RIGHT, NORMALISED>(C, e, k);
    return e;
}

// ranks of the lane's RF_ITEMS cumulative values c[k] (NORMALISED) or C0 + q[0] + ... + q[k] (integer weights): the
// eight fast paths are independent instruction streams the scheduler can interleave; the exact comparisons (rare) sit
// behind ONE branch
template <bool POW2, bool TIES_RIGHT, bool NORMALISED>
__device__ __forceinline__ void ranks_of(double C0, const double (&q)[RF_ITEMS], const RankConsts& rc, int (&e)[RF_ITEMS]) {
    unsigned int flags = 0;
    double C = C0;
#pragma unroll
    for (int k = 0; k < RF_ITEMS; ++k) {
        C = NORMALISED ? q[k] : C + q[k];
        bool near_integer;
        e[k] = rank_fast<NORMALISED>(C, rc, near_integer);
        flags |= near_integer ? (1u << k) : 0u;
    }
    if (flags) {
        C = C0;
#pragma unroll
        for (int k = 0; k < RF_ITEMS; ++k) {
            C = NORMALISED ? q[k] : C + q[k];
            if (flags & (1u << k)) e[k] = rank_slow<POW2, TIES_RIGHT, NORMALISED>(C, e[k], rc);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Streaming fill of one warp's contiguous sub-run of sources.  Source s of the sub-run owns the outputs
// [start_s, e_s), start_s = e_{s-1}.  The warp keeps a ring of two 256-output windows in shared memory, aligned to 256
// outputs globally: every source WITH offspring drops the marker (its 1-based index in the sub-run) at its first output;
// once the ranks of a tile reach past the end of the oldest window, that window is complete: a prefix maximum spreads
// the markers over the runs (carry = the source covering the window's first output) and the window leaves as two
// coalesced 128-bit stores per lane.  Only the first and the last window of a sub-run are partial (masked scalar stores).
// ------------------------------------------------------------------------------------------------
// shard table of the SHARDED kernels in shared memory (indexing the kernel parameter with a run-time index would push
// the whole argument struct into local memory)
struct ShardTable {
    int nshards;
    int home_lo, home_hi;          // the calling rank's own slots: most windows land there
    int32_t* home_idx;
    int lo[GSE_MAX_SHARDS + 1];
    int32_t* idx[GSE_MAX_SHARDS];
};
__device__ __forceinline__ int shard_of_output(const ShardTable& st, int j) {
    int t = 0;
#pragma unroll
    for (int u = 1; u < GSE_MAX_SHARDS; ++u) t += (u < st.nshards && j >= st.lo[u]) ? 1 : 0;
    return t;
}

struct WarpFill {
    int wb;            // first output of the oldest window not yet flushed (multiple of RF_WIN)
    int carry_m;       // marker of the source that covers output wb (0: none yet -- only below the sub-run's first output)
    int mlo;           // outputs below mlo are not this sub-run's (or not this launch's)
    int hi_al;         // no window at or beyond this output is needed (out_hi rounded up to a window)
    int vbase;         // stored ancestor = vbase + marker
    int* ring;
    int32_t* out;      // idx_out - out_lo
    bool vec_ok;       // out + (multiple of 4) is 16-byte aligned
    const ShardTable* st;   // SHARDED kernels: where each output slot lives

    __device__ __forceinline__ void begin(int E0, const FusedArgs& a, int vbase_, int* ring_, int lane,
                                          const ShardTable* st_ = NULL) {
        st = st_;
        ring = ring_;
        vbase = vbase_;
        out = a.idx_out - a.out_lo;
        vec_ok = (((uintptr_t)out) & 15u) == 0;
        mlo = max(E0, a.out_lo);
        wb = mlo & ~(RF_WIN - 1);
        hi_al = (a.out_hi + RF_WIN - 1) & ~(RF_WIN - 1);
        carry_m = 0;
#pragma unroll
        for (int v = 0; v < RF_RING / 128; ++v) reinterpret_cast<int4*>(ring)[32 * v + lane] = make_int4(0, 0, 0, 0);
    }

    // complete window [wb, wb + RF_WIN): prefix maximum of its markers, store, clear, advance.  Lane l owns the outputs
    // wb + 4 l + {0..3} and wb + 128 + 4 l + {0..3}: both 128-bit stores of the warp are fully coalesced.
    template <bool SHARDED>
    __device__ __forceinline__ void flush(int lane, int mhi, const FusedArgs& fa) {
        int4* r4 = reinterpret_cast<int4*>(ring + (wb & (RF_RING - 1)));
        int4 a = r4[lane], b = r4[32 + lane];
        r4[lane] = make_int4(0, 0, 0, 0);
        r4[32 + lane] = make_int4(0, 0, 0, 0);
        a.y = max(a.x, a.y); a.z = max(a.y, a.z); a.w = max(a.z, a.w);
        b.y = max(b.x, b.y); b.z = max(b.y, b.z); b.w = max(b.z, b.w);
        int ia = a.w, ib = b.w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia = max(ia, ta); ib = max(ib, tb); }
        }
        const int tot_a = max(__shfl_sync(0xffffffffu, ia, 31), carry_m);
        int ea = __shfl_up_sync(0xffffffffu, ia, 1), eb = __shfl_up_sync(0xffffffffu, ib, 1);
        if (lane == 0) { ea = 0; eb = 0; }
        ea = max(ea, carry_m);
        eb = max(eb, tot_a);
        const int va = vbase;
        a = make_int4(max(a.x, ea) + va, max(a.y, ea) + va, max(a.z, ea) + va, max(a.w, ea) + va);
        b = make_int4(max(b.x, eb) + va, max(b.y, eb) + va, max(b.z, eb) + va, max(b.w, eb) + va);
        carry_m = max(tot_a, __shfl_sync(0xffffffffu, ib, 31));
        if (SHARDED) {
            // the window usually lies inside one shard's slots: two coalesced 128-bit stores into that shard's buffer
            // (over NVLink when it is a peer's); windows across a shard boundary or the sub-run's ends go element-wise
            // (the home shard's bounds and buffer are kernel parameters: constant-bank operands, no shared-memory reads)
            const bool whole = wb >= mlo && wb + RF_WIN <= mhi;
            if (whole && wb >= fa.out_home_lo && wb + RF_WIN <= fa.out_home_hi) {
                int32_t* o = fa.idx_out + (wb - fa.out_home_lo) + 4 * lane;
                *reinterpret_cast<int4*>(o) = a;
                *reinterpret_cast<int4*>(o + 128) = b;
                wb += RF_WIN;
                return;
            }
            const int t0 = shard_of_output(*st, wb), t1 = shard_of_output(*st, wb + RF_WIN - 1);
            if (t0 == t1 && whole) {
                int32_t* o = st->idx[t0] + (wb - st->lo[t0]) + 4 * lane;
                *reinterpret_cast<int4*>(o) = a;
                *reinterpret_cast<int4*>(o + 128) = b;
            } else {
                const int vals[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int p = wb + 4 * lane + (i & 3) + (i >> 2) * 128;
                    if (p >= mlo && p < mhi) {
                        const int t = shard_of_output(*st, p);
                        st->idx[t][p - st->lo[t]] = vals[i];
                    }
                }
            }
            wb += RF_WIN;
            return;
        }
        int32_t* o = out + wb + 4 * lane;
        if (vec_ok && wb >= mlo && wb + RF_WIN <= mhi) {
            *reinterpret_cast<int4*>(o) = a;
            *reinterpret_cast<int4*>(o + 128) = b;
        } else {
            const int p = wb + 4 * lane;
            if (p + 0 >= mlo && p + 0 < mhi) o[0] = a.x;
            if (p + 1 >= mlo && p + 1 < mhi) o[1] = a.y;
            if (p + 2 >= mlo && p + 2 < mhi) o[2] = a.z;
            if (p + 3 >= mlo && p + 3 < mhi) o[3] = a.w;
            if (p + 128 >= mlo && p + 128 < mhi) o[128] = b.x;
            if (p + 129 >= mlo && p + 129 < mhi) o[129] = b.y;
            if (p + 130 >= mlo && p + 130 < mhi) o[130] = b.z;
            if (p + 131 >= mlo && p + 131 < mhi) o[131] = b.w;
        }
        wb += RF_WIN;
    }

    // one tile: lane's sources have ranks e[0..7], the source before the lane's first has rank ep; the tile's outputs are
    // [E0, E1); mbase = marker of the tile's first source
    template <bool SHARDED>
    __device__ __forceinline__ void tile(const int (&e)[RF_ITEMS], int ep, int E0, int E1, int mbase, const FusedArgs& a,
                                         int lane) {
        const int E1c = min(E1, hi_al);
        if (E1c <= max(E0, wb)) return;                           // no outputs of this launch in the tile (warp-uniform)
        if (E0 < wb && carry_m == 0) {
            // the launch's outputs start inside this tile's range: the source covering output wb started below it
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < RF_ITEMS; ++k) cnt += (e[k] <= wb) ? 1 : 0;
            carry_m = mbase + warp_sum_i32(cnt);
        }
        const bool maybe_heavy = (E1 - E0) > RF_HEAVY_MIN;
        for (;;) {
            __syncwarp();
            const int limit = wb + RF_RING;
            const unsigned int ring_s = (unsigned int)__cvta_generic_to_shared(ring);
            const int m0 = mbase + lane * RF_ITEMS;
#pragma unroll
            for (int k = 0; k < RF_ITEMS; ++k) {
                const int start = k ? e[k - 1] : ep;
                const bool take = e[k] > start && (unsigned int)(start - wb) < (unsigned int)RF_RING;   // start < wb wraps around
                sts_if(ring_s + 4u * (unsigned int)(start & (RF_RING - 1)), m0 + k, take);
            }
            __syncwarp();
            bool again = E1c > limit;
            while (wb + RF_WIN <= min(E1c, limit)) {
                flush<SHARDED>(lane, a.out_hi, a);
                if (maybe_heavy && carry_m >= mbase) {
                    // the source covering the new window is one of this tile's: does its run go on for long?
                    // (one shuffle per k, then a uniform select: indexing e[] with li would push the array into local memory)
                    const int li = carry_m - mbase;
                    int end_c = 0;
#pragma unroll
                    for (int k = 0; k < RF_ITEMS; ++k) {
                        const int v = __shfl_sync(0xffffffffu, e[k], li >> 3);
                        if ((li & (RF_ITEMS - 1)) == k) end_c = v;
                    }
                    const int skip_to = min(end_c, hi_al) & ~(RF_WIN - 1);
                    if (skip_to - wb > RF_HEAVY_MIN) {            // hand [wb, skip_to) to the whole grid
                        if (lane == 0) {
                            const int lo_q = max(wb, a.out_lo), hi_q = min(skip_to, a.out_hi);
                            const int pieces = hi_q > lo_q ? (hi_q - lo_q + RF_PIECE - 1) / RF_PIECE : 0;
                            const int slot = pieces ? (int)atomicAdd(a.counters + 2, (unsigned int)pieces) : 0;
                            for (int p = 0; p < pieces; ++p) {
                                if (slot + p < a.queue_cap)
                                    a.queue[slot + p] = make_int4(lo_q + p * RF_PIECE, min(lo_q + (p + 1) * RF_PIECE, hi_q),
                                                                  vbase + carry_m, 0);
                                else
                                    atomicOr(a.err, GSE_ERR_QUEUE_OVERFLOW);
                            }
                        }
                        wb = skip_to;                             // the ring is all zero here: no source starts inside a run
                        again = E1c > wb;
                        break;
                    }
                }
            }
            if (!again) break;
        }
    }

    // end of the sub-run: its last outputs sit in a partial window
    template <bool SHARDED>
    __device__ __forceinline__ void finish(int E1, const FusedArgs& a, int lane) {
        const int mhi = min(E1, a.out_hi);
        __syncwarp();
        if (wb < mhi) flush<SHARDED>(lane, mhi, a);
    }
};

// [p, end) of one destination buffer <- val: head up to 16-byte alignment, 128-bit body, tail
__device__ __forceinline__ void warp_fill_const(int32_t* out, int p, int end, int val, int lane) {
    const int head = min(end, p + (int)((4 - (((uintptr_t)(out + p) >> 2) & 3)) & 3));
    if (p + lane < head) out[p + lane] = val;
    p = head;
    const int4 v4 = make_int4(val, val, val, val);
    for (int q = p + 4 * lane; q + 4 <= end; q += 128) *reinterpret_cast<int4*>(out + q) = v4;
    const int body_end = p + ((end - p) & ~3);
    if (body_end + lane < end) out[body_end + lane] = val;
}

// drain of the heavy-run queue by every warp of the grid (after all CTAs have finished phase 3)
template <bool SHARDED>
__device__ __forceinline__ void drain_queue(const FusedArgs& a, const ShardTable* st, int vb, int nblocks, int lane, int wid) {
    const int qn = min((int)ld_status32(a.counters + 2), a.queue_cap);
    for (int i = vb * RF_WARPS + wid; i < qn; i += nblocks * RF_WARPS) {
        const int4 ent = __ldcg(a.queue + i);
        if (!SHARDED) {
            warp_fill_const(a.idx_out - a.out_lo, ent.x, ent.y, ent.z, lane);
        } else {
            for (int t = shard_of_output(*st, ent.x); t < st->nshards && st->lo[t] < ent.y; ++t) {
                const int lo = max(ent.x, st->lo[t]), hi = min(ent.y, st->lo[t + 1]);
                if (hi > lo) warp_fill_const(st->idx[t] - st->lo[t], lo, hi, ent.z, lane);
            }
        }
    }
}

template <bool HAS_LL, bool HAS_BASE, bool POW2, int MINB, bool SHARDED, bool MEAN>
__global__ void __launch_bounds__(RF_THREADS, MINB)
k_resample_fused(const __grid_constant__ FusedArgs a) {
    __shared__ __align__(16) int s_ring[RF_WARPS][RF_RING];
    __shared__ int2 s_list[MEAN ? RF_WARPS : 1][MEAN ? RF_TILE : 1];     // (row, offspring) of the tile's rows that have any
    __shared__ double s_mean[MEAN ? RF_WARPS : 1][RF_NMEAN];
    __shared__ double s_wsum[RF_WARPS];
    __shared__ uint64_t s_part[RF_WARPS][2];
    __shared__ unsigned int s_vb;
    __shared__ ShardTable s_st;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nblocks = gridDim.x;
    if (tid == 0) s_vb = atomicAdd(a.counters, 1u);               // CTAs are numbered in the order they start
    if (SHARDED && tid == 32) {
        s_st.nshards = a.nshards;
        s_st.home_lo = a.out_home_lo;
        s_st.home_hi = a.out_home_hi;
        s_st.home_idx = a.idx_out;
#pragma unroll
        for (int t = 0; t <= GSE_MAX_SHARDS; ++t) s_st.lo[t] = a.shard_lo[t];
#pragma unroll
        for (int t = 0; t < GSE_MAX_SHARDS; ++t) s_st.idx[t] = a.shard_idx[t];
    }
    __syncthreads();
    const int vb = (int)s_vb;
    if (a.trace && tid == 0) { a.trace[8 * vb] = global_timer_ns(); unsigned int smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); a.trace[8 * vb + 4] = smid; }
    const float nm = HAS_LL ? weight_exp_offset((float)a.stats[0]) : 0.0f;       // -M log2(e), see weight_exp()
    const int sexp = quantisation_exponent(a.stats[1]);
    const float scale_f = __int_as_float((127 + sexp) << 23);     // 2^sexp, 0 <= sexp <= 52
    const double scale_d = ldexp(1.0, sexp);
    // this CTA owns rows [b0, b1), warp w the contiguous sub-run [w0, w1) of it (whole tiles)
    const int64_t b0 = (int64_t)vb * a.rows_per_block, b1 = min(b0 + a.rows_per_block, a.n_src);
    const int64_t rows_per_warp = a.rows_per_block / RF_WARPS;
    const int64_t w0 = min(b0 + wid * rows_per_warp, b1), w1 = min(w0 + rows_per_warp, b1);

    // ---- phase 1: sum of the warp's sub-run (four tiles of loads in flight per lane) --------------------------
    double sum = 0.0;
    if (HAS_LL && !HAS_BASE) {
        for (int64_t t0 = w0; t0 < w1; t0 += 4 * RF_TILE) {
            float l[4][RF_ITEMS];
#pragma unroll
            for (int u = 0; u < 4; ++u) load_loglik(a.loglik, t0 + u * RF_TILE + (int64_t)lane * RF_ITEMS, w1, l[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                double q[RF_ITEMS];
                quantise_loglik(l[u], nm, scale_f, q);
                sum += ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
            }
        }
    } else {
        for (int64_t t0 = w0; t0 < w1; t0 += RF_TILE) {
            double q[RF_ITEMS];
            quantise_base<HAS_LL>(a.loglik, a.base, nm, scale_d, t0 + (int64_t)lane * RF_ITEMS, w1, q);
            sum += ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
        }
    }
    sum = warp_sum_f64(sum);
    if (lane == 0) s_wsum[wid] = sum;
    __syncthreads();
    if (a.trace && tid == 0) a.trace[8 * vb + 1] = global_timer_ns();

    // ---- phase 2: publish the CTA aggregate; the LAST CTA to arrive scans all of them and hands every CTA its exclusive
    //      prefix and the total.  One thread per CTA polls one word: polling by everyone for everything (nblocks^2 loads per
    //      round) floods the L2 and slows the CTAs that are still streaming their rows.
    uint64_t* const prefix = a.status + RF_MAX_BLOCKS;            // [i] exclusive prefix of CTA i, [nblocks] the total
    if (tid == 0) {
        double agg = 0.0;
#pragma unroll
        for (int w = 0; w < RF_WARPS; ++w) agg += s_wsum[w];
        st_status(a.status + vb, status_pack(ST_AGGREGATE, 0u, (uint64_t)__double2ll_rn(agg)));
        __threadfence();
        s_vb = atomicAdd(a.counters + 4, 1u);                     // arrival ticket
    }
    __syncthreads();
    if (s_vb == (unsigned int)nblocks - 1u) {                     // every aggregate is visible now
        __threadfence();
        // block-wide exclusive scan of nblocks values: thread t owns a contiguous chunk
        const int chunk = (nblocks + RF_THREADS - 1) / RF_THREADS;
        const int i0 = tid * chunk;
        uint64_t part = 0;
        for (int k = 0; k < chunk; ++k)
            if (i0 + k < nblocks) part += ld_status(a.status + i0 + k) & ((1ull << 54) - 1ull);
        const uint64_t incl_w = warp_inclusive_scan_u64(part, lane);
        if (lane == 31) s_part[wid][0] = incl_w;
        __syncthreads();
        uint64_t base = 0, local_total = 0;
#pragma unroll
        for (int w = 0; w < RF_WARPS; ++w) { base += (w < wid) ? s_part[w][0] : 0ull; local_total += s_part[w][0]; }
        uint64_t shard_off = 0, total = local_total;
        if (SHARDED) {
            // the shards' totals cross NVLink here: exclusive offset of this shard and the global total
            if (wid == 0) {
                unsigned long long r0;
                mbox_exchange_tagged(a.mb, a.rank, a.nshards, a.epoch_totals, local_total, lane, r0, a.err);
                const uint64_t incl = warp_inclusive_scan_u64(r0, lane);
                const uint64_t mine = __shfl_sync(0xffffffffu, incl - r0, a.rank);
                const uint64_t all = __shfl_sync(0xffffffffu, incl, a.nshards - 1);
                if (lane == 0) { s_part[0][1] = mine; s_part[1][1] = all; }
            }
            __syncthreads();
            shard_off = s_part[0][1];
            total = s_part[1][1];
        }
        uint64_t run = shard_off + base + incl_w - part;
        for (int k = 0; k < chunk; ++k) {
            if (i0 + k < nblocks) {
                st_status(prefix + i0 + k, status_pack(ST_PREFIX, 0u, run));
                run += ld_status(a.status + i0 + k) & ((1ull << 54) - 1ull);
            }
        }
        if (tid == 0) st_status(prefix + nblocks, status_pack(ST_PREFIX, 0u, total));
    }
    if (tid == 0) {
        uint64_t w0_ = ld_status(prefix + vb);
        while ((w0_ >> 62) == 0ull) { __nanosleep(200); w0_ = ld_status(prefix + vb); }
        uint64_t w1_ = ld_status(prefix + nblocks);
        while ((w1_ >> 62) == 0ull) { __nanosleep(100); w1_ = ld_status(prefix + nblocks); }
        s_part[0][0] = w0_ & ((1ull << 54) - 1ull);
        s_part[0][1] = w1_ & ((1ull << 54) - 1ull);
    }
    __syncthreads();
    const uint64_t excl_u = s_part[0][0], tot_u = s_part[0][1];
    if (tid == 0 && vb == 0 && a.total_out) *a.total_out = tot_u;          // (sharded: the GLOBAL total)
    if (a.trace && tid == 0) a.trace[8 * vb + 2] = global_timer_ns();

    // ---- phase 3: scan + rank + fill, warp-autonomous ---------------------------------------------------------
    RankConsts rc;
    rank_consts(rc, __ull2double_rn(tot_u), a);
    const bool degenerate = (tot_u == 0ull);                      // all weights zero: everything descends from the last row
    double carry = __ull2double_rn(excl_u);
#pragma unroll
    for (int w = 0; w < RF_WARPS; ++w) carry += (w < wid) ? s_wsum[w] : 0.0;
    double macc[RF_NMEAN];
#pragma unroll
    for (int c = 0; c < RF_NMEAN; ++c) macc[c] = 0.0;
    if (w0 < w1) {
        int carry_rank = 0;
        if (!(a.first_shard && w0 == 0) && !degenerate) carry_rank = rank_of<POW2, false, false>(carry, rc);
        if (degenerate && lane == 0 && vb == 0 && wid == 0) atomicOr(a.err, GSE_ERR_ZERO_WEIGHTS);
        WarpFill wf;
        wf.begin(carry_rank, a, a.src_row0 + (int)w0 - 1, s_ring[wid], lane, &s_st);
        float lcur[RF_ITEMS];
        if (HAS_LL && !HAS_BASE) load_loglik(a.loglik, w0 + (int64_t)lane * RF_ITEMS, w1, lcur);
        const int ntiles = (int)((w1 - w0 + RF_TILE - 1) / RF_TILE);
        for (int t = 0; t < ntiles; ++t) {
            const int64_t row0 = w0 + (int64_t)t * RF_TILE + (int64_t)lane * RF_ITEMS;
            double q[RF_ITEMS];
            if (HAS_LL && !HAS_BASE) quantise_loglik(lcur, nm, scale_f, q);
            else quantise_base<HAS_LL>(a.loglik, a.base, nm, scale_d, row0, w1, q);
            const double lane_sum = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
            double incl = lane_sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            double C = carry + (incl - lane_sum);                // cumulative weight before the lane's first row
            carry += __shfl_sync(0xffffffffu, incl, 31);
            int e[RF_ITEMS];
            if (!degenerate) {
                ranks_of<POW2, false, false>(C, q, rc, e);
            } else {
#pragma unroll
                for (int k = 0; k < RF_ITEMS; ++k)          // only the last row of the whole population has offspring
                    e[k] = (row0 + k >= a.n_src - 1 && a.src_row0 + a.n_src >= (int64_t)a.n_total_i) ? a.n_total_i : 0;
            }
            // next tile's rows: in flight while this tile's outputs are filled
            if (HAS_LL && !HAS_BASE && t + 1 < ntiles) load_loglik(a.loglik, row0 + RF_TILE, w1, lcur);
            int ep = __shfl_up_sync(0xffffffffu, e[RF_ITEMS - 1], 1);
            if (lane == 0) ep = carry_rank;
            const int E1 = __shfl_sync(0xffffffffu, e[RF_ITEMS - 1], 31);
            if (MEAN) {
                // list the rows with offspring (about one in twenty once the weights have spread), then the whole warp reads
                // them at once: a lane that read its own rows one after the other would wait out a memory round trip for each
                int2* const list = s_list[wid];
                const unsigned int below = (1u << lane) - 1u;
                int prev = ep, cnt = 0;
#pragma unroll
                for (int k = 0; k < RF_ITEMS; ++k) {
                    const int c = e[k] - prev;
                    prev = e[k];
                    const unsigned int m = __ballot_sync(0xffffffffu, c > 0);
                    if (c > 0) list[cnt + __popc(m & below)] = make_int2((int)row0 + k, c);
                    cnt += __popc(m);
                }
                __syncwarp();
                for (int i = lane; i < cnt; i += 32) {
                    const int2 en = list[i];
                    const float* const xr = a.mean_state + en.x;
                    const double cd = (double)en.y;
                    float xv[5];
#pragma unroll
                    for (int c = 0; c < 5; ++c) xv[c] = __ldg(xr + c * a.mean_ld);
#pragma unroll
                    for (int c = 0; c < 5; ++c) macc[c] = fma(cd, (double)xv[c], macc[c]);     // exact products
                    macc[5] += cd;
                }
                __syncwarp();
            }
            wf.template tile<SHARDED>(e, ep, carry_rank, E1, t * RF_TILE + 1, a, lane);
            carry_rank = E1;
        }
        wf.template finish<SHARDED>(carry_rank, a, lane);
    }
    if (MEAN) {
#pragma unroll
        for (int c = 0; c < RF_NMEAN; ++c) {
            const double v = warp_sum_f64(macc[c]);
            if (lane == 0) s_mean[wid][c] = v;
        }
    }

    // ---- tail: wait for every CTA, drain the heavy-run queue, reset the launch state --------------------------
    __syncthreads();
    if (a.trace && tid == 0) a.trace[8 * vb + 3] = global_timer_ns();
    if (MEAN && tid >= 32 && tid < 32 + RF_NMEAN) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < RF_WARPS; ++w) t += s_mean[w][tid - 32];
        a.mean_partials[(size_t)vb * RF_NMEAN + (tid - 32)] = t;
        __threadfence();                                          // read by the last CTA out, two barriers from here
    }
    if (tid == 0) {
        __threadfence();
        atomicAdd(a.counters + 1, 1u);
        while (ld_status32(a.counters + 1) < (unsigned int)nblocks) __nanosleep(100);
        __threadfence();
    }
    __syncthreads();
    drain_queue<SHARDED>(a, &s_st, vb, nblocks, lane, wid);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_vb = atomicAdd(a.counters + 3, 1u);
    }
    __syncthreads();
    if (s_vb == (unsigned int)nblocks - 1u) {                     // last CTA out
        if (SHARDED) {
            // every ancestor index this rank sources has been written (into local and peer buffers).  Tell the peers
            // and wait until they have said the same: when this kernel ends, this shard's index buffer is complete.
            if (wid == 0) {
                __threadfence_system();
                mbox_signal_wait(a.mb, a.rank, a.nshards, a.epoch_done, lane, a.err);
            }
            __syncthreads();
        }
        if (MEAN) {
            __threadfence();
            if (wid < RF_NMEAN) {                                 // fixed summation order: the estimate is reproducible
                double t = 0.0;
                for (int b = lane; b < nblocks; b += 32) t += __ldcg(a.mean_partials + (size_t)b * RF_NMEAN + wid);
                t = warp_sum_f64(t);
                if (lane == 0) a.mean_out[wid == 5 ? 0 : 1 + wid] = t;
                if (lane == 1 && wid < 5) a.mean_out[21 + wid] = 0.0;
            }
            if (tid == 0) { a.mean_out[41] = 0.0; a.mean_out[42] = a.n_total; }      // (M, S) of the uniform weights
        }
        if (a.stats_reset && tid == 0) { a.stats_reset[0] = 0.0; a.stats_reset[1] = a.n_total; }
        // leave everything zero for the next launch
        for (int i = tid; i < nblocks; i += RF_THREADS) { a.status[i] = 0ull; a.status[RF_MAX_BLOCKS + i] = 0ull; }
        if (tid == 0) a.status[RF_MAX_BLOCKS + nblocks] = 0ull;
        if (tid < 5) a.counters[tid] = 0u;
    }
}

// ------------------------------------------------------------------------------------------------
// rank + fill on a caller's float64 cumulative sum: no scan, no inter-CTA dependency (every warp reads the element
// before its chunk for e_{k-1}).  TIES_RIGHT = the reference's GPU kernel (`cumsum[k] > u` walk, particle.py:223-263,
// searchsorted side='right'); otherwise the CPU loop (`cumsum[k] < u`, :96-100, side='left').
// ------------------------------------------------------------------------------------------------
template <bool POW2, bool TIES_RIGHT>
__global__ void __launch_bounds__(RF_THREADS, 4)
k_resample_search_f64(const __grid_constant__ FusedArgs a, int64_t tiles_per_warp) {
    __shared__ __align__(16) int s_ring[RF_WARPS][RF_RING];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    RankConsts rc;
    rank_consts(rc, a.normalise ? a.cumsum[a.n_src - 1] : 1.0, a);
    const int64_t w0 = ((int64_t)blockIdx.x * RF_WARPS + wid) * tiles_per_warp * RF_TILE;
    const int64_t w1 = min(w0 + tiles_per_warp * RF_TILE, a.n_src);
    if (w0 >= w1) return;
    int carry_rank = 0;
    if (w0 > 0) {
        double cp = a.cumsum[w0 - 1];
        if (a.normalise) cp = __ddiv_rn(cp, rc.Td);
        carry_rank = rank_of<POW2, TIES_RIGHT, true>(cp, rc);
    }
    WarpFill wf;
    wf.begin(carry_rank, a, a.src_row0 + (int)w0 - 1, s_ring[wid], lane);
    for (int64_t t0 = w0; t0 < w1; t0 += RF_TILE) {
        const int64_t row0 = t0 + (int64_t)lane * RF_ITEMS;
        double c[RF_ITEMS];
        if (row0 + RF_ITEMS <= a.n_src) {
#pragma unroll
            for (int v = 0; v < RF_ITEMS / 4; ++v) ld_f64x4(a.cumsum + row0 + 4 * v, c + 4 * v);
        } else {
            const double last = a.cumsum[a.n_src - 1];           // padding rows repeat the last value: no outputs
#pragma unroll
            for (int k = 0; k < RF_ITEMS; ++k) c[k] = (row0 + k < a.n_src) ? a.cumsum[row0 + k] : last;
        }
        if (a.normalise) {
#pragma unroll
            for (int k = 0; k < RF_ITEMS; ++k) c[k] = __ddiv_rn(c[k], rc.Td);        // cumsum /= cumsum[-1]  (:90)
        }
        int e[RF_ITEMS];
        ranks_of<POW2, TIES_RIGHT, true>(0.0, c, rc, e);
        if (row0 + RF_ITEMS >= a.n_src) {
            // the last row takes whatever is left (TIES_RIGHT with u = 1.0: the reference kernel would index one past
            // the end, particle.py:259-263; a cumsum that does not end at 1.0)
#pragma unroll
            for (int k = 0; k < RF_ITEMS; ++k) if (row0 + k >= a.n_src - 1) e[k] = a.n_total_i;
        }
        int ep = __shfl_up_sync(0xffffffffu, e[RF_ITEMS - 1], 1);
        if (lane == 0) ep = carry_rank;
        const int E1 = __shfl_sync(0xffffffffu, e[RF_ITEMS - 1], 31);
        wf.template tile<false>(e, ep, carry_rank, E1, (int)(t0 - w0) + 1, a, lane);
        carry_rank = E1;
    }
    wf.template finish<false>(carry_rank, a, lane);
}

// queued heavy runs of k_resample_search_f64 (its CTAs do not wait for one another): a second, tiny launch
__global__ void __launch_bounds__(RF_THREADS)
k_resample_drain(const __grid_constant__ FusedArgs a) {
    drain_queue<false>(a, NULL, blockIdx.x, gridDim.x, threadIdx.x & 31, threadIdx.x >> 5);
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
    a.magic_minus_r = RF_MAGIC16 - r;
    a.r_dev = ctx->step_params ? &ctx->step_params->r : NULL;
    a.n_total = (double)n_total;
    a.inv_n = 1.0 / (double)n_total;
    a.n_total_i = (int)n_total;
    a.out_lo = (int)out0;
    a.out_hi = (int)(out0 + n_out);
    a.idx_out = idx_out_dev;
    a.src_row0 = (int)src_row0;
}

template <bool LL, bool BASE, bool POW2, int MINB, bool SHARDED, bool MEAN = false>
static int launch_fused(gse_ctx* ctx, FusedArgs& a, int variant, cudaStream_t s) {
    // every CTA must be resident at once (phase 2 and the tail wait on the other CTAs)
    if (ctx->fused_resident[variant] == 0) {
        int per_sm = 0;
        GSE_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_resample_fused<LL, BASE, POW2, MINB, SHARDED, MEAN>,
                                                                     RF_THREADS, 0));
        GSE_REQUIRE(per_sm >= 1, "fused resample kernel does not fit an SM");
        ctx->fused_resident[variant] = per_sm * ctx->num_sms;
    }
    const int64_t group = (int64_t)RF_WARPS * RF_TILE;               // rows of one tile per warp
    const int64_t groups = gse_div_up(a.n_src, group);
    int64_t blocks = groups < ctx->fused_resident[variant] ? groups : ctx->fused_resident[variant];
    a.rows_per_block = gse_div_up(groups, blocks) * group;
    blocks = gse_div_up(a.n_src, a.rows_per_block);
    GSE_REQUIRE(blocks <= ctx->max_tiles && blocks < RF_MAX_BLOCKS, "workspace too small");
    k_resample_fused<LL, BASE, POW2, MINB, SHARDED, MEAN><<<(unsigned)blocks, RF_THREADS, 0, s>>>(a);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}

extern "C" int gse_resample_fused(gse_ctx* ctx, const float* loglik_dev, const double* base_dev, const double* stats_dev,
                                  int64_t n_src, double r, int64_t n_total, int64_t out0, int64_t n_out,
                                  int64_t src_row0, int32_t* idx_out_dev, uint64_t* total_dev, const float* state_dev,
                                  int64_t ld, double* moments_dev, int reset_stats, void* stream) {
    GSE_REQUIRE(ctx != NULL && stats_dev != NULL && idx_out_dev != NULL, "ctx / stats / idx is NULL");
    GSE_REQUIRE(moments_dev == NULL || (state_dev != NULL && ld >= n_src && loglik_dev != NULL && base_dev == NULL &&
                                        out0 == 0 && n_out == n_total),
                "the in-kernel estimate needs the state, log-likelihood weights and the whole output range");
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
    a.stats_reset = reset_stats ? const_cast<double*>(stats_dev) : NULL;
    a.trace = ctx->fused_trace;
    fill_common(ctx, a, r, n_total, out0, n_out, idx_out_dev, src_row0);
    cudaStream_t s = (cudaStream_t)stream;
    const bool pow2 = (n_total & (n_total - 1)) == 0;
    if (moments_dev) {
        a.mean_state = state_dev;
        a.mean_ld = ld;
        a.mean_partials = ctx->red_partials;
        a.mean_out = moments_dev;
        if (pow2) return launch_fused<true, false, true, 3, false, true>(ctx, a, 16, s);
        return launch_fused<true, false, false, 3, false, true>(ctx, a, 17, s);
    }
#define FUSED_CASE(LL, BASE, V)                                                           \
    do {                                                                                  \
        if (pow2 && ctx->fused_minb == 3) return launch_fused<LL, BASE, true, 3, false>(ctx, a, 6 + (V), s);   \
        if (pow2) return launch_fused<LL, BASE, true, 4, false>(ctx, a, 2 * (V), s);      \
        return launch_fused<LL, BASE, false, 3, false>(ctx, a, 2 * (V) + 1, s);           \
    } while (0)
    if (loglik_dev && base_dev) FUSED_CASE(true, true, 0);
    else if (loglik_dev) FUSED_CASE(true, false, 1);
    else FUSED_CASE(false, true, 2);
#undef FUSED_CASE
    return GSE_OK;
}

// Sharded population: the same kernel on every rank, two mailbox exchanges inside it (shard totals after the local
// scan, "all my writes are out" at the end).  Scan, search, the collectives and the exchange of ancestor indices are ONE
// launch per rank; nothing goes through the host or NCCL.
extern "C" int gse_resample_fused_sharded(gse_ctx* ctx, const float* loglik_dev, const double* base_dev,
                                          const double* stats_dev, double r, const gse_shards* sh,
                                          void* const mailboxes[GSE_MAX_SHARDS], int rank, unsigned int epoch_totals,
                                          unsigned int epoch_done, uint64_t* total_dev, const float* state_dev,
                                          int64_t ld, double* moments_dev, int reset_stats, void* stream) {
    GSE_REQUIRE(ctx != NULL && stats_dev != NULL && sh != NULL, "ctx / stats / shards is NULL");
    GSE_REQUIRE(moments_dev == NULL || (state_dev != NULL && loglik_dev != NULL && base_dev == NULL),
                "the in-kernel estimate needs the state and log-likelihood weights");
    GSE_REQUIRE(loglik_dev != NULL || base_dev != NULL, "need loglik or base weights");
    GSE_REQUIRE(loglik_dev == NULL || aligned32(loglik_dev), "loglik must be 32-byte aligned");
    GSE_REQUIRE(base_dev == NULL || aligned32(base_dev), "base must be 32-byte aligned");
    GSE_REQUIRE(sh->nshards >= 1 && sh->nshards <= GSE_MAX_SHARDS && rank >= 0 && rank < sh->nshards, "bad shard table");
    GSE_REQUIRE(epoch_totals != 0 && epoch_done != 0 && epoch_totals != epoch_done, "bad mailbox epochs");
    const int64_t n_total = sh->rows[sh->nshards];
    GSE_REQUIRE(sh->rows[0] == 0 && n_total >= 1 && n_total <= 0x7ffffff0ll, "global row count out of range (int32 index)");
    const int64_t n_src = sh->rows[rank + 1] - sh->rows[rank];
    GSE_REQUIRE(n_src >= 1 && n_src <= ctx->n_max, "n_src out of range for this context");
    // this rank may source every output of the population: the heavy-run queue must hold n_total / 4096 runs in pieces
    GSE_REQUIRE(n_total / RF_HEAVY_MIN + n_total / RF_PIECE + 2 <= (int64_t)ctx->heavy_queue_cap,
                "workspace too small (create the context with workspace rows >= global rows)");
    GSE_REQUIRE(r >= 0.0 && r < 1.0, "r must be in [0, 1)");
    gse_device_guard guard(ctx->device);
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    a.loglik = loglik_dev;
    a.base = base_dev;
    a.stats = stats_dev;
    a.n_src = n_src;
    a.first_shard = (rank == 0) ? 1 : 0;
    a.total_out = total_dev;
    a.stats_reset = reset_stats ? const_cast<double*>(stats_dev) : NULL;
    a.trace = ctx->fused_trace;
    a.nshards = sh->nshards;
    a.rank = rank;
    a.epoch_totals = epoch_totals;
    a.epoch_done = epoch_done;
    for (int t = 0; t <= sh->nshards; ++t) a.shard_lo[t] = (int)sh->rows[t];
    for (int t = 0; t < sh->nshards; ++t) {
        GSE_REQUIRE(sh->idx_dev[t] != NULL && (((uintptr_t)sh->idx_dev[t]) & 15u) == 0 && sh->rows[t] % 4 == 0,
                    "shard index buffers must be non-NULL and 16-byte aligned, shards cut at multiples of four rows");
        a.shard_idx[t] = sh->idx_dev[t];
    }
    int rc = gse_build_mailboxes(mailboxes, rank, sh->nshards, &a.mb);
    if (rc) return rc;
    fill_common(ctx, a, r, n_total, 0, n_total, sh->idx_dev[rank], sh->rows[rank]);
    a.out_home_lo = (int)sh->rows[rank];
    a.out_home_hi = (int)sh->rows[rank + 1];
    cudaStream_t s = (cudaStream_t)stream;
    const bool pow2 = (n_total & (n_total - 1)) == 0;
    if (moments_dev) {
        GSE_REQUIRE(ld >= n_src, "ld < n_src");
        a.mean_state = state_dev;
        a.mean_ld = ld;
        a.mean_partials = ctx->red_partials;
        a.mean_out = moments_dev;
        if (pow2) return launch_fused<true, false, true, 3, true, true>(ctx, a, 18, s);
        return launch_fused<true, false, false, 3, true, true>(ctx, a, 19, s);
    }
#define SHARDED_CASE(LL, BASE, V)                                                         \
    do {                                                                                  \
        if (pow2) return launch_fused<LL, BASE, true, 3, true>(ctx, a, 9 + 2 * (V), s);   \
        return launch_fused<LL, BASE, false, 3, true>(ctx, a, 10 + 2 * (V), s);           \
    } while (0)
    if (loglik_dev && base_dev) SHARDED_CASE(true, true, 0);
    else if (loglik_dev) SHARDED_CASE(true, false, 1);
    else SHARDED_CASE(false, true, 2);
#undef SHARDED_CASE
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
    const int64_t ntiles = gse_div_up(n_src, RF_TILE);
    const int64_t max_warps = (int64_t)ctx->num_sms * 8 * RF_WARPS;
    const int64_t tiles_per_warp = gse_div_up(ntiles, max_warps);
    const int64_t blocks = gse_div_up(gse_div_up(ntiles, tiles_per_warp), RF_WARPS);
    const bool pow2 = (n_total & (n_total - 1)) == 0;
#define F64_CASE(P, T) k_resample_search_f64<P, T><<<(unsigned)blocks, RF_THREADS, 0, s>>>(a, tiles_per_warp)
    if (pow2) { if (ties_right) F64_CASE(true, true); else F64_CASE(true, false); }
    else { if (ties_right) F64_CASE(false, true); else F64_CASE(false, false); }
#undef F64_CASE
    GSE_CHECK_LAUNCH(ctx);
    k_resample_drain<<<(unsigned)ctx->num_sms, RF_THREADS, 0, s>>>(a);
    GSE_CHECK_LAUNCH(ctx);
    return GSE_OK;
}
