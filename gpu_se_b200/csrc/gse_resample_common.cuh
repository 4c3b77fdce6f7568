// Helpers shared by the resampling kernels (gse_resample.cu: scan / merge-path search / gather;
// gse_resample_fused.cu: scan + rank + fill in one launch).
#pragma once

#include "gse_common.cuh"

static inline bool aligned32(const void* p) { return ((uintptr_t)p & 31u) == 0; }

#ifdef __CUDACC__
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

// The one expression every scan kernel evaluates for a weight exp(l - M): ex2(l * log2(e) - M * log2(e)), the product
// exact inside one FMA.  fl(M log2 e) is common to all rows, so its rounding scales every weight by the same factor --
// invisible after normalisation.  (One instruction less per row than __expf(l - M), and no cancellation error in l - M.)
#define GSE_LOG2E 1.4426950408889634f
__device__ __forceinline__ float weight_exp_offset(float M) { return -__fmul_rn(M, GSE_LOG2E); }
__device__ __forceinline__ float weight_exp(float l, float neg_m_log2e) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmaf_rn(l, GSE_LOG2E, neg_m_log2e)));
    return r;
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

// status words of the single-pass kernels: 2 bits state | 8 bits epoch | 54 bits value (values < 2^53), read and
// written with volatile 64-bit accesses so that publishing needs no fence
#define ST_AGGREGATE 1ull
#define ST_PREFIX 2ull
__device__ __forceinline__ uint64_t status_pack(uint64_t state, unsigned int epoch, uint64_t value) {
    return (state << 62) | ((uint64_t)(epoch & 0xffu) << 54) | value;
}
__device__ __forceinline__ uint64_t ld_status(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint64_t* p, uint64_t v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// first global output index i in [dbase, dend] with u_i > g (TIES_RIGHT: u_i >= g), g = fl(cd / Td) -- or cd itself when
// Td == 1 -- starting from a guess.  u_i = (i + r) / N exactly as the reference evaluates it (particle.py:97).
// Out of line and fed with scalars only: passing an argument struct would make every thread spill it.
template <bool POW2, bool TIES_RIGHT>
__device__ __noinline__ double rank_exact_g(double r, double n_total, double inv_n, double g, double di, double dbase,
                                            double dend) {
    di = fmin(fmax(di, dbase), dend);
    if (TIES_RIGHT) {
        while (di > dbase && gse_sample_position(di - 1.0, r, n_total, inv_n, POW2) >= g) di -= 1.0;
        while (di < dend && !(gse_sample_position(di, r, n_total, inv_n, POW2) >= g)) di += 1.0;
    } else {
        while (di > dbase && gse_sample_position(di - 1.0, r, n_total, inv_n, POW2) > g) di -= 1.0;
        while (di < dend && !(gse_sample_position(di, r, n_total, inv_n, POW2) > g)) di += 1.0;
    }
    return di;
}
template <bool POW2>
__device__ __forceinline__ double rank_exact(double r, double n_total, double inv_n, double cd, double Td, double di,
                                             double dbase, double dend) {
    return rank_exact_g<POW2, false>(r, n_total, inv_n, __ddiv_rn(cd, Td), di, dbase, dend);   // cumsum / cumsum[-1]  (:90)
}
#endif
