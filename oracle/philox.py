"""Specification (numpy restatement) of the device sampler.  TEST INFRASTRUCTURE ONLY.

This is not a restatement of reference code -- the reference draws its noise with
numpy/cupy ``random.choice`` + ``multivariate_normal`` (MultivariateGaussianSum.py:79-95), which a
counter-based in-kernel generator cannot and need not reproduce.  It restates
THIS repo's sampler (csrc/gse_common.cuh: philox4x32_10, box_muller, draw_mixture5) so that the
kernels' draws can be checked value by value:

* Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11;
  pinned by the Random123 known-answer vectors in tests/test_philox.py.
* uniform: (x + 0.5) * 2^-32 in float32; Box-Muller with the angle folded to (-pi, pi].
* grouped layout (particle-filter predict): see ``grouped_normals5``.
* per-row layout (initial draw, GS-UKF sigma points, MultivariateGaussianSum.draw) for row ``index`` at
  ``step``, subsequence ``sub``:
    A = philox(ctr=(index_lo, index_hi, step, 2*sub),   key=(seed_lo, seed_hi))
    B = philox(ctr=(index_lo, index_hi, step, 2*sub+1), key=(seed_lo, seed_hi))
    (z0, z1) = BM(A0, A1), (z2, z3) = BM(A2, A3), (z4, _) = BM(B0, B1), component from B2.
"""
import numpy

M0, M1 = numpy.uint64(0xD2511F53), numpy.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = numpy.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; counters are uint32 arrays (broadcastable), key two Python ints."""
    c0, c1, c2, c3 = (numpy.asarray(c, dtype=numpy.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = numpy.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> numpy.uint64(32)) ^ c1 ^ numpy.uint64(k0)
        n2 = (p0 >> numpy.uint64(32)) ^ c3 ^ numpy.uint64(k1)
        c1 = p1 & MASK
        c3 = p0 & MASK
        c0, c2 = n0 & MASK, n2 & MASK
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(numpy.uint32), c1.astype(numpy.uint32), c2.astype(numpy.uint32), c3.astype(numpy.uint32))


def u32_to_unit(x):
    x = numpy.asarray(x, dtype=numpy.uint32).astype(numpy.float32)     # round-to-nearest like I2F
    return (x.astype(numpy.float64) * 2.0 ** -32 + 2.0 ** -33).astype(numpy.float32)   # one fma rounding


def box_muller(a, b):
    u1 = u32_to_unit(a).astype(numpy.float64)
    u2 = u32_to_unit(b).astype(numpy.float64)
    r = numpy.sqrt(-2.0 * numpy.log(u1))
    th = (u2 * numpy.float64(numpy.float32(6.2831853071795865))
          - numpy.float64(numpy.float32(3.1415926535897932))).astype(numpy.float32).astype(numpy.float64)
    return r * numpy.cos(th), r * numpy.sin(th)


def standard_normals5(index, step, sub, seed):
    """(n, 5) float64 standard normals and (n,) float32 component uniforms for rows ``index``."""
    index = numpy.asarray(index, dtype=numpy.uint64)
    lo, hi = index & MASK, index >> numpy.uint64(32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    A = philox4x32_10(lo, hi, numpy.uint64(step & 0xFFFFFFFF), numpy.uint64(2 * sub), k0, k1)
    B = philox4x32_10(lo, hi, numpy.uint64(step & 0xFFFFFFFF), numpy.uint64(2 * sub + 1), k0, k1)
    z0, z1 = box_muller(A[0], A[1])
    z2, z3 = box_muller(A[2], A[3])
    z4, _ = box_muller(B[0], B[1])
    return numpy.stack([z0, z1, z2, z3, z4], axis=-1), u32_to_unit(B[2])


def grouped_normals5(index, step, seed):
    """(n, 5) float64 standard normals and (n,) float32 component uniforms of the GROUPED layout the
    particle-filter predict kernel uses (csrc/gse_common.cuh: draw_mixture5_x4): rows 4g..4g+3 share
    six Philox calls  P_j = philox(ctr=(g_lo, g_hi, step, 0x80000000 + j)),  j = 0..5;  words
    w[4j..4j+3] = P_j; normals (n[2p], n[2p+1]) = BM(w[2p], w[2p+1]) for p = 0..9; row r of the group
    takes n[5r..5r+4] and the selector word w[20 + r]."""
    index = numpy.asarray(index, dtype=numpy.uint64)
    group, lane = index >> numpy.uint64(2), (index & numpy.uint64(3)).astype(numpy.int64)
    lo, hi = group & MASK, group >> numpy.uint64(32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    words = []
    for j in range(6):
        words += list(philox4x32_10(lo, hi, numpy.uint64(step & 0xFFFFFFFF), numpy.uint64(0x80000000 + j), k0, k1))
    w = numpy.stack(words, axis=-1)                                   # (n, 24)
    z = numpy.empty((len(index), 20))
    for p in range(10):
        z[:, 2 * p], z[:, 2 * p + 1] = box_muller(w[:, 2 * p], w[:, 2 * p + 1])
    rows = numpy.arange(len(index))
    normals = numpy.stack([z[rows, 5 * lane + k] for k in range(5)], axis=-1)
    return normals, u32_to_unit(w[rows, 20 + lane])


def sigma_grouped_normals(index, step, seed):
    """(n, 11, 5) float64 standard normals and (n, 11) float32 component uniforms of the layout the GS-UKF predict kernel
    uses for the eleven sigma points of component ``index`` (csrc/gse_common.cuh: SigmaNoise): 17 Philox calls
    P_j = philox(ctr=(i_lo, i_hi, step, 0x40000000 + j)), j = 0..16; words w[4j..4j+3] = P_j; normals
    (n[2p], n[2p+1]) = BM(w[2p], w[2p+1]) for p = 0..27; sigma point s takes n[5s..5s+4] and the selector word w[56 + s]."""
    index = numpy.asarray(index, dtype=numpy.uint64)
    lo, hi = index & MASK, index >> numpy.uint64(32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    words = []
    for j in range(17):
        words += list(philox4x32_10(lo, hi, numpy.uint64(step & 0xFFFFFFFF), numpy.uint64(0x40000000 + j), k0, k1))
    w = numpy.stack(words, axis=-1)                                   # (n, 68)
    z = numpy.empty((len(index), 56))
    for p in range(28):
        z[:, 2 * p], z[:, 2 * p + 1] = box_muller(w[:, 2 * p], w[:, 2 * p + 1])
    return z[:, :55].reshape(len(index), 11, 5), u32_to_unit(w[:, 56:67])


def draw_mixture5_sigma(means, covariances, weights, index, step, seed):
    """(n, 11, 5) float64 state-noise samples of the GS-UKF predict kernel (one independent draw per sigma point)."""
    z, uc = sigma_grouped_normals(index, step, seed)
    out = [_mix(means, covariances, weights, z[:, s, :], uc[:, s])[0] for s in range(11)]
    return numpy.stack(out, axis=1)


def draw_mixture5_grouped(means, covariances, weights, index, step, seed):
    """As draw_mixture5, with the grouped layout of the predict kernel."""
    return _mix(means, covariances, weights, *grouped_normals5(index, step, seed))


def draw_mixture5(means, covariances, weights, index, step, sub, seed):
    """(n, 5) float64 samples the device sampler produces for these rows (up to MUFU rounding)."""
    means = numpy.asarray(means, dtype=numpy.float32).astype(numpy.float64)
    covs = numpy.asarray(covariances, dtype=numpy.float32).astype(numpy.float64)
    w = numpy.asarray(weights, dtype=numpy.float32).astype(numpy.float64)
    cdf = (numpy.cumsum(w / w.sum())).astype(numpy.float32)
    cdf[-1] = 1.0
    return _mix(means, covariances, weights, *standard_normals5(index, step, sub, seed))


def _mix(means, covariances, weights, z, uc):
    means = numpy.asarray(means, dtype=numpy.float32).astype(numpy.float64)
    covs = numpy.asarray(covariances, dtype=numpy.float32).astype(numpy.float64)
    w = numpy.asarray(weights, dtype=numpy.float32).astype(numpy.float64)
    cdf = (numpy.cumsum(w / w.sum())).astype(numpy.float32)
    cdf[-1] = 1.0
    comp = numpy.zeros(z.shape[0], dtype=numpy.int64)
    for d in range(len(w) - 1):
        comp += (uc > cdf[d]).astype(numpy.int64)
    L = numpy.linalg.cholesky(covs).astype(numpy.float32).astype(numpy.float64)
    return means[comp] + numpy.einsum('nij,nj->ni', L[comp], z), comp
