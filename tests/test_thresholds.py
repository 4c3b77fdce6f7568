"""Exactness of the integer resampling thresholds (csrc/gse_common.cuh: gse_threshold) against the
reference's own comparison  cumsum[k]/cumsum[-1] < (i + r)/N  (particle.py:89-98) evaluated with
numpy float64 on exact Python integers.  Host build of the same inline function the kernels use.
CPU only."""
import random

import numpy

from gpu_se_b200 import _lib

L = _lib.lib


def g(C, T):
    return numpy.float64(C) / numpy.float64(T)      # int -> float64 is round-to-nearest-even


def test_threshold_is_the_exact_inverse_of_the_reference_comparison():
    random.seed(1)
    for _ in range(40000):
        k = random.choice([1, 3, 10, 20, 30, 40, 50, 51, 52, 53])      # totals stay below 2^53 by construction
        T = random.randrange(max(1, 2 ** (k - 1)), 2 ** k)
        mode = random.random()
        if mode < 0.3:
            C = random.randrange(0, T + 1)
            u = float(g(C, T))
            if random.random() < 0.6:
                u = float(numpy.nextafter(u, random.choice([0.0, 2.0])))
        elif mode < 0.6:
            N = random.choice([16, 1000, 2 ** 20, 2 ** 24, 12345])
            u = (random.randrange(N) + random.random()) / N
        elif mode < 0.7:
            u = 2.0 ** -random.randrange(0, 60)                      # powers of two (the half-ulp boundary below)
            if random.random() < 0.5:
                u = float(numpy.nextafter(u, random.choice([0.0, 2.0])))
        else:
            u = random.random()
        u = min(max(u, 0.0), 1.0)
        q = L.gse_threshold_u64(u, T)
        assert q <= T and g(q, T) >= u and (q == 0 or g(q - 1, T) < u), (T, u, q)


def test_count_outputs_below_matches_searchsorted():
    rng = numpy.random.default_rng(3)
    for N in (7, 64, 1000):
        w = rng.integers(0, 1 << 30, N).astype(object)
        C = numpy.cumsum(w)
        T = int(C[-1])
        r = float(rng.random())
        cn = numpy.array([float(c) for c in C]) / float(T)
        u = (numpy.arange(N) + r) / N
        idx = numpy.searchsorted(cn, u, side="left")
        for k in (0, N // 3, N - 1):
            # outputs sourced at or below source k == number of idx <= k
            assert L.gse_count_outputs_below(int(C[k]), T, r, N) == int((idx <= k).sum())
