"""-m gpu parity tests of the resampling contract (BASELINE.json north_star: "resampled index arrays must be
bit-exact against the reference when both are fed identical weights and uniform offset"):

 (i)   weights with exact partial sums          -> bit-exact (tests/test_gpu_particle.py)
 (ii)  the reference's own float64 cumulative sum fed to ``resample_from_cumsum`` -> bit-exact, any N   (here)
 (iii) generic float64 weights: the device's exact fixed-point scan against numpy's sequentially rounded
       ``cumsum`` -> a handful of +-1 differences at 2^22 / 2^24, counted and bounded                 (here)
plus the fused scan + rank + fill kernel against the two-stage scan / merge-path search and against numpy on
its own cumulative weights, including degenerate weight vectors (the heavy-run queue).
"""
import json
import os

import numpy
import pytest

from conftest import ROOT, golden
from gpu_common import consistent_measurement, expected_indices_from_cumsum, make_pf
from oracle import particle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gpu_se_b200
    return gpu_se_b200


def _record(name, value):
    """Counts the DESIGN.md tables quote: appended to gpurun_out/resample_parity.json when the directory exists."""
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        path = os.path.join(out, "resample_parity.json")
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[name] = value
        json.dump(data, open(path, "w"), indent=1, sort_keys=True)


# ---- (ii) resample_from_cumsum ------------------------------------------------------------------------------
def test_f64_cumsum_matches_the_reference_gpu_kernel(g):
    """tests/golden/nicely_cudasim.npz: the reference's hand-written _parallel_resample (particle.py:223-263) run
    through numba's CUDA simulator on its own normalised cumsum.  side='right' is that kernel, side='left' the
    reference CPU loop on the same array."""
    from gpu_se_b200.filter import resample_from_cumsum
    z = golden("nicely_cudasim.npz")
    for t in ("a", "b"):
        c, r = z["cumsum_" + t], float(z["r_" + t])
        got = resample_from_cumsum(c, r, side="right").cpu().numpy()
        assert numpy.array_equal(got, z["idx_" + t]), t
        left = resample_from_cumsum(c, r, side="left").cpu().numpy()
        assert numpy.array_equal(left, particle.systematic_indices(c, r)), t


@pytest.mark.parametrize("log2n", [10, 16, 20, 22, 24])
@pytest.mark.parametrize("normalise", [False, True])
def test_f64_cumsum_bit_exact_against_numpy(g, log2n, normalise):
    """numpy.cumsum of random float64 weights (particle.py:89-90), fed back: 0 mismatches at every N."""
    from gpu_se_b200.filter import resample_from_cumsum
    n = 1 << log2n
    rng = numpy.random.default_rng(log2n)
    w = rng.random(n)
    w /= w.sum()                                   # pf_run_seq.py:124-125
    c = numpy.cumsum(w)
    cn = c / c[-1]
    for r in (0.0, float(rng.random()), 1.0 - 2.0 ** -53):
        got = resample_from_cumsum(c if normalise else cn, r, normalise=normalise).cpu().numpy()
        u = (numpy.arange(n, dtype=numpy.float64) + r) / n
        expect = numpy.searchsorted(cn, u, side="left")
        bad = int((got != expect).sum())
        assert bad == 0, (log2n, r, bad)
        right = resample_from_cumsum(cn, r, side="right").cpu().numpy()
        assert numpy.array_equal(right, numpy.minimum(numpy.searchsorted(cn, u, side="right"), n - 1)), (log2n, r)


@pytest.mark.parametrize("n,n_out", [(1, 1), (5, 17), (1000, 37), (37, 1000), (4097, 100003), (100003, 4097)])
def test_f64_cumsum_ragged_ties_and_subsampling(g, n, n_out):
    """Repeated values (zero weights), n_out != n, non-power-of-two N (u = (i + r) / N with a real division)."""
    from gpu_se_b200.filter import resample_from_cumsum
    rng = numpy.random.default_rng(n * 7 + n_out)
    w = rng.integers(0, 1 << 12, n).astype(numpy.float64)
    w[rng.random(n) < 0.4] = 0.0
    w[-1] += 1.0
    c = numpy.cumsum(w)
    cn = c / c[-1]
    for r in (0.0, 0.5, float(rng.random())):
        u = (numpy.arange(n_out, dtype=numpy.float64) + r) / n_out
        got = resample_from_cumsum(cn, r, n_out=n_out).cpu().numpy()
        assert numpy.array_equal(got, numpy.searchsorted(cn, u, side="left")), (n, n_out, r)
        got = resample_from_cumsum(c, r, n_out=n_out, normalise=True, side="right").cpu().numpy()
        assert numpy.array_equal(got, numpy.minimum(numpy.searchsorted(cn, u, side="right"), n - 1)), (n, n_out, r)


# ---- (iii) generic float64 weights ---------------------------------------------------------------------------
@pytest.mark.parametrize("log2n,bound", [(16, 0), (20, 2), (22, 12), (24, 80)])
def test_generic_f64_weights_mismatch_count(g, log2n, bound):
    """A parallel scan cannot reproduce numpy.cumsum's sequential rounding in general: the device sums the
    quantised weights exactly, numpy rounds every partial sum.  Where a sample position falls between the two
    cumulative values the ancestor differs by one.  Expected from the rounding-error random walk: ~0 / 2 / 25 at
    2^20 / 2^22 / 2^24 (emulation gave 0 / 2 / 25).  Every difference must be exactly +-1."""
    n = 1 << log2n
    total, worst = 0, 0
    for seed in range(3):
        rng = numpy.random.default_rng(100 * log2n + seed)
        w = rng.random(n)
        w /= w.sum()
        r = float(rng.random())
        pf = make_pf(g, n, particles=numpy.zeros((n, 5), dtype=numpy.float32))
        pf.weights = w
        idx = pf.resample(r=r, return_index=True).cpu().numpy()
        c = numpy.cumsum(w)
        c /= c[-1]
        expect = numpy.searchsorted(c, (numpy.arange(n, dtype=numpy.float64) + r) / n, side="left")
        d = idx - expect
        assert numpy.abs(d).max() <= 1
        bad = int((d != 0).sum())
        total += bad
        worst = max(worst, bad)
        del pf
    _record("generic_f64_mismatches_2p%d" % log2n, {"per_seed_max": worst, "sum_over_3_seeds": total})
    print("generic float64 weights, N = 2^%d: %d mismatches over 3 seeds (max %d per resample), all +-1"
          % (log2n, total, worst))
    assert worst <= bound


# ---- fused kernel ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [1, 7, 255, 256, 257, 2047, 2049, 5000, 65536, 100003, 1 << 20, (1 << 22) + 12345])
def test_fused_equals_two_stage_and_numpy_after_update(g, N, monkeypatch):
    """After a real update: fused scan + rank + fill == scan + merge-path search == searchsorted on the device's own
    integer cumulative weights."""
    from gpu_se_b200.filter import _base
    rng = numpy.random.default_rng(N)
    u = numpy.array([0.06, 0.2])
    z = consistent_measurement(u, 1.0, rng)
    r = float(rng.random())
    res = {}
    for mode in (True, False):
        monkeypatch.setattr(_base, "FUSED_RESAMPLE", mode)
        pf = make_pf(g, N, seed=5)
        pf.predict(u, 1.0)
        pf.update(u, z)
        c, total = pf.cumulative_weights()
        idx = pf.resample(r=r, return_index=True).cpu().numpy()
        assert numpy.array_equal(idx, expected_indices_from_cumsum(c, r)), (N, mode)
        res[mode] = idx
    assert numpy.array_equal(res[True], res[False])


@pytest.mark.parametrize("N", [1 << 12, 1 << 20, 1 << 24])
def test_fused_degenerate_weights_use_the_heavy_queue(g, N):
    """All the weight on one, two or three particles: runs of up to N outputs of one source go through the queue."""
    rng = numpy.random.default_rng(N + 1)
    for case in range(4):
        w = numpy.zeros(N)
        if case == 0:
            w[N // 3] = 1.0
        elif case == 1:
            w[0], w[N - 1] = 0.25, 0.75
        elif case == 2:
            w[[5, N // 2, N - 2]] = [0.5, 0.125, 0.375]
        else:                                      # heavy runs next to ordinary ones
            w[:] = rng.integers(0, 4, N)
            w[rng.integers(0, N, 6)] = N / 2.0
        pf = make_pf(g, N, particles=numpy.zeros((N, 5), dtype=numpy.float32))
        pf.weights = w
        for r in (0.0, 0.37):
            c = numpy.cumsum(w)
            c /= c[-1]
            expect = particle.systematic_indices(c, r) if N <= (1 << 12) else \
                numpy.searchsorted(c, (numpy.arange(N, dtype=numpy.float64) + r) / N, side="left")
            idx = pf.resample(r=r, return_index=True).cpu().numpy()
            assert numpy.array_equal(idx, expect), (N, case, r)
            pf.weights = w
        del pf


def test_fused_assigned_weights_times_likelihood(g):
    """weights assigned by the caller AND updated afterwards (base * exp(loglik)): fused == two-stage."""
    from gpu_se_b200.filter import _base
    N = 300007
    rng = numpy.random.default_rng(3)
    w = rng.random(N)
    u = numpy.array([0.05, 0.25])
    z = consistent_measurement(u, 0.5, rng)
    out = []
    for mode in (True, False):
        _base.FUSED_RESAMPLE = mode
        try:
            pf = make_pf(g, N, seed=9)
            pf.weights = w
            pf.update(u, z)
            c, total = pf.cumulative_weights()
            idx = pf.resample(r=0.123, return_index=True).cpu().numpy()
            assert numpy.array_equal(idx, expected_indices_from_cumsum(c, 0.123))
            out.append(idx)
        finally:
            _base.FUSED_RESAMPLE = True
    assert numpy.array_equal(out[0], out[1])


def test_all_zero_weights_raise_at_the_next_synchronisation(g):
    N = 4096
    pf = make_pf(g, N)
    pf.weights = numpy.zeros(N)
    pf.resample(r=0.5)
    with pytest.raises(FloatingPointError):
        pf.point_estimate()


@pytest.mark.parametrize("N,out0,n_out", [(5000, 0, 5000), (5000, 123, 517), (100003, 4096, 50000), (100003, 99999, 4),
                                          (1 << 20, 777, 300001)])
def test_fused_abi_output_subrange(g, N, out0, n_out):
    """gse_resample_fused called directly with a clipped output range [out0, out0 + n_out) (what a shard of a
    multi-GPU population asks for) and an unaligned destination."""
    import ctypes
    import torch
    from gpu_se_b200 import _lib
    rng = numpy.random.default_rng(N + out0)
    w = rng.random(N) ** 8                     # skewed
    w[rng.random(N) < 0.5] = 0.0
    pf = make_pf(g, N, particles=numpy.zeros((N, 5), dtype=numpy.float32))
    pf.weights = w
    c, total = pf.cumulative_weights()
    r = 0.618
    expect = expected_indices_from_cumsum(c, r)[out0:out0 + n_out]
    idx = torch.full((n_out + 64,), -7, dtype=torch.int32, device=pf.device)
    ll, base, stats = pf._weight_sources()
    _lib.check(_lib.lib.gse_resample_fused(pf._ctx.handle, ll, base, stats.data_ptr(), N, r, N, out0, n_out, 0,
                                           idx.data_ptr(), None, None, 0, None, 0, pf._stream()))
    got = idx.cpu().numpy()
    assert numpy.array_equal(got[:n_out], expect)
    assert (got[n_out:] == -7).all()           # nothing written past the range


# ---- the estimate of the resampled population out of the resample kernel --------------------------------------
def _cycle(pf, rng, with_estimate=True):
    u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
    pf.predict(u, 0.5)
    pf.update(u, consistent_measurement(u, 0.5, rng))
    pf.resample(r=float(rng.uniform()))
    return pf.point_estimate() if with_estimate else None


@pytest.mark.parametrize("n", [1000, 4096, 100003, 1 << 20, (1 << 21) + 12])
def test_estimate_inside_the_resample_kernel(g, n):
    """point_estimate() straight after resample() (particle.py:105-108 after :85-103, every filter loop of the
    reference): from the second cycle on the resample kernel itself leaves sum_k offspring_k x_k behind.  It must equal
    the mean of the gathered rows -- the products are exact in float64, only the summation order differs -- and the
    filter must go on exactly as one that never asked for an estimate."""
    pf, plain = make_pf(g, n, seed=5), make_pf(g, n, seed=5)
    rng, rng2 = numpy.random.default_rng(8), numpy.random.default_rng(8)
    for c in range(4):
        est = _cycle(pf, rng)
        _cycle(plain, rng2, with_estimate=False)
        in_kernel = pf._mom_valid and not pf._mom_unused and pf._est_hint and c > 0
        assert in_kernel == (c > 0)
        rows = pf.particles.get().astype(numpy.float64)               # materialises the pending gather
        # (the first estimate comes from k_means: float32 sums of four rows, then float64)
        assert numpy.allclose(est, rows.mean(axis=0), rtol=1e-12 if c > 0 else 1e-6, atol=0)
        assert numpy.array_equal(plain.particles.get(), rows.astype(numpy.float32))
    # the caller stops reading estimates: one unread estimate later the kernel is the plain one again
    _cycle(pf, rng, with_estimate=False)
    assert pf._mom_unused
    _cycle(pf, rng, with_estimate=False)
    assert not pf._est_hint and not pf._mom_from_resample
    # the covariance needs the second moments: the separate moments kernel takes over
    est = _cycle(pf, rng)
    cov = pf.point_covariance()
    rows = pf.particles.get().astype(numpy.float64)
    assert numpy.allclose(est, rows.mean(axis=0), rtol=1e-6)
    assert cov == pytest.approx(numpy.linalg.svd(numpy.cov(rows.T, bias=True), compute_uv=False)[0], rel=1e-6)
    est = _cycle(pf, rng)
    assert not pf._mom_from_resample
    assert numpy.allclose(est, pf.particles.get().astype(numpy.float64).mean(axis=0), rtol=1e-6)


def test_estimate_inside_the_resample_kernel_one_heavy_row(g):
    """All outputs descend from one row (heavy-run queue): its offspring count is the whole population."""
    from oracle import bioreactor
    n = 1 << 16
    rng = numpy.random.default_rng(2)
    rows0 = bioreactor.X_STEADY[None, :] * (1.0 + 1e-3 * rng.normal(size=(n, 5)))
    rows0[12345, 0] += 2.0                                   # glucose far from everyone else's: 360 apart in the output
    pf = make_pf(g, n, seed=3, particles=rows0)
    pf.resample(r=0.5)                                       # uniform weights: the identity
    pf.point_estimate()                                      # ... and the pattern the filter looks for
    assert pf._est_hint
    u = numpy.array([0.06, 0.2])
    pf.predict(u, 0.5)
    x = pf.particles.get()
    pf.update(u, numpy.array([x[12345, 0] * 180.0 + 0.1, x[12345, 2] * 116.0]))
    idx = pf.resample(r=0.37, return_index=True).cpu().numpy()
    est = pf.point_estimate()
    assert pf._mom_valid and not pf._mom_unused
    assert (idx == 12345).all()
    assert numpy.allclose(est, x[12345].astype(numpy.float64), rtol=1e-12, atol=0)


def test_estimate_inside_the_resample_kernel_graph_replay(g):
    n = 1 << 14
    pf, eager = make_pf(g, n, seed=9), make_pf(g, n, seed=9)
    pf.enable_graphs()
    rng, rng2 = numpy.random.default_rng(4), numpy.random.default_rng(4)
    for c in range(6):
        a, b = _cycle(pf, rng), _cycle(eager, rng2)
        assert numpy.allclose(a, b, rtol=1e-12 if c > 0 else 1e-6, atol=0)
    assert pf.graph_replays >= 4
    assert numpy.array_equal(pf.particles.get(), eager.particles.get())
