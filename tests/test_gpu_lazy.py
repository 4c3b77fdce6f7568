"""-m gpu tests of the lazy resample: ``resample()`` leaves the ancestor index pending and the next
``predict`` / moments kernel reads its rows through it.  Every lazy path must give exactly what the
materialised path (particles[sample_index] first, particle.py:102) gives."""
import ctypes

import numpy
import pytest

from conftest import ulp32
from gpu_common import consistent_measurement, expected_indices_from_cumsum, make_gsf, make_pf
from oracle import mixture, particle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gpu_se_b200
    return gpu_se_b200


def _cycle_inputs(N, seed):
    rng = numpy.random.default_rng(seed)
    u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
    z = consistent_measurement(u, 0.1, rng)
    noise = (rng.normal(size=(N, 5)) * numpy.sqrt(numpy.diag(mixture.STATE_COVS[0]))).astype(numpy.float32)
    return u, z, noise, float(rng.random())


@pytest.mark.parametrize("N", [1, 3, 257, 4099, 65536, 1000003])
@pytest.mark.parametrize("host_noise", [True, False])
def test_lazy_predict_equals_materialised(g, N, host_noise):
    a, b = make_pf(g, N, seed=21), make_pf(g, N, seed=21)
    for c in range(3):
        u, z, noise, r = _cycle_inputs(N, 100 * c + 7)
        nz = noise if host_noise else None
        for pf in (a, b):
            pf.predict(u, 0.1, noise=nz)
            pf.update(u, z)
            pf.resample(r=r)
        assert a._pending and b._pending
        b.particles                                   # materialise b, leave a lazy
        assert a._pending and not b._pending
        assert numpy.array_equal(a.point_estimate(), b.point_estimate())          # gathered moments
        assert a.point_covariance() == b.point_covariance()
        assert a._pending
    u, z, noise, r = _cycle_inputs(N, 999)
    a.predict(u, 0.1, noise=noise if host_noise else None)                        # gathering predict
    b.predict(u, 0.1, noise=noise if host_noise else None)
    a._flush()                                        # (a predict is recorded until the next call: it may fuse with update)
    b._flush()
    assert not a._pending
    assert numpy.array_equal(a.particles.get(), b.particles.get())


def test_lazy_predict_against_oracle(g, noise_pdfs):
    state, meas = noise_pdfs
    N = 5000
    rng = numpy.random.default_rng(3)
    numpy.random.seed(8)
    init = mixture.benchmark_x0(particle.bioreactor.X_STEADY).draw(N)
    pf = make_pf(g, N, particles=init)
    o = particle.ParticleFilterOracle(N, None, state, meas, particles=init)
    u, z, noise, r = _cycle_inputs(N, 5)
    pf.update(u, z)
    c, total = pf.cumulative_weights()
    idx = pf.resample(r=r, return_index=True).cpu().numpy()
    assert numpy.array_equal(idx, expected_indices_from_cumsum(c, r))
    o.particles = init[idx]
    pf.predict(u, 0.1, noise=noise)                   # reads init[idx] through the pending index
    o.predict(u, 0.1, noise=noise)
    assert ulp32(pf.particles.get().astype(numpy.float64) - o.particles, o.particles).max() <= 4.0


def test_update_and_second_resample_after_resample(g):
    N = 3001
    a, b = make_pf(g, N, seed=2), make_pf(g, N, seed=2)
    u, z, noise, r = _cycle_inputs(N, 11)
    for pf in (a, b):
        pf.predict(u, 0.1)
        pf.update(u, z)
        pf.resample(r=r)
    b.particles
    u2, z2, _, r2 = _cycle_inputs(N, 12)
    a.update(u2, z2)                                  # update straight after resample: materialises first
    b.update(u2, z2)
    assert numpy.array_equal(a.weights.get(), b.weights.get())
    ia = a.resample(r=r2, return_index=True).cpu().numpy()
    ib = b.resample(r=r2, return_index=True).cpu().numpy()
    assert numpy.array_equal(ia, ib)
    ia = a.resample(r=0.5, return_index=True).cpu().numpy()       # resample of a pending uniform population
    assert numpy.array_equal(ia, numpy.arange(N))
    assert numpy.array_equal(a.particles.get(), b.particles.get())


def test_weights_after_resample_are_uniform_without_touching_loglik(g):
    N = 777
    pf = make_pf(g, N, seed=4)
    u, z, _, r = _cycle_inputs(N, 1)
    pf.predict(u, 0.1)
    pf.update(u, z)
    pf.resample(r=r)
    assert pf._loglik_zero
    assert numpy.array_equal(pf.weights.get(), numpy.full(N, 1 / N))
    pf.predict(u, 0.1)
    pf.update(u, z)                                   # first update after the reset must not read stale loglik
    w = pf.weights.get()
    fresh = make_pf(g, N, particles=pf.particles.get())
    fresh.update(u, z)
    assert numpy.array_equal(w, fresh.weights.get())


@pytest.mark.parametrize("n_src,out0,n_out,n_total", [(1000, 0, 1000, 1000), (1000, 123, 517, 4000),
                                                       (5000, 4093, 2049, 9000), (7, 0, 40, 40), (4096, 1, 1, 3)])
def test_search_and_gather_abi_subranges(g, n_src, out0, n_out, n_total):
    """gse_resample_search / gse_gather_rows on arbitrary output sub-ranges with a shard offset, as
    the sharded driver calls them (unaligned destinations included)."""
    import torch
    from gpu_se_b200 import _lib
    from gpu_se_b200.filter._base import Context
    rng = numpy.random.default_rng(n_src + out0)
    w = rng.integers(0, 1 << 16, n_src).astype(numpy.uint64)
    w[rng.random(n_src) < 0.2] = 0
    c = numpy.cumsum(w, dtype=numpy.uint64)
    below, above = int(w.sum()) // 3 + 5, int(w.sum()) // 2 + 11
    offset, total = below, below + int(c[-1]) + above
    # outputs [out0, out0 + n_out) of n_total must be sourced by this shard for idx to be in range:
    # shift the range to where that holds
    r = 0.37
    lo_i = int(_lib.lib.gse_count_outputs_below(offset, total, r, n_total))
    hi_i = int(_lib.lib.gse_count_outputs_below(offset + int(c[-1]), total, r, n_total))
    if hi_i - lo_i <= 0:
        pytest.skip("no output falls into this shard")
    out0 = lo_i + min(out0, hi_i - lo_i - 1)
    n_out = min(n_out, hi_i - out0)
    dev = torch.device("cuda", 0)
    ctx = Context(dev, max(n_src, n_out), None, None)
    cs = torch.as_tensor(c.view(numpy.int64), device=dev)
    offtot = torch.tensor([offset, total], dtype=torch.int64, device=dev)
    idx = torch.full((n_out + 8,), -7, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib.gse_resample_search(ctx.handle, cs.data_ptr(), n_src, offtot.data_ptr(), r, n_total, out0,
                                            n_out, idx.data_ptr(), stream))
    got = idx.cpu().numpy()
    cn = (c.astype(numpy.float64) + numpy.float64(offset)) / numpy.float64(total)
    uu = (numpy.arange(out0, out0 + n_out, dtype=numpy.float64) + r) / n_total
    assert numpy.array_equal(got[:n_out], numpy.searchsorted(cn, uu, side="left"))
    assert (got[n_out:] == -7).all()
    # gather into an unaligned window of a wider destination
    src = torch.as_tensor(rng.normal(size=(3, n_src)).astype(numpy.float32), device=dev)
    dst = torch.full((3, n_out + 11), 9.0, dtype=torch.float32, device=dev)
    for shift in (0, 3):
        dst.fill_(9.0)
        _lib.check(_lib.lib.gse_gather_rows(ctx.handle, idx.data_ptr(), n_out, src.data_ptr(), n_src,
                                            dst.data_ptr() + 4 * shift, n_out + 11, 3, None, stream))
        d = dst.cpu().numpy()
        assert numpy.array_equal(d[:, shift:shift + n_out], src.cpu().numpy()[:, got[:n_out]])
        assert (d[:, :shift] == 9.0).all() and (d[:, shift + n_out:] == 9.0).all()
    ctx.close()


def test_gsukf_lazy_predict_equals_materialised(g):
    N = 1500
    a, b = make_gsf(g, N, seed=6), make_gsf(g, N, seed=6)
    for c in range(2):
        u, z, _, r = _cycle_inputs(N, 40 + c)
        for f in (a, b):
            f.predict(u, 0.1)
            f.update(u, z)
            f.resample(r=r)
        b.means
        assert a._pending and not b._pending
        assert numpy.array_equal(a.point_estimate(), b.point_estimate())
        assert a.point_covariance() == b.point_covariance()
    u, z, _, r = _cycle_inputs(N, 77)
    a.predict(u, 0.1)
    b.predict(u, 0.1)
    assert numpy.array_equal(a.means.get(), b.means.get())
    assert numpy.array_equal(a.covariances.get(), b.covariances.get())


def test_result_block_and_wait_abi(g):
    """gse_ctx_result_block / gse_ctx_wait through the C ABI: gse_resample_fused leaves the estimate of the resampled
    population (particle.py:318-320 right after :296-316) in the context's host-mapped block -- no copy of its own -- and
    it equals the mean of the gathered rows; the public classes read their estimate the same way."""
    import ctypes
    import torch
    from gpu_se_b200 import _lib
    from gpu_se_b200.filter._base import Context
    dev = torch.device("cuda", 0)
    n, ld = 50001, 50048
    rng = numpy.random.default_rng(12)
    ctx = Context(dev, n, None, None)
    host, devp = _lib.c_dbl_p(), _lib.c_dbl_p()
    _lib.check(_lib.lib.gse_ctx_result_block(ctx.handle, ctypes.byref(host), ctypes.byref(devp)))
    assert ctypes.cast(devp, ctypes.c_void_p).value == ctx.result_dev and ctx.result_np.shape == (64,)
    x = torch.as_tensor(rng.normal(size=(5, ld)).astype(numpy.float32), device=dev)
    ll = torch.as_tensor((rng.normal(size=ld) * 3).astype(numpy.float32), device=dev)
    stats = torch.tensor([float(ll[:n].max()), float(torch.exp(ll[:n].double() - ll[:n].max().double()).sum()), 0.0, 0.0],
                         dtype=torch.float64, device=dev)
    idx = torch.zeros(ld, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    ctx.result_np[:] = -1.0
    _lib.check(_lib.lib.gse_resample_fused(ctx.handle, ll.data_ptr(), None, stats.data_ptr(), n, 0.625, n, 0, n, 0,
                                           idx.data_ptr(), total.data_ptr(), x.data_ptr(), ld, ctx.result_dev, 1, stream))
    bits = ctypes.c_uint(99)
    _lib.check(_lib.lib.gse_ctx_wait(ctx.handle, stream, ctypes.byref(bits)))
    assert bits.value == 0
    mom = ctx.result_np[:48].copy()
    anc = idx[:n].cpu().numpy()
    assert (numpy.diff(anc) >= 0).all() and anc.min() >= 0 and anc.max() < n
    want = x.cpu().numpy()[:, anc].astype(numpy.float64).sum(axis=1)
    assert mom[0] == n and numpy.allclose(mom[1:6], want, rtol=1e-12, atol=1e-9)
    assert (mom[21:26] == 0.0).all() and mom[41] == 0.0 and mom[42] == n          # pivot 0, uniform (M, S)
    assert stats.cpu().numpy()[:2].tolist() == [0.0, float(n)]                      # reset_stats
    ctx.wait(stream)                                                                 # the Python wrapper of the same call
    ctx.close()
    assert ctx.result_np is None
    # the classes: point_estimate() straight after resample() comes out of that block and equals the moments kernel's
    pf = make_pf(g, 30000)
    u = numpy.array([0.06, 0.2])
    z = consistent_measurement(u, 0.1, rng)
    for k in range(3):
        pf.predict(u, 0.1)
        pf.update(u, z)
        pf.resample(r=0.3 + 0.2 * k)
        est = pf.point_estimate()
        if k > 0:
            assert pf._est_hint and not pf._mom_unused
        ref = pf.particles.get().astype(numpy.float64).mean(axis=0)
        assert numpy.allclose(est, ref, rtol=1e-6)
