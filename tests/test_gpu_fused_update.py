"""-m gpu: predict() directly followed by update() can run as ONE kernel (gse_pf_predict_update; GSE_FUSE_UPDATE=1 -- off
by default, it is no faster, see filter/particle.py).  It must be the two
separate kernels bit for bit: same rows, same log-likelihoods, same maximum; the sum of exp differs in the last float32
bits only (other partition).  Reference pattern: particle.py:265-294 called back to back by every filter loop."""
import numpy
import pytest

from gpu_common import consistent_measurement, make_pf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gpu_se_b200
    return gpu_se_b200


def _pair(g, n, **kw):
    a, b = make_pf(g, n, seed=11, **kw), make_pf(g, n, seed=11, **kw)
    assert a.can_fuse_update()
    a._fuse = True                        # (off by default: GSE_FUSE_UPDATE=1)
    b._fuse = False                       # the two-kernel path
    return a, b


def _same(a, b, n):
    assert numpy.array_equal(a.particles.get(), b.particles.get())
    assert numpy.array_equal(a._loglik[:n].cpu().numpy(), b._loglik[:n].cpu().numpy())
    sa, sb = a._stats.cpu().numpy(), b._stats.cpu().numpy()
    assert sa[0] == sb[0]
    assert sa[1] == pytest.approx(sb[1], rel=2e-6)


@pytest.mark.parametrize("n", [1000, 4099, 65536, 300001, (1 << 20) + 3])
def test_fused_predict_update_equals_the_two_kernels(g, n):
    a, b = _pair(g, n)
    rng = numpy.random.default_rng(5)
    l0 = (a._ctx.launches, b._ctx.launches)
    for c in range(5):
        u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
        z = consistent_measurement(u, 0.5, rng)
        for f in (a, b):
            f.predict(u, 0.5)
            f.update(u, z)
        if c == 2:
            # no resample: the next update accumulates onto this log-likelihood (loglik_in != NULL), rows in place
            _same(a, b, n)
            continue
        r = float(rng.uniform())
        ia = a.resample(r=r, return_index=True).cpu().numpy()
        ib = b.resample(r=r, return_index=True).cpu().numpy()
        assert numpy.array_equal(ia, ib)
    for f in (a, b):
        f.predict(u, 0.5)                 # through the pending ancestor index
        f.update(u, z)
    _same(a, b, n)
    # the fused kernel and its one-CTA merge of the block partials: as many launches as the two-kernel path
    assert (b._ctx.launches - l0[1]) == (a._ctx.launches - l0[0])


def test_a_call_between_predict_and_update_runs_them_separately(g):
    n = 5000
    a, b = _pair(g, n)
    u, z = numpy.array([0.06, 0.2]), numpy.array([88.0, 66.0])
    for f in (a, b):
        f.predict(u, 0.5)
    pa = a.particles.get()                # flushes the recorded predict
    assert numpy.array_equal(pa, b.particles.get())
    for f in (a, b):
        f.update(u, z)
    _same(a, b, n)
    for f in (a, b):
        f.predict(u, 0.5)
        f.predict(u, 0.25)                # two predicts in a row: the first runs alone, the second fuses
        f.update(u, z)
        f.update(u, z + 0.5)              # and a second update is a plain update
    _same(a, b, n)
    est_a, est_b = a.point_estimate(normalised=True), b.point_estimate(normalised=True)
    assert numpy.allclose(est_a, est_b, rtol=1e-6)


def test_host_noise_and_sub_steps_do_not_fuse(g):
    from oracle import mixture
    n = 4096
    pf = make_pf(g, n, seed=2, n_sub=2)
    assert not pf.can_fuse_update()
    pf = make_pf(g, n, seed=2)
    pf._fuse = True
    noise = numpy.random.default_rng(0).normal(size=(n, 5)).astype(numpy.float32) * 1e-3
    ref = make_pf(g, n, seed=2)
    ref._fuse = False
    u, z = numpy.array([0.06, 0.2]), numpy.array([88.0, 66.0])
    for f in (pf, ref):
        f.predict(u, 0.5, noise=noise)    # host-supplied noise: the predict runs at once
        f.update(u, z)
    _same(pf, ref, n)
    del mixture
