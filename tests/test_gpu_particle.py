"""-m gpu parity tests of the particle-filter path: CUDA kernels (through the public classes and
the C ABI) against the oracle and the committed golden vectors from the reference."""
import numpy
import pytest

from conftest import golden, ulp32
from gpu_common import consistent_measurement, expected_indices_from_cumsum, make_pf
from oracle import mixture, particle, philox

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gpu_se_b200
    return gpu_se_b200


@pytest.mark.parametrize("name", ["pf_n256.npz", "pf_n1024_dt1.npz"])
def test_golden_cycles(g, noise_pdfs, name):
    """predict / update / resample / estimates against the reference's own outputs and the oracle,
    re-seeded from the reference state each stage so that errors do not compound."""
    state, meas = noise_pdfs
    gv = golden(name)
    N, dt = int(gv["N"]), float(gv["dt"])
    pf = make_pf(g, N, particles=gv["particles0"])
    o = particle.ParticleFilterOracle(N, None, state, meas, particles=gv["particles0"])
    for c in range(int(gv["n_cycles"])):
        u, z, noise = gv["u_%d" % c], gv["z_%d" % c], gv["noise_%d" % c]
        pf.predict(u, dt, noise=noise)
        o.predict(u, dt, noise=noise)
        got = pf.particles.get()
        ref = gv["particles_pred_%d" % c]
        assert ulp32(got.astype(numpy.float64) - o.particles, o.particles).max() <= 4.0      # vs oracle
        assert ulp32(got.astype(numpy.float64) - ref, ref).max() <= 6.0                      # vs reference
        pf.particles = ref
        o.particles = ref.copy()
        assert numpy.allclose(pf.point_estimate(), gv["est_pred_%d" % c], rtol=1e-6)
        assert pf.point_covariance() == pytest.approx(float(gv["cov_pred_%d" % c]), rel=2e-5)

        w_before = o.weights.astype(numpy.float64).copy()
        ll = o.log_likelihood(u, z)
        pf.update(u, z)
        o.update(u, z)
        rw = gv["weights_upd_%d" % c].astype(numpy.float64)
        got_w = pf.weights.get()
        # log-weights: |d log w| <= 1e-6 max(1, |log w|)  
        lw_ref = numpy.log(w_before) + ll
        lw_got = numpy.log(got_w)
        assert (numpy.abs(lw_got - lw_ref) <= 1e-6 * numpy.maximum(1.0, numpy.abs(lw_ref))).all()
        assert (numpy.abs(got_w - rw) <= 2e-5 * rw).all()                                    # vs reference
        assert numpy.allclose(pf.point_estimate(), gv["est_upd_%d" % c], rtol=2e-6)
        assert pf.point_covariance() == pytest.approx(float(gv["cov_upd_%d" % c]), rel=1e-4)

        # resample of the reference's own weights: indices bit-exact
        pf.weights = gv["weights_upd_%d" % c]
        o.weights = gv["weights_upd_%d" % c].copy()
        r = float(gv["r_%d" % c])
        idx = pf.resample(r=r, return_index=True).cpu().numpy()
        oidx = o.resample(r=r)
        assert numpy.array_equal(idx, oidx)
        assert numpy.array_equal(pf.particles.get(), gv["particles_res_%d" % c])
        assert numpy.array_equal(pf.weights.get(), gv["weights_res_%d" % c])
        assert numpy.allclose(pf.point_estimate(), gv["est_res_%d" % c], rtol=1e-6)
        assert pf.point_covariance() == pytest.approx(float(gv["cov_res_%d" % c]), rel=2e-5)


def test_resample_indices_bit_exact_on_reference_weights(g):
    """pf_run_seq.py:123-128: a fresh host float64 weight vector is assigned, then resample()."""
    r = golden("resample.npz")
    for t in ("a", "b", "c", "skew", "dyadic"):
        w = r["weights_" + t]
        N = len(w)
        tagged = numpy.zeros((N, 5), dtype=numpy.float32)
        tagged[:, 0] = numpy.arange(N)
        pf = make_pf(g, N, particles=tagged)
        pf.weights = w
        assert numpy.array_equal(pf.weights.get(), w)
        idx = pf.resample(r=float(r["r_" + t]), return_index=True).cpu().numpy()
        assert numpy.array_equal(idx, r["idx_" + t]), t
        assert numpy.array_equal(pf.particles.get()[:, 0].astype(numpy.int64), r["idx_" + t]), t
        assert numpy.array_equal(pf.weights.get(), numpy.full(N, 1 / N))


@pytest.mark.parametrize("N", [1, 2, 3, 5, 31, 33, 1023, 1025, 4095, 4097, 8191, 12289, 100003])
def test_resample_ragged_sizes_and_ties(g, N):
    """Ragged N (tile / vector tails), zero weights (ties), r = 0 and r close to 1; weights with
    exactly representable partial sums must reproduce the reference indices bit for bit."""
    rng = numpy.random.default_rng(N)
    w = rng.integers(0, 1 << 20, N).astype(numpy.float64)
    w[rng.random(N) < 0.3] = 0.0
    if w.sum() == 0:
        w[-1] = 1.0
    tagged = numpy.zeros((N, 5), dtype=numpy.float32)
    tagged[:, 1] = numpy.arange(N) % 1000
    tagged[:, 0] = numpy.arange(N) // 1000
    for r in (0.0, float(rng.random()), 1.0 - 2.0 ** -53):
        pf = make_pf(g, N, particles=tagged)
        pf.weights = w
        idx = pf.resample(r=r, return_index=True).cpu().numpy()
        c = numpy.cumsum(w)
        c /= c[-1]
        assert numpy.array_equal(idx, particle.systematic_indices(c, r)), (N, r)
        got = pf.particles.get()
        assert numpy.array_equal((got[:, 0] * 1000 + got[:, 1]).astype(numpy.int64), idx)


def test_uniform_weights_resample_is_identity_like(g):
    N = 5000
    pf = make_pf(g, N)
    before = pf.particles.get().copy()
    idx = pf.resample(r=0.5, return_index=True).cpu().numpy()
    assert numpy.array_equal(idx, numpy.arange(N))
    assert numpy.array_equal(pf.particles.get(), before)


def test_update_resample_indices_follow_device_cumsum(g, noise_pdfs):
    """After a real update the device scans exp(loglik - max) in fixed point; the indices must be
    the reference comparison applied to exactly those cumulative weights, and the quantised
    weights must agree with the oracle's likelihood."""
    state, meas = noise_pdfs
    N = 20000
    rng = numpy.random.default_rng(5)
    pf = make_pf(g, N, seed=3)
    u = numpy.array([0.06, 0.2])
    pf.predict(u, 0.1)
    z = consistent_measurement(u, 0.1, rng)
    x = pf.particles.get().copy()
    pf.update(u, z)
    c, total = pf.cumulative_weights()
    o = particle.ParticleFilterOracle(N, None, state, meas, particles=x)
    ll = o.log_likelihood(u, z)
    wq = numpy.diff(numpy.concatenate([[0], c.astype(numpy.float64)]))
    wn = numpy.exp(ll - ll.max())
    big = wn > 1e-6
    assert numpy.allclose(wq[big] / wq.max(), wn[big], rtol=2e-5)
    r = 0.371
    idx = pf.resample(r=r, return_index=True).cpu().numpy()
    assert numpy.array_equal(idx, expected_indices_from_cumsum(c, r))
    assert numpy.array_equal(pf.particles.get(), x[idx])
    # the oracle resampling its own float64 weights picks the same ancestors except where a
    # threshold falls within the float32 log-weight rounding (~1e-6 of the total) of a boundary:
    # expected fraction ~ N * 1e-6, and a mismatch moves to a neighbouring ancestor
    o.update(u, z)
    oidx = o.resample(r=r)
    assert (idx != oidx).mean() < 0.05
    assert numpy.abs(numpy.sort(x[idx][:, 0]) - numpy.sort(x[oidx][:, 0])).max() < 0.05


def test_philox_noise_matches_specification(g):
    """In-kernel Philox + Box-Muller draws against the numpy specification (oracle/philox.py)."""
    N = 4096
    seed = 0x1234567
    pf = make_pf(g, N, seed=seed)
    x_before = pf.particles.get().astype(numpy.float64)
    # initial draw: step 0xFFFFFFFF, x0 mixture
    from oracle import bioreactor
    x0_ref, _ = philox.draw_mixture5(mixture.STATE_MEANS + bioreactor.X_STEADY[None, :], mixture.STATE_COVS,
                                     mixture.STATE_WEIGHTS, numpy.arange(N), 0xFFFFFFFF, 0, seed)
    sd = numpy.sqrt(numpy.diag(mixture.STATE_COVS[1]))
    assert (numpy.abs(x_before - x0_ref) <= 1e-4 * sd + 2 * numpy.spacing(numpy.float32(30.0))).all()
    # predict with u chosen so that the noise can be isolated: noise = x_after - (x + f(x))
    u = numpy.array([0.06, 0.2])
    pf.predict(u, 0.1)
    x_after = pf.particles.get().astype(numpy.float64)
    stepped = (x_before + bioreactor.increment(x_before.astype(numpy.float32), u, 0.1)).astype(numpy.float32)
    noise_ref, comp = philox.draw_mixture5_grouped(mixture.STATE_MEANS, mixture.STATE_COVS, mixture.STATE_WEIGHTS,
                                                   numpy.arange(N), 0, seed)
    err = numpy.abs(x_after - (stepped.astype(numpy.float64) + noise_ref))
    assert (err <= 1e-4 * sd + 6 * numpy.spacing(numpy.float32(30.0))).all()
    assert abs((comp == 0).mean() - 0.75) < 0.03


def test_philox_stream_is_independent_of_shape(g):
    """Counter-based noise: a particle's draw depends on (seed, index, step) only."""
    a = make_pf(g, 1000, seed=9, particles=numpy.zeros((1000, 5), numpy.float32) + 1.0)
    b = make_pf(g, 3000, seed=9, particles=numpy.zeros((3000, 5), numpy.float32) + 1.0)
    u = numpy.array([0.05, 0.1])
    a.predict(u, 0.1)
    b.predict(u, 0.1)
    assert numpy.array_equal(a.particles.get(), b.particles.get()[:1000])
    a.predict(u, 0.1)
    b.predict(u, 0.1)
    assert numpy.array_equal(a.particles.get(), b.particles.get()[:1000])


def test_noise_statistics_large(g):
    N = 1 << 20
    pf = make_pf(g, N, seed=11, particles=numpy.zeros((N, 5), numpy.float32))
    x0 = pf.particles.get().astype(numpy.float64)
    pf.predict(numpy.array([0.0, 0.0]), 0.0)          # dt = 0: x += noise only
    d = pf.particles.get().astype(numpy.float64) - x0
    var = 0.75 * numpy.diag(mixture.STATE_COVS[0]) + 0.25 * numpy.diag(mixture.STATE_COVS[1])
    assert numpy.allclose(d.var(axis=0), var, rtol=0.02)
    assert (numpy.abs(d.mean(axis=0)) < 5 * numpy.sqrt(var / N)).all()
    kurt = (d[:, 0] ** 4).mean() / var[0] ** 2           # mixture kurtosis 3 * E[s^4] / E[s^2]^2
    expect = 3 * (0.75 * 1e-8 + 0.25 * 1e-6) / (0.75 * 1e-4 + 0.25 * 1e-3) ** 2
    assert kurt == pytest.approx(expect, rel=0.05)


def test_multi_step_vs_oracle_with_host_noise(g, noise_pdfs):
    """10 predict/update/resample cycles at N = 2^16 fed identical noise and offsets: state means
    and covariances track the oracle (the tolerance widens with steps)."""
    state, meas = noise_pdfs
    N = 1 << 16
    rng = numpy.random.default_rng(17)
    x0 = mixture.benchmark_x0(particle.bioreactor.X_STEADY)
    numpy.random.seed(4)
    init = x0.draw(N)
    pf = make_pf(g, N, particles=init)
    o = particle.ParticleFilterOracle(N, None, state, meas, particles=init)
    for c in range(10):
        u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
        noise = (rng.normal(size=(N, 5)) * numpy.sqrt(numpy.diag(mixture.STATE_COVS[0]))).astype(numpy.float32)
        z = consistent_measurement(u, 0.1, rng)
        pf.predict(u, 0.1, noise=noise)
        o.predict(u, 0.1, noise=noise)
        assert ulp32(pf.particles.get().astype(numpy.float64) - o.particles, o.particles).max() <= 4.0 * (c + 1)
        pf.update(u, z)
        o.update(u, z)
        wn = o.weights.astype(numpy.float64) / o.weights.sum(dtype=numpy.float64)
        assert numpy.allclose(pf.point_estimate(normalised=True), wn @ o.particles.astype(numpy.float64),
                              rtol=2e-6, atol=1e-6)
        r = float(rng.random())
        pf.resample(r=r)
        o.resample(r=r)
        assert numpy.allclose(pf.point_estimate(), o.point_estimate(), rtol=1e-4, atol=1e-3)   # ~1% of the posterior spread
        # re-seed the oracle with the device state: ancestors may differ at a handful of boundaries
        o.particles = pf.particles.get().copy()


def test_weights_setter_scaling_and_getter(g):
    pf = make_pf(g, 100)
    w = numpy.random.default_rng(1).random(100)
    pf.weights = w
    pf.weights /= 1e3                       # gsf_power.py:85 style in-place scaling
    assert numpy.allclose(pf.weights.get(), w / 1e3, rtol=1e-15)
    est = pf.point_estimate()
    assert numpy.allclose(est, (w / 1e3) @ pf.particles.get().astype(numpy.float64), rtol=1e-9)


@pytest.mark.parametrize("N", [1 << 20, 1 << 24])
def test_full_size_invariants(g, N):
    """BASELINE.json sizes: one full cycle, checked through size-independent properties."""
    rng = numpy.random.default_rng(N % 1000)
    pf = make_pf(g, N, seed=5)
    u = numpy.array([0.06, 0.2])
    pf.predict(u, 1.0)
    z = consistent_measurement(u, 1.0, rng)
    pf.update(u, z)
    est_w = pf.point_estimate(normalised=True)
    c, total = pf.cumulative_weights()
    assert (numpy.diff(c.astype(numpy.int64)) >= 0).all() and 0 < total < 2 ** 62
    x = pf.particles.get().copy()
    r = float(rng.random())
    idx = pf.resample(r=r, return_index=True).cpu().numpy()
    assert (numpy.diff(idx) >= 0).all() and idx.min() >= 0 and idx.max() < N          # sortedness
    assert numpy.array_equal(idx, expected_indices_from_cumsum(c, r))                # exact comparison
    assert numpy.array_equal(pf.particles.get(), x[idx])                             # gather
    counts = numpy.bincount(idx, minlength=N)
    wq = numpy.diff(numpy.concatenate([[0], c.astype(numpy.float64)])) / float(total)
    assert numpy.abs(counts - N * wq).max() <= 1.0 + 1e-6                            # systematic: |n_k - N w_k| < 1
    est_r = pf.point_estimate()
    sd = numpy.sqrt(numpy.maximum(pf.covariance_matrix(normalised=True).diagonal(), 1e-30))
    assert (numpy.abs(est_r - est_w) <= 6 * sd / numpy.sqrt(1000) + 1e-5).all()      # unbiased resample
    idx2 = pf.resample(r=0.25, return_index=True).cpu().numpy()                      # idempotence on uniform weights
    assert numpy.array_equal(idx2, numpy.arange(N))
