"""-m gpu parity tests of the GS-UKF path against the oracle and the reference's golden vectors."""
import numpy
import pytest

from conftest import golden, ulp32
from gpu_common import consistent_measurement, expected_indices_from_cumsum, make_gsf
from oracle import gs_ukf, mixture, philox

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gpu_se_b200
    return gpu_se_b200


def cov_close(a, b, rel):
    return numpy.abs(a - b).max() <= rel * numpy.abs(b).max()


TRIL = numpy.tril_indices(5)


def lower(c):
    """The device stores the lower triangle of each covariance (the reference stores the full 5x5,
    whose two halves differ in the last float32 bit); bit-exact comparisons use that triangle."""
    return numpy.asarray(c)[:, TRIL[0], TRIL[1]]


@pytest.mark.parametrize("name", ["gsukf_n64.npz", "gsukf_n256_dt1.npz"])
def test_golden_cycles(g, noise_pdfs, name):
    state, meas = noise_pdfs
    gv = golden(name)
    N, dt = int(gv["N"]), float(gv["dt"])
    gf = make_gsf(g, N, means=gv["means0"])
    o = gs_ukf.GSUKFOracle(N, None, state, meas, means=gv["means0"])
    assert numpy.array_equal(gf._w_sigma, gv["w_sigma"])
    assert numpy.array_equal(gf.covariances.get(), gv["covariances0"])
    assert numpy.array_equal(gf.covariances.get(), gf.covariances.get().swapaxes(1, 2))
    assert numpy.abs(gf._get_sigma_points().get() - gv["sigmas0"]).max() <= 2e-7
    for c in range(int(gv["n_cycles"])):
        u, z, noise = gv["u_%d" % c], gv["z_%d" % c], gv["noise_%d" % c]
        gf.predict(u, dt, noise=noise)
        o.predict(u, dt, noise=noise)
        rm, rc = gv["means_pred_%d" % c], gv["covs_pred_%d" % c]
        gm, gc = gf.means.get(), gf.covariances.get()
        assert ulp32(gm.astype(numpy.float64) - o.means, o.means).max() <= 4.0          # vs oracle
        assert cov_close(gc, o.covariances, 5e-6)
        assert ulp32(gm.astype(numpy.float64) - rm, rm).max() <= 8.0                    # vs reference
        assert cov_close(gc, rc, 1e-5)
        gf.means, gf.covariances = rm, rc
        o.means, o.covariances = rm.copy(), rc.copy()

        w_before = o.weights.astype(numpy.float64).copy()
        gf.update(u, z)
        o.update(u, z)
        rm, rc, rw = gv["means_upd_%d" % c], gv["covs_upd_%d" % c], gv["weights_upd_%d" % c].astype(numpy.float64)
        gm, gc, gw = gf.means.get(), gf.covariances.get(), gf.weights.get()
        assert ulp32(gm.astype(numpy.float64) - o.means, o.means).max() <= 4.0
        assert cov_close(gc, o.covariances, 5e-6)
        assert ulp32(gm.astype(numpy.float64) - rm, rm).max() <= 8.0
        assert cov_close(gc, rc, 1e-5)
        ow = o.weights.astype(numpy.float64)
        assert (numpy.abs(gw - ow) <= 2e-4 * ow).all()          # weights amplify few-ulp mean differences
        assert (numpy.abs(gw - rw) <= 2e-4 * rw).all()
        gf.means, gf.covariances = rm, rc
        gf.weights = gv["weights_upd_%d" % c]
        o.means, o.covariances, o.weights = rm.copy(), rc.copy(), gv["weights_upd_%d" % c].copy()
        assert numpy.allclose(gf.point_estimate(), gv["est_upd_%d" % c], rtol=1e-6)
        assert gf.point_covariance() == pytest.approx(float(gv["cov_upd_%d" % c]), rel=1e-5)
        r = float(gv["r_%d" % c])
        idx = gf.resample(r=r, return_index=True).cpu().numpy()
        assert numpy.array_equal(idx, o.resample(r=r))
        assert numpy.array_equal(gf.means.get(), gv["means_res_%d" % c])
        assert numpy.array_equal(lower(gf.covariances.get()), lower(gv["covs_res_%d" % c]))
        assert cov_close(gf.covariances.get(), gv["covs_res_%d" % c], 1e-6)
        assert numpy.allclose(gf.point_estimate(), gv["est_res_%d" % c], rtol=1e-6)
        assert gf.point_covariance() == pytest.approx(float(gv["cov_res_%d" % c]), rel=1e-5)


def test_same_cpu_gpu_like_the_reference_test(g, noise_pdfs):
    """The reference's only cross-implementation check (tests/GSUKF_test.py:48-99): with
    deterministic noise, after update and after resample the averages of the signed differences of
    means / covariances / normalised weights vanish (1e-7 / 1e-10 / 1e-7)."""
    state, meas = noise_pdfs
    N = 7
    rng = numpy.random.default_rng(0)
    from oracle import bioreactor
    # like the reference's double, the cached stream is a draw of x0 (rows near the steady state)
    stream = (bioreactor.X_STEADY[None, :] + rng.normal(size=(800, 5)) * 1e-2).astype(numpy.float32).ravel()
    g.DeterministicGaussianSum.set_stream(stream)
    dstate = g.DeterministicGaussianSum(mixture.STATE_MEANS, mixture.STATE_COVS, mixture.STATE_WEIGHTS)
    dx0 = g.DeterministicGaussianSum(mixture.STATE_MEANS + bioreactor.X_STEADY[None, :], mixture.STATE_COVS,
                                     mixture.STATE_WEIGHTS)
    dmeas = g.MultivariateGaussianSum(mixture.MEAS_MEANS, mixture.MEAS_COVS, mixture.MEAS_WEIGHTS)
    pgf = g.ParallelGaussianSumUnscentedKalmanFilter(g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs, N,
                                                     dx0, dstate, dmeas)
    ostate = mixture.FixedNoise(mixture.STATE_MEANS, mixture.STATE_COVS, mixture.STATE_WEIGHTS, stream)
    ox0 = mixture.FixedNoise(mixture.STATE_MEANS + bioreactor.X_STEADY[None, :], mixture.STATE_COVS,
                             mixture.STATE_WEIGHTS, stream)
    gf = gs_ukf.GSUKFOracle(N, ox0, ostate, meas)
    assert numpy.array_equal(pgf.means.get(), gf.means)

    def same():
        assert numpy.average(pgf.means.get() - gf.means) == pytest.approx(0, abs=1e-7)
        assert numpy.average(pgf.covariances.get() - gf.covariances) == pytest.approx(0, abs=1e-10)
        a = pgf.weights.get()
        b = gf.weights.astype(numpy.float64)
        assert numpy.average(a / a.sum() - b / b.sum()) == pytest.approx(0, abs=1e-7)

    u = numpy.array([0.06, 0.2])
    same()
    pgf.predict(u, 0.1)       # the reference leaves this comparison commented out (:87-91)
    gf.predict(u, 0.1)
    assert ulp32(pgf.means.get().astype(numpy.float64) - gf.means, gf.means).max() <= 4.0
    gf.means, gf.covariances = pgf.means.get().copy(), pgf.covariances.get().copy()
    z = bioreactor.outputs(gf.means.astype(numpy.float64).mean(axis=0)) + numpy.array([0.1, -0.2])
    gf.update(u, z)
    pgf.update(u, z)
    same()
    numpy.random.seed(3)
    pgf.resample()
    numpy.random.seed(3)
    gf.resample()
    same()


def test_philox_sigma_noise_matches_specification(g):
    """Every sigma point gets an independent draw from the component's (index, step) stream (gs_ukf.py:99)."""
    N = 512
    seed = 77
    gf = make_gsf(g, N, seed=seed)
    m0 = gf.means.get().astype(numpy.float64)
    sig = gf._get_sigma_points().get()
    u = numpy.array([0.06, 0.2])
    gf.predict(u, 0.1)
    from oracle import bioreactor
    stepped = (sig.astype(numpy.float64) + bioreactor.increment(sig, u, 0.1)).astype(numpy.float32).astype(numpy.float64)
    noise = philox.draw_mixture5_sigma(mixture.STATE_MEANS, mixture.STATE_COVS, mixture.STATE_WEIGHTS, numpy.arange(N), 0, seed)
    sg = (stepped + noise).astype(numpy.float32).astype(numpy.float64)
    w = gs_ukf.sigma_weights().astype(numpy.float64)
    mean = numpy.einsum("nsj,s->nj", sg, w) / w.sum()
    d = sg - mean.astype(numpy.float32)[:, None, :]
    cov = numpy.einsum("nsi,nsj,s->nij", d, d, w)
    assert numpy.abs(gf.means.get() - mean).max() <= 2e-5
    assert cov_close(gf.covariances.get(), cov, 2e-3)
    assert numpy.abs(m0 - mean).max() < 1.0


@pytest.mark.parametrize("N", [1, 5, 127, 129, 1 << 16])
def test_cycle_invariants(g, N):
    rng = numpy.random.default_rng(N)
    gf = make_gsf(g, N, seed=2)
    u = numpy.array([0.06, 0.2])
    for _ in range(2):
        gf.predict(u, 0.1)
        z = consistent_measurement(u, 0.1, rng)
        gf.update(u, z)
        covs = gf.covariances.get()
        assert numpy.isfinite(covs).all() and numpy.isfinite(gf.means.get()).all()
        eig = numpy.linalg.eigvalsh(covs.astype(numpy.float64))
        assert eig.min() > -1e-6                               # stays (numerically) positive semi-definite
        c, total = gf.cumulative_weights()
        m, P = gf.means.get().copy(), covs.copy()
        r = float(rng.random())
        idx = gf.resample(r=r, return_index=True).cpu().numpy()
        assert numpy.array_equal(idx, expected_indices_from_cumsum(c, r))
        assert numpy.array_equal(gf.means.get(), m[idx]) and numpy.array_equal(gf.covariances.get(), P[idx])
        est = gf.point_estimate()
        assert numpy.allclose(est, gf.means.get().astype(numpy.float64).mean(axis=0), rtol=1e-6)
        assert gf.point_covariance() > 0
