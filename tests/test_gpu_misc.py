"""-m gpu: pieces of the path the round-1 review found untested -- the stand-alone mixture pdf / sampler
(gse_mixture_pdf, gse_mixture_draw for any dimension), Euler sub-steps (n_sub > 1), reference-style mixture objects
handed to the drop-in classes, device-side error reporting, and a filter on a device that is not the current one."""
import numpy
import pytest

from conftest import golden, ulp32
from gpu_common import consistent_measurement, make_gsf, make_pdfs, make_pf
from oracle import bioreactor, gs_ukf, mixture, particle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gpu_se_b200
    return gpu_se_b200


def test_mixture_pdf_against_the_reference(g, noise_pdfs):
    """gse_mixture_pdf (MultivariateGaussianSum.pdf, :39-63) against the reference's own values (mixture.npz) and
    against the float64 oracle evaluated at the float32 points the device receives."""
    ostate, omeas = noise_pdfs
    x0, state, meas = make_pdfs(g)
    mx = golden("mixture.npz")
    for pdf, opdf, pts, ref in ((meas, omeas, mx["e"], mx["pdf_meas"]), (state, ostate, mx["xs"], mx["pdf_state"])):
        got = pdf.pdf(pts)
        assert got.shape == ref.shape and got.dtype == numpy.float64
        p32 = pts.astype(numpy.float32).astype(numpy.float64)
        want = opdf.pdf(p32)
        ok = want > 1e-300
        assert (numpy.abs(got[ok] - want[ok]) <= 1e-6 * want[ok]).all()          # same inputs: float32 constants only
        nz = ref > 1e-290
        # the reference saw the float64 points: rounding them to float32 moves log pdf by |x / sigma^2| * 6e-8 |x|
        lg, lr = numpy.log(got[nz]), numpy.log(ref[nz])
        assert (numpy.abs(lg - lr) <= 1e-5 * numpy.maximum(1.0, numpy.abs(lr))).all()
        assert (got[~nz] < 1e-280).all()
        import torch
        dev = pdf.pdf(torch.as_tensor(pts, dtype=torch.float32, device="cuda"))      # device in, device out
        assert numpy.array_equal(dev.get(), got)


def test_mixture_sampler_any_dimension(g):
    """MultivariateGaussianSum.draw (:65-97) on the device for the 2-d measurement mixture (non-diagonal second
    component) and a 3-d one: moments of 2^20 draws against the mixture's."""
    rng = numpy.random.default_rng(0)
    A = rng.normal(size=(3, 3))
    cases = [(mixture.MEAS_MEANS, mixture.MEAS_COVS, mixture.MEAS_WEIGHTS),
             (numpy.array([[0.0, 1.0, -1.0], [2.0, 0.0, 0.5]]),
              numpy.array([A @ A.T + 0.1 * numpy.eye(3), numpy.diag([0.5, 2.0, 0.01])]), numpy.array([0.3, 0.7]))]
    n = 1 << 20
    for means, covs, w in cases:
        w = numpy.asarray(w, dtype=numpy.float64) / numpy.sum(w)
        pdf = g.MultivariateGaussianSum(means, covs, w, seed=5)
        x = pdf.draw(n).get().astype(numpy.float64)
        assert x.shape == (n, means.shape[1])
        mean = (w[:, None] * means).sum(0)
        cov = sum(w[d] * (covs[d] + numpy.outer(means[d], means[d])) for d in range(len(w))) - numpy.outer(mean, mean)
        sd = numpy.sqrt(numpy.diag(cov))
        assert (numpy.abs(x.mean(0) - mean) <= 6 * sd / numpy.sqrt(n)).all()
        got = numpy.cov(x.T)
        assert (numpy.abs(got - cov) <= 0.02 * numpy.outer(sd, sd)).all()
        again = pdf.draw(16).get()
        assert not numpy.array_equal(again, x[:16].astype(numpy.float32))            # the draw counter advances
        same = g.MultivariateGaussianSum(means, covs, w, seed=5).draw(n).get()
        assert numpy.array_equal(same.astype(numpy.float64), x)                      # a pure function of (seed, draw, row)
        assert pdf.draw((3, 4)).get().shape == (3, 4, means.shape[1])


@pytest.mark.parametrize("n_sub", [2, 4])
def test_euler_sub_steps(g, noise_pdfs, n_sub):
    """north_star: "Euler sub-steps across the sample period".  n_sub explicit-Euler steps of dt / n_sub, each the
    reference's increment (BioreactorModel.py:170-231), then the noise: against the same loop in the float64 oracle."""
    N = 3000
    rng = numpy.random.default_rng(n_sub)
    numpy.random.seed(n_sub)
    init = mixture.benchmark_x0(bioreactor.X_STEADY).draw(N)
    init[: N // 2] += (rng.normal(size=(N // 2, 5)) * [0.5, 1.0, 0.5, 0.01, 1.0]).astype(numpy.float32)   # away from steady state
    noise = (rng.normal(size=(N, 5)) * numpy.sqrt(numpy.diag(mixture.STATE_COVS[0]))).astype(numpy.float32)
    u, dt = numpy.array([0.07, 0.15]), 1.0
    pf = make_pf(g, N, particles=init, n_sub=n_sub)
    pf.predict(u, dt, noise=noise)
    x = init.astype(numpy.float64)
    for _ in range(n_sub):
        x = (x + bioreactor.increment(x.astype(numpy.float32), u, dt / n_sub)).astype(numpy.float32).astype(numpy.float64)
    want = (x.astype(numpy.float32) + noise).astype(numpy.float64)
    got = pf.particles.get().astype(numpy.float64)
    assert ulp32(got - want, want).max() <= 4.0 * n_sub
    one = make_pf(g, N, particles=init)
    one.predict(u, dt, noise=noise)
    assert not numpy.array_equal(one.particles.get(), pf.particles.get())            # it does something
    # Philox mode: the noise of step t does not depend on n_sub
    a, b = make_pf(g, N, particles=init, n_sub=n_sub, seed=3), make_pf(g, N, particles=init, seed=3)
    a.predict(u, 1e-9)
    b.predict(u, 1e-9)
    assert numpy.allclose(a.particles.get(), b.particles.get(), rtol=0, atol=1e-6)


class _ReferenceStyleMixture:
    """What the reference's own MultivariateGaussianSum exposes (MultivariateGaussianSum.py:27-37): float32 means /
    covariances / weights and the float64 information only through ``_inverse_covariances``."""

    def __init__(self, means, covariances, weights):
        self.means = numpy.asarray(means, dtype=numpy.float32)
        self.covariances = numpy.asarray(covariances, dtype=numpy.float32)
        self.weights = numpy.asarray(weights, dtype=numpy.float32)
        self._inverse_covariances = numpy.linalg.inv(numpy.asarray(covariances, dtype=numpy.float64))
        self._Nd, self._Nx = self.means.shape


def test_reference_style_mixture_objects_are_accepted(g):
    """The drop-in classes take the reference's own mixture objects (duck-typed): same results as with this
    package's MultivariateGaussianSum, both filters."""
    N = 2048
    rng = numpy.random.default_rng(1)
    numpy.random.seed(1)
    init = mixture.benchmark_x0(bioreactor.X_STEADY).draw(N)
    u = numpy.array([0.06, 0.2])
    z = consistent_measurement(u, 0.5, rng)
    rstate = _ReferenceStyleMixture(mixture.STATE_MEANS, mixture.STATE_COVS, mixture.STATE_WEIGHTS)
    rmeas = _ReferenceStyleMixture(mixture.MEAS_MEANS, mixture.MEAS_COVS, mixture.MEAS_WEIGHTS)
    rx0 = _ReferenceStyleMixture(mixture.STATE_MEANS + bioreactor.X_STEADY[None, :], mixture.STATE_COVS, mixture.STATE_WEIGHTS)
    f, gg = g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs
    ours = make_pf(g, N, seed=4, particles=init)
    theirs = g.ParticleFilter(f, gg, N, rx0, rstate, rmeas, seed=4, particles=init)
    for pf in (ours, theirs):
        pf.predict(u, 0.5)
        pf.update(u, z)
    assert numpy.array_equal(ours.particles.get(), theirs.particles.get())
    wa, wb = ours.weights.get(), theirs.weights.get()
    assert (numpy.abs(wa - wb) <= 1e-6 * wa).all()             # inv(inv(C)) is C up to float64 rounding
    ia = ours.resample(r=0.3, return_index=True).cpu().numpy()
    ib = theirs.resample(r=0.3, return_index=True).cpu().numpy()
    assert (ia != ib).mean() < 1e-3
    # x0 drawn on the device from the reference-style object, GS-UKF construction through the same path
    drawn = g.ParticleFilter(f, gg, N, rx0, rstate, rmeas, seed=4)
    assert numpy.abs(drawn.particles.get().mean(0) - bioreactor.X_STEADY).max() < 0.01
    ga = make_gsf(g, 256, seed=4, means=init[:256])
    gb = g.GaussianSumUnscentedKalmanFilter(f, gg, 256, rx0, rstate, rmeas, seed=4, means=init[:256])
    for gf in (ga, gb):
        gf.predict(u, 0.5)
        gf.update(u, z)
    assert numpy.array_equal(ga.means.get(), gb.means.get())
    assert numpy.array_equal(ga.covariances.get(), gb.covariances.get())


def test_gsukf_rank_deficient_covariance_takes_the_retry_path(g, noise_pdfs):
    """numpy.linalg.cholesky raises on a positive SEMI-definite covariance and the reference retries with
    + 1e-10 I in float64 (gs_ukf.py:72-75): same here, against the oracle; no error is reported."""
    ostate, omeas = noise_pdfs
    N = 64
    numpy.random.seed(2)
    means0 = mixture.benchmark_x0(bioreactor.X_STEADY).draw(N)
    v = numpy.array([1.0, 0.0, 0.0, 0.0, 0.0]) * 1e-2
    cov = numpy.repeat(numpy.outer(v, v)[None], N, axis=0).astype(numpy.float32)      # rank one: zero pivots
    cov[N // 2:] = numpy.diag([1e-4, 0.0, 1e-3, 1e-3, 0.0]).astype(numpy.float32)      # rank three
    gf = make_gsf(g, N, means=means0)
    gf.covariances = cov
    o = gs_ukf.GSUKFOracle(N, None, ostate, omeas, means=means0)
    o.covariances = cov.copy()
    got = gf._get_sigma_points().get()
    want = o._get_sigma_points()
    assert numpy.isfinite(got).all()
    assert ulp32(got.astype(numpy.float64) - want, want).max() <= 1.0                  # sqrt(1e-10) = 1e-5 columns
    assert numpy.abs(got[:, 2, 1] - got[:, 0, 1]).min() > 0                            # the jitter column is there
    u = numpy.array([0.06, 0.2])
    noise = numpy.zeros((N, 11, 5), dtype=numpy.float32)
    gf.predict(u, 0.1, noise=noise)
    assert numpy.isfinite(gf.point_estimate()).all()                                    # synchronises, polls the error word


def test_gsukf_indefinite_covariance_raises_like_the_reference(g):
    """A covariance with a negative eigenvalue fails the retry as well: the reference raises LinAlgError
    (gs_ukf.py:75); here the kernel flags it and the next synchronising call raises."""
    N = 32
    gf = make_gsf(g, N)
    cov = numpy.repeat(numpy.diag([1e-4, 1e-7, -1e-3, 1e-3, 1e-7])[None], N, axis=0).astype(numpy.float32)
    gf.covariances = cov
    gf.predict(numpy.array([0.06, 0.2]), 0.1)
    with pytest.raises(numpy.linalg.LinAlgError):
        gf.point_estimate()
    # the error word was cleared: a healthy filter on the same device works
    ok = make_gsf(g, N)
    ok.predict(numpy.array([0.06, 0.2]), 0.1)
    assert numpy.isfinite(ok.point_estimate()).all()


def test_filter_on_a_device_that_is_not_current(g):
    """Every C-ABI entry selects the context's device and restores the caller's: a filter built with
    device=cuda:1 while cuda:0 is current runs, and the current device is unchanged afterwards."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    N = 5000
    u = numpy.array([0.06, 0.2])
    rng = numpy.random.default_rng(0)
    z = consistent_measurement(u, 0.5, rng)
    numpy.random.seed(0)
    init = mixture.benchmark_x0(bioreactor.X_STEADY).draw(N)
    res = []
    for dev in ("cuda:0", "cuda:1"):
        pf = make_pf(g, N, seed=7, particles=init, device=dev)
        pf.predict(u, 0.5)
        pf.update(u, z)
        idx = pf.resample(r=0.25, return_index=True).cpu().numpy()
        res.append((pf.particles.get(), idx, pf.point_estimate()))
        assert torch.cuda.current_device() == 0
        del pf
    assert numpy.array_equal(res[0][0], res[1][0]) and numpy.array_equal(res[0][1], res[1][1])
    assert numpy.array_equal(res[0][2], res[1][2])
