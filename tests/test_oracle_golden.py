"""The oracle (oracle/) against the golden vectors produced by the reference's own CPU classes
(tests/golden/make_golden.py).  CPU only.  Tolerances are the measured reference-vs-float64
deviations (the reference's arithmetic is a float32/float64 promotion accident)
with head-room, stated per assertion."""
import numpy
import pytest

from conftest import golden, ulp32
from oracle import bioreactor, gs_ukf, mixture, particle


def test_model_increment_and_outputs():
    m = golden("model_fg.npz")
    x = m["x"]
    assert numpy.allclose(m["x_ss"], bioreactor.X_STEADY, rtol=0, atol=1e-12)
    for k in range(len(m["dts"])):
        inc = bioreactor.increment(x, m["us"][k], m["dts"][k])
        # increments: the reference evaluates part of f in float32 -> <= 4 ulp32 of the STATE
        assert ulp32(inc - m["incs"][k], x).max() <= 4.0
        stored = (x.astype(numpy.float64) + inc).astype(numpy.float32)
        assert ulp32(stored.astype(numpy.float64) - m["stored"][k], x).max() <= 4.0
        # outputs are float32-rounded products in the reference: exact
        assert numpy.array_equal(bioreactor.outputs(x), m["ys"][k])
        for i in (0, 17, 40, 70, 100, 140, 170, 300):
            assert numpy.allclose(bioreactor.increment_scalar(x[i].astype(numpy.float64), m["us"][k], m["dts"][k]),
                                  bioreactor.increment(x[i], m["us"][k], m["dts"][k]), rtol=1e-13, atol=1e-15)


def test_mixture_pdf(noise_pdfs):
    state, meas = noise_pdfs
    mx = golden("mixture.npz")
    ref = mx["pdf_meas"]
    assert (numpy.abs(meas.pdf(mx["e"]) - ref) <= 5e-7 * ref).all()          # float32 constants in the reference
    ref = mx["pdf_state"]
    assert (numpy.abs(state.pdf(mx["xs"]) - ref) <= 5e-7 * ref).all()
    nz = mx["pdf_meas"] > 1e-290
    assert numpy.allclose(meas.logpdf(mx["e"])[nz], numpy.log(mx["pdf_meas"][nz]), rtol=0, atol=5e-7)
    assert numpy.allclose(meas._constants, mx["meas_constants"], rtol=3e-7)
    assert numpy.allclose(state._constants, mx["state_constants"], rtol=3e-7)


def test_mixture_draw_is_the_reference_call_sequence(noise_pdfs):
    state, meas = noise_pdfs
    mx = golden("mixture.npz")
    numpy.random.seed(11)
    assert numpy.array_equal(state.draw(1000), mx["draw_state_1000_seed11"])
    numpy.random.seed(12)
    assert numpy.array_equal(state.draw((20, 11)), mx["draw_state_20x11_seed12"])
    numpy.random.seed(13)
    assert numpy.array_equal(meas.draw(64), mx["draw_meas_64_seed13"])


def test_fixed_noise_double(noise_pdfs):
    state, _ = noise_pdfs
    stream = numpy.arange(200, dtype=numpy.float32)
    fx = mixture.FixedNoise(state.means, state.covariances, state.weights, stream)
    a, b = fx.draw(10), fx.draw((2, 5))
    assert numpy.sum(a - b.reshape(10, 5)) == 0          # DeterministicGaussianSum_test.py:16-19
    assert a.shape == (10, 5) and b.shape == (2, 5, 5)


@pytest.mark.parametrize("name", ["pf_n256.npz", "pf_n1024_dt1.npz"])
def test_particle_filter_cycles(noise_pdfs, name):
    state, meas = noise_pdfs
    g = golden(name)
    o = particle.ParticleFilterOracle(int(g["N"]), None, state, meas, particles=g["particles0"])
    assert numpy.array_equal(o.weights, g["weights0"]) and o.weights.dtype == numpy.float32
    dt = float(g["dt"])
    for c in range(int(g["n_cycles"])):
        u, z = g["u_%d" % c], g["z_%d" % c]
        o.predict(u, dt, noise=g["noise_%d" % c])
        ref = g["particles_pred_%d" % c]
        assert ulp32(o.particles.astype(numpy.float64) - ref, ref).max() <= 4.0     # SURVEY §7 bound
        o.particles = ref.copy()                     # re-seed so that errors do not compound
        assert numpy.allclose(o.point_estimate(), g["est_pred_%d" % c], rtol=1e-6)
        assert o.point_covariance() == pytest.approx(float(g["cov_pred_%d" % c]), rel=1e-5)
        ll = o.log_likelihood(u, z)
        w_before = o.weights.astype(numpy.float64)
        o.update(u, z)
        rw = g["weights_upd_%d" % c]
        assert o.weights.dtype == rw.dtype           # float32 first cycle, float64 afterwards (Q3)
        assert (numpy.abs(o.weights - rw) <= 1e-6 * rw).all()
        nz = rw > 0
        assert numpy.allclose((ll + numpy.log(w_before))[nz], numpy.log(rw[nz].astype(numpy.float64)), rtol=0, atol=1e-6)
        o.weights = rw.copy()
        assert numpy.allclose(o.point_estimate(), g["est_upd_%d" % c], rtol=1e-6)
        o.resample(r=float(g["r_%d" % c]))
        assert numpy.array_equal(o.particles, g["particles_res_%d" % c])           # bit-exact gather
        assert numpy.array_equal(o.weights, g["weights_res_%d" % c]) and o.weights.dtype == numpy.float64
        assert numpy.allclose(o.point_estimate(), g["est_res_%d" % c], rtol=1e-7)
        assert o.point_covariance() == pytest.approx(float(g["cov_res_%d" % c]), rel=1e-6)


def test_particle_loop_port_equals_vectorised(noise_pdfs):
    """The loop-faithful port (timed as the CPU baseline) and the vectorised oracle agree."""
    state, meas = noise_pdfs
    g = golden("pf_n256.npz")
    a = particle.ParticleFilterOracle(256, None, state, meas, particles=g["particles0"])
    b = particle.ParticleFilterOracle(256, None, state, meas, particles=g["particles0"])
    u, z, noise = g["u_0"], g["z_0"], g["noise_0"]
    a.predict(u, 0.1, noise=noise)
    b.predict_loop(u, 0.1, noise=noise)
    assert ulp32(a.particles.astype(numpy.float64) - b.particles, a.particles).max() <= 2.0
    assert ulp32(b.particles.astype(numpy.float64) - g["particles_pred_0"], b.particles).max() <= 2.0
    b.particles = a.particles.copy()
    a.update(u, z)
    b.update_loop(u, z)
    assert numpy.allclose(a.weights, b.weights, rtol=1e-6, atol=0)
    ia = a.resample(r=0.3)
    ib = b.resample(r=0.3, loop=True)
    assert numpy.array_equal(ia, ib)


def test_resample_indices_bit_exact():
    r = golden("resample.npz")
    for t in ("a", "b", "c", "skew", "dyadic"):
        w = r["weights_" + t]
        c = numpy.cumsum(w)
        c /= c[-1]
        assert numpy.array_equal(particle.systematic_indices(c, float(r["r_" + t])), r["idx_" + t]), t
        assert numpy.array_equal(particle.systematic_indices_loop(c, float(r["r_" + t])), r["idx_" + t]), t


def test_nicely_kernel_semantics():
    """The reference's hand-written CUDA kernel, executed by numba's simulator (golden), equals
    searchsorted(side='right'); the CPU sweep is side='left'; they agree away from exact ties."""
    n = golden("nicely_cudasim.npz")
    for t in ("a", "b"):
        c, r = n["cumsum_" + t], n["r_" + t]
        assert numpy.array_equal(particle.nicely_indices(c, r), n["idx_" + t])
        assert numpy.array_equal(particle.systematic_indices(c, r), n["idx_" + t])
    c = numpy.array([0.25, 0.5, 0.75, 1.0])
    assert particle.systematic_indices(c, 0.0).tolist() == [0, 0, 1, 2]
    assert particle.nicely_indices(c, 0.0).tolist() == [0, 1, 2, 3]


@pytest.mark.parametrize("name", ["gsukf_n64.npz", "gsukf_n256_dt1.npz"])
def test_gsukf_cycles(noise_pdfs, name):
    state, meas = noise_pdfs
    g = golden(name)
    o = gs_ukf.GSUKFOracle(int(g["N"]), None, state, meas, means=g["means0"])
    assert numpy.array_equal(o._w_sigma, g["w_sigma"])
    assert numpy.array_equal(o.covariances, g["covariances0"])
    assert numpy.abs(o._get_sigma_points() - g["sigmas0"]).max() <= 1e-7
    dt = float(g["dt"])
    for c in range(int(g["n_cycles"])):
        u, z = g["u_%d" % c], g["z_%d" % c]
        o.predict(u, dt, noise=g["noise_%d" % c])
        rm, rc = g["means_pred_%d" % c], g["covs_pred_%d" % c]
        assert ulp32(o.means.astype(numpy.float64) - rm, rm).max() <= 8.0
        assert numpy.abs(o.covariances - rc).max() <= 5e-6 * numpy.abs(rc).max()
        o.means, o.covariances = rm.copy(), rc.copy()
        o.update(u, z)
        rm, rc, rw = g["means_upd_%d" % c], g["covs_upd_%d" % c], g["weights_upd_%d" % c]
        assert ulp32(o.means.astype(numpy.float64) - rm, rm).max() <= 8.0
        assert numpy.abs(o.covariances - rc).max() <= 5e-6 * numpy.abs(rc).max()
        assert (numpy.abs(o.weights - rw) <= 1e-4 * rw).all()   # weights amplify the 4-ulp mean differences
        o.means, o.covariances, o.weights = rm.copy(), rc.copy(), rw.copy()
        assert numpy.allclose(o.point_estimate(), g["est_upd_%d" % c], rtol=1e-6)
        assert o.point_covariance() == pytest.approx(float(g["cov_upd_%d" % c]), rel=1e-6)
        o.resample(r=float(g["r_%d" % c]))
        assert numpy.array_equal(o.means, g["means_res_%d" % c])
        assert numpy.array_equal(o.covariances, g["covs_res_%d" % c])
        assert numpy.allclose(o.point_estimate(), g["est_res_%d" % c], rtol=1e-7)
        assert o.point_covariance() == pytest.approx(float(g["cov_res_%d" % c]), rel=1e-6)
