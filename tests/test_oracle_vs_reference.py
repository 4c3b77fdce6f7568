"""Live cross-check of the oracle against the reference's own CPU classes, on seeds the golden vectors
do not contain.  Runs only where /root/reference exists (the build container); in a subprocess,
because importing the reference registers stub ``cupy`` / ``osqp`` modules and top-level packages
named ``filter`` and ``model``.  CPU only."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import ref_loader

SCRIPT = r'''
import sys, warnings
import numpy
sys.path.insert(0, %(root)r)
warnings.simplefilter("ignore")
from oracle import ref_loader, bioreactor, mixture, particle, gs_ukf
R = ref_loader.load()
MGS = R.gaussian_sum_dist.MultivariateGaussianSum
f_ref, g_ref = R.model.Bioreactor.homeostatic_DEs, R.model.Bioreactor.static_outputs

def ulp32(err, ref):
    scale = numpy.spacing(numpy.maximum(numpy.abs(ref), 1).astype(numpy.float32)).astype(numpy.float64)
    return numpy.abs(err) / scale

def ref_noise(shift=None):
    means = numpy.zeros((2, 5)) if shift is None else numpy.zeros((2, 5)) + shift[None, :]
    return (MGS(means, mixture.STATE_COVS, numpy.array([0.75, 0.25]), library=numpy),
            MGS(mixture.MEAS_MEANS, mixture.MEAS_COVS, numpy.array([0.85, 0.15]), library=numpy))

X_SS = R.model.Bioreactor.find_SS(numpy.array([0.06, 0.2]), numpy.array([260 / 180, 640 / 24.6, 1000 / 116, 0, 0]))
assert numpy.allclose(X_SS, bioreactor.X_STEADY, rtol=0, atol=1e-12)
ostate, omeas = mixture.benchmark_noise()
seed = %(seed)d
rng = numpy.random.default_rng(seed)

# ---- particle filter (filter/particle.py:43-114) ------------------------------------------------
state, meas = ref_noise()
x0, _ = ref_noise(X_SS)
N, dt = 300, 0.5
numpy.random.seed(seed)
pf = R.filter.ParticleFilter(f_ref, g_ref, N, x0, state, meas)
o = particle.ParticleFilterOracle(N, None, ostate, omeas, particles=pf.particles.copy())
x_true = X_SS.copy()
for c in range(3):
    u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
    x_true = x_true + numpy.asarray(f_ref(x_true, u, dt), dtype=numpy.float64)
    z = numpy.asarray(g_ref(x_true, u)) + rng.normal(size=2) * numpy.array([0.2, 0.25])
    numpy.random.seed(seed + 10 + c); noise = state.draw(N).copy()
    numpy.random.seed(seed + 10 + c); pf.predict(u, dt)
    o.predict(u, dt, noise=noise)
    assert ulp32(o.particles.astype(numpy.float64) - pf.particles, pf.particles).max() <= 6.0, "pf predict"
    o.particles = pf.particles.copy()
    pf.update(u, z); o.update(u, z)
    assert numpy.allclose(o.weights, pf.weights, rtol=2e-5, atol=0), "pf update"
    assert numpy.allclose(o.point_estimate(), pf.point_estimate(), rtol=1e-5), "pf estimate"
    assert abs(o.point_covariance() / pf.point_covariance() - 1) < 1e-4, "pf covariance"
    o.weights = pf.weights.copy()
    numpy.random.seed(seed + 50 + c); r = numpy.random.rand()
    numpy.random.seed(seed + 50 + c); pf.resample()
    o.resample(r=r, loop=bool(c %% 2))
    assert numpy.array_equal(o.particles, pf.particles), "pf resample"
    assert numpy.array_equal(o.weights, pf.weights), "pf weights after resample"

# ---- GS-UKF (filter/gs_ukf.py:45-183) -----------------------------------------------------------
M = 40
numpy.random.seed(seed + 1)
gf = R.filter.GaussianSumUnscentedKalmanFilter(f_ref, g_ref, M, x0, state, meas)
og = gs_ukf.GSUKFOracle(M, None, ostate, omeas, means=gf.means.copy())
for c in range(3):
    u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
    z = numpy.asarray(g_ref(X_SS, u)) + rng.normal(size=2) * numpy.array([0.2, 0.25])
    numpy.random.seed(seed + 100 + c); noise = state.draw((M, 11)).copy()
    numpy.random.seed(seed + 100 + c); gf.predict(u, 0.1)
    og.predict(u, 0.1, noise=noise)
    assert ulp32(og.means.astype(numpy.float64) - gf.means, gf.means).max() <= 8.0, "gsukf predict means"
    assert numpy.abs(og.covariances - gf.covariances).max() <= 2e-5 * numpy.abs(gf.covariances).max(), "gsukf predict cov"
    og.means, og.covariances = gf.means.copy(), gf.covariances.copy()
    gf.update(u, z); og.update(u, z)
    assert ulp32(og.means.astype(numpy.float64) - gf.means, gf.means).max() <= 8.0, "gsukf update means"
    assert numpy.abs(og.covariances - gf.covariances).max() <= 2e-5 * numpy.abs(gf.covariances).max(), "gsukf update cov"
    assert numpy.allclose(og.weights, gf.weights, rtol=5e-5, atol=0), "gsukf weights"
    og.means, og.covariances, og.weights = gf.means.copy(), gf.covariances.copy(), gf.weights.copy()
    numpy.random.seed(seed + 150 + c); r = numpy.random.rand()
    numpy.random.seed(seed + 150 + c); gf.resample()
    og.resample(r=r)
    assert numpy.array_equal(og.means, gf.means) and numpy.array_equal(og.covariances, gf.covariances), "gsukf resample"
print("LIVE-OK")
'''


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("seed", [2024, 31337])
def test_oracle_matches_the_reference_live(seed):
    res = subprocess.run([sys.executable, "-W", "ignore", "-c", SCRIPT % {"root": ROOT, "seed": seed}],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "LIVE-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
