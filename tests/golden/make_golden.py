#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own CPU classes.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no golden vectors of its own, so these are produced by importing
its unmodified numpy code paths (oracle/ref_loader.py registers empty ``cupy``/``osqp`` stubs)
under numpy %(numpy)s.  Inputs are seeded; noise draws are captured by re-seeding
``numpy.random`` before the reference call that consumes them, so no reference code is patched.
The one hand-written GPU kernel of the reference (``_parallel_resample``) is executed through
numba's CUDA simulator in a subprocess (NUMBA_ENABLE_CUDASIM=1).
"""
import os
import subprocess
import sys
import warnings

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from oracle import ref_loader  # noqa: E402

R = ref_loader.load()
MGS = R.gaussian_sum_dist.MultivariateGaussianSum
f_ref = R.model.Bioreactor.homeostatic_DEs
g_ref = R.model.Bioreactor.static_outputs

STATE_COVS = numpy.array([numpy.diag([1e-4, 1e-7, 1e-3, 1e-3, 1e-7]),
                          numpy.diag([1e-3, 1e-6, 1e-2, 1e-2, 1e-6])])
MEAS_MEANS = numpy.array([[1e-1, 0], [0, -1e-1]])
MEAS_COVS = numpy.array([[[6e-2, 0], [0, 8e-2]], [[500, 100], [100, 700]]])


def noise_objects(x_shift=None):
    """sim_base.get_noise (sim_base.py:141-160) with library=numpy."""
    means = numpy.zeros((2, 5)) if x_shift is None else numpy.zeros((2, 5)) + x_shift[None, :]
    state = MGS(means, STATE_COVS, numpy.array([0.75, 0.25]), library=numpy)
    meas = MGS(MEAS_MEANS, MEAS_COVS, numpy.array([0.85, 0.15]), library=numpy)
    return state, meas


def save(name, **arrays):
    path = os.path.join(HERE, name)
    numpy.savez_compressed(path, **arrays)
    print("wrote %s (%d bytes)" % (name, os.path.getsize(path)))


X_SS = R.model.Bioreactor.find_SS(numpy.array([0.06, 0.2]),
                                  numpy.array([260 / 180, 640 / 24.6, 1000 / 116, 0, 0]))


# ---------------------------------------------------------------------------------------------
def gen_model():
    rng = numpy.random.default_rng(1)
    n = 512
    x = (X_SS[None, :] + rng.normal(size=(n, 5)) * numpy.array([0.3, 1.0, 0.5, 0.05, 1.0])).astype(numpy.float32)
    # rows that exercise the clamps and every min/max branch of the rate logic
    x[:32, 0] = rng.uniform(-0.2, 0.05, 32)          # Cg <= 0 and tiny
    x[32:64, 3] = rng.uniform(-0.1, 0.1, 32)         # Ce around 0
    x[64:96, 4] = rng.uniform(-400, 400, 32)         # Ch large: r_theta1_req <0 / > max
    x[96:128, 0] = rng.uniform(1.0, 4.0, 32)
    x[128:160, 1] = rng.uniform(-1.0, 1.0, 32)       # Cx around 0
    x[160:192, 2] = rng.uniform(-1.0, 1.0, 32)
    cases = []
    for u, dt in (((0.06, 0.2), 0.1), ((0.06, 0.2), 1.0), ((0.013, 0.17), 1.0), ((0.0, 0.0), 0.5)):
        u = numpy.array(u)
        inc = numpy.array([numpy.asarray(f_ref(row, u, dt), dtype=numpy.float64) for row in x])
        # what the reference stores: float32 row += tuple  (particle.py:66)
        stored = x.copy()
        for i, row in enumerate(stored):
            stored[i] += f_ref(row, u, dt)
        y = numpy.array([numpy.asarray(g_ref(row, u), dtype=numpy.float64) for row in x])
        cases.append((u, dt, inc, stored, y))
    save("model_fg.npz", x=x, x_ss=X_SS,
         us=numpy.array([c[0] for c in cases]), dts=numpy.array([c[1] for c in cases]),
         incs=numpy.array([c[2] for c in cases]), stored=numpy.array([c[3] for c in cases]),
         ys=numpy.array([c[4] for c in cases]))


def gen_mixture():
    state, meas = noise_objects()
    rng = numpy.random.default_rng(2)
    e = numpy.concatenate([rng.normal(size=(200, 2)) * 0.3, rng.normal(size=(200, 2)) * 30.0,
                           rng.normal(size=(112, 2)) * 300.0])
    pdf_meas = meas.pdf(e)
    xs = (rng.normal(size=(256, 5)) * numpy.sqrt(numpy.diag(STATE_COVS[1]))[None, :] * 1.5)
    pdf_state = state.pdf(xs)
    numpy.random.seed(11)
    d1 = state.draw(1000)
    numpy.random.seed(12)
    d2 = state.draw((20, 11))
    numpy.random.seed(13)
    d3 = meas.draw(64)
    save("mixture.npz", e=e, pdf_meas=pdf_meas, xs=xs, pdf_state=pdf_state,
         draw_state_1000_seed11=d1, draw_state_20x11_seed12=d2, draw_meas_64_seed13=d3,
         meas_constants=numpy.asarray(meas._constants), state_constants=numpy.asarray(state._constants))


def gen_pf(name, N, n_cycles, dt, seed):
    state, meas = noise_objects()
    x0, _ = noise_objects(X_SS)
    numpy.random.seed(seed)
    pf = R.filter.ParticleFilter(f_ref, g_ref, N, x0, state, meas)
    out = {"N": numpy.int64(N), "dt": numpy.float64(dt), "particles0": pf.particles.copy(),
           "weights0": pf.weights.copy()}
    x_true = X_SS.copy()
    rng = numpy.random.default_rng(seed + 100)
    for c in range(n_cycles):
        u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
        # plant truth follows the same model so that z is physically consistent (finite weights)
        x_true = x_true + numpy.asarray(f_ref(x_true, u, dt), dtype=numpy.float64)
        z = numpy.asarray(g_ref(x_true, u)) + rng.normal(size=2) * numpy.array([0.2, 0.25])
        s = seed * 1000 + c
        numpy.random.seed(s)
        noise = state.draw(N).copy()
        numpy.random.seed(s)
        pf.predict(u, dt)
        out["u_%d" % c] = u
        out["z_%d" % c] = z
        out["noise_%d" % c] = noise
        out["particles_pred_%d" % c] = pf.particles.copy()
        out["est_pred_%d" % c] = numpy.asarray(pf.point_estimate())
        out["cov_pred_%d" % c] = numpy.float64(pf.point_covariance())
        pf.update(u, z)
        out["weights_upd_%d" % c] = pf.weights.copy()
        out["est_upd_%d" % c] = numpy.asarray(pf.point_estimate())
        out["cov_upd_%d" % c] = numpy.float64(pf.point_covariance())
        numpy.random.seed(s + 500)
        r = numpy.random.rand()
        numpy.random.seed(s + 500)
        pf.resample()
        out["r_%d" % c] = numpy.float64(r)
        out["particles_res_%d" % c] = pf.particles.copy()
        out["weights_res_%d" % c] = pf.weights.copy()
        out["est_res_%d" % c] = numpy.asarray(pf.point_estimate())
        out["cov_res_%d" % c] = numpy.float64(pf.point_covariance())
    out["n_cycles"] = numpy.int64(n_cycles)
    save(name, **out)


def gen_resample():
    """pf_run_seq.py:123-128 style: fresh host float64 weights assigned, then resample()."""
    state, meas = noise_objects()
    x0, _ = noise_objects(X_SS)
    out = {}
    for tag, N, seed in (("a", 1000, 5), ("b", 4096, 6), ("c", 37, 7)):
        numpy.random.seed(seed)
        pf = R.filter.ParticleFilter(f_ref, g_ref, N, x0, state, meas)
        pf.particles[:, 0] = numpy.arange(N, dtype=numpy.float32)     # tag rows to recover indices
        w = numpy.random.random(size=N)
        w /= numpy.sum(w)
        pf.weights = w.copy()
        numpy.random.seed(seed + 50)
        r = numpy.random.rand()
        numpy.random.seed(seed + 50)
        pf.resample()
        out["weights_" + tag] = w
        out["r_" + tag] = numpy.float64(r)
        out["idx_" + tag] = pf.particles[:, 0].astype(numpy.int64)
    # skewed weights (a few dominant ancestors) and weights with exactly representable partial sums
    N = 2048
    numpy.random.seed(8)
    pf = R.filter.ParticleFilter(f_ref, g_ref, N, x0, state, meas)
    pf.particles[:, 0] = numpy.arange(N, dtype=numpy.float32)
    w = numpy.exp(-numpy.random.random(N) * 60.0)
    pf.weights = w.copy()
    numpy.random.seed(58)
    r = numpy.random.rand()
    numpy.random.seed(58)
    pf.resample()
    out["weights_skew"], out["r_skew"], out["idx_skew"] = w, numpy.float64(r), pf.particles[:, 0].astype(numpy.int64)
    pf = R.filter.ParticleFilter(f_ref, g_ref, N, x0, state, meas)
    pf.particles[:, 0] = numpy.arange(N, dtype=numpy.float32)
    w = numpy.random.randint(0, 1 << 12, N).astype(numpy.float64)    # includes zeros -> ties
    pf.weights = w.copy()
    pf_r = 0.5
    numpy.random.seed(59)
    r = numpy.random.rand()
    numpy.random.seed(59)
    pf.resample()
    out["weights_dyadic"], out["r_dyadic"], out["idx_dyadic"] = w, numpy.float64(r), pf.particles[:, 0].astype(numpy.int64)
    save("resample.npz", **out)


def gen_gsukf(name, N, n_cycles, dt, seed):
    state, meas = noise_objects()
    x0, _ = noise_objects(X_SS)
    numpy.random.seed(seed)
    gf = R.filter.GaussianSumUnscentedKalmanFilter(f_ref, g_ref, N, x0, state, meas)
    out = {"N": numpy.int64(N), "dt": numpy.float64(dt), "means0": gf.means.copy(),
           "covariances0": gf.covariances.copy(), "w_sigma": gf._w_sigma.copy(),
           "sigmas0": gf._get_sigma_points().copy()}
    x_true = X_SS.copy()
    rng = numpy.random.default_rng(seed + 100)
    for c in range(n_cycles):
        u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
        x_true = x_true + numpy.asarray(f_ref(x_true, u, dt), dtype=numpy.float64)
        z = numpy.asarray(g_ref(x_true, u)) + rng.normal(size=2) * numpy.array([0.2, 0.25])
        s = seed * 1000 + c
        numpy.random.seed(s)
        noise = state.draw((N, 11)).copy()
        numpy.random.seed(s)
        gf.predict(u, dt)
        out["u_%d" % c], out["z_%d" % c], out["noise_%d" % c] = u, z, noise
        out["means_pred_%d" % c] = gf.means.copy()
        out["covs_pred_%d" % c] = gf.covariances.copy()
        gf.update(u, z)
        out["means_upd_%d" % c] = gf.means.copy()
        out["covs_upd_%d" % c] = gf.covariances.copy()
        out["weights_upd_%d" % c] = gf.weights.copy()
        out["est_upd_%d" % c] = numpy.asarray(gf.point_estimate())
        out["cov_upd_%d" % c] = numpy.float64(gf.point_covariance())
        numpy.random.seed(s + 500)
        r = numpy.random.rand()
        numpy.random.seed(s + 500)
        gf.resample()
        out["r_%d" % c] = numpy.float64(r)
        out["means_res_%d" % c] = gf.means.copy()
        out["covs_res_%d" % c] = gf.covariances.copy()
        out["est_res_%d" % c] = numpy.asarray(gf.point_estimate())
        out["cov_res_%d" % c] = numpy.float64(gf.point_covariance())
    out["n_cycles"] = numpy.int64(n_cycles)
    save(name, **out)


NICELY_SNIPPET = r"""
import sys, numpy, warnings
warnings.simplefilter('ignore')
sys.path.insert(0, %(root)r)
from oracle import ref_loader
R = ref_loader.load()
P = R.filter.ParallelParticleFilter
out = {}
for tag, N, seed in (('a', 256, 1), ('b', 1000, 2)):
    rng = numpy.random.default_rng(seed)
    w = rng.random(N)
    if tag == 'b':
        w = numpy.exp(-rng.random(N) * 30)
    c = numpy.cumsum(w); c /= c[-1]
    r = numpy.float64(rng.random())
    idx = numpy.zeros(N, numpy.int64)
    tpb = 1024 if N >= 1024 else 32 * ((N - 1) // 32 + 1)
    P._parallel_resample[(N - 1) // tpb + 1, tpb](c, idx, r, N)
    out['cumsum_' + tag], out['r_' + tag], out['idx_' + tag] = c, r, idx
numpy.savez_compressed(%(path)r, **out)
"""


def gen_nicely():
    path = os.path.join(HERE, "nicely_cudasim.npz")
    env = dict(os.environ, NUMBA_ENABLE_CUDASIM="1")
    subprocess.run([sys.executable, "-c", NICELY_SNIPPET % {"root": ROOT, "path": path}], check=True, env=env)
    print("wrote nicely_cudasim.npz (%d bytes)" % os.path.getsize(path))


if __name__ == "__main__":
    gen_model()
    gen_mixture()
    gen_pf("pf_n256.npz", 256, 3, 0.1, seed=21)
    gen_pf("pf_n1024_dt1.npz", 1024, 2, 1.0, seed=22)       # BASELINE.json configs[0] scale
    gen_resample()
    gen_gsukf("gsukf_n64.npz", 64, 3, 0.1, seed=31)
    gen_gsukf("gsukf_n256_dt1.npz", 256, 2, 1.0, seed=32)
    gen_nicely()
