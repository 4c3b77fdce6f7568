"""Shared builders for the -m gpu parity tests (everything goes through the public classes, which
call the C ABI)."""
import numpy

from oracle import bioreactor, mixture


def make_pdfs(g):
    state = g.MultivariateGaussianSum(mixture.STATE_MEANS, mixture.STATE_COVS, mixture.STATE_WEIGHTS)
    meas = g.MultivariateGaussianSum(mixture.MEAS_MEANS, mixture.MEAS_COVS, mixture.MEAS_WEIGHTS)
    x0 = g.MultivariateGaussianSum(mixture.STATE_MEANS + bioreactor.X_STEADY[None, :], mixture.STATE_COVS,
                                   mixture.STATE_WEIGHTS)
    return x0, state, meas


def make_pf(g, N, seed=1, particles=None, **kw):
    x0, state, meas = make_pdfs(g)
    return g.ParticleFilter(g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs, N, x0, state, meas,
                            seed=seed, particles=particles, **kw)


def make_gsf(g, N, seed=1, means=None, **kw):
    x0, state, meas = make_pdfs(g)
    return g.GaussianSumUnscentedKalmanFilter(g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs, N, x0,
                                              state, meas, seed=seed, means=means, **kw)


def consistent_measurement(u, dt, rng):
    """z = g(x_true) + v with x_true one model step from the steady state (keeps weights finite)."""
    x = bioreactor.X_STEADY + bioreactor.increment(bioreactor.X_STEADY, u, dt)
    return bioreactor.outputs(x, round32=False) + rng.normal(size=2) * numpy.array([0.2, 0.25])


def expected_indices_from_cumsum(c_u64, r, n_out=None):
    """The reference comparison (particle.py:89-98) applied to the device's integer cumulative
    weights: cumsum/cumsum[-1] in float64, searchsorted left."""
    c = c_u64.astype(numpy.float64)          # uint64 -> float64 rounds to nearest even
    cn = c / c[-1]
    n = len(c) if n_out is None else n_out
    u = (numpy.arange(n, dtype=numpy.float64) + numpy.float64(r)) / n
    return numpy.searchsorted(cn, u, side="left")
