"""Oracle restatement of the bootstrap particle filter.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference filter/particle.py:43-114 (the CPU ``ParticleFilter``):
  * ``__init__``        :43-52   particles = x0.draw(N) (N,5) float32; weights = full(N, 1/N, float32)
  * ``predict``         :54-67   x_i += f(x_i,u,dt) per particle, then X += state_pdf.draw(N)
  * ``update``          :69-83   w_i *= measurement_pdf.pdf(z - g(x_i,u))
  * ``resample``        :85-103  cumsum, normalise, sequential systematic sweep
                                 (== searchsorted(side='left')), gather, weights = full(N, 1/N) float64
  * ``point_estimate``  :105-107 weights @ particles (un-normalised weights, quirk Q4)
  * ``point_covariance``:109-114 max singular value of the weighted scatter

Two forms of each stage: vectorised float64 numpy (large-N parity checks) and ``*_loop`` methods
that keep the reference's per-particle Python loops (the cost structure timed as the CPU
baseline).  Storage dtypes follow the reference: particles float32; weights float32 until the
first resample, float64 afterwards (quirk Q3).
"""
import numpy

from . import bioreactor


def systematic_indices(cumsum_normalised, r, n_out=None):
    """particle.py:92-100: idx_i = min{k : cumsum[k] >= (i + r)/N}; sequential two-pointer sweep
    restated with searchsorted(side='left') (identical comparisons, identical float64 operands)."""
    n = cumsum_normalised.shape[0] if n_out is None else int(n_out)
    u = (numpy.arange(n, dtype=numpy.float64) + numpy.float64(r)) / n
    return numpy.searchsorted(cumsum_normalised, u, side='left').astype(numpy.int64)


def systematic_indices_loop(cumsum_normalised, r):
    """The literal loop of particle.py:92-100 (small N only)."""
    n = cumsum_normalised.shape[0]
    out = numpy.zeros(n, dtype=numpy.int64)
    k = 0
    for i in range(n):
        u = (i + r) / n
        while cumsum_normalised[k] < u:
            k += 1
        out[i] = k
    return out


def nicely_indices(cumsum_normalised, r):
    """What the reference's GPU kernel computes (particle.py:223-263): the up/down walk ends at
    max{k : cumsum[k] <= u} + 1 == searchsorted(side='right').  Differs from the CPU sweep only
    on exact ties cumsum[k] == u."""
    n = cumsum_normalised.shape[0]
    u = (numpy.arange(n, dtype=numpy.float64) + numpy.float64(r)) / n
    return numpy.searchsorted(cumsum_normalised, u, side='right').astype(numpy.int64)


class ParticleFilterOracle:
    def __init__(self, N_particles, x0, state_pdf, measurement_pdf, particles=None):
        self.N_particles = int(N_particles)
        self.particles = (x0.draw(self.N_particles) if particles is None
                          else numpy.array(particles, dtype=numpy.float32))       # :49
        self.weights = numpy.full(self.N_particles, 1 / self.N_particles, dtype=numpy.float32)  # :50
        self.state_pdf = state_pdf
        self.measurement_pdf = measurement_pdf

    # -- predict -------------------------------------------------------------------------
    def predict(self, u, dt, noise=None):
        inc = bioreactor.increment(self.particles, u, dt)                        # :65-66
        self.particles = (self.particles.astype(numpy.float64) + inc).astype(numpy.float32)
        if noise is None:
            noise = self.state_pdf.draw(self.N_particles)
        self.particles += numpy.asarray(noise, dtype=numpy.float32)              # :67

    def predict_loop(self, u, dt, noise=None):
        for i, particle in enumerate(self.particles):                           # :65
            self.particles[i] += bioreactor.increment_scalar(particle, u, dt)    # :66
        if noise is None:
            noise = self.state_pdf.draw(self.N_particles)
        self.particles += noise

    # -- update --------------------------------------------------------------------------
    def likelihood(self, u, z):
        y = bioreactor.outputs(self.particles, u)                                # :81
        e = numpy.asarray(z, dtype=numpy.float64)[None, :] - y                   # :82
        return self.measurement_pdf.pdf(e)

    def log_likelihood(self, u, z):
        y = bioreactor.outputs(self.particles, u)
        e = numpy.asarray(z, dtype=numpy.float64)[None, :] - y
        return self.measurement_pdf.logpdf(e)

    def update(self, u, z):
        self.weights *= self.likelihood(u, z)                                    # :83

    def update_loop(self, u, z):
        z = numpy.asarray(z, dtype=numpy.float64)
        for i, particle in enumerate(self.particles):                           # :80
            y = bioreactor.outputs_scalar(particle, u)                           # :81
            e = z - y                                                            # :82
            self.weights[i] *= self.measurement_pdf.pdf(e)[0]                    # :83

    # -- resample ------------------------------------------------------------------------
    def resample(self, r=None, loop=False):
        cumsum = numpy.cumsum(self.weights)                                      # :89
        cumsum /= cumsum[-1]                                                     # :90
        if r is None:
            r = numpy.random.rand()                                              # :93
        idx = (systematic_indices_loop if loop else systematic_indices)(cumsum, r)
        self.particles = self.particles[idx]                                     # :102
        self.weights = numpy.full(self.N_particles, 1 / self.N_particles)        # :103 (float64)
        return idx

    # -- estimates -----------------------------------------------------------------------
    def point_estimate(self):
        return self.weights @ self.particles                                     # :107

    def covariance_matrix(self):
        dist = self.particles - (self.weights @ self.particles)                  # :111
        return dist.T @ (dist * self.weights[:, None])                           # :112

    def point_covariance(self):
        return numpy.linalg.svd(self.covariance_matrix(), compute_uv=False)[0]   # :113-114
