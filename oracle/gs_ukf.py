"""Oracle restatement of the Gaussian-sum unscented Kalman filter.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference filter/gs_ukf.py:45-183 (the CPU ``GaussianSumUnscentedKalmanFilter``):
  * ``__init__``          :45-67   means = x0.draw(N); covariances = repeat(state_pdf.covariances[0]);
                                   sigma weights w0 = 1/(1+5Nx/4), wi = 1/(2Nx+8/5) (float32)
  * ``_get_sigma_points`` :69-80   L = chol(P) (retry with +1e-10 I, :72-75); mean, mean + L[:, j],
                                   mean - L[:, j]  -- NO sqrt(n+lambda) scaling (quirk Q6)
  * ``predict``           :82-103  f on every sigma point, an INDEPENDENT noise draw per sigma point
                                   (:99), numpy.average mean (divides by sum of weights), weighted
                                   scatter with the raw weights
  * ``update``            :105-149 sigma points of the current (m, P); g; eta mean; P_xy, P_yy;
                                   K = P_xy pinv(P_yy); m += K (z - eta_mean); P -= K P_yy K';
                                   w *= pdf(z - g(m))
  * ``resample``          :151-171 as the particle filter, gathers means and covariances
  * ``point_estimate``    :173-175; ``point_covariance`` :177-183

Arithmetic float64, storage float32 for means / covariances (the reference keeps them float32
through in-place ``+=`` / ``-=``; ``predict`` re-creates them from float32 sigma points).
"""
import numpy

from . import bioreactor
from .particle import systematic_indices


def sigma_weights(Nx=5):
    w = numpy.full(2 * Nx + 1, 1 / (2 * Nx + 8 / 5), dtype=numpy.float32)   # :66
    w[0] = 1 / (1 + 5 / 4 * Nx)                                             # :67
    return w


class GSUKFOracle:
    def __init__(self, N_particles, x0, state_pdf, measurement_pdf, means=None):
        self.N_particles = int(N_particles)
        self.means = (x0.draw(self.N_particles) if means is None
                      else numpy.array(means, dtype=numpy.float32))                     # :50
        self.covariances = numpy.repeat(state_pdf.covariances[0][None, :, :],
                                        self.N_particles, axis=0).astype(numpy.float32)  # :52
        self.weights = numpy.full(self.N_particles, 1 / self.N_particles, dtype=numpy.float32)  # :54
        self.state_pdf = state_pdf
        self.measurement_pdf = measurement_pdf
        self._Nx = self.means.shape[1]
        self._Ny = 2
        self._N_sigmas = 2 * self._Nx + 1
        self._w_sigma = sigma_weights(self._Nx)

    def cholesky(self):
        """:71-75. Per-component retry (the reference retries the whole batch; identical result
        for components whose first attempt succeeds up to the 1e-10 jitter, which is below
        float32 resolution of every covariance entry on this model)."""
        P = self.covariances.astype(numpy.float64)
        try:
            return numpy.linalg.cholesky(P)
        except numpy.linalg.LinAlgError:
            return numpy.linalg.cholesky(P + 1e-10 * numpy.eye(self._Nx))

    def _get_sigma_points(self):
        stds = self.cholesky().swapaxes(1, 2)                                           # :72
        sigmas = numpy.repeat(self.means[:, None, :].astype(numpy.float64), self._N_sigmas, axis=1)
        sigmas[:, 1:self._Nx + 1, :] += stds                                            # :77
        sigmas[:, self._Nx + 1:, :] -= stds                                             # :78
        return sigmas.astype(numpy.float32)   # reference sigma array is float32 (repeat of means)

    def predict(self, u, dt, noise=None):
        sigmas = self._get_sigma_points()
        inc = bioreactor.increment(sigmas, u, dt)                                       # :95-97
        sigmas = (sigmas.astype(numpy.float64) + inc).astype(numpy.float32)
        if noise is None:
            noise = self.state_pdf.draw((self.N_particles, self._N_sigmas))
        sigmas += numpy.asarray(noise, dtype=numpy.float32).reshape(sigmas.shape)       # :99
        w = self._w_sigma.astype(numpy.float64)
        s64 = sigmas.astype(numpy.float64)
        self.means = (numpy.einsum('nsj,s->nj', s64, w) / w.sum()).astype(numpy.float32)  # :101
        d = (sigmas - self.means[:, None, :]).astype(numpy.float64)                     # :102 (f32 subtract)
        self.covariances = numpy.einsum('nsi,nsj,s->nij', d, d, w).astype(numpy.float32)  # :103

    def update(self, u, z):
        z = numpy.asarray(z, dtype=numpy.float64)
        w = self._w_sigma.astype(numpy.float64)
        sigmas = self._get_sigma_points()
        etas = bioreactor.outputs(sigmas, u)                                            # :118-123
        eta_means = numpy.einsum('nsj,s->nj', etas, w) / w.sum()                        # :126
        ds = (sigmas - self.means[:, None, :]).astype(numpy.float64)                    # :127
        de = etas - eta_means[:, None, :]                                               # :128
        P_xys = numpy.einsum('nsi,nsj,s->nij', ds, de, w)                               # :130
        P_yys = numpy.einsum('nsi,nsj,s->nij', de, de, w)                               # :131
        P_yy_invs = numpy.linalg.pinv(P_yys)                                            # :132
        Ks = P_xys @ P_yy_invs                                                          # :133
        es = z[None, :] - eta_means                                                     # :136
        self.means = (self.means + numpy.einsum('nij,nj->ni', Ks, es)).astype(numpy.float32)  # :137
        self.covariances = (self.covariances
                            - Ks @ P_yys @ Ks.swapaxes(1, 2)).astype(numpy.float32)     # :139
        y_means = bioreactor.outputs(self.means, u)                                     # :145-146
        glob_es = z[None, :] - y_means                                                  # :148
        self.weights *= self.measurement_pdf.pdf(glob_es)                               # :149

    def log_likelihood_of_means(self, u, z):
        z = numpy.asarray(z, dtype=numpy.float64)
        return self.measurement_pdf.logpdf(z[None, :] - bioreactor.outputs(self.means, u))

    def resample(self, r=None):
        cumsum = numpy.cumsum(self.weights)                                             # :155
        cumsum /= cumsum[-1]                                                            # :156
        if r is None:
            r = numpy.random.rand()                                                     # :159
        idx = systematic_indices(cumsum, r)                                             # :162-166
        self.means = self.means[idx]                                                    # :168
        self.covariances = self.covariances[idx]                                        # :169
        self.weights = numpy.full(self.N_particles, 1 / self.N_particles)               # :170
        return idx

    def point_estimate(self):
        return self.weights @ self.means                                                # :175

    def covariance_matrix(self):
        cov_cov = numpy.sum(self.weights[:, None, None] * self.covariances, axis=0)     # :179
        dist = self.means - (self.weights @ self.means)                                 # :180
        return cov_cov + dist.T @ (dist * self.weights[:, None])                        # :181-182

    def point_covariance(self):
        return numpy.linalg.svd(self.covariance_matrix(), compute_uv=False)[0]          # :183
