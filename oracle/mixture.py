"""Oracle restatement of the Gaussian-sum (mixture) distribution.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference gaussian_sum_dist/MultivariateGaussianSum.py:
  * ``__init__`` :27-37  parameters stored float32; inverse covariances float64 from the float64
    input (:33); normalising constants from the float32 covariances (:36-37)
  * ``pdf``      :39-63  sum_d w_d c_d exp(-1/2 (x-mu_d)' S_d^-1 (x-mu_d)), float64
  * ``draw``     :65-97  component counts by ``random.choice`` + ``bincount``, then
    ``multivariate_normal`` per component, rows ORDERED BY COMPONENT (quirk Q5)
and gaussian_sum_dist/DeterministicGaussianSum.py:32-65 (``FixedNoise``: the same prefix of one
cached stream on every call).
"""
import numpy


class MixtureOracle:
    def __init__(self, means, covariances, weights):
        means = numpy.asarray(means, dtype=numpy.float64)
        covariances = numpy.asarray(covariances, dtype=numpy.float64)
        weights = numpy.asarray(weights, dtype=numpy.float64)
        # :29-31  stored parameters are float32
        self.means = means.astype(numpy.float32)
        self.weights = weights.astype(numpy.float32)
        self.covariances = covariances.astype(numpy.float32)
        # :33 inverse of the float64 input, never cast
        self._inverse_covariances = numpy.linalg.inv(covariances)
        self._Nd, self._Nx = means.shape
        # :36-37 (the reference evaluates det/sqrt on the float32 copy; float64 here, the
        # difference is <= 1 ulp32 of the constant and is inside the stated pdf tolerance)
        self._constants = (2 * numpy.pi) ** (-self._Nx / 2) / numpy.sqrt(
            numpy.linalg.det(self.covariances.astype(numpy.float64)))

    # -- pdf -----------------------------------------------------------------------------
    def quadratic_forms(self, x):
        """(m, Nd) array of (x-mu_d)' S_d^-1 (x-mu_d)   (:52-56)."""
        x = numpy.atleast_2d(numpy.asarray(x, dtype=numpy.float64))
        es = x[:, None, :] - self.means[None, :, :].astype(numpy.float64)
        return numpy.einsum('mdi,dij,mdj->md', es, self._inverse_covariances, es)

    def pdf(self, x):
        q = self.quadratic_forms(x)
        r = numpy.exp(-0.5 * q)                                              # :57
        return numpy.sum(self._constants * self.weights.astype(numpy.float64) * r, axis=1)  # :60

    def logpdf(self, x):
        """log of :meth:`pdf`, evaluated without underflow (what the device path accumulates)."""
        q = self.quadratic_forms(x)
        a = numpy.log(self._constants * self.weights.astype(numpy.float64)) - 0.5 * q
        m = a.max(axis=1)
        return m + numpy.log(numpy.exp(a - m[:, None]).sum(axis=1))

    # -- draw ----------------------------------------------------------------------------
    def draw(self, shape=(1,)):
        """Same numpy.random call sequence as the reference (:79-95): with the same global seed
        it returns the same samples as the reference's numpy path."""
        if not isinstance(shape, tuple):
            shape = (shape,)
        size = int(numpy.prod(shape))
        bins = numpy.bincount(
            numpy.random.choice(numpy.arange(self._Nd), size, p=self.weights),
            minlength=self._Nd)
        out = numpy.empty((size, self._Nx), dtype=numpy.float32)
        index = 0
        for n, mean, cov in zip(bins, self.means, self.covariances):
            out[index:index + n] = numpy.random.multivariate_normal(mean, cov, int(n))
            index += n
        return out.reshape(shape + (self._Nx,))


class FixedNoise(MixtureOracle):
    """Test double in the spirit of DeterministicGaussianSum (DeterministicGaussianSum.py:32-65):
    ``draw`` returns the same prefix of one fixed stream every time.  The stream is per instance
    here (the reference shares one class-level buffer between all instances, :27)."""

    def __init__(self, means, covariances, weights, stream):
        super().__init__(means, covariances, weights)
        self._values = numpy.asarray(stream, dtype=numpy.float32).ravel()

    def draw(self, shape=(1,)):
        if not isinstance(shape, tuple):
            shape = (shape,)
        size = int(numpy.prod(shape)) * self._Nx
        if self._values.size < size:
            raise ValueError("fixed noise stream too short")
        return self._values[:size].reshape(shape + (self._Nx,))


# --- the benchmark's distributions (sim_base.py:141-160) ---------------------------------------
STATE_MEANS = numpy.zeros((2, 5))
STATE_COVS = numpy.array([numpy.diag([1e-4, 1e-7, 1e-3, 1e-3, 1e-7]),
                          numpy.diag([1e-3, 1e-6, 1e-2, 1e-2, 1e-6])])
STATE_WEIGHTS = numpy.array([0.75, 0.25])
MEAS_MEANS = numpy.array([[1e-1, 0], [0, -1e-1]])
MEAS_COVS = numpy.array([[[6e-2, 0], [0, 8e-2]], [[500, 100], [100, 700]]])
MEAS_WEIGHTS = numpy.array([0.85, 0.15])


def benchmark_noise():
    """(state_pdf, measurement_pdf) of sim_base.get_noise (sim_base.py:141-160)."""
    return (MixtureOracle(STATE_MEANS, STATE_COVS, STATE_WEIGHTS),
            MixtureOracle(MEAS_MEANS, MEAS_COVS, MEAS_WEIGHTS))


def benchmark_x0(x_steady):
    """x0 of sim_base.get_parts (sim_base.py:102-104): the state-noise mixture shifted to the
    steady state."""
    return MixtureOracle(STATE_MEANS + numpy.asarray(x_steady)[None, :], STATE_COVS, STATE_WEIGHTS)
