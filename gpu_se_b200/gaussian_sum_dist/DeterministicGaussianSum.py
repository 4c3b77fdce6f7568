"""Deterministic test double (gaussian_sum_dist/DeterministicGaussianSum.py:6-65 of the reference).

``draw`` returns the same prefix of one cached stream on every call, so two filters fed by it see
identical "noise".  The filters detect this class and hand its values to the kernels through the
host-noise argument of gse_pf_predict / gse_gsf_predict instead of drawing with Philox.
The cache is shared class-wide, as in the reference (:27).
"""
import numpy
import torch

from gpu_se_b200 import _device
from gpu_se_b200.gaussian_sum_dist.MultivariateGaussianSum import MultivariateGaussianSum


class DeterministicGaussianSum(MultivariateGaussianSum):
    _values = numpy.array([], dtype=numpy.float32)

    def __init__(self, means, covariances, weights, library=None, seed=None, device=None):
        super().__init__(means, covariances, weights, library, seed=seed, device=device)

    @classmethod
    def set_stream(cls, values):
        """Install an explicit value stream (lets a test share it with the oracle)."""
        cls._values = numpy.asarray(values, dtype=numpy.float32).ravel().copy()

    def draw_host(self, shape=(1,)):
        """numpy (*shape, Nx) prefix of the cached stream (:48-63)."""
        if not isinstance(shape, tuple):
            shape = (shape,)
        size = int(numpy.prod(shape)) * self._Nx
        cls = DeterministicGaussianSum
        if cls._values.size < size:
            missing = size - cls._values.size
            rows = -(-missing // self._Nx)
            drawn = _device.to_numpy(MultivariateGaussianSum.draw(self, rows)).reshape(-1)
            cls._values = numpy.hstack([cls._values, drawn.astype(numpy.float32)])
        return cls._values[:size].reshape(shape + (self._Nx,))

    def draw(self, shape=(1,)):
        out = self.draw_host(shape)
        dev = _device.resolve_device(self._device)
        return _device.wrap(torch.as_tensor(numpy.ascontiguousarray(out), device=dev))
