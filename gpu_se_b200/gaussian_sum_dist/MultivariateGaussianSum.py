"""Gaussian-sum distribution with device-side pdf and sampler.

Mirrors gaussian_sum_dist/MultivariateGaussianSum.py:7-97 of the reference: same constructor
(``means, covariances, weights, library``), same ``means / covariances / weights`` float32
attributes, ``pdf(x)`` and ``draw(shape)``.  ``library`` is accepted and ignored -- there is one
backend.  ``pdf`` runs gse_mixture_pdf (float64), ``draw`` runs the Philox sampler
gse_mixture_draw for any dimension up to 5 (state noise, measurement noise, x0); rows of a draw are NOT
grouped by component.
"""
import numpy
import torch

from gpu_se_b200 import _device, _lib


class MultivariateGaussianSum:
    _seed_counter = 0

    def __init__(self, means, covariances, weights, library=None, seed=None, device=None):
        self.lib = library
        means = _device.to_numpy(means)
        covariances = _device.to_numpy(covariances)
        weights = _device.to_numpy(weights)
        # the reference stores float32 copies (:29-31) and keeps the float64 covariances only
        # through their inverse (:33); both precisions are kept here for the same purpose
        self.means = numpy.array(means, dtype=numpy.float32)
        self.weights = numpy.array(weights, dtype=numpy.float32)
        self.covariances = numpy.array(covariances, dtype=numpy.float32)
        self._covariances64 = numpy.array(covariances, dtype=numpy.float64)
        self._Nd, self._Nx = self.means.shape
        if seed is None:
            MultivariateGaussianSum._seed_counter += 1
            seed = 0x5EED0000 + MultivariateGaussianSum._seed_counter
        self._seed = int(seed)
        self._draws = 0
        self._device = device
        self._ctx = None

    # -- C-ABI views ---------------------------------------------------------------------
    def as_gse_mixture(self):
        """Current parameters as a gse_mixture (means may have been shifted after construction,
        sim_base.py:103-104)."""
        return _lib.make_mixture(numpy.asarray(self.means, dtype=numpy.float64), self._covariances64,
                                 numpy.asarray(self.weights, dtype=numpy.float64))

    def _context(self):
        # a private context so that pdf / draw work without a filter; mixtures passed to
        # gse_ctx_create are placeholders of the right dimension
        if self._ctx is None:
            from gpu_se_b200.filter._base import Context
            dev = _device.resolve_device(self._device)
            self._ctx = Context(dev, 1, None, None)
        return self._ctx

    # -- pdf -----------------------------------------------------------------------------
    def pdf(self, x):
        """(m,) float64 pdf values at the (m, Nx) points x (:39-63).  Returns a numpy array for
        host input, a device tensor for device input."""
        on_device = isinstance(x, torch.Tensor) and x.is_cuda
        ctx = self._context()
        xt = torch.as_tensor(_device.to_numpy(x) if not on_device else x, dtype=torch.float32, device=ctx.device)
        xt = torch.atleast_2d(xt)
        if xt.shape[1] != self._Nx:
            raise ValueError("pdf expects (m, %d) points" % self._Nx)
        m = xt.shape[0]
        soa = xt.t().contiguous()
        out = torch.empty(m, dtype=torch.float64, device=ctx.device)
        mix = self.as_gse_mixture()
        _lib.check(_lib.lib.gse_mixture_pdf(ctx.handle, mix, soa.data_ptr(), m, m, out.data_ptr(), 0,
                                            _device.stream_ptr(ctx.device)))
        return _device.wrap(out) if on_device else out.cpu().numpy()

    # -- draw ----------------------------------------------------------------------------
    def draw(self, shape=(1,)):
        """(*shape, Nx) float32 samples on the device (:65-97), any Nx <= 5."""
        if not isinstance(shape, tuple):
            shape = (shape,)
        size = int(numpy.prod(shape))
        ctx = self._context()
        ld = _device.round_up(max(size, 1), 4)
        buf = torch.empty((self._Nx, ld), dtype=torch.float32, device=ctx.device)
        mix = self.as_gse_mixture()
        _lib.check(_lib.lib.gse_mixture_draw(ctx.handle, mix, buf.data_ptr(), ld, size, self._seed, self._draws, 0,
                                             _device.stream_ptr(ctx.device)))
        self._draws += 1
        return _device.wrap(buf[:, :size].t().reshape(shape + (self._Nx,)))
