"""Gaussian-sum unscented Kalman filter on B200: the reference's
``ParallelGaussianSumUnscentedKalmanFilter`` interface (filter/gs_ukf.py:223-449) over the sm_100a
kernels in csrc/gse_gsf.cu.

Component state on the device is one SoA float32 buffer ``(20, ld)``: rows 0-4 the means, rows
5-19 the lower triangle of the covariance (the reference stores the full 5x5, gs_ukf.py:52).
``means`` is the ``(N, 5)`` view, ``covariances`` materialises ``(N, 5, 5)`` on demand.
"""
import numpy
import torch

from gpu_se_b200 import _device, _lib
from gpu_se_b200.filter._base import WeightedEnsemble, mixture_view
from gpu_se_b200.model.BioreactorModel import model_id_for

_TRI = [(i, j) for i in range(5) for j in range(i + 1)]


class ParallelGaussianSumUnscentedKalmanFilter(WeightedEnsemble):
    """GS-UKF running on the GPU; parameters as the reference (gs_ukf.py:187-221) plus keyword-only
    ``device``, ``seed``, ``means`` (initial means instead of ``x0.draw(N)``) and, for one shard of a larger population,
    ``index0`` / ``workspace_rows`` / ``peer`` as in ParallelParticleFilter."""

    NCOLS = 20

    def __init__(self, f, g, N_particles, x0, state_pdf, measurement_pdf, *, device=None, seed=None, means=None,
                 index0=0, workspace_rows=None, peer=False):
        self.f = f
        self.g = g
        self._model_id = model_id_for(f, g)
        self._init_ensemble(N_particles, state_pdf, measurement_pdf, device, seed, workspace_rows, peer)
        self._index0 = int(index0)
        n = self.N_particles
        self._Nx, self._Ny, self._N_sigmas = 5, 2, 11
        self._w_sigma = numpy.full(11, 1 / (2 * 5 + 8 / 5), dtype=numpy.float32)   # gs_ukf.py:66
        self._w_sigma[0] = 1 / (1 + 5 / 4 * 5)                                      # gs_ukf.py:67
        if means is not None:
            self.means = means
        elif hasattr(x0, "draw_host"):
            self.means = x0.draw_host(n)
        else:
            mix = mixture_view(x0).as_gse_mixture()
            _lib.check(_lib.lib.gse_mixture_draw(self._ctx.handle, mix, self._state.data_ptr(), self._ld, n,
                                                 self._seed, 0xFFFFFFFF, self._index0, self._stream()))
        # covariances = repeat(state_pdf.covariances[0])  (gs_ukf.py:52), float32
        cov0 = numpy.asarray(_device.to_numpy(self._state_mix.covariances)[0], dtype=numpy.float32)
        self.covariances = numpy.repeat(cov0[None], n, axis=0)

    # -- state attributes --------------------------------------------------------------
    def _mean_ptr(self):
        return self._state.data_ptr()

    def _cov_ptr(self):
        return self._state.data_ptr() + 5 * self._ld * 4

    @property
    def means(self):
        self._materialise()
        return _device.wrap(self._state[:5, :self.N_particles].t())

    @means.setter
    def means(self, value):
        n = self.N_particles
        v = torch.as_tensor(numpy.ascontiguousarray(_device.to_numpy(value), dtype=numpy.float32), device=self.device)
        if tuple(v.shape) != (n, 5):
            raise ValueError("means must have shape (%d, 5)" % n)
        self._materialise()
        self._state[:5, :n].copy_(v.t())
        self._touch()

    @property
    def covariances(self):
        n = self.N_particles
        self._materialise()
        full = torch.empty((n, 5, 5), dtype=torch.float32, device=self.device)
        for t, (i, j) in enumerate(_TRI):
            full[:, i, j] = self._state[5 + t, :n]
            full[:, j, i] = self._state[5 + t, :n]
        return _device.wrap(full)

    @covariances.setter
    def covariances(self, value):
        n = self.N_particles
        v = torch.as_tensor(numpy.ascontiguousarray(_device.to_numpy(value), dtype=numpy.float32), device=self.device)
        if tuple(v.shape) != (n, 5, 5):
            raise ValueError("covariances must have shape (%d, 5, 5)" % n)
        self._materialise()
        for t, (i, j) in enumerate(_TRI):
            self._state[5 + t, :n].copy_(v[:, i, j])
        self._touch()

    # -- stages ------------------------------------------------------------------------
    def _get_sigma_points(self):
        """(N, 11, 5) sigma points (gs_ukf.py:332-346)."""
        n = self.N_particles
        self._materialise()
        out = torch.empty((55, self._ld), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.gse_gsf_sigma_points(self._ctx.handle, self._mean_ptr(), self._cov_ptr(), self._ld, n,
                                                 out.data_ptr(), self._ld, self._stream()))
        return _device.wrap(out[:, :n].t().reshape(n, 11, 5))

    def _predict_now(self, u, dt, noise=None):
        """gs_ukf.py:348-367.  ``noise`` (N, 11, 5): host-supplied draws for cross-checks."""
        n = self.N_particles
        if noise is None:
            noise = self._host_noise(self.state_pdf, (n, 11))
        nz_ptr, ld_nz, nz = None, 0, None
        if noise is not None:
            nz = torch.zeros((55, self._ld), dtype=torch.float32, device=self.device)
            host = numpy.ascontiguousarray(_device.to_numpy(noise), dtype=numpy.float32).reshape(n, 55)
            nz[:, :n].copy_(torch.as_tensor(host, device=self.device).t())
            nz_ptr, ld_nz = nz.data_ptr(), self._ld
        src_mean, src_cov = self._mean_ptr(), self._cov_ptr()
        idx = self._idx_ptr()
        if self._pending:                      # apply the pending resample on the way in
            self._state, self._state_alt = self._state_alt, self._state
            self._pending = False
        _lib.check(_lib.lib.gse_gsf_predict(self._ctx.handle, src_mean, src_cov, self._ld, idx, self._mean_ptr(),
                                            self._cov_ptr(), self._ld, n, _lib.as_double2(u), float(dt), self._seed,
                                            self._step, self._index0, nz_ptr, ld_nz, self._stream()))
        self._step += 1
        self._touch()

    def _update_now(self, u, z):
        """gs_ukf.py:369-407."""
        self._materialise()
        _lib.check(_lib.lib.gse_gsf_update(self._ctx.handle, self._mean_ptr(), self._cov_ptr(), self._ld,
                                           self.N_particles, self._loglik_ptr(), self._loglik.data_ptr(),
                                           _lib.as_double2(u), _lib.as_double2(z), self._stats.data_ptr(),
                                           self._stream()))
        self._after_update()

    # resample(): WeightedEnsemble.resample gathers all 20 rows (gs_ukf.py:409-436)

    # -- estimates -----------------------------------------------------------------------
    def _launch_moments(self, mean_only=False):
        _lib.check(_lib.lib.gse_gsf_moments(
            self._ctx.handle, self._mean_ptr(), self._cov_ptr(), self._ld, self.N_particles, self._idx_ptr(),
            self._loglik_ptr(), self._base.data_ptr() if self._base is not None else None,
            self._stats.data_ptr(), self._mom.data_ptr(), self._stream()))

    def covariance_matrix(self, normalised=False):
        """cov_cov + cov_mean (gs_ukf.py:442-447)."""
        cov_mean, mom, A = self._scatter_about(normalised)
        return A * self._unpack_sym(mom[26:41]) + cov_mean

    def point_covariance(self, normalised=False):
        return float(numpy.linalg.svd(self.covariance_matrix(normalised), compute_uv=False)[0])


GaussianSumUnscentedKalmanFilter = ParallelGaussianSumUnscentedKalmanFilter
