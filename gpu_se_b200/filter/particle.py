"""Bootstrap particle filter on B200: the reference's ``ParallelParticleFilter`` interface
(filter/particle.py:151-327; constructor ``(f, g, N_particles, x0, state_pdf, measurement_pdf)``,
``predict(u, dt)``, ``update(u, z)``, ``resample()``, ``point_estimate()``,
``point_covariance()``, attributes ``particles`` / ``weights`` / ``N_particles``) over the
hand-written sm_100a kernels in csrc/ (through the C ABI of include/gse.h).

Particles live on the device as SoA float32 columns ``(5, ld)``; the ``particles`` attribute is the
``(N, 5)`` transposed view.  See filter/_base.py for the weight representation.
"""
import os

import numpy
import torch

from gpu_se_b200 import _device, _lib
from gpu_se_b200.filter._base import WeightedEnsemble, mixture_view
from gpu_se_b200.model.BioreactorModel import model_id_for


class ParallelParticleFilter(WeightedEnsemble):
    """Particle filter running on the GPU.

    Parameters are those of the reference (particle.py:118-149).  Extras, keyword-only:

    device : torch device (default: current CUDA device)
    seed : int, Philox key for the in-kernel process noise and the initial draw
    n_sub : int, explicit-Euler sub-steps per ``predict`` (the reference takes one, quirk Q1)
    particles : (N, 5) array, initial particles instead of ``x0.draw(N)``
    index0 : int, global index of row 0 when this filter is one shard of a larger population
    workspace_rows : int, size the library workspace for this many rows (>= N_particles)
    peer : bool, keep state and cumulative weights in memory other ranks can map (sharded runs)
    """

    NCOLS = 5

    def __init__(self, f, g, N_particles, x0, state_pdf, measurement_pdf, *, device=None, seed=None,
                 n_sub=1, particles=None, index0=0, workspace_rows=None, peer=False):
        self.f = f
        self.g = g
        self._model_id = model_id_for(f, g)
        self._init_ensemble(N_particles, state_pdf, measurement_pdf, device, seed, workspace_rows, peer)
        self._index0 = int(index0)     # global index of local row 0 (sharded runs): keys the Philox stream
        self._n_sub = int(n_sub)
        self._fuse = None              # predict + update as one kernel: asked of the library on first use
        n = self.N_particles
        if particles is not None:
            self.particles = particles
        elif hasattr(x0, "draw_host"):
            self.particles = x0.draw_host(n)                       # deterministic test double
        else:
            # particles = x0.draw(N)  (particle.py:49) -- drawn on the device, straight into SoA
            mix = mixture_view(x0).as_gse_mixture()
            _lib.check(_lib.lib.gse_mixture_draw(self._ctx.handle, mix, self._state.data_ptr(), self._ld, n,
                                                 self._seed, 0xFFFFFFFF, self._index0, self._stream()))

    # -- state attribute ---------------------------------------------------------------
    @property
    def particles(self):
        self._materialise()
        return _device.wrap(self._state[:, :self.N_particles].t())

    @particles.setter
    def particles(self, value):
        self._flush()
        n = self.N_particles
        if isinstance(value, torch.Tensor):
            v = value.detach().as_subclass(torch.Tensor).to(device=self.device, dtype=torch.float32)
        else:
            v = torch.as_tensor(numpy.ascontiguousarray(_device.to_numpy(value), dtype=numpy.float32),
                                device=self.device)
        if tuple(v.shape) != (n, 5):
            raise ValueError("particles must have shape (%d, 5)" % n)
        self._pending = False
        self._state[:, :n].copy_(v.t())
        self._touch()

    # -- the three stages ----------------------------------------------------------------
    def _predict_now(self, u, dt, noise=None):
        """particle.py:265-277.  ``noise`` (N, 5): host-supplied draws for cross-checks; by default
        the state noise comes from the in-kernel Philox stream (or from the state pdf itself when
        that is a DeterministicGaussianSum)."""
        n = self.N_particles
        if noise is None:
            noise = self._host_noise(self.state_pdf, n)
        nz_ptr, ld_nz, nz = None, 0, None
        if noise is not None:
            nz = torch.zeros((5, self._ld), dtype=torch.float32, device=self.device)
            nz[:, :n].copy_(torch.as_tensor(numpy.ascontiguousarray(_device.to_numpy(noise), dtype=numpy.float32)
                                            .reshape(n, 5), device=self.device).t())
            nz_ptr, ld_nz = nz.data_ptr(), self._ld
        # a pending resample is applied on the way in: read state[:, idx], write the other buffer
        dst = self._state_alt if self._pending else self._state
        _lib.check(_lib.lib.gse_pf_predict(self._ctx.handle, self._state.data_ptr(), self._ld, self._idx_ptr(),
                                           dst.data_ptr(), self._ld, n, _lib.as_double2(u), float(dt), self._n_sub,
                                           self._seed, self._step, self._index0, nz_ptr, ld_nz, self._stream()))
        if self._pending:
            self._state, self._state_alt = self._state_alt, self._state
            self._pending = False
        self._step += 1
        self._touch()

    def _fuses_update(self):
        """predict + update as one kernel.  Built, tested bit-identical to the two kernels, and OFF by default: at 2^24
        rows it takes 177 us + a 4 us merge against 136 + 47 us for the two -- the update is bound by instruction issue
        (log-pdf, max / sum-exp), which fusing does not remove, and predict has no issue slots to spare (DESIGN.md
        section 4).  GSE_FUSE_UPDATE=1 turns it on (it saves the re-read of the two measured columns: 8 B per row)."""
        if self._fuse is None:
            self._fuse = os.environ.get("GSE_FUSE_UPDATE", "0") == "1" and self.can_fuse_update()
        return self._fuse

    def can_fuse_update(self):
        return bool(_lib.lib.gse_pf_can_fuse_update(self._ctx.handle, self._n_sub, self._index0))

    def _predict_update_now(self, u, dt, z):
        """predict (particle.py:265-277) and update (:279-294) in one pass over the rows."""
        n = self.N_particles
        dst = self._state_alt if self._pending else self._state
        _lib.check(_lib.lib.gse_pf_predict_update(
            self._ctx.handle, self._state.data_ptr(), self._ld, self._idx_ptr(), dst.data_ptr(), self._ld, n,
            _lib.as_double2(u), float(dt), self._n_sub, self._seed, self._step, self._index0, _lib.as_double2(z),
            self._loglik_ptr(), self._loglik.data_ptr(), self._stats.data_ptr(), self._stream()))
        if self._pending:
            self._state, self._state_alt = self._state_alt, self._state
            self._pending = False
        self._step += 1
        self._after_update()

    def _update_now(self, u, z):
        """particle.py:279-294."""
        self._materialise()
        _lib.check(_lib.lib.gse_pf_update(self._ctx.handle, self._state.data_ptr(), self._ld, self.N_particles,
                                          self._loglik_ptr(), self._loglik.data_ptr(), _lib.as_double2(u),
                                          _lib.as_double2(z), self._stats.data_ptr(), self._stream()))
        self._after_update()

    # predict() / update() / resample(): WeightedEnsemble (eager, or recorded for a CUDA-graph replay)

    # -- estimates -----------------------------------------------------------------------
    MEAN_ONLY_KERNEL = True

    def _launch_moments(self, mean_only=False):
        _lib.check(_lib.lib.gse_pf_moments(
            self._ctx.handle, self._state.data_ptr(), self._ld, self.N_particles, self._idx_ptr(), self._loglik_ptr(),
            self._base.data_ptr() if self._base is not None else None, self._stats.data_ptr(), int(mean_only),
            self._mom.data_ptr(), self._stream()))

    # point_estimate(): WeightedEnsemble.point_estimate  (particle.py:318-320)

    def covariance_matrix(self, normalised=False):
        return self._scatter_about(normalised)[0]

    def point_covariance(self, normalised=False):
        """Largest singular value of the weighted scatter (particle.py:322-327)."""
        cov = self.covariance_matrix(normalised)
        return float(numpy.linalg.svd(cov, compute_uv=False)[0])


ParticleFilter = ParallelParticleFilter
