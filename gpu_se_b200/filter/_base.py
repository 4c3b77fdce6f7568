"""Shared host-side plumbing of the two filters: the library context, the weight representation, the systematic
resample (one fused scan + rank + fill kernel; scan -> merge-path search as the two-stage alternative) and the read-back
of the estimates.

Weights.  The reference keeps linear-domain weights (float32, float64 after the first resample,
SURVEY.md quirk Q3) and multiplies pdf values into them, which underflows for informative
measurements.  Here a weight is  ``base_scale * base_k * exp(loglik_k)``:

* ``loglik`` -- float32 device array, the log pdf accumulated by ``update`` since the last reset;
* ``base``   -- ``None`` (all ones) or a float64 device array holding weights a caller ASSIGNED
  (``p.weights = w``, results/pf_openloop/pf_run_seq.py:124-125) exactly as given, so that a
  resample of assigned weights works on the caller's float64 values bit for bit;
* ``base_scale`` -- Python float, ``1/N`` after construction / resample (particle.py:50,103).

The ``weights`` attribute materialises the reference's un-normalised linear weights on demand.

Lazy resampling.  ``resample()`` only produces the int32 ancestor index (the reference's
``sample_index``, particle.py:100): ``_pending`` is set and the next ``predict`` / moments kernel
reads row ``idx[i]`` of the pre-resample state for row ``i``, so ``particles[sample_index]``
(particle.py:102) costs no pass of its own through HBM.  Anything that needs the rows in place
(``particles`` / ``means`` attributes, ``update`` right after ``resample``) calls ``_materialise()``.
``_loglik_zero`` marks the accumulated log-likelihood as all zero (weights just reset, :103) so that
``update`` does not read it and nothing has to zero-fill it.

CUDA graphs.  Below ~2^18 rows a predict -> update -> resample cycle is launch-bound (six kernels of
a few microseconds each behind three Python calls).  ``enable_graphs()`` switches the three calls to
deferred execution: ``predict`` and ``update`` only record their arguments, and ``resample`` replays
a CUDA graph of the whole cycle captured once per state-buffer parity.  The per-step scalars (u, dt,
z, r, the Philox step counter) live in a 64-byte device block the kernels read
(``gse_ctx_set_step_params``), refreshed by one small H2D copy per step.  Any other access
(``particles``, ``point_estimate`` between the calls, host-supplied noise ...) flushes the recorded
calls eagerly first, so results are bit-identical with and without graphs.
"""
import ctypes
import math

import numpy
import torch

import os

from gpu_se_b200 import _device, _lib

# GSE_RESAMPLE=unfused: scan + merge-path search as two stages (A/B measurements, the sharded path's kernels)
FUSED_RESAMPLE = os.environ.get("GSE_RESAMPLE", "fused") != "unfused"


class Context:
    """Owns a gse_ctx (include/gse.h: gse_ctx_create / gse_ctx_destroy)."""

    def __init__(self, device, n_max, state_pdf, measurement_pdf):
        self.device = _device.resolve_device(device)
        state = (state_pdf.as_gse_mixture() if state_pdf is not None else
                 _lib.make_mixture(numpy.zeros((1, 5)), numpy.eye(5)[None], numpy.ones(1)))
        meas = (measurement_pdf.as_gse_mixture() if measurement_pdf is not None else
                _lib.make_mixture(numpy.zeros((1, 2)), numpy.eye(2)[None], numpy.ones(1)))
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.gse_ctx_create(self.device.index, _lib.GSE_MODEL_BIOREACTOR, int(n_max),
                                               ctypes.byref(state), ctypes.byref(meas), ctypes.byref(handle)))
        self.handle = handle
        self.n_max = int(n_max)
        # 64 doubles of host-mapped memory a kernel can leave a moment block in (no copy of its own on the way back)
        host, devp = _lib.c_dbl_p(), _lib.c_dbl_p()
        _lib.check(_lib.lib.gse_ctx_result_block(handle, ctypes.byref(host), ctypes.byref(devp)))
        self.result_np = numpy.ctypeslib.as_array(host, shape=(64,))
        self.result_dev = ctypes.cast(devp, ctypes.c_void_p).value
        self._err_bits = ctypes.c_uint(0)

    @property
    def launches(self):
        return int(_lib.lib.gse_launch_count(self.handle))

    def wait(self, stream):
        """Synchronise ``stream`` and raise what ``check_device_errors`` would (one library call for both)."""
        _lib.check(_lib.lib.gse_ctx_wait(self.handle, stream, ctypes.byref(self._err_bits)))
        if self._err_bits.value:
            self._raise_device_errors(self._err_bits.value)

    def check_device_errors(self):
        """Raise if a kernel of this context flagged an error since the last check (include/gse.h GSE_ERR_*).
        Call after a synchronisation: the word lives in host-mapped memory and is read without one."""
        bits = int(_lib.lib.gse_ctx_errors(self.handle, 1))
        if bits:
            self._raise_device_errors(bits)

    def _raise_device_errors(self, bits):
        if bits:
            msg = _lib.lib.gse_last_error().decode()
            if bits & (_lib.GSE_ERR_CHOLESKY | _lib.GSE_ERR_SINGULAR_PYY):
                raise numpy.linalg.LinAlgError(msg)          # what the reference raises (gs_ukf.py:72-75)
            if bits & _lib.GSE_ERR_ZERO_WEIGHTS:
                raise FloatingPointError(msg)
            raise _lib.GseError(msg)

    def close(self):
        if getattr(self, "handle", None):
            self.result_np = None                 # a view of memory the library frees now
            _lib.lib.gse_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mixture_view(pdf):
    """Accept the reference's own MultivariateGaussianSum objects (or anything with
    means / covariances / weights) as well as this package's."""
    if hasattr(pdf, "as_gse_mixture"):
        return pdf
    from gpu_se_b200.gaussian_sum_dist import MultivariateGaussianSum
    cov = getattr(pdf, "_covariances64", None)
    if cov is None:
        inv = getattr(pdf, "_inverse_covariances", None)      # the reference keeps float64 only here (:33)
        cov = numpy.linalg.inv(_device.to_numpy(inv)) if inv is not None else _device.to_numpy(pdf.covariances)
    return MultivariateGaussianSum(_device.to_numpy(pdf.means), cov, _device.to_numpy(pdf.weights))


class WeightedEnsemble:
    """N weighted rows of an SoA float32 state of NCOLS columns, with systematic resampling."""

    NCOLS = 5

    def _init_ensemble(self, N, state_pdf, measurement_pdf, device, seed, workspace_rows=None, peer=False):
        self.N_particles = int(N)
        if self.N_particles < 1:
            raise ValueError("N_particles must be >= 1")
        self.state_pdf = state_pdf
        self.measurement_pdf = measurement_pdf
        self._state_mix = mixture_view(state_pdf)
        self._meas_mix = mixture_view(measurement_pdf)
        self._ctx = Context(device, max(self.N_particles, int(workspace_rows or 0)), self._state_mix, self._meas_mix)
        self.device = self._ctx.device
        n = self.N_particles
        self._ld = _device.round_up(n, 64)
        dev = self.device
        self._peer_bufs = {}
        if peer:        # sharded run: buffers other ranks read live in IPC-exportable memory (gpu_se_b200/_peer.py)
            from gpu_se_b200 import _peer
            self._state, self._peer_bufs["state"] = _peer.peer_zeros(dev, (self.NCOLS, self._ld), torch.float32)
            self._state_alt, self._peer_bufs["state_alt"] = _peer.peer_zeros(dev, (self.NCOLS, self._ld), torch.float32)
        else:
            self._state = torch.zeros((self.NCOLS, self._ld), dtype=torch.float32, device=dev)
            self._state_alt = torch.zeros((self.NCOLS, self._ld), dtype=torch.float32, device=dev)
        self._loglik = torch.zeros(self._ld, dtype=torch.float32, device=dev)
        self._loglik_zero = True       # logically all zero; the buffer need not be
        self._idx = torch.zeros(self._ld, dtype=torch.int32, device=dev)      # ancestor index of the last resample
        self._pending = False          # rows must be read through self._idx
        self._base = None
        self._base_max = None
        self._base_scale = 1.0 / n
        self._loglik_dirty = False
        self._stats = torch.tensor([0.0, float(n), 0.0, 0.0], dtype=torch.float64, device=dev)
        self._stats_uniform = self._stats.clone()
        self._cumsum = None            # uint64 cumulative weights: only the two-stage scan / search path materialises them
        self._offtot = torch.zeros(2, dtype=torch.int64, device=dev)             # [offset, total]
        self._mom = torch.zeros(48, dtype=torch.float64, device=dev)
        self._mom_host = torch.zeros(48, dtype=torch.float64).pin_memory()
        self._mom_valid = False
        self._mom_full = False         # the cached moments include the second moments
        self._want_cov = False         # a caller asked for point_covariance before: compute both in one pass
        # point_estimate() straight after resample() -- every filter loop of the reference -- is folded into the resample
        # kernel once the filter has seen the pattern (it needs no second pass over the ancestor index)
        self._est_hint = False         # the caller reads the estimate of the resampled population
        self._fresh_resample = False   # nothing has touched the population since the last resample
        self._mom_from_resample = False    # the context's host-mapped result block holds that estimate
        self._mom_unused = False       # ... and nobody has asked for it yet
        self._seed = int(seed) if seed is not None else int(numpy.random.randint(0, 2 ** 31 - 1))
        self._step = 0
        self.last_sample_index = None
        self._stage_hook = None        # bench.py: called with a label between the kernels of resample()
        self._graph_mode = False
        self._deferred = []            # recorded ("predict", u, dt) / ("update", u, z) calls
        self._graphs = {}              # current state buffer -> captured cycle
        self.graph_replays = 0

    # -- CUDA graphs ---------------------------------------------------------------------
    def enable_graphs(self, enable=True):
        """Deferred execution + CUDA-graph replay of predict -> update -> resample (see the module
        docstring).  Worth it for launch-bound sizes; harmless otherwise."""
        self._flush()
        if enable and not self._graph_mode:
            self._params = _lib.gse_step_params()
        self._graph_mode = bool(enable)

    def _fuses_update(self):
        """predict() directly followed by update() runs as ONE kernel (the rows never leave the registers in between)."""
        return False

    def _predict_update_now(self, u, dt, z):
        raise NotImplementedError

    def predict(self, u, dt, noise=None):
        # recorded, not run: the call that follows decides whether it becomes part of a fused predict + update kernel
        # (or of a graph replay); anything else first runs it on its own (_flush)
        if ((self._graph_mode or self._fuses_update()) and noise is None and not self._deferred
                and not hasattr(self.state_pdf, "draw_host")):
            self._deferred = [("predict", (float(u[0]), float(u[1])), float(dt))]      # values, not references
            return
        self._flush()
        self._predict_now(u, dt, noise)

    def update(self, u, z):
        if self._graph_mode and len(self._deferred) == 1:
            self._deferred.append(("update", (float(u[0]), float(u[1])), (float(z[0]), float(z[1]))))
            return
        if len(self._deferred) == 1 and self._fuses_update():
            (_, up, dt), self._deferred = self._deferred[0], []
            self._predict_update_now(up, dt, (float(z[0]), float(z[1])))
            return
        self._flush()
        self._update_now(u, z)

    def _flush(self):
        """Run the recorded calls eagerly (something other than the graphed cycle is happening)."""
        if self._deferred:
            ops, self._deferred = self._deferred, []
            if len(ops) == 2 and self._fuses_update():
                self._predict_update_now(ops[0][1], ops[0][2], ops[1][2])
                return
            for op in ops:
                if op[0] == "predict":
                    self._predict_now(op[1], op[2], None)
                else:
                    self._update_now(op[1], op[2])

    def _replay_cycle(self, r):
        """predict -> update -> resample as one graph launch."""
        (_, u, dt), (_, _, z) = self._deferred
        self._deferred = []
        p = self._params
        p.u[0], p.u[1] = u
        p.dt = dt
        p.z[0], p.z[1] = z
        p.r = r
        p.step = self._step
        _lib.check(_lib.lib.gse_ctx_upload_step_params(self._ctx.handle, ctypes.byref(p), self._stream()))
        want_mean = self._mean_in_resample()
        key = (self._state.data_ptr(), want_mean)
        g = self._graphs.get(key)
        if g is None:
            # capture: the eager code path runs once into the capture stream (its host-side state
            # transitions are the real ones), reading the per-step scalars from the device block
            _lib.check(_lib.lib.gse_ctx_use_step_params(self._ctx.handle, 1))
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g):
                    if self._fuses_update():
                        self._predict_update_now(u, dt, z)
                    else:
                        self._predict_now(u, dt, None)
                        self._update_now(u, z)
                    self._resample_now(r, False)
            finally:
                _lib.check(_lib.lib.gse_ctx_use_step_params(self._ctx.handle, 0))
            self._graphs[key] = g
            g.replay()
        else:
            g.replay()
            # the host-side transitions of the cycle: the gathering predict swapped the buffers and
            # advanced the Philox step; resample left the weights uniform and the index pending
            self._state, self._state_alt = self._state_alt, self._state
            self._step += 1
            self._touch()
            self._fresh_resample = True
            self._mom_from_resample = self._mom_unused = want_mean
        self.graph_replays += 1

    # -- helpers -------------------------------------------------------------------------
    def _stream(self):
        return _device.stream_ptr(self.device)

    def _touch(self):
        self._mom_valid = False
        self._fresh_resample = False
        self._mom_from_resample = False

    def _idx_ptr(self):
        return self._idx.data_ptr() if self._pending else None

    def _loglik_ptr(self):
        """Device pointer of the accumulated log-likelihood, None while it is all zero."""
        return None if self._loglik_zero else self._loglik.data_ptr()

    def _ensure_loglik_buffer(self):
        if self._loglik_zero:
            self._loglik.zero_()
            self._loglik_zero = False
        return self._loglik.data_ptr()

    def _materialise(self):
        """Apply a pending resample: state <- state[:, idx]  (particles[sample_index], particle.py:102)."""
        self._flush()
        if self._pending:
            n = self.N_particles
            _lib.check(_lib.lib.gse_gather_rows(self._ctx.handle, self._idx.data_ptr(), n, self._state.data_ptr(),
                                                self._ld, self._state_alt.data_ptr(), self._ld, self.NCOLS, None,
                                                self._stream()))
            self._state, self._state_alt = self._state_alt, self._state
            self._pending = False

    def _host_noise(self, pdf, shape):
        """Noise rows for the host-noise mode when the state pdf is a deterministic test double."""
        if hasattr(pdf, "draw_host"):
            return numpy.asarray(pdf.draw_host(shape), dtype=numpy.float32)
        return None

    # -- weights -------------------------------------------------------------------------
    @property
    def weights(self):
        self._flush()
        n = self.N_particles
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib.gse_weights_linear(
            self._ctx.handle, self._loglik_ptr(), self._base.data_ptr() if self._base is not None else None,
            n, float(self._base_scale), out.data_ptr(), self._stream()))
        return _device.wrap(out)

    @weights.setter
    def weights(self, w):
        self._flush()
        n = self.N_particles
        if isinstance(w, torch.Tensor):
            w = w.detach().as_subclass(torch.Tensor).to(device=self.device, dtype=torch.float64).reshape(-1)
        else:
            w = torch.as_tensor(numpy.ascontiguousarray(_device.to_numpy(w), dtype=numpy.float64).reshape(-1),
                                device=self.device)
        if w.numel() != n:
            raise ValueError("weights must have %d entries" % n)
        self._base = w.contiguous().clone()
        self._base_max = self._base.max()
        self._base_scale = 1.0
        self._loglik_zero = True
        self._loglik_dirty = False
        self._stats[0] = 0.0
        self._stats[1] = self._base.sum()
        self._touch()

    def _reset_uniform(self, stats_done=False):
        """Uniform weights; ``stats_done``: the resample kernel has already stored (M, S) = (0, N)."""
        self._base = None
        self._base_max = None
        self._base_scale = 1.0 / self.N_particles
        self._loglik_dirty = False
        if not stats_done:
            self._stats.copy_(self._stats_uniform)

    def _after_update(self):
        self._loglik_dirty = True
        self._loglik_zero = False
        self._touch()

    # -- resample ------------------------------------------------------------------------
    def _scan(self):
        """cumsum of the fixed-point weights -> self._cumsum, total -> self._offtot[1]."""
        n = self.N_particles
        if self._cumsum is None:
            self._cumsum = torch.zeros(self._ld, dtype=torch.int64, device=self.device)  # uint64 payload
        ll, base, stats = self._weight_sources()
        _lib.check(_lib.lib.gse_scan_weights(self._ctx.handle, ll, base, stats.data_ptr(), n, self._cumsum.data_ptr(),
                                             self._offtot.data_ptr() + 8, self._stream()))

    def resample(self, r=None, return_index=False):
        """Systematic resample (particle.py:85-103 / gs_ukf.py:151-171).  ``r`` defaults to
        ``numpy.random.rand()`` exactly as the reference's CPU path draws it (:93), so seeding
        numpy reproduces the reference's offset.  Produces the ancestor index only; the rows move
        when the next kernel reads them (see the module docstring)."""
        if r is None:
            r = numpy.random.rand()
        r = float(r)
        if not (0.0 <= r < 1.0):
            raise ValueError("r must be in [0, 1)")
        if (self._graph_mode and len(self._deferred) == 2 and not return_index and self._pending
                and self._base is None and self._stage_hook is None):
            self._replay_cycle(r)
            return None
        self._flush()
        return self._resample_now(r, return_index)

    def _weight_sources(self):
        """(loglik pointer or None, base pointer or None, stats tensor) the scan kernels quantise."""
        use_loglik = self._base is None or self._loglik_dirty
        stats = self._stats
        if self._base is not None and self._loglik_dirty:
            # S bounds sum exp(loglik - M); the scale must bound sum base * exp(loglik - M)
            stats = self._stats.clone()
            stats[1] = stats[1] * self._base_max
        return (self._ensure_loglik_buffer() if use_loglik else None,
                self._base.data_ptr() if self._base is not None else None, stats)

    def _mean_in_resample(self):
        """Should the next resample also produce the estimate of the resampled population?  Yes once a caller has asked
        for ``point_estimate()`` right after a resample, until an estimate computed that way goes unread."""
        if self._mom_unused:
            self._est_hint = False
        return (FUSED_RESAMPLE and self.MEAN_ONLY_KERNEL and self._est_hint and not self._want_cov
                and self._base is None)

    def _resample_now(self, r, return_index):
        n = self.N_particles
        self._materialise()                       # a second resample without a predict in between
        want_mean = stats_done = False
        if FUSED_RESAMPLE:
            # scan + rank + fill in one launch: the cumulative weights never reach HBM (csrc/gse_resample_fused.cu)
            want_mean = self._mean_in_resample()
            ll, base, stats = self._weight_sources()
            stats_done = stats is self._stats     # the kernel leaves (M, S) = (0, N) behind: no launch for that
            _lib.check(_lib.lib.gse_resample_fused(
                self._ctx.handle, ll, base, stats.data_ptr(), n, r, n, 0, n, 0, self._idx.data_ptr(),
                self._offtot.data_ptr() + 8, self._state.data_ptr() if want_mean else None, self._ld,
                self._ctx.result_dev if want_mean else None, int(stats_done), self._stream()))      # estimate -> host-mapped block
        else:
            self._scan()
            if self._stage_hook is not None:
                self._stage_hook("scan")
            _lib.check(_lib.lib.gse_resample_search(
                self._ctx.handle, self._cumsum.data_ptr(), n, self._offtot.data_ptr(), r, n, 0, n,
                self._idx.data_ptr(), self._stream()))
        self._pending = True
        self._loglik_zero = True                  # weights = 1/N  (:103 / :316)
        self._reset_uniform(stats_done)
        self._touch()
        self._fresh_resample = True
        self._mom_from_resample = self._mom_unused = want_mean
        idx = self._idx[:n].to(torch.int64) if return_index else None
        self.last_sample_index = idx
        return idx

    def cumulative_weights(self):
        """(cumsum uint64 as numpy, total) of the current weights -- test / diagnostics hook."""
        self._flush()
        self._scan()
        c = self._cumsum[:self.N_particles].cpu().numpy().view(numpy.uint64)
        return c, int(c[-1])

    # -- moments -------------------------------------------------------------------------
    MEAN_ONLY_KERNEL = False

    def _launch_moments(self, mean_only=False):
        raise NotImplementedError

    def _moments(self, need_cov=False):
        """Weighted moments of the population, cached until the state changes.  point_estimate alone
        runs the means-only kernel; once a caller has asked for point_covariance (sim_base.py:291-295
        does both every step) both are computed in one pass."""
        self._flush()
        if need_cov:
            self._want_cov = True
        if not self._mom_valid or (need_cov and not self._mom_full):
            full = need_cov or self._want_cov or not self.MEAN_ONLY_KERNEL
            if self._fresh_resample and not full:
                self._est_hint = True
            if self._mom_from_resample and not full:
                # the resample kernel has written the estimate (and M, S) straight into the context's host-mapped
                # result block: nothing to launch, nothing to copy -- wait for the stream and read it
                self._mom_unused = False
                self._ctx.wait(self._stream())
                self._mom_np = self._ctx.result_np[:48].copy()
            else:
                self._launch_moments(mean_only=not full)
                self._mom[41:43].copy_(self._stats[0:2])    # M, S ride along in the same read-back
                self._mom_host.copy_(self._mom, non_blocking=True)
                self._ctx.wait(self._stream())
                self._mom_np = self._mom_host.numpy().copy()
            self._mom_full = full
            self._mom_valid = True
        return self._mom_np

    def _weight_prefactor(self, mom):
        """A with  true weight_k = A * w_k,  w_k = base_k * exp(loglik_k - M)  the kernel's weights."""
        return self._base_scale * math.exp(self._M_host(mom))

    def _M_host(self, mom):
        return float(mom[41])

    @staticmethod
    def _unpack_sym(v):
        m = numpy.zeros((5, 5))
        t = 0
        for i in range(5):
            for j in range(i + 1):
                m[i, j] = m[j, i] = v[t]
                t += 1
        return m

    def _estimate_parts(self, need_cov=False):
        mom = self._moments(need_cov)
        S0, S1, S2 = mom[0], mom[1:6], (self._unpack_sym(mom[6:21]) if need_cov else None)
        p = mom[21:26]
        A = self._weight_prefactor(mom)
        return mom, S0, S1, S2, p, A

    def point_estimate(self, normalised=False):
        """``weights @ particles`` with the reference's un-normalised weights (quirk Q4); pass
        ``normalised=True`` for the weighted mean."""
        mom, S0, S1, S2, p, A = self._estimate_parts()
        if normalised:
            return p + S1 / S0
        return A * (S0 * p + S1)

    def _scatter_about(self, normalised):
        mom, S0, S1, S2, p, A = self._estimate_parts(need_cov=True)
        if normalised:
            d = S1 / S0
            return (S2 / S0 - numpy.outer(d, d)), mom, 1.0 / S0
        mu = A * (S0 * p + S1)                 # the reference centres on the un-normalised estimate (:111)
        d = mu - p
        cov = A * (S2 - numpy.outer(S1, d) - numpy.outer(d, S1) + S0 * numpy.outer(d, d))
        return cov, mom, A
