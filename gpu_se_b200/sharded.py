"""Particle filter / GS-UKF sharded over the GPUs of one node: one process per GPU (``torch.distributed`` for the
rendezvous), contiguous shards of the global row index space, and a globally consistent systematic resample.

The reference has no multi-GPU path; this is the north-star design:

* ``predict`` / ``update`` / moments touch local rows only.  The Philox stream is keyed by the GLOBAL row index, so
  a sharded run draws exactly the noise a single-GPU run of the same seed draws.
* after ``update`` the local ``(M_s, S_s)`` = (max log-likelihood, sum exp(loglik - M_s)) are all-gathered and merged
  so that every shard quantises its weights with the same fixed-point scale -- the cumulative weights are integers,
  hence independent of how the rows are split over GPUs.
* ``resample`` (``exchange="peer"``, the default): ONE kernel per rank (``gse_resample_fused_sharded``).  Every rank
  scans its own rows; the shard totals cross NVLink inside the kernel (peer mailboxes); every rank then ranks its rows
  against the global total and WRITES the global ancestor row of each output it sources straight into the index
  buffer of the shard that owns the output slot (peer stores); a second mailbox exchange at the end of the kernel
  makes every index buffer complete when its kernel completes.  The rows themselves move lazily: the next ``predict``
  / moments kernel pulls row ``idx[i]`` out of whichever GPU holds it.  No NCCL call, no host synchronisation.
* estimates: every rank reduces its shard to a 48-double moment block; the blocks are all-gathered through the
  mailboxes and merged on the device (``gse_peer_allgather_moments``) straight into the context's host-mapped result
  block: the read-back is one stream synchronisation, no copy.

  ``exchange="slabs"`` is the host-planned alternative for GPUs without peer access (and the cross-check of the peer
  path): all-gather of the totals over NCCL -> closed-form output ranges per source shard
  (``gse_count_outputs_below``) -> two-stage scan / search per range -> grouped NCCL send/recv of contiguous column
  slabs received in place.  It needs one device-to-host read per resample.

Ordering between ranks in peer mode.  Every rank owns two state buffers X (current) and Y and one ancestor-index
buffer I.  Peers read X (through I) in ``predict`` / moments and write I in ``resample``.  A rank leaves a mailbox
exchange only after every peer has entered it, and a peer enters it after everything its stream ran before.  Every
``resample`` contains two exchanges (totals; done) and every ``update`` one (stats), hence: (1) a peer writes this
rank's I only after the totals exchange of that resample, which this rank enters after its own earlier readers of I
(predict, moments, gather) have finished; (2) I is complete when the resample kernel completes (the done exchange);
(3) this rank's X is overwritten by its predict after next, which follows a resample whose exchanges every peer
only joins once its own predict -- the last reader of X -- has finished.  No sequence of calls (resample -> resample,
``set_global_weights(); resample()`` repeated, ...) needs an extra barrier.

``plan_resample`` and ``exchange_columns`` are pure host / ``torch.distributed`` code and run on CPU tensors over
``gloo`` as well (tests/test_sharded_cpu.py).
"""
import ctypes

import numpy
import torch
import torch.distributed as dist

from gpu_se_b200 import _device, _lib
from gpu_se_b200.filter.gs_ukf import ParallelGaussianSumUnscentedKalmanFilter
from gpu_se_b200.filter.particle import ParallelParticleFilter


def shard_bounds(n_total, world):
    """Contiguous split of [0, n_total) into shards of whole groups of four rows (the predict kernel draws
    the noise of a group of four global rows together), as even as that allows; the last shard takes
    the remainder."""
    n_total, world = int(n_total), int(world)
    groups = n_total // 4
    base, rem = divmod(groups, world)
    bounds, lo = [], 0
    for s in range(world):
        hi = lo + 4 * (base + (1 if s < rem else 0))
        if s == world - 1:
            hi = n_total
        bounds.append((lo, hi))
        lo = hi
    if any(b <= a for a, b in bounds):
        raise ValueError("too few rows (%d) for %d shards" % (n_total, world))
    return bounds


class ResamplePlan:
    """Who sources which global outputs, and which slabs move between shards.

    offsets[s], total : exclusive prefix of the shard totals and their sum (Python ints)
    src_ranges[s]     : (a_s, b_s) -- outputs sourced by shard s
    transfers         : list of (src, dst, start, stop) global output ranges, src-major, ascending
    """

    def __init__(self, totals, r, n_total, bounds):
        totals = [int(t) for t in totals]
        world = len(totals)
        self.offsets, acc = [], 0
        for t in totals:
            self.offsets.append(acc)
            acc += t
        self.total = acc
        if self.total <= 0:
            raise FloatingPointError("all weights are zero: cannot resample")
        cnt = _lib.lib.gse_count_outputs_below
        self.src_ranges = []
        for s in range(world):
            a = 0 if s == 0 else int(cnt(self.offsets[s], self.total, float(r), int(n_total)))
            b = int(n_total) if s == world - 1 else int(cnt(self.offsets[s] + totals[s], self.total, float(r),
                                                            int(n_total)))
            self.src_ranges.append((a, max(a, b)))
        self.transfers = []
        for s, (a, b) in enumerate(self.src_ranges):
            for t, (lo, hi) in enumerate(bounds):
                start, stop = max(a, lo), min(b, hi)
                if stop > start:
                    self.transfers.append((s, t, start, stop))

    def exchanged_rows(self):
        return sum(stop - start for s, t, start, stop in self.transfers if s != t)


def plan_resample(totals, r, n_total, bounds):
    return ResamplePlan(totals, r, n_total, bounds)


def exchange_columns(plan, rank, bounds, staging, staging_start, dst, ncols, group=None):
    """Ship the slabs of ``plan`` that cross shards.  ``staging`` (ncols, >= rows sourced here for
    other shards) holds the gathered rows of the remote-bound outputs in ascending output order,
    starting with output ``staging_start[t]`` at column offset ``staging_off[t]`` per destination
    (see ``staging_layout``); ``dst`` (ncols, ld) is the local destination state whose column j
    holds global output ``bounds[rank][0] + j``.  Receives land in place."""
    ops = []
    lo = bounds[rank][0]
    for s, t, start, stop in plan.transfers:
        if s == t:
            continue
        cnt = stop - start
        if s == rank:
            off = staging_start[t]
            for c in range(ncols):
                ops.append(dist.P2POp(dist.isend, staging[c, off:off + cnt], t, group))
        elif t == rank:
            for c in range(ncols):
                ops.append(dist.P2POp(dist.irecv, dst[c, start - lo:stop - lo], s, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def staging_layout(plan, rank, align=64):
    """Column offset in the staging buffer of each remote destination's slab, and the total width."""
    offs, width = {}, 0
    for s, t, start, stop in plan.transfers:
        if s == rank and t != rank:
            offs[t] = width
            width += _device.round_up(stop - start, align)
    return offs, width


class _ShardedEnsemble:
    """A weighted ensemble (particles, or Gaussian components) over all ranks of ``group``: the constructor and the
    ``predict / update / resample / point_estimate / point_covariance`` calls of the single-GPU class, every rank
    calling each method collectively with identical arguments.  ``N_particles`` is the GLOBAL count; the state
    attributes are this rank's shard.  "Collectively" includes reading the state attributes: that applies a pending
    resample, which flips the state buffers every rank's kernels read, so all ranks must do it at the same point of
    the call sequence."""

    LOCAL_CLS = None

    def __init__(self, f, g, N_particles, x0, state_pdf, measurement_pdf, *, device=None, seed=0, group=None,
                 exchange="peer", **local_kw):
        if not dist.is_initialized():
            raise RuntimeError("a sharded filter needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.N_particles = int(N_particles)
        if self.N_particles < self.world:
            raise ValueError("need at least one row per shard")
        self.bounds = shard_bounds(self.N_particles, self.world)
        lo, hi = self.bounds[self.rank]
        if exchange not in ("peer", "slabs"):
            raise ValueError("exchange must be 'peer' or 'slabs'")
        if self.world > _lib.GSE_MAX_SHARDS:
            raise ValueError("at most %d shards" % _lib.GSE_MAX_SHARDS)
        self.exchange = exchange
        for key in ("particles", "means"):                    # initial state given for the whole population
            if local_kw.get(key) is not None:
                local_kw[key] = _device.to_numpy(local_kw[key])[lo:hi]
        # the fused resample may source every output of the population from this shard (heavy-run queue)
        self.local = self.LOCAL_CLS(f, g, hi - lo, x0, state_pdf, measurement_pdf, device=device, seed=seed, index0=lo,
                                    workspace_rows=self.N_particles, peer=(exchange == "peer"), **local_kw)
        self.device = self.local.device
        self._set_uniform(init=True)
        self._pending = False
        if exchange == "peer":
            self._open_peers()
        self.last_plan = None
        self.exchanged_rows = 0
        self._stage_hook = None
        self._want_cov = False
        self._deferred_predict = None        # (u, dt) of a predict that may still fuse with the update that follows

    # -- peer memory -----------------------------------------------------------------------------
    def _open_peers(self):
        """Exchange the IPC handles of (state, state_alt, ancestor index, mailbox) and map every other rank's."""
        from gpu_se_b200 import _peer
        loc = self.local
        if self.N_particles > 2 ** 31 - 17:
            raise ValueError("peer exchange indexes global rows with int32")
        self._mailbox = _peer.PeerBuffer(self.device, _lib.GSE_MAILBOX_BYTES)
        self._idx_global, self._idx_buf = _peer.peer_zeros(self.device, (_device.round_up(loc.N_particles, 64),),
                                                           torch.int32)
        mine = (loc._peer_bufs["state"].handle, loc._peer_bufs["state_alt"].handle, self._idx_buf.handle, loc._ld,
                self._mailbox.handle)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        self._mappings = []
        self._peer_ptr = []                  # per rank: (state, state_alt, idx) device pointers, ld
        boxes = (ctypes.c_void_p * _lib.GSE_MAX_SHARDS)()
        for s, (h0, h1, hi_, ld, hm) in enumerate(everyone):
            if s == self.rank:
                self._peer_ptr.append((loc._peer_bufs["state"].ptr, loc._peer_bufs["state_alt"].ptr, self._idx_buf.ptr, ld))
                boxes[s] = self._mailbox.ptr
            else:
                maps = [_peer.PeerMapping(self.device, h) for h in (h0, h1, hi_, hm)]
                self._mappings += maps
                self._peer_ptr.append((maps[0].ptr, maps[1].ptr, maps[2].ptr, ld))
                boxes[s] = maps[3].ptr
        self._boxes = boxes
        self._epoch = 0                      # mailbox exchanges so far: identical on every rank
        self._parity = 0                     # which of (state, state_alt) is current -- flips on every rank together
        self._shards = []
        for parity in (0, 1):
            sh = _lib.gse_shards()
            sh.nshards = self.world
            sh.rank = self.rank
            for s, (a, b) in enumerate(self.bounds):
                sh.rows[s] = a
                sh.state_dev[s] = self._peer_ptr[s][parity]
                sh.idx_dev[s] = self._peer_ptr[s][2]
                sh.ld[s] = self._peer_ptr[s][3]
            sh.rows[self.world] = self.N_particles
            self._shards.append(sh)
        dist.barrier(group=self.group)

    def _next_epoch(self):
        """Sequence number of the next mailbox exchange: 1, 2, ..., 0xFFFFFFFE, 1, ... (never 0; the parity of
        consecutive numbers alternates across the wrap, which the double-buffered slots rely on)."""
        self._epoch = self._epoch % 0xFFFFFFFE + 1
        return self._epoch

    def close(self):
        """Unmap the other ranks' buffers (collective: every rank must call it before any frees its own)."""
        self._flush()
        if getattr(self, "_mappings", None):
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            for m in self._mappings:
                m.close()
            self._mappings = []
            dist.barrier(group=self.group)

    # -- resample --------------------------------------------------------------------------------
    def _resample_peer(self, r, return_index):
        loc = self.local
        want_mean = loc._mean_in_resample()      # the estimate of the resampled population out of the same kernel
        ll, base, stats = loc._weight_sources()
        stats_done = stats is loc._stats
        e1, e2 = self._next_epoch(), self._next_epoch()
        sh = self._shards[self._parity]
        _lib.check(_lib.lib.gse_resample_fused_sharded(loc._ctx.handle, ll, base, stats.data_ptr(), r, ctypes.byref(sh),
                                                       self._boxes, self.rank, e1, e2, loc._offtot.data_ptr() + 8,
                                                       loc._state.data_ptr() if want_mean else None, loc._ld,
                                                       loc._mom.data_ptr() if want_mean else None, int(stats_done),
                                                       loc._stream()))
        # lazy, as on one GPU: the rows move when the next kernel reads them (predict / moments pull them
        # out of the owning shard's memory through the global ancestor index)
        self._pending = True
        loc._loglik_zero = True
        loc._reset_uniform(stats_done)
        self._set_uniform()
        loc._touch()
        loc._fresh_resample = True
        loc._mom_from_resample = loc._mom_unused = want_mean
        self.exchanged_rows = None                                # known on the device only: see rows_from_peers()
        return self._idx_global[:loc.N_particles].to(torch.int64) if return_index else None

    def _materialise(self):
        """Apply a pending sharded resample: pull the ancestors' rows into this shard's other buffer."""
        self._flush()
        if self._pending:
            loc = self.local
            sh = self._shards[self._parity]
            _lib.check(_lib.lib.gse_gather_rows_sharded(loc._ctx.handle, ctypes.byref(sh),
                                                        self._idx_global.data_ptr(), loc.N_particles,
                                                        loc._state_alt.data_ptr(), loc._ld, loc.NCOLS, loc._stream()))
            self._swap()

    def _swap(self):
        loc = self.local
        loc._state, loc._state_alt = loc._state_alt, loc._state
        self._parity ^= 1
        self._pending = False
        loc._touch()

    def rows_from_peers(self):
        """Rows of the last resample whose ancestor lives on another GPU (device-to-host read)."""
        if self.exchange != "peer":
            return self.exchanged_rows
        lo, hi = self.bounds[self.rank]
        idx = self._idx_global[:self.local.N_particles]
        return int(((idx < lo) | (idx >= hi)).sum().item())

    # -- weights ---------------------------------------------------------------------------------
    def _set_uniform(self, init=False):
        """Uniform weights over the WHOLE population (the local class's reset has just assumed its own row count)."""
        loc = self.local
        loc._base_scale = 1.0 / self.N_particles
        if init:
            loc._stats_uniform[1] = float(self.N_particles)       # from here on loc._reset_uniform() copies the global S
            loc._stats.copy_(loc._stats_uniform)

    @property
    def weights(self):
        self._flush()
        return self.local.weights

    def set_global_weights(self, w):
        """Assign weights from the full (N,) global array (every rank passes the same array)."""
        self._flush()
        lo, hi = self.bounds[self.rank]
        w = numpy.ascontiguousarray(_device.to_numpy(w), dtype=numpy.float64).reshape(-1)
        if w.size != self.N_particles:
            raise ValueError("weights must have %d entries" % self.N_particles)
        full = torch.as_tensor(w, device=self.device)
        self.local.weights = full[lo:hi]
        self.local._stats[1] = full.sum()              # the same bound on every shard -> one scale
        self.local._base_max = full.max()

    def _allreduce_stats(self):
        """Global (M, S) from the shards' (M_s, S_s): one all-gather of a pair + the merge, one tiny kernel."""
        loc = self.local
        if self.world == 1:
            return
        if self.exchange == "peer":
            _lib.check(_lib.lib.gse_peer_allgather_stats(loc._ctx.handle, self._boxes, self.rank, self.world,
                                                         self._next_epoch(), loc._stats.data_ptr(), loc._stream()))
            return
        if not hasattr(self, "_stat_pairs"):
            self._stat_pairs = torch.zeros(2 * self.world, dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(self._stat_pairs, loc._stats[0:2], group=self.group)
        _lib.check(_lib.lib.gse_merge_stats(loc._ctx.handle, self._stat_pairs.data_ptr(), self.world,
                                            loc._stats.data_ptr(), loc._stream()))

    # -- the three stages ------------------------------------------------------------------------
    def _predict_pending(self, u, dt, noise_ptr, ld_noise):
        raise NotImplementedError

    def _noise_rows(self, noise):
        raise NotImplementedError

    def _fuses_update(self):
        return False

    def _predict_update_pending(self, u, dt, z):
        raise NotImplementedError

    def _flush(self):
        """Run a recorded predict on its own: something other than ``update`` followed it."""
        d, self._deferred_predict = self._deferred_predict, None
        if d is not None:
            self._predict_now(d[0], d[1], None)

    def predict(self, u, dt, noise=None):
        self._flush()
        if noise is None and self._fuses_update() and not hasattr(self.local.state_pdf, "draw_host"):
            self._deferred_predict = ((float(u[0]), float(u[1])), float(dt))
            return
        self._predict_now(u, dt, noise)

    def _predict_now(self, u, dt, noise):
        lo, hi = self.bounds[self.rank]
        if noise is not None:
            noise = _device.to_numpy(noise)[lo:hi]
        loc = self.local
        if not self._pending:
            loc._flush()
            loc._predict_now(u, dt, noise)
            return
        # pending sharded resample: read row idx[i] out of whichever GPU holds it, write the other buffer
        nz_ptr, ld_nz, keep = self._noise_rows(noise)
        self._predict_pending(u, dt, nz_ptr, ld_nz)
        del keep
        loc._step += 1
        self._swap()

    def update(self, u, z):
        d, self._deferred_predict = self._deferred_predict, None
        loc = self.local
        if d is not None:
            # predict + update in one pass over the rows (through the pending ancestor index, if there is one)
            if self._pending:
                self._predict_update_pending(d[0], d[1], z)
                loc._step += 1
                self._swap()
                loc._after_update()
            else:
                loc._flush()
                loc._predict_update_now(d[0], d[1], (float(z[0]), float(z[1])))
        else:
            self._materialise()
            loc._flush()
            loc._update_now(u, z)
        self._allreduce_stats()

    def resample(self, r=None, return_index=False):
        loc = self.local
        self._flush()
        self._materialise()
        if r is None:
            rt = torch.tensor([numpy.random.rand()], dtype=torch.float64, device=self.device)
            dist.broadcast(rt, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                           group=self.group)
            r = float(rt.item())
        r = float(r)
        if not (0.0 <= r < 1.0):
            raise ValueError("r must be in [0, 1)")
        if self.exchange == "peer":
            return self._resample_peer(r, return_index)
        n_loc = loc.N_particles
        lo, hi = self.bounds[self.rank]
        loc._scan()                                               # local cumsum, T_s -> _offtot[1]
        if self._stage_hook is not None:
            self._stage_hook("scan")
        mine = loc._offtot[1:2].clone()
        allt = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allt, mine, group=self.group)
        totals = [int(t) for t in torch.cat(allt).cpu().tolist()]
        plan = plan_resample(totals, r, self.N_particles, self.bounds)
        self.last_plan = plan
        loc._offtot.copy_(torch.tensor([plan.offsets[self.rank], plan.total], dtype=torch.int64))
        offs, width = staging_layout(plan, self.rank)
        ncols = loc.NCOLS
        staging = torch.empty((ncols, max(width, 1)), dtype=torch.float32, device=self.device)
        idx = torch.full((n_loc,), -1, dtype=torch.int64, device=self.device) if return_index else None
        widest = max([stop - start for s, t, start, stop in plan.transfers if s == self.rank] + [4])
        scratch = torch.empty(_device.round_up(widest, 64), dtype=torch.int32, device=self.device)
        for s, t, start, stop in plan.transfers:
            if s != self.rank:
                continue
            cnt = stop - start
            if t == self.rank:
                dst_ptr, ld_dst = loc._state_alt.data_ptr() + 4 * (start - lo), loc._ld
            else:
                dst_ptr, ld_dst = staging.data_ptr() + 4 * offs[t], staging.shape[1]
            _lib.check(_lib.lib.gse_resample_search(
                loc._ctx.handle, loc._cumsum.data_ptr(), n_loc, loc._offtot.data_ptr(), r, self.N_particles, start,
                cnt, scratch.data_ptr(), loc._stream()))
            _lib.check(_lib.lib.gse_gather_rows(
                loc._ctx.handle, scratch.data_ptr(), cnt, loc._state.data_ptr(), loc._ld, dst_ptr, ld_dst, ncols,
                None, loc._stream()))
            if idx is not None and t == self.rank:
                idx[start - lo:stop - lo] = scratch[:cnt].to(torch.int64) + lo      # global ancestor index
        exchange_columns(plan, self.rank, self.bounds, staging, offs, loc._state_alt, ncols, self.group)
        self.exchanged_rows = plan.exchanged_rows()
        loc._state, loc._state_alt = loc._state_alt, loc._state
        loc._loglik_zero = True
        loc._reset_uniform()
        self._set_uniform()
        loc._touch()
        return idx

    # -- estimates -------------------------------------------------------------------------------
    def _launch_local_moments(self, mean_only):
        raise NotImplementedError

    def _global_moments(self, need_cov=True):
        """(mom, S0, S1, S2, p, A) of the WHOLE population, identical on every rank."""
        loc = self.local
        self._flush()
        if need_cov:
            self._want_cov = True
        mean_only = not (need_cov or self._want_cov)
        loc._want_cov = self._want_cov
        if loc._fresh_resample and mean_only and loc.MEAN_ONLY_KERNEL:
            loc._est_hint = True
        if loc._mom_from_resample and mean_only:
            loc._mom_unused = False          # this rank's block (the outputs it sources) came out of the resample kernel
            loc._mom_from_resample = False   # ... and the merge below overwrites it
        else:
            self._launch_local_moments(mean_only)
        if self.exchange == "peer":
            # the shards' moment blocks cross NVLink inside one single-warp kernel that also merges them and leaves the
            # result (with the global M, S) in the context's host-mapped result block: no copy of its own on the way back
            ctx = loc._ctx
            if self.world > 1:
                _lib.check(_lib.lib.gse_peer_allgather_moments(ctx.handle, self._boxes, self.rank, self.world,
                                                               self._next_epoch(), loc._mom.data_ptr(),
                                                               loc._stats.data_ptr(), ctx.result_dev, loc._stream()))
                ctx.wait(loc._stream())
                mom = ctx.result_np[:48].copy()
            else:
                loc._mom[41:43].copy_(loc._stats[0:2])
                loc._mom_host.copy_(loc._mom, non_blocking=True)
                ctx.wait(loc._stream())
                mom = loc._mom_host.numpy().copy()
            S0, S1, p = mom[0], mom[1:6].copy(), mom[21:26].copy()
            S2 = loc._unpack_sym(mom[6:21]) if not mean_only else None
        else:
            loc._mom[41:43].copy_(loc._stats[0:2])
            allm = [torch.empty_like(loc._mom) for _ in range(self.world)]
            dist.all_gather(allm, loc._mom, group=self.group)
            moms = torch.stack(allm).cpu().numpy()
            loc._ctx.check_device_errors()
            p = moms[0, 21:26].copy()                             # common pivot: rank 0's row 0
            S0, S1, S2, X = 0.0, numpy.zeros(5), numpy.zeros((5, 5)), numpy.zeros(15)
            for m in moms:
                s0, s1, s2 = m[0], m[1:6], loc._unpack_sym(m[6:21])
                d = m[21:26] - p
                S2 = S2 + s2 + numpy.outer(s1, d) + numpy.outer(d, s1) + s0 * numpy.outer(d, d)
                S1 = S1 + s1 + s0 * d
                S0 = S0 + s0
                X = X + m[26:41]
            mom = moms[0].copy()
            mom[26:41] = X
        A = loc._base_scale * float(numpy.exp(mom[41]))
        return mom, S0, S1, S2, p, A

    def point_estimate(self, normalised=False):
        mom, S0, S1, S2, p, A = self._global_moments(need_cov=False)
        return p + S1 / S0 if normalised else A * (S0 * p + S1)

    def _scatter(self, normalised):
        mom, S0, S1, S2, p, A = self._global_moments()
        if normalised:
            d = S1 / S0
            return S2 / S0 - numpy.outer(d, d), mom, 1.0 / S0
        d = A * (S0 * p + S1) - p
        return A * (S2 - numpy.outer(S1, d) - numpy.outer(d, S1) + S0 * numpy.outer(d, d)), mom, A

    def covariance_matrix(self, normalised=False):
        return self._scatter(normalised)[0]

    def point_covariance(self, normalised=False):
        return float(numpy.linalg.svd(self.covariance_matrix(normalised), compute_uv=False)[0])

    @property
    def _ctx(self):
        return self.local._ctx


class ShardedParticleFilter(_ShardedEnsemble):
    """``ParallelParticleFilter`` (filter/particle.py:151-327) over all ranks of ``group``; ``particles`` / ``weights``
    are this rank's shard."""

    LOCAL_CLS = ParallelParticleFilter

    def __init__(self, f, g, N_particles, x0, state_pdf, measurement_pdf, *, device=None, seed=0, n_sub=1, group=None,
                 particles=None, exchange="peer"):
        super().__init__(f, g, N_particles, x0, state_pdf, measurement_pdf, device=device, seed=seed, group=group,
                         exchange=exchange, n_sub=n_sub, particles=particles)

    @property
    def particles(self):
        self._materialise()
        return self.local.particles

    def _noise_rows(self, noise):
        loc = self.local
        n = loc.N_particles
        if noise is None:
            noise = loc._host_noise(loc.state_pdf, n)
        if noise is None:
            return None, 0, None
        nz = torch.zeros((5, loc._ld), dtype=torch.float32, device=self.device)
        nz[:, :n].copy_(torch.as_tensor(numpy.ascontiguousarray(_device.to_numpy(noise), dtype=numpy.float32)
                                        .reshape(n, 5), device=self.device).t())
        return nz.data_ptr(), loc._ld, nz

    def _predict_pending(self, u, dt, nz_ptr, ld_nz):
        loc = self.local
        sh = self._shards[self._parity]
        _lib.check(_lib.lib.gse_pf_predict_sharded(loc._ctx.handle, ctypes.byref(sh), self._idx_global.data_ptr(),
                                                   loc._state_alt.data_ptr(), loc._ld, loc.N_particles,
                                                   _lib.as_double2(u), float(dt), loc._n_sub, loc._seed, loc._step,
                                                   loc._index0, nz_ptr, ld_nz, loc._stream()))

    def _fuses_update(self):
        return self.exchange == "peer" and self.local._fuses_update()

    def _predict_update_pending(self, u, dt, z):
        loc = self.local
        sh = self._shards[self._parity]
        _lib.check(_lib.lib.gse_pf_predict_update_sharded(
            loc._ctx.handle, ctypes.byref(sh), self._idx_global.data_ptr(), loc._state_alt.data_ptr(), loc._ld,
            loc.N_particles, _lib.as_double2(u), float(dt), loc._n_sub, loc._seed, loc._step, loc._index0,
            _lib.as_double2(z), loc._loglik_ptr(), loc._loglik.data_ptr(), loc._stats.data_ptr(), loc._stream()))

    def _launch_local_moments(self, mean_only):
        loc = self.local
        if self._pending:
            sh = self._shards[self._parity]
            _lib.check(_lib.lib.gse_pf_moments_sharded(
                loc._ctx.handle, ctypes.byref(sh), self._idx_global.data_ptr(), loc.N_particles, loc._loglik_ptr(),
                loc._base.data_ptr() if loc._base is not None else None, loc._stats.data_ptr(), int(mean_only),
                loc._mom.data_ptr(), loc._stream()))
        else:
            loc._launch_moments(mean_only=mean_only)


class ShardedGaussianSumUnscentedKalmanFilter(_ShardedEnsemble):
    """``ParallelGaussianSumUnscentedKalmanFilter`` (filter/gs_ukf.py:223-449) over all ranks of ``group``: the Gaussian
    components shard like particles do (BASELINE.json north_star); a resampled component's mean and covariance (20
    floats) are pulled from the owning GPU by the next predict.  ``means`` / ``covariances`` are this rank's shard."""

    LOCAL_CLS = ParallelGaussianSumUnscentedKalmanFilter

    def __init__(self, f, g, N_particles, x0, state_pdf, measurement_pdf, *, device=None, seed=0, group=None, means=None,
                 exchange="peer"):
        super().__init__(f, g, N_particles, x0, state_pdf, measurement_pdf, device=device, seed=seed, group=group,
                         exchange=exchange, means=means)

    @property
    def means(self):
        self._materialise()
        return self.local.means

    @property
    def covariances(self):
        self._materialise()
        return self.local.covariances

    def _noise_rows(self, noise):
        loc = self.local
        n = loc.N_particles
        if noise is None:
            noise = loc._host_noise(loc.state_pdf, (n, 11))
        if noise is None:
            return None, 0, None
        nz = torch.zeros((55, loc._ld), dtype=torch.float32, device=self.device)
        host = numpy.ascontiguousarray(_device.to_numpy(noise), dtype=numpy.float32).reshape(n, 55)
        nz[:, :n].copy_(torch.as_tensor(host, device=self.device).t())
        return nz.data_ptr(), loc._ld, nz

    def _predict_pending(self, u, dt, nz_ptr, ld_nz):
        loc = self.local
        sh = self._shards[self._parity]
        dst = loc._state_alt
        _lib.check(_lib.lib.gse_gsf_predict_sharded(loc._ctx.handle, ctypes.byref(sh), self._idx_global.data_ptr(),
                                                    dst.data_ptr(), dst.data_ptr() + 5 * loc._ld * 4, loc._ld,
                                                    loc.N_particles, _lib.as_double2(u), float(dt), loc._seed, loc._step,
                                                    loc._index0, nz_ptr, ld_nz, loc._stream()))

    def _launch_local_moments(self, mean_only):
        loc = self.local
        if self._pending:
            sh = self._shards[self._parity]
            _lib.check(_lib.lib.gse_gsf_moments_sharded(
                loc._ctx.handle, ctypes.byref(sh), self._idx_global.data_ptr(), loc.N_particles, loc._loglik_ptr(),
                loc._base.data_ptr() if loc._base is not None else None, loc._stats.data_ptr(), loc._mom.data_ptr(),
                loc._stream()))
        else:
            loc._launch_moments(mean_only=False)

    def covariance_matrix(self, normalised=False):
        """cov_cov + cov_mean (gs_ukf.py:442-447)."""
        cov_mean, mom, A = self._scatter(normalised)
        return A * self.local._unpack_sym(mom[26:41]) + cov_mean
