"""-m gpu tests of the sharded particle filter and GS-UKF (gpu_se_b200/sharded.py), one process per GPU.

* world size 1 (always runs): the sharded driver must reproduce the single-GPU filter bit for bit.
* both exchange modes: "peer" (kernels read the other GPUs' memory over NVLink) and "slabs" (NCCL send/recv)
* world size 2 (needs two GPUs; skipped otherwise): two shards of one population must reproduce
  the single-GPU run of the same seed bit for bit -- Philox is keyed by the global row index and the
  cumulative weights are integers, so neither the noise nor the resample depends on the split.
"""
import os
import socket
import sys

import numpy
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_cycles(pf, n_cycles, seed, set_weights=None):
    rng = numpy.random.default_rng(seed)
    out = []
    from oracle import bioreactor
    x = bioreactor.X_STEADY.copy()
    for c in range(n_cycles):
        u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
        x = x + bioreactor.increment(x, u, 0.5)
        z = bioreactor.outputs(x, round32=False) + rng.normal(size=2) * numpy.array([0.2, 0.25])
        r = float(rng.uniform())
        pf.predict(u, 0.5)
        pf.update(u, z)
        est_u = pf.point_estimate(normalised=True)
        pf.resample(r=r)
        out.append((est_u, pf.point_estimate(), pf.point_covariance()))      # moments through the pending index
    pf.update(u, z)                           # update straight after a resample: the pending gather is applied first
    out.append((pf.point_estimate(normalised=True), pf.point_estimate(), pf.point_covariance()))
    # sequences without an update in between (nothing but the resamples themselves orders the ranks):
    # resample -> resample, assigned weights -> resample twice (the reference's pf_run_seq pattern), resample -> predict
    pf.resample(r=0.11)
    pf.resample(r=0.93)
    out.append((pf.point_estimate(normalised=True), pf.point_estimate(), pf.point_covariance()))
    if set_weights is not None:
        for k in range(2):
            set_weights(pf, k)
            pf.resample(r=0.5 + 0.25 * k)
        out.append((pf.point_estimate(normalised=True), pf.point_estimate(), pf.point_covariance()))
    pf.predict(u, 0.5)
    return out


def _run_estimates(pf, n_cycles, seed):
    """The reference's filter loop: predict, update, resample, point_estimate -- from the second cycle on the estimate
    comes out of the (sharded) resample kernel, one moment block per rank merged over the mailboxes."""
    rng = numpy.random.default_rng(seed)
    out = []
    for c in range(n_cycles):
        u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
        z = numpy.array([rng.uniform(85, 95), rng.uniform(60, 75)])
        pf.predict(u, 0.5)
        pf.update(u, z)
        pf.resample(r=float(rng.uniform()))
        out.append(pf.point_estimate())
        if c == n_cycles - 2:
            out.append(pf.point_estimate())            # asked twice: the second answer must not count the blocks twice
    return out


def _global_weights(n, k):
    w = numpy.random.default_rng(1000 + k).random(n) ** 6
    return w / w.sum()


def _worker(rank, world, port, n, exchange, kind, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import gpu_se_b200 as g
        from gpu_common import make_pdfs
        from gpu_se_b200.sharded import ShardedGaussianSumUnscentedKalmanFilter, ShardedParticleFilter
        x0, state, meas = make_pdfs(g)
        f, gg = g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs
        gsf = kind == "gsf"
        sharded_cls = ShardedGaussianSumUnscentedKalmanFilter if gsf else ShardedParticleFilter
        single_cls = g.GaussianSumUnscentedKalmanFilter if gsf else g.ParticleFilter
        spf = sharded_cls(f, gg, n, x0, state, meas, device=dev, seed=77, exchange=exchange)
        if kind == "pf_fused":                 # predict + update as one kernel through the global ancestor index
            assert spf.local.can_fuse_update()
            spf.local._fuse = True
        est_only = kind == "pf_estimates"
        if est_only:
            got = _run_estimates(spf, 5, seed=3)
            assert spf.local._est_hint and not spf.local._mom_unused
        else:
            got = _run_cycles(spf, 4, seed=3, set_weights=lambda p, k: p.set_global_weights(_global_weights(n, k)))
        exchanged = spf.rows_from_peers()
        ncol = 20 if gsf else 5
        parts = [torch.empty((b - a, ncol), dtype=torch.float32, device=dev) for a, b in spf.bounds]
        for s in range(world):
            if s == rank:
                if gsf:
                    spf.means                                    # applies the pending gather
                    parts[s].copy_(spf.local._state[:, :spf.local.N_particles].t())
                else:
                    parts[s].copy_(spf.particles)
            dist.broadcast(parts[s], src=s)
        full = torch.cat(parts).cpu().numpy()
        if rank == 0:
            pf = single_cls(f, gg, n, x0, state, meas, device=dev, seed=77)

            def assign(p, k):
                p.weights = _global_weights(n, k)
            if est_only:
                ref = _run_estimates(pf, 5, seed=3)
                for i, (x, y) in enumerate(zip(got, ref)):
                    # first cycle and the repeated question (entry 4) go through the float32 group sums of k_means
                    assert numpy.allclose(x, y, rtol=1e-12 if i not in (0, 4) else 1e-6, atol=0), i
                got = ref = []
            else:
                ref = _run_cycles(pf, 4, seed=3, set_weights=assign)
            if gsf:
                pf.means
                want = pf._state[:, :n].t().cpu().numpy()
            else:
                want = pf.particles.get()
            assert numpy.array_equal(full, want), "sharded state differs from the single-GPU run"
            # identical rows and weights; only the float64 summation order of the moments differs
            for (a0, a1, a2), (b0, b1, b2) in zip(got, ref):
                assert numpy.allclose(a0, b0, rtol=1e-10, atol=1e-10)
                assert numpy.allclose(a1, b1, rtol=1e-10, atol=1e-10)
                assert a2 == pytest.approx(b2, rel=1e-8)
        spf.close()
        q.put((rank, "ok", exchanged))
        dist.destroy_process_group()
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc() + repr(e)))


def _launch(world, n, exchange, kind="pf"):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rank, world, port, n, exchange, kind, q)) for rank in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", "rank %d: %s" % (rank, info)
    return results


@pytest.mark.parametrize("exchange", ["peer", "slabs"])
@pytest.mark.parametrize("n", [4096, 100003])
def test_world1_equals_single_gpu(n, exchange):
    _launch(1, n, exchange)


@pytest.mark.parametrize("n", [4096, 300007])
def test_world1_fused_predict_update(n):
    _launch(1, n, "peer", kind="pf_fused")


@pytest.mark.parametrize("world,n", [(2, 1000003), (4, 1000003)])
def test_worldN_fused_predict_update(world, n):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _launch(world, n, "peer", kind="pf_fused")


@pytest.mark.parametrize("n", [4096, 300007])
def test_world1_estimates_out_of_the_resample_kernel(n):
    _launch(1, n, "peer", kind="pf_estimates")


@pytest.mark.parametrize("world,n", [(2, 8192), (2, 1000003), (4, 1000003), (8, 2000003)])
def test_worldN_estimates_out_of_the_resample_kernel(world, n):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _launch(world, n, "peer", kind="pf_estimates")


@pytest.mark.parametrize("exchange,n", [("peer", 5000), ("slabs", 5000), ("peer", 70001)])
def test_world1_gsukf_equals_single_gpu(n, exchange):
    _launch(1, n, exchange, kind="gsf")


@pytest.mark.parametrize("world,exchange,n", [(2, "peer", 8192), (2, "peer", 1000003), (2, "slabs", 8192),
                                              (2, "slabs", 1000003), (4, "peer", 1000003), (4, "slabs", 100003),
                                              (8, "peer", 2000003)])
def test_worldN_equals_single_gpu(world, n, exchange):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    results = _launch(world, n, exchange)
    assert any(info > 0 for _, _, info in results), "informative measurement: shards must exchange rows"


@pytest.mark.parametrize("world,exchange,n", [(2, "peer", 8192), (2, "peer", 200003), (2, "slabs", 8192), (4, "peer", 200003),
                                              (8, "peer", 400003)])
def test_worldN_gsukf_equals_single_gpu(world, n, exchange):
    """BASELINE.json north_star: "particles and Gaussian components shard naturally across the 8 B200s"."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _launch(world, n, exchange, kind="gsf")
