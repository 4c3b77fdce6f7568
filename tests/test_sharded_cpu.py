"""World-size-2 and -3 ``gloo`` tests (CPU) of the sharded resample's host side: the output-range
plan (gpu_se_b200/sharded.py: plan_resample) and the slab exchange (exchange_columns), with the
oracle standing in for the device search + gather.  The concatenated shards must equal the
single-process reference resample (filter/particle.py:85-103 via oracle.particle) bit for bit."""
import os
import socket
import sys

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _weights(case, n, rng):
    if case == "uniform":
        return numpy.full(n, 4096, dtype=numpy.uint64)
    if case == "random":
        return rng.integers(0, 1 << 20, size=n).astype(numpy.uint64)
    if case == "skewed":          # nearly all mass on a few rows of the last shard
        w = rng.integers(0, 4, size=n).astype(numpy.uint64)
        w[-7] = 1 << 30
        w[-3] = 1 << 29
        return w
    if case == "front":           # all mass on shard 0, zeros elsewhere
        w = numpy.zeros(n, dtype=numpy.uint64)
        w[: n // 5] = rng.integers(1, 1 << 12, size=n // 5).astype(numpy.uint64)
        return w
    raise ValueError(case)


def _worker(rank, world, port, n, case, r, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from gpu_se_b200 import sharded
        from oracle import particle

        rng = numpy.random.default_rng(5)
        w = _weights(case, n, rng)                                  # integer weights: every sum is exact
        x = rng.normal(size=(n, 5)).astype(numpy.float32)
        bounds = sharded.shard_bounds(n, world)
        lo, hi = bounds[rank]
        c_loc = numpy.cumsum(w[lo:hi], dtype=numpy.uint64)
        mine = torch.tensor([int(c_loc[-1]) if hi > lo else 0], dtype=torch.int64)
        allt = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine)
        totals = [int(t) for t in torch.cat(allt).tolist()]
        plan = sharded.plan_resample(totals, r, n, bounds)

        # the ranges tile [0, n) in shard order
        assert plan.src_ranges[0][0] == 0 and plan.src_ranges[-1][1] == n
        for s in range(1, world):
            assert plan.src_ranges[s][0] == plan.src_ranges[s - 1][1]

        offs, width = sharded.staging_layout(plan, rank)
        staging = torch.zeros((5, max(width, 1)), dtype=torch.float32)
        dst = torch.full((5, hi - lo), float("nan"), dtype=torch.float32)
        O, T = plan.offsets[rank], plan.total
        cn = (c_loc.astype(numpy.float64) + numpy.float64(O)) / numpy.float64(T)     # (O_s + c_k) / T, exact ints
        for s, t, start, stop in plan.transfers:
            if s != rank:
                continue
            u = (numpy.arange(start, stop, dtype=numpy.float64) + r) / n           # particle.py:97
            k = numpy.searchsorted(cn, u, side="left")
            assert k.min() >= 0 and k.max() < hi - lo, "plan sent an output to a shard that does not source it"
            rows = torch.from_numpy(x[lo:hi][k].T.copy())
            if t == rank:
                dst[:, start - lo:stop - lo] = rows
            else:
                staging[:, offs[t]:offs[t] + stop - start] = rows
        sharded.exchange_columns(plan, rank, bounds, staging, offs, dst, 5)

        parts = [torch.empty((5, b - a), dtype=torch.float32) for a, b in bounds]
        dist.all_gather(parts, dst) if len({b - a for a, b in bounds}) == 1 else _gather_uneven(parts, dst, rank, world)
        got = torch.cat(parts, dim=1).numpy().T
        o = particle.ParticleFilterOracle(n, None, None, None, particles=x.copy())
        o.weights = w.astype(numpy.float64)
        o.resample(r=r)
        assert numpy.array_equal(got, o.particles), "sharded resample differs from the single-process reference"
        q.put((rank, "ok", plan.exchanged_rows()))
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc() + repr(e)))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def _gather_uneven(parts, dst, rank, world):
    for s in range(world):
        if s == rank:
            parts[s].copy_(dst)
        dist.broadcast(parts[s], src=s)


@pytest.mark.parametrize("world,n,case,r", [
    (2, 1024, "uniform", 0.5),
    (2, 1000, "random", 0.123456789),
    (2, 1024, "skewed", 0.75),
    (2, 1000, "front", 0.0),
    (3, 1001, "random", 0.999),
    (3, 999, "skewed", 0.3),
])
def test_sharded_resample_gloo(world, n, case, r):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rank, world, port, n, case, r, q)) for rank in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", "rank %d: %s" % (rank, info)
    if case == "uniform":
        assert all(info == 0 for _, _, info in results)       # equal weights: nothing crosses shards
    if case in ("skewed", "front"):
        assert any(info > 0 for _, _, info in results)


def test_shard_bounds():
    from gpu_se_b200 import sharded
    assert sharded.shard_bounds(10, 2) == [(0, 4), (4, 10)]
    assert sharded.shard_bounds(34, 3) == [(0, 12), (12, 24), (24, 34)]
    assert sharded.shard_bounds(32, 8) == [(4 * i, 4 * i + 4) for i in range(8)]
    for n, g in ((1000003, 2), (2 ** 24, 8), (999, 3)):
        b = sharded.shard_bounds(n, g)
        assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert all(lo % 4 == 0 for lo, _ in b)
