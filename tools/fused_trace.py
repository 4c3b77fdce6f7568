#!/usr/bin/env python
"""Per-CTA phase timing of the fused resample kernel (GSE_FUSED_TRACE=1): where the grid-wide waits go.

    GSE_FUSED_TRACE=1 python tools/fused_trace.py [log2n] [--sharded]     (--sharded: the one-rank sharded driver)
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GSE_FUSED_TRACE", "1")

import numpy  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import gpu_se_b200 as g  # noqa: E402
from gpu_se_b200 import _lib  # noqa: E402
from gpu_se_b200.model.BioreactorModel import X_STEADY  # noqa: E402


def main():
    sharded = "--sharded" in sys.argv
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    log2n = int(args[0]) if args else 24
    n = 1 << log2n
    sm, sc = numpy.zeros((2, 5)), numpy.array([numpy.diag([1e-4, 1e-7, 1e-3, 1e-3, 1e-7]),
                                               numpy.diag([1e-3, 1e-6, 1e-2, 1e-2, 1e-6])])
    state = g.MultivariateGaussianSum(sm, sc, [0.75, 0.25])
    meas = g.MultivariateGaussianSum([[1e-1, 0], [0, -1e-1]], [[[6e-2, 0], [0, 8e-2]], [[500, 100], [100, 700]]], [0.85, 0.15])
    x0 = g.MultivariateGaussianSum(sm + numpy.array(X_STEADY)[None, :], sc, [0.75, 0.25])
    if sharded:
        import torch.distributed as dist
        from gpu_se_b200.sharded import ShardedParticleFilter
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        torch.cuda.set_device(0)
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        pf = ShardedParticleFilter(g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs, n, x0, state, meas,
                                   device=torch.device("cuda", 0), seed=1234)
    else:
        pf = g.ParticleFilter(g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs, n, x0, state, meas, seed=1234)
    us, zs = bench.trajectory(12, seed=7)
    buf = numpy.zeros(8 * 4096, dtype=numpy.uint64)
    for k in range(12):
        pf.predict(us[k], 1.0)
        pf.update(us[k], zs[k])
        w = pf.weights.as_subclass(torch.Tensor)
        ess = float(w.sum() ** 2 / (w * w).sum())
        idx = pf.resample(r=0.37, return_index=True)
        distinct = int(torch.unique_consecutive(idx).numel())
        print("step %d: ESS %.0f of %d (%.3f%%), distinct ancestors %d (%.2f%%)" % (k, ess, n, 100 * ess / n, distinct, 100.0 * distinct / n))
        torch.cuda.synchronize()
        got = _lib.lib.gse_ctx_read_trace(pf._ctx.handle, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
        t = buf[:got].reshape(-1, 8).astype(numpy.int64)
        t = t[t[:, 0] > 0]
        if k < 8:
            continue
        t0 = t[:, 0].min()
        rel = (t - t0) / 1e3
        d1, d2, d3 = rel[:, 1] - rel[:, 0], rel[:, 2] - rel[:, 1], rel[:, 3] - rel[:, 2]
        smid = t[:, 4]
        per_sm = numpy.bincount(smid)
        crowd = per_sm[smid]
        for c in sorted(set(crowd)):
            print("   CTAs on SMs hosting %d CTAs: n=%d phase1 %.1f phase3 mean %.1f max %.1f" % (c, (crowd == c).sum(), d1[crowd == c].mean() if (crowd == c).any() else 0, d3[crowd == c].mean(), d3[crowd == c].max()))
        order = numpy.argsort(d3)
        print("   slowest phase-3 CTAs (vb, smid, us):", [(int(i), int(smid[i]), round(float(d3[i]), 1)) for i in order[-6:]])
        print("step %d: %d CTAs; start spread %.1f us | phase1 mean %.1f max %.1f | wait+collect mean %.1f max %.1f | "
              "phase3 mean %.1f min %.1f max %.1f | phase1 done at mean %.1f max %.1f | phase3 done at mean %.1f max %.1f us"
              % (k, len(t), rel[:, 0].max(), d1.mean(), d1.max(), d2.mean(), d2.max(), d3.mean(), d3.min(), d3.max(),
                 rel[:, 1].mean(), rel[:, 1].max(), rel[:, 3].mean(), rel[:, 3].max()))
    if sharded:
        pf.close()


if __name__ == "__main__":
    main()
