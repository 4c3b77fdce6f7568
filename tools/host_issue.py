"""Where does the host spend its time while it issues a filter step?  perf_counter after every public call, no
synchronisation in between: a call that takes as long as a kernel is a call that waits for the device.

    python tools/host_issue.py [--log2n 24] [--sharded]
"""
import argparse
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=float, default=24)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--sharded", action="store_true")
    ap.add_argument("--estimate", action="store_true", help="read point_estimate() every step (the end-to-end loop)")
    args = ap.parse_args()
    import torch
    import gpu_se_b200 as g
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from gpu_common import make_pdfs
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    x0, state, meas = make_pdfs(g)
    f, gg = g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs
    n = int(round(2 ** args.log2n))
    if args.sharded:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29544")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        from gpu_se_b200.sharded import ShardedParticleFilter
        pf = ShardedParticleFilter(f, gg, n, x0, state, meas, device=dev, seed=1)
    else:
        pf = g.ParticleFilter(f, gg, n, x0, state, meas, device=dev, seed=1)
    u, z = numpy.array([0.06, 0.2]), numpy.array([90.0, 70.0])
    rs = numpy.random.default_rng(0).random(args.steps + 5)
    names = ["predict", "update", "resample"] + (["point_estimate"] if args.estimate else [])
    acc = {k: [] for k in names}
    for k in range(args.steps + 5):
        if k == 5:
            torch.cuda.synchronize(dev)
            t_begin = time.perf_counter()
        t = [time.perf_counter()]
        pf.predict(u, 1.0)
        t.append(time.perf_counter())
        pf.update(u, z)
        t.append(time.perf_counter())
        pf.resample(r=float(rs[k]))
        t.append(time.perf_counter())
        if args.estimate:
            pf.point_estimate()
            t.append(time.perf_counter())
        if k >= 5:
            for i, name in enumerate(names):
                acc[name].append(t[i + 1] - t[i])
    t_issued = time.perf_counter()
    torch.cuda.synchronize(dev)
    t_done = time.perf_counter()
    print("steps %d  issue %.1f us/step  until the device is done %.1f us/step" % (
        args.steps, (t_issued - t_begin) * 1e6 / args.steps, (t_done - t_begin) * 1e6 / args.steps))
    for name in names:
        v = numpy.array(acc[name]) * 1e6
        print("  %-15s median %7.1f us   min %7.1f   max %7.1f" % (name, numpy.median(v), v.min(), v.max()))
    if hasattr(pf, "close"):
        pf.close()


if __name__ == "__main__":
    main()
