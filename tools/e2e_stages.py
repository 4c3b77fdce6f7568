"""Per-stage device times of the END-TO-END loop (host u / z in, point_estimate() read back every step): CUDA events
around every public call plus the host-side wall clock of the whole loop.

    python tools/e2e_stages.py [--log2n 24] [--steps 100]
"""
import argparse
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=float, default=24)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--no-events", action="store_true")
    args = ap.parse_args()
    import torch
    import bench
    import gpu_se_b200 as g
    from gpu_se_b200.model.BioreactorModel import X_STEADY
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n = int(round(2 ** args.log2n))
    state_means, state_covs = numpy.zeros((2, 5)), numpy.array([numpy.diag([1e-4, 1e-7, 1e-3, 1e-3, 1e-7]),
                                                                 numpy.diag([1e-3, 1e-6, 1e-2, 1e-2, 1e-6])])
    state = g.MultivariateGaussianSum(state_means, state_covs, [0.75, 0.25])
    meas = g.MultivariateGaussianSum([[1e-1, 0], [0, -1e-1]], [[[6e-2, 0], [0, 8e-2]], [[500, 100], [100, 700]]],
                                     [0.85, 0.15])
    x0 = g.MultivariateGaussianSum(state_means + numpy.array(X_STEADY)[None, :], state_covs, [0.75, 0.25])
    pf = g.ParticleFilter(g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs, n, x0, state, meas, device=dev,
                          seed=1234)
    K, W = args.steps, 10
    us, zs = bench.trajectory(K + W, seed=7)
    rs = numpy.random.default_rng(99).random(K + W)
    stream = torch.cuda.current_stream(dev)
    evs = []

    def mark():
        if args.no_events:
            return
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        evs.append(e)

    host = {"predict": 0.0, "update": 0.0, "resample": 0.0, "estimate": 0.0}
    for k in range(K + W):
        if k == W:
            torch.cuda.synchronize(dev)
            evs.clear()
            for key in host:
                host[key] = 0.0
            t0 = time.perf_counter()
        ta = time.perf_counter()
        mark()
        pf.predict(us[k], 1.0)
        tb = time.perf_counter()
        mark()
        pf.update(us[k], zs[k])
        tc = time.perf_counter()
        mark()
        pf.resample(r=float(rs[k]))
        td = time.perf_counter()
        mark()
        pf.point_estimate()
        te = time.perf_counter()
        mark()
        host["predict"] += tb - ta
        host["update"] += tc - tb
        host["resample"] += td - tc
        host["estimate"] += te - td
    torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - t0) / K * 1e6
    print("e2e %.1f us/step (wall)" % wall)
    print("host  : " + "  ".join("%s %.1f" % (k, v / K * 1e6) for k, v in host.items()))
    if not args.no_events:
        names = ["predict", "update", "resample", "estimate", "gap"]
        acc = {k: 0.0 for k in names}
        for s in range(K):
            e = evs[5 * s:5 * s + 5]
            for i in range(4):
                acc[names[i]] += e[i].elapsed_time(e[i + 1])
            if s + 1 < K:
                acc["gap"] += e[4].elapsed_time(evs[5 * s + 5])
        print("device: " + "  ".join("%s %.1f" % (k, v / K * 1e3) for k, v in acc.items()))


if __name__ == "__main__":
    main()
