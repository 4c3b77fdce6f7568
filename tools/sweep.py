#!/usr/bin/env python
"""Run-sequence sweep (BASELINE.json configs[1] and configs[3]): per-stage device times of the particle
filter for N = 2^10 ... 2^24 and of the GS-UKF for N = 2^8 ... 2^20 on one B200, the way
results/pf_openloop/pf_run_seq.py:328-351 and results/gsf_openloop/gsf_run_seq.py:474-497 sweep N
(median of the run sequence, 10 % / 90 % quantiles), but timed with CUDA events on the launching
stream instead of unsynchronised time.time().

    python tools/sweep.py [--runs 100] [--out gpurun_out/sweep.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402  (trajectory, peak)


PF_STEP_BYTES = bench.STAGE_BYTES["predict"] + bench.STAGE_BYTES["update"] + bench.STAGE_BYTES["resample"]     # 64 B
GSF_STEP_BYTES = sum(bench.GSF_STAGE_BYTES.values())                                                            # 336 B


def build(kind, n, dev):
    import gpu_se_b200 as g
    from gpu_se_b200.model.BioreactorModel import X_STEADY
    means = numpy.zeros((2, 5))
    covs = numpy.array([numpy.diag([1e-4, 1e-7, 1e-3, 1e-3, 1e-7]), numpy.diag([1e-3, 1e-6, 1e-2, 1e-2, 1e-6])])
    state = g.MultivariateGaussianSum(means, covs, [0.75, 0.25])
    meas = g.MultivariateGaussianSum([[1e-1, 0], [0, -1e-1]], [[[6e-2, 0], [0, 8e-2]], [[500, 100], [100, 700]]],
                                     [0.85, 0.15])
    x0 = g.MultivariateGaussianSum(means + numpy.array(X_STEADY)[None, :], covs, [0.75, 0.25])
    cls = g.ParticleFilter if kind == "pf" else g.GaussianSumUnscentedKalmanFilter
    return cls(g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs, n, x0, state, meas, device=dev, seed=11)


def time_filter(kind, n, runs, dev, dt, graphs=False):
    f = build(kind, n, dev)
    if graphs:
        f.enable_graphs()
    us, zs = bench.trajectory(runs + 10, seed=3)
    rs = numpy.random.default_rng(1).random(runs + 10)
    stream = torch.cuda.current_stream(dev)
    rec = {"predict": [], "update": [], "resample": []}
    for k in range(runs + 10):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        f.predict(us[k], dt)
        if not graphs:
            ev[1].record(stream)
        f.update(us[k], zs[k])
        if not graphs:
            ev[2].record(stream)
        f.resample(r=float(rs[k]))
        ev[3].record(stream)
        if graphs:                     # the three calls are one graph launch: only the whole step can be timed
            ev[1] = ev[2] = ev[0]
        if k >= 10:
            rec["predict"].append((ev[0], ev[1]))
            rec["update"].append((ev[1], ev[2]))
            rec["resample"].append((ev[2], ev[3]))
    torch.cuda.synchronize(dev)
    out = {"N": n}
    total = numpy.zeros(runs)
    for stage, pairs in rec.items():
        t = numpy.array([a.elapsed_time(b) for a, b in pairs])
        total += t
        out[stage + "_ms"] = {"median": float(numpy.median(t)), "q10": float(numpy.quantile(t, 0.1)),
                              "q90": float(numpy.quantile(t, 0.9))}
    out["step_ms_median"] = float(numpy.median(total))
    out["units_per_s"] = n / (float(numpy.median(total)) * 1e-3)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--pf-max", type=int, default=24)
    ap.add_argument("--gsf-max", type=int, default=20)
    ap.add_argument("--graphs", action="store_true", help="CUDA-graph replay of the cycle (stage times collapse into 'resample')")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak, src = bench.measured_peak_gbs()
    res = {"hbm_peak_gbs": peak, "peak_source": src, "runs": a.runs, "dt": 1.0, "pf": [], "gsf": [], "cuda_graphs": a.graphs,
           "pf_bytes_per_particle_step": PF_STEP_BYTES, "gsf_bytes_per_comp_step": GSF_STEP_BYTES}
    for p in range(10, a.pf_max + 1, 2):
        r = time_filter("pf", 1 << p, a.runs, dev, 1.0, a.graphs)
        r["hbm_frac"] = PF_STEP_BYTES * r["N"] / (r["step_ms_median"] * 1e-3) / 1e9 / peak
        res["pf"].append(r)
        print(json.dumps(r), flush=True)
    for p in list(range(8, min(a.gsf_max, 16) + 1, 2)) + ([18, 20] if a.gsf_max >= 20 else []):
        r = time_filter("gsf", 1 << p, a.runs, dev, 1.0, a.graphs)
        r["hbm_frac"] = res["gsf_bytes_per_comp_step"] * r["N"] / (r["step_ms_median"] * 1e-3) / 1e9 / peak
        res["gsf"].append(r)
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
