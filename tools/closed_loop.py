#!/usr/bin/env python
"""BASELINE.json configs[4]: closed-loop bioreactor, 2^20-particle filter at the thesis control
period (dt_control = dt_predict = 0.1 min = 6 s, results/pf_closedloop/bioreactor_performance_pf.py:105).
Reports the filter's share of the control period (utilisation, :157), the ISE and the per-step
filter time.

    python tools/closed_loop.py [--log2n 20] [--end-time 50] [--gsf] [--out gpurun_out/closed_loop.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--end-time", type=int, default=50)
    ap.add_argument("--gsf", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "closed_loop.json"))
    a = ap.parse_args()
    from gpu_se_b200.sim_base import Simulation
    n = 1 << a.log2n
    sim = Simulation(n, dt_control=0.1, dt_predict=0.1, end_time=a.end_time, pf=not a.gsf, seed=1)
    sim.simulate()
    fs = numpy.asarray(sim.filter_seconds)[5:]
    err = numpy.abs(sim.ys_f - sim.ys[:, list(sim.OUTPUTS)])[20:]
    res = {"filter": "gsf" if a.gsf else "pf", "N": n, "steps": int(len(sim.ts) - 1), "dt_control_min": 0.1,
           "control_period_s": 6.0, "filter_ms_per_step": {"median": float(numpy.median(fs) * 1e3),
                                                           "q10": float(numpy.quantile(fs, 0.1) * 1e3),
                                                           "q90": float(numpy.quantile(fs, 0.9) * 1e3)},
           "utilisation": sim.utilisation(), "ise": sim.performance,
           "median_abs_output_error_mg_per_L": [float(numpy.median(err[:, 0])), float(numpy.median(err[:, 1]))],
           "calls_per_step": "predict, update, resample, point_estimate (x2), point_covariance",
           "controller": type(sim.K).__name__}
    print(json.dumps(res))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
