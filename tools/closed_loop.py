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
    ap.add_argument("--dt-control", type=float, default=0.1, help="control period in minutes (thesis: 0.1)")
    ap.add_argument("--controller", default="mpc", choices=["mpc", "pi"],
                    help="mpc: the reference's linear MPC (gpu_se_b200/controller.py, host ADMM); pi: cheap stand-in")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "closed_loop.json"))
    a = ap.parse_args()
    from gpu_se_b200.sim_base import Simulation
    n = 1 << a.log2n
    import time
    t0 = time.perf_counter()
    sim = Simulation(n, dt_control=a.dt_control, dt_predict=0.1, end_time=a.end_time, pf=not a.gsf, seed=1,
                     controller=a.controller)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    sim.simulate()
    t_sim = time.perf_counter() - t0
    fs = numpy.asarray(sim.filter_seconds)[5:]
    err = numpy.abs(sim.ys_f - sim.ys[:, list(sim.OUTPUTS)])[20:]
    res = {"filter": "gsf" if a.gsf else "pf", "N": n, "steps": int(len(sim.ts) - 1), "dt_control_min": a.dt_control,
           "control_period_s": 60.0 * a.dt_control, "filter_ms_per_step": {"median": float(numpy.median(fs) * 1e3),
                                                           "q10": float(numpy.quantile(fs, 0.1) * 1e3),
                                                           "q90": float(numpy.quantile(fs, 0.9) * 1e3)},
           "utilisation": sim.utilisation(), "ise": sim.performance,
           "median_abs_output_error_mg_per_L": [float(numpy.median(err[:, 0])), float(numpy.median(err[:, 1]))],
           "calls_per_step": "predict, update, resample, point_estimate (x2), point_covariance",
           "controller": type(sim.K).__name__, "build_s": t_build, "simulate_s": t_sim,
           "final_outputs_mg_per_L": [float(v) for v in sim.ys[-1][list(sim.OUTPUTS)]]}
    if hasattr(sim.K, "mpc_frac"):
        its = numpy.asarray(sim.K.iterations) if sim.K.iterations else numpy.zeros(1)
        res["mpc"] = {"mpc_frac": sim.K.mpc_frac, "admm_iterations_median": float(numpy.median(its)),
                      "admm_iterations_max": int(its.max()), "variables": int(sim.K.K.H.shape[0]),
                      "constraints": int(sim.K.K.A_matrix.shape[0]), "P": sim.K.K.P, "M": sim.K.K.M}
    print(json.dumps(res))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
