#!/usr/bin/env python
"""Energy per filter step: the reference's PowerMeasurement harness
(decorators.py:94-206, results/pf_openloop/pf_power.py) on NVML.

The reference spawns a side process that polls ``nvidia-smi power.draw`` and ``psutil.cpu_percent`` every 0.2 s and
integrates with the trapezoid rule (decorators.py:135-137, 190-206).  Here the GPU energy comes from NVML's own
energy counter (``nvmlDeviceGetTotalEnergyConsumption``, millijoules since driver load: no sampling error), with a
50 ms power-draw sampler beside it as the cross-check and the fallback; the CPU share is estimated as the reference
does, utilisation x an assumed full-load power (``--cpu-max-power``, decorators.py:107).

    python tools/power.py [--t-run 2.0] [--pf-max 24] [--gsf-max 20] [--out gpurun_out/power.json]

For every N the filter runs predict -> update -> resample cycles for ``t_run`` seconds (pf_power.py:17-50); reported:
steps run, joules per step (GPU, CPU estimate), average watts, particle-steps per joule.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from tools import sweep  # noqa: E402


class EnergyMeter:
    """GPU joules between start() and stop(): NVML energy counter, plus a sampled-power integral."""

    def __init__(self, index=0, period=0.05):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.period = period
        try:
            import psutil
            self.psutil = psutil
        except ImportError:
            self.psutil = None

    def _sample(self):
        while not self._stop.is_set():
            self._t.append(time.perf_counter())
            self._w.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
            if self.psutil is not None:
                self._cpu.append(self.psutil.cpu_percent() / 100.0)
            self._stop.wait(self.period)

    def start(self):
        self._t, self._w, self._cpu = [], [], []
        self._stop = threading.Event()
        try:
            self._e0 = self.nv.nvmlDeviceGetTotalEnergyConsumption(self.h)
        except Exception:
            self._e0 = None
        if self.psutil is not None:
            self.psutil.cpu_percent()
        self._t0 = time.perf_counter()
        self._thread = threading.Thread(target=self._sample, daemon=True)
        self._thread.start()

    def stop(self):
        t1 = time.perf_counter()
        e1 = self.nv.nvmlDeviceGetTotalEnergyConsumption(self.h) if self._e0 is not None else None
        self._stop.set()
        self._thread.join()
        t, w = numpy.asarray(self._t), numpy.asarray(self._w)
        sampled = float(numpy.sum(0.5 * (w[1:] + w[:-1]) * numpy.diff(t))) if len(t) > 1 else float("nan")   # trapezoid, as :135
        out = {"seconds": t1 - self._t0, "gpu_joules_sampled": sampled, "gpu_watts_max": float(w.max()) if len(w) else None,
               "samples": int(len(t))}
        out["gpu_joules"] = (e1 - self._e0) / 1e3 if e1 is not None else sampled
        out["gpu_joules_source"] = "nvml energy counter" if e1 is not None else "sampled power, trapezoid"
        if self._cpu:
            out["cpu_fraction_mean"] = float(numpy.mean(self._cpu))
        return out


def run_for(kind, n, t_run, dev, meter, cpu_max_power):
    f = sweep.build(kind, n, dev)
    us, zs = bench.trajectory(64, seed=3)
    rs = numpy.random.default_rng(1).random(64)
    for k in range(8):                                   # warm-up
        f.predict(us[k], 1.0); f.update(us[k], zs[k]); f.resample(r=float(rs[k]))
    torch.cuda.synchronize(dev)
    meter.start()
    steps, t0, stopped = 0, time.perf_counter(), None
    while time.perf_counter() - t0 < t_run:
        for _ in range(16):                              # keep the queue fed; the estimate read-back bounds the run-ahead
            k = steps % 64
            f.predict(us[k], 1.0); f.update(us[k], zs[k]); f.resample(r=float(rs[k]))
            steps += 1
        try:
            f.point_estimate()
        except numpy.linalg.LinAlgError as e:
            # tens of thousands of GS-UKF cycles on a looped trajectory: some component's float32 covariance stops
            # being positive definite even after the jitter retry -- where the reference raises too (gs_ukf.py:72-75).
            # The energy of the steps run so far stands.
            stopped = "LinAlgError after %d steps: %s" % (steps, e)
            break
    torch.cuda.synchronize(dev)
    m = meter.stop()
    cpu_j = m.get("cpu_fraction_mean", float("nan")) * cpu_max_power * m["seconds"]
    return {"N": n, "steps": steps, "seconds": m["seconds"], "gpu_joules_per_step": m["gpu_joules"] / steps,
            "gpu_watts_mean": m["gpu_joules"] / m["seconds"], "gpu_watts_max": m["gpu_watts_max"],
            "gpu_joules_sampled_per_step": m["gpu_joules_sampled"] / steps, "cpu_joules_per_step_estimate": cpu_j / steps,
            "units_per_gpu_joule": n * steps / m["gpu_joules"], "units_per_s": n * steps / m["seconds"],
            "gpu_joules_source": m["gpu_joules_source"], "power_samples": m["samples"], "stopped_early": stopped}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--t-run", type=float, default=2.0)
    ap.add_argument("--pf-max", type=int, default=24)
    ap.add_argument("--gsf-max", type=int, default=20)
    ap.add_argument("--cpu-max-power", type=float, default=30.0, help="assumed CPU power at 100 %% (decorators.py:107)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "power.json"))
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    meter = EnergyMeter(0)
    res = {"t_run_s": a.t_run, "cpu_max_power_w": a.cpu_max_power, "pf": [], "gsf": [],
           "note": "energy of predict+update+resample cycles, estimate read back every 16 steps; idle power is included"}
    for p in range(10, a.pf_max + 1, 2):
        r = run_for("pf", 1 << p, a.t_run, dev, meter, a.cpu_max_power)
        res["pf"].append(r)
        print(json.dumps(r), flush=True)
    for p in list(range(8, min(a.gsf_max, 16) + 1, 4)) + ([20] if a.gsf_max >= 20 else []):
        r = run_for("gsf", 1 << p, a.t_run, dev, meter, a.cpu_max_power)
        res["gsf"].append(r)
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
