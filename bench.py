#!/usr/bin/env python
"""Headline benchmark: open-loop bioreactor particle filter, particle-steps/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n 24]

One "step" = predict(u, dt=1.0) -> update(u, z) -> resample() on the whole particle population
(the three calls the reference times, results/pf_openloop/pf_run_seq.py:45-49,84-88,123-128).
Workload at N = 1: BASELINE.json configs[1] at its largest size, 2^24 particles on one B200
(north_star target size).  N > 1 (torchrun, one rank per GPU): 2^24 particles per GPU, sharded
with the global systematic resample of gpu_se_b200/sharded.py (weak scaling).

Prints ONE JSON line (rank 0).  `value` is device-timed with the particles resident in HBM; `e2e`
goes through the public predict/update/resample/point_estimate API with host u, z and the
estimate read back every step; `roofline` is the dominant kernel against the measured HBM peak;
`cpu_baseline` is the reference's own `filter.ParticleFilter` (the unmodified copy under oracle/_ref)
timed on one host core; `--impl reference` times it alone (the oracle port only if the copy is absent).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy  # noqa: E402

METRIC = "particle-steps/sec (predict+update+resample)"
UNIT = "particle-steps/s"
DT = 1.0                      # pf_run_seq.py:48  p.predict(u, 1.)
U_NOMINAL = numpy.array([0.06, 0.2])
# algorithmic bytes per particle and stage (SURVEY.md §8(d) re-cut for the lazy resample --
# predict reads its rows through the int32 ancestor index, so K5's gather rides inside K1):
#   predict 4 R idx + 20 R + 20 W, update 8 R + 4 W, scan 4 R + 8 W, search 8 R + 4 W
#   fused resample: 4 R loglik + 4 W ancestor index (the cumulative weights stay on chip); two-stage (GSE_RESAMPLE=unfused and
#   the sharded path): scan 4 R + 8 W, search 8 R + 4 W
#   predict + update as one kernel (the default when update() directly follows predict()): predict's 44 + 4 W loglik -- the
#   two measured columns are not read back
STAGE_BYTES = {"predict": 44, "update": 12, "predict+update": 48, "resample": 8, "scan": 12, "search": 12}
# sharded run: the same, plus the all-gather of the shard totals ("offsets", no HBM traffic to speak of)
STAGE_BYTES_SHARDED = {"predict": 44, "update": 12, "predict+update": 48, "resample": 8}


NCU_CAPTURE = os.path.join(ROOT, "profiles", "r2_pf_step_2p24_ncu_full.csv")
NCU_KERNEL_OF_STAGE = {"predict": "k_pf_predict", "update": "k_pf_update", "resample": "k_resample_fused"}


def ncu_traffic_bytes_per_row():
    """(bytes per row by stage, source) from the committed ncu capture of this command at 2^24 rows: the mean of
    dram__bytes_read.sum + dram__bytes_write.sum over the captured launches of the stage's kernel.  Falls back to the
    constants above (copied from the same capture) for stages the file does not hold."""
    import csv
    out, src = dict(NCU_TRAFFIC_BYTES_PER_ROW), "constants copied from profiles/r2_pf_step_2p24_ncu_full.csv"
    try:
        rows = list(csv.reader(open(NCU_CAPTURE)))
        hdr = rows[0]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        ur, uw = scale[rows[1][ir]], scale[rows[1][iw]]
        for stage, kern in NCU_KERNEL_OF_STAGE.items():
            v = [float(r[ir]) * ur + float(r[iw]) * uw for r in rows[2:] if len(r) == len(hdr) and kern in r[0]]
            if v:
                out[stage] = sum(v) / len(v) / 2 ** 24
        src = "read from profiles/r2_pf_step_2p24_ncu_full.csv"
    except (OSError, ValueError, KeyError):
        pass
    return out, src


def workload_name(log2n):
    return ("pf_openloop predict+update+resample, BioreactorModel, 2^%g particles per GPU, dt=1.0, "
            "in-kernel Philox noise" % log2n)


# DRAM traffic per launch of each kernel at 2^24 particles (dram__bytes_read.sum + dram__bytes_write.sum of
# the committed `ncu --set full` capture, profiles/r2_pf_step_2p24_ncu_full.csv; scan / search: the two-stage path,
# profiles/r1_final_pf_step_2p24_ncu_full.csv), in bytes per row
NCU_TRAFFIC_BYTES_PER_ROW = {"predict": (335.63e6 + 285.28e6) / 2 ** 24, "update": (134.24e6 + 31.28e6) / 2 ** 24,
                             "resample": (91.94e6 + 21.66e6) / 2 ** 24,
                             "scan": (110.59e6 + 84.81e6) / 2 ** 24,
                             "search": (72.14e6 + 1.71e6 + 134.33e6 + 46.80e6) / 2 ** 24}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=float, default=24.0, help="log2 of particles per GPU")
    ap.add_argument("--cpu-log2n", type=int, default=13, help="log2 of the CPU-baseline sample size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="pf", choices=["pf", "gsf"],
                    help="gsf: GS-UKF components/s at 2^--log2n components (BASELINE configs[3]; give --log2n 16)")
    ap.add_argument("--no-gsf", action="store_true", help="skip the GS-UKF sub-object of the default line")
    ap.add_argument("--graphs", action="store_true", help="CUDA-graph replay of the cycle (launch-bound sizes)")
    ap.add_argument("--peer-buffers", action="store_true",
                    help="single-GPU filter with its state in IPC-exportable memory (diagnostic for the sharded path)")
    ap.add_argument("--sharded", action="store_true",
                    help="use the sharded driver even on one GPU (profiling the peer-memory kernels under ncu)")
    return ap.parse_args()


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------
# synthetic open-loop trajectory: inputs as sim_base.get_random_io draws them (sim_base.py:196-199),
# measurements physically consistent with a host plant following the same model
# ----------------------------------------------------------------------------------------------
def trajectory(n_steps, seed=0, model=None):
    """`model` = (f, g, x_steady): host evaluations of the plant.  Our arm uses gpu_se_b200.model.Bioreactor's plain-Python
    static methods; the reference arm uses the reference's own (or the oracle's) so that it never loads libgse_b200.so."""
    if model is None:
        from gpu_se_b200.model.BioreactorModel import Bioreactor, X_STEADY
        model = (Bioreactor.homeostatic_DEs, Bioreactor.static_outputs, X_STEADY)
    f, g, x_steady = model
    rng = numpy.random.default_rng(seed)
    x = numpy.array(x_steady, dtype=numpy.float64)
    us, zs = [], []
    for _ in range(n_steps):
        u = numpy.array([rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.3)])
        x = x + numpy.array(f(x, u, DT), dtype=numpy.float64)
        z = numpy.array(g(x, u), dtype=numpy.float64) + rng.normal(size=2) * numpy.array([0.2, 0.25])
        us.append(u)
        zs.append(z)
    return us, zs


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, B200_PROFILING.md clocks line)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                clk, cmax = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            if t0 - 0.05 <= t <= t1 + 0.01:
                sm.append(clk)
                mx.append(cmax)
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:      # region shorter than the sampling period: use whatever was seen
            for t, line in self.lines:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except Exception:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm.  The reference's OWN classes (filter.ParticleFilter / GaussianSumUnscentedKalmanFilter, numpy code path,
# filter/particle.py:9-114, filter/gs_ukf.py:9-183) imported unmodified from /root/reference or from the git-ignored copy
# oracle/_ref that oracle/make_ref.py makes and gpurun ships; only when neither is present, the loop-faithful oracle port
# of the same per-particle algorithm.  Single-threaded Python either way: 1 core.
# ----------------------------------------------------------------------------------------------
def _oracle_model():
    from oracle import bioreactor
    return (bioreactor.increment_scalar, bioreactor.outputs_scalar, bioreactor.X_STEADY)


def _reference_parts():
    """(R, f, g, state, meas, x0) built from the reference's own classes, or None when it is not available."""
    import warnings
    from oracle import mixture, ref_loader
    if not ref_loader.available():
        return None
    warnings.simplefilter("ignore")
    R = ref_loader.load()
    MGS = R.gaussian_sum_dist.MultivariateGaussianSum
    f, g = R.model.Bioreactor.homeostatic_DEs, R.model.Bioreactor.static_outputs
    from oracle import bioreactor
    x_ss = numpy.asarray(bioreactor.X_STEADY)      # = Bioreactor.find_SS(...) (sim_base.py:46-53; tests pin the equality)
    state = MGS(numpy.zeros((2, 5)), mixture.STATE_COVS, numpy.array([0.75, 0.25]), library=numpy)
    meas = MGS(mixture.MEAS_MEANS, mixture.MEAS_COVS, numpy.array([0.85, 0.15]), library=numpy)
    x0 = MGS(numpy.zeros((2, 5)) + x_ss[None, :], mixture.STATE_COVS, numpy.array([0.75, 0.25]), library=numpy)
    return R, f, g, state, meas, x0, ("reference copy oracle/_ref" if ref_loader.is_copy() else "/root/reference")


def cpu_reference_run(log2n, steps, warmup, gsf=False):
    """predict(u, 1.) -> update(u, z) -> resample() of the reference's own CPU class; returns None if unavailable."""
    parts = _reference_parts()
    if parts is None:
        return None
    R, f, g, state, meas, x0, where = parts
    n = 1 << log2n
    numpy.random.seed(0)
    cls = R.filter.GaussianSumUnscentedKalmanFilter if gsf else R.filter.ParticleFilter
    flt = cls(f, g, n, x0, state, meas)
    us, zs = trajectory(steps + warmup, seed=1, model=(f, g, numpy.asarray(x0.means[0], dtype=numpy.float64)))
    times = []
    for k in range(steps + warmup):
        t = time.perf_counter()
        flt.predict(us[k], DT)
        flt.update(us[k], zs[k])
        flt.resample()
        dtm = time.perf_counter() - t
        if k >= warmup:
            times.append(dtm)
    total = sum(times)
    return n * len(times) / total, 1e3 * total / len(times), n, where


def cpu_port_run(log2n, steps, warmup, vectorised=False):
    from oracle import bioreactor, mixture, particle
    n = 1 << log2n
    state, meas = mixture.benchmark_noise()
    numpy.random.seed(0)
    pf = particle.ParticleFilterOracle(n, mixture.benchmark_x0(bioreactor.X_STEADY), state, meas)
    us, zs = trajectory(steps + warmup, seed=1, model=_oracle_model())
    times = []
    for k in range(steps + warmup):
        t = time.perf_counter()
        if vectorised:
            pf.predict(us[k], DT)
            pf.update(us[k], zs[k])
            pf.resample()
        else:
            pf.predict_loop(us[k], DT)
            pf.update_loop(us[k], zs[k])
            pf.resample(loop=True)
        dtm = time.perf_counter() - t
        if k >= warmup:
            times.append(dtm)
        # keep the population alive (weights of an open-loop run degenerate; the reference's
        # benchmark re-creates weights each run, pf_run_seq.py:124-125)
        if not numpy.isfinite(pf.particles).all():
            pf.particles = mixture.benchmark_x0(bioreactor.X_STEADY).draw(n)
    total = sum(times)
    return n * len(times) / total, 1e3 * total / len(times), n


def cpu_baseline(log2n, steps=10):
    """~10 s of the reference's own ParticleFilter on one host core (rank 0, N = 1 only)."""
    ref = cpu_reference_run(log2n, steps=steps, warmup=1)
    if ref is not None:
        value, ms, n, where = ref
        out = {"value": value, "unit": UNIT, "cores": 1, "kind": "reference",
               "sample": "the reference's own filter.ParticleFilter (numpy path, filter/particle.py:9-114; %s), 2^%d particles "
                         "x %d steps, %.1f ms/step; host has %d cores, the reference is single-threaded Python"
                         % (where, log2n, steps, ms, os.cpu_count())}
    else:
        value, ms, n = cpu_port_run(log2n - 1, steps=4, warmup=1)
        out = {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "oracle loop port (per-particle Python loops as filter/particle.py:54-103; the reference copy oracle/_ref "
                         "is absent), 2^%d particles x 4 steps, %.1f ms/step; host has %d cores" % (log2n - 1, ms, os.cpu_count())}
    vec_value, vec_ms, vec_n = cpu_port_run(20, steps=2, warmup=1, vectorised=True)
    out["vectorised_numpy_value"] = vec_value
    out["vectorised_numpy_sample"] = "oracle vectorised float64 numpy port, 2^20 particles x 2 steps, %.1f ms/step" % vec_ms
    return out


def cpu_baseline_gsf(log2n=9, steps=8):
    """GS-UKF on the host: the reference's own GaussianSumUnscentedKalmanFilter (N x 11 Python calls of f per predict,
    gs_ukf.py:82-103); the vectorised float64 oracle port when the reference is not available."""
    ref = cpu_reference_run(log2n, steps=steps, warmup=1, gsf=True)
    if ref is not None:
        value, ms, n, where = ref
        return {"value": value, "unit": "components/s", "cores": 1, "kind": "reference",
                "sample": "the reference's own filter.GaussianSumUnscentedKalmanFilter (numpy path, filter/gs_ukf.py:9-183; %s), "
                          "2^%d components x %d steps, %.1f ms/step" % (where, log2n, steps, ms)}
    from oracle import bioreactor, gs_ukf, mixture
    n = 1 << 10
    state, meas = mixture.benchmark_noise()
    numpy.random.seed(0)
    f = gs_ukf.GSUKFOracle(n, mixture.benchmark_x0(bioreactor.X_STEADY), state, meas)
    us, zs = trajectory(6, seed=1, model=_oracle_model())
    times = []
    for k in range(6):
        t = time.perf_counter()
        f.predict(us[k], DT)
        f.update(us[k], zs[k])
        f.resample()
        if k:
            times.append(time.perf_counter() - t)
    total = sum(times)
    return {"value": n * len(times) / total, "unit": "components/s", "cores": 1, "kind": "port",
            "sample": "oracle vectorised float64 numpy port of filter/gs_ukf.py:45-171, 2^10 components x %d steps, "
                      "%.1f ms/step" % (len(times), 1e3 * total / len(times))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    log2n = args.cpu_log2n
    t0 = time.perf_counter()
    ref = cpu_reference_run(log2n, args.steps, max(args.warmup, 1))
    if ref is not None:
        value, ms, n, where = ref
        kind = "reference"
        sample = ("the reference's own filter.ParticleFilter (numpy path, filter/particle.py:9-114; %s): predict(u, 1.) -> "
                  "update(u, z) -> resample(), 2^%d particles x %d steps" % (where, log2n, args.steps))
    else:
        value, ms, n = cpu_port_run(log2n, args.steps, max(args.warmup, 1))
        kind = "port"
        sample = ("oracle loop port of filter/particle.py:54-103 (reference copy oracle/_ref absent), 2^%d particles x %d steps"
                  % (log2n, args.steps))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.log2n),
                       "sample": "bounded CPU sample of 2^%d particles per step (per-particle Python loops: cost is linear in N)" % log2n,
                       "particles_per_step": n},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    emit(line)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    elif args.sharded:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)

    import gpu_se_b200 as g
    from gpu_se_b200.model.BioreactorModel import X_STEADY

    n_local = int(round(2 ** args.log2n))
    n_total = n_local * world
    state_means, state_covs = numpy.zeros((2, 5)), numpy.array([numpy.diag([1e-4, 1e-7, 1e-3, 1e-3, 1e-7]),
                                                                 numpy.diag([1e-3, 1e-6, 1e-2, 1e-2, 1e-6])])
    state = g.MultivariateGaussianSum(state_means, state_covs, [0.75, 0.25])
    meas = g.MultivariateGaussianSum([[1e-1, 0], [0, -1e-1]], [[[6e-2, 0], [0, 8e-2]], [[500, 100], [100, 700]]],
                                     [0.85, 0.15])
    x0 = g.MultivariateGaussianSum(state_means + numpy.array(X_STEADY)[None, :], state_covs, [0.75, 0.25])
    f, gg = g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs
    parity = None
    if world > 1 or args.sharded:
        from gpu_se_b200.sharded import ShardedParticleFilter
        parity = sharded_parity(g, dev, world, rank)
        pf = ShardedParticleFilter(f, gg, n_total, x0, state, meas, device=dev, seed=1234)
    elif args.workload == "gsf":
        pf = g.GaussianSumUnscentedKalmanFilter(f, gg, n_total, x0, state, meas, device=dev, seed=1234)
    else:
        pf = g.ParticleFilter(f, gg, n_total, x0, state, meas, device=dev, seed=1234, peer=args.peer_buffers)
    if args.graphs:
        pf.enable_graphs()
    from gpu_se_b200.filter import _base
    two_stage = not _base.FUSED_RESAMPLE and not (world > 1 or args.sharded)

    K, W = args.steps, max(args.warmup, 3)
    us, zs = trajectory(2 * (K + W), seed=7)
    numpy.random.seed(99)                       # resample offsets r: numpy.random.rand() as the reference
    rs = numpy.random.rand(2 * (K + W))
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-timed region: particles resident in HBM, per-stage CUDA events ----------------
    stage_events = []

    def hook(label):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream)
        stage_events.append((label, ev))

    fused_update = bool(pf._fuses_update())      # predict() is recorded and runs inside update()'s kernel

    def step(k, record):
        record = record and not args.graphs     # a graph replay is one launch: only the whole step can be timed
        if record:
            hook("start")
        pf.predict(us[k], DT)
        if record and not fused_update:
            hook("predict")
        pf.update(us[k], zs[k])
        if record:
            hook("predict+update" if fused_update else "update")
        pf._stage_hook = hook if record else None
        pf.resample(r=float(rs[k]))
        pf._stage_hook = None
        if record:
            hook("search" if two_stage else "resample")

    for k in range(W):
        step(k, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def fill_until_sample(timeout):
        """Keep the sampled GPU under load (streaming passes over a scratch buffer: nothing of the filter is touched, its
        trajectory stays what it was) until the clock sampler has just delivered a line.  Every nvidia-smi query stalls
        the launches of the sampled GPU for a millisecond or two -- 4 % of a 50 ms region, landing in ONE step.  The timed
        region therefore starts right behind a sample: a region shorter than the sampling period has samples under load
        on both sides of it and none of these stalls inside; longer regions simply contain them.  Only rank 0 is
        sampled; the other ranks wait for it in the barrier that follows."""
        if rank != 0 or sampler.proc is None:
            return
        n0, t_begin = len(sampler.lines), time.perf_counter()
        while len(sampler.lines) == n0 and time.perf_counter() - t_begin < timeout:
            for _ in range(8):
                scratch.mul_(1.0)
            torch.cuda.synchronize(dev)

    scratch = torch.ones(1 << 26, dtype=torch.float32, device=dev) if rank == 0 else None      # 256 MB: larger than L2
    fill_until_sample(3.0)                      # (nvidia-smi takes a few hundred milliseconds to deliver its first line)
    barrier()                                   # every rank enters the timed region together
    launches0 = pf._ctx.launches
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for k in range(W, W + K):
        step(k, True)
    pf._materialise()                           # the last resample's pending gather belongs to the timed region
    ev1.record(stream)
    t_issued = time.perf_counter()              # host side done issuing: close to the device time = launch-bound
    barrier()
    t_wall1 = time.perf_counter()
    launches = pf._ctx.launches - launches0
    dev_ms = ev0.elapsed_time(ev1)
    fill_until_sample(0.5)                      # ... and the sample that closes the bracket is taken under load as well
    torch.cuda.synchronize(dev)
    clocks = sampler.stop(t_wall0, time.perf_counter()) if rank == 0 else None

    # per-stage durations
    stage_ms = {}
    prev = None
    for label, ev in stage_events:
        if label != "start" and prev is not None:
            stage_ms.setdefault(label, []).append(prev.elapsed_time(ev))
        prev = ev
    stage_avg = {k: sum(v) / len(v) for k, v in stage_ms.items()}

    # ---- end-to-end region: public API, host u / z in, estimate out, every step ---------------
    barrier()
    t0 = time.perf_counter()
    est = None
    for k in range(W + K, W + 2 * K):
        pf.predict(us[k], DT)
        pf.update(us[k], zs[k])
        pf.resample(r=float(rs[k]))
        est = pf.point_estimate()               # D2H read of the step's result (synchronises)
    barrier()
    e2e_s = time.perf_counter() - t0

    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    lsum = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(lsum, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms = float(times[0]), float(times[1])
    if hasattr(pf, "close"):
        pf.close()                              # collective: unmap peer buffers before anyone frees them
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    traffic_row, traffic_src = ncu_traffic_bytes_per_row()
    STAGE_BYTES = globals()["STAGE_BYTES_SHARDED" if world > 1 else "STAGE_BYTES"]
    if args.workload == "gsf":      # 20 floats per component, lazy resample
        STAGE_BYTES = {"predict": 4 + 80 + 80, "update": 80 + 80 + 4, "resample": 8, "scan": 12, "search": 12}
    STEP_BYTES = sum(v for k, v in STAGE_BYTES.items() if k in stage_avg) if stage_avg else 64
    value = n_total * K / (dev_ms * 1e-3)
    dom = max(stage_avg, key=stage_avg.get) if stage_avg else "predict"
    dom_ms = stage_avg.get(dom, dev_ms / K)
    achieved = STAGE_BYTES.get(dom, STEP_BYTES) * n_local / (dom_ms * 1e-3) / 1e9
    stages = {k: {"ms": round(v, 4), "bytes_per_particle": STAGE_BYTES.get(k),
                  "gbs": round(STAGE_BYTES.get(k, 0) * n_local / (v * 1e-3) / 1e9, 1),
                  "frac": round(STAGE_BYTES.get(k, 0) * n_local / (v * 1e-3) / 1e9 / peak, 4)}
              for k, v in stage_avg.items()}
    line = {
        "metric": METRIC if args.workload == "pf" else "GSF comps/sec (predict+update+resample)", "value": value,
        "unit": UNIT if args.workload == "pf" else "components/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.log2n) if args.workload == "pf" else
                   "gsf_openloop predict+update+resample, BioreactorModel, 2^%g components, dt=1.0%s"
                   % (args.log2n, ", CUDA-graph replay" if args.graphs else ""),
                   "particles_total": n_total, "particles_per_gpu": n_local,
                   "parallelism": "shard%d" % world if world > 1 else "single",
                   "l2": "inputs larger than L2 (state %.0f MB per GPU vs 126 MB L2)" % (n_local * 20 / 1e6)},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (traffic_row[dom] * n_local if dom in traffic_row else None),
                     "traffic_source": "ncu --set full capture of this command at 2^24 rows (%s), scaled by rows; "
                                       "per launch of the stage's kernels" % traffic_src,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_particle": STAGE_BYTES.get(dom),
                     "whole_step_frac": STEP_BYTES * n_local / (dev_ms / K * 1e-3) / 1e9 / peak,
                     "whole_step_bytes_per_particle": STEP_BYTES,
                     "survey_step_frac": 128 * n_local / (dev_ms / K * 1e-3) / 1e9 / peak},
        "stages": stages,
        "e2e": {"value": n_total * K / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / K,
                "h2d_bytes_per_step": 48, "d2h_bytes_per_step": 48 * 8,
                "note": "u, z are host arrays passed per call; particles stay resident (as in the reference's GPU "
                        "class); point_estimate() read back every step"},
        "host_issue_ms_per_step": (t_issued - t_wall0) * 1e3 / K,
        "gpu_launches": int(lsum[0]),
        "sharded_parity": parity,
        "clocks": clocks,
        "last_estimate": [float(v) for v in est],
    }
    if world == 1 and args.workload == "pf" and not args.no_gsf and not args.sharded:
        line["gsf"] = gsf_block(g, dev, (x0, state, meas), peak, with_cpu=not args.no_cpu_baseline)
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args.cpu_log2n + 1) if args.workload == "pf" else cpu_baseline_gsf()
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# algorithmic bytes per Gaussian component and stage: state = 5 means + 15 covariance entries (lower
# triangle) in float32, lazy resample (the gather of 80 B rides inside predict through the int32 ancestor index)
GSF_STAGE_BYTES = {"predict": 4 + 80 + 80, "update": 80 + 80 + 4, "resample": 8}


def run_gsf(g, dev, pdfs, log2n, steps, warmup, graphs):
    """GS-UKF predict -> update -> resample at 2^log2n components (BASELINE.json metric "GSF comps/sec", configs[3]):
    device-timed components/s and, without graphs, the per-stage times."""
    import torch
    x0, state, meas = pdfs
    n = 1 << log2n
    f, gg = g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs
    gf = g.GaussianSumUnscentedKalmanFilter(f, gg, n, x0, state, meas, device=dev, seed=4321)
    if graphs:
        gf.enable_graphs()
    us, zs = trajectory(steps + warmup, seed=17)
    rs = numpy.random.default_rng(3).random(steps + warmup)
    stream = torch.cuda.current_stream(dev)
    events = []

    def hook(label):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream)
        events.append((label, ev))

    for k in range(steps + warmup):
        if k == warmup:
            torch.cuda.synchronize(dev)
            launches0 = gf._ctx.launches
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
        rec = k >= warmup and not graphs
        if rec:
            hook("start")
        gf.predict(us[k], DT)
        if rec:
            hook("predict")
        gf.update(us[k], zs[k])
        if rec:
            hook("update")
        gf.resample(r=float(rs[k]))
        if rec:
            hook("resample")
    gf._materialise()
    ev1 = torch.cuda.Event(enable_timing=True)
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1) / steps
    est = gf.point_estimate()
    out = {"components": n, "value": n / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "cuda_graphs": bool(graphs),
           "gpu_launches": int(gf._ctx.launches - launches0), "last_estimate": [float(v) for v in est]}
    if events:
        acc, prev = {}, None
        for label, ev in events:
            if label != "start" and prev is not None:
                acc.setdefault(label, []).append(prev.elapsed_time(ev))
            prev = ev
        out["stages"] = {k: round(sum(v) / len(v), 4) for k, v in acc.items()}
    del gf
    torch.cuda.empty_cache()
    return out


def gsf_block(g, dev, pdfs, peak, with_cpu):
    """The "gsf" sub-object of the default bench line: BASELINE configs[3] at its top size (2^16 components, CUDA-graph
    replay: the sweep is launch-bound) and at 2^20 (out of the launch-bound regime) with the roofline of the step."""
    small = run_gsf(g, dev, pdfs, 16, 300, 20, graphs=True)
    small_eager = run_gsf(g, dev, pdfs, 16, 100, 10, graphs=False)
    big = run_gsf(g, dev, pdfs, 20, 60, 10, graphs=False)
    n = big["components"]
    st = big.get("stages", {})
    hbm_ms = {k: GSF_STAGE_BYTES[k] * n / (peak * 1e9) * 1e3 for k in GSF_STAGE_BYTES}
    # issue-slot floor: warp-instructions per component measured with ncu (profiles/r2_gsf_2p20_ncu_full.csv) at one
    # instruction per SM sub-partition and clock: 148 SMs x 4 x 1.965 GHz
    issue_ms = {k: GSF_WARP_INSTR_PER_COMPONENT[k] * n / (148 * 4 * 1.965e9) * 1e3 for k in GSF_WARP_INSTR_PER_COMPONENT}
    bound_ms = {k: max(hbm_ms[k], issue_ms.get(k, 0.0)) for k in hbm_ms}
    block = {"metric": "GSF comps/sec (predict+update+resample)", "unit": "components/s",
             "config": {"workload": "gsf_openloop predict+update+resample, BioreactorModel, dt=1.0, in-kernel Philox noise",
                        "sizes": "2^16 components (configs[3] top size; CUDA-graph replay and eager), 2^20 components (eager)"},
             "2p16_cuda_graphs": small, "2p16_eager": small_eager, "2p20": big,
             "roofline": {"at": "2^20 components", "bytes_per_component": GSF_STAGE_BYTES,
                          "hbm_floor_ms": {k: round(v, 4) for k, v in hbm_ms.items()},
                          "issue_floor_ms": {k: round(v, 4) for k, v in issue_ms.items()},
                          "warp_instructions_per_component": GSF_WARP_INSTR_PER_COMPONENT,
                          "bound": {k: ("issue" if issue_ms.get(k, 0.0) > hbm_ms[k] else "hbm") for k in hbm_ms},
                          "frac_of_bound": {k: round(bound_ms[k] / st[k], 4) for k in bound_ms if k in st},
                          "step_frac_of_bound": round(sum(bound_ms.values()) / big["ms_per_step"], 4),
                          "step_frac_of_hbm": round(sum(hbm_ms.values()) / big["ms_per_step"], 4), "peak_gbs": peak}}
    if with_cpu:
        block["cpu_baseline"] = cpu_baseline_gsf()
    return block


# warp-instructions executed per component (smsp__inst_executed.sum / components, ncu at 2^20:
# profiles/r2_gsf_2p20_ncu_full.csv -- 78.68 M and 31.73 M warp-instructions for 2^20 components)
GSF_WARP_INSTR_PER_COMPONENT = {"predict": 75.03, "update": 30.26}


def sharded_parity(g, dev, world, rank, n=(1 << 20) + 8):
    """World >= 2 parity cannot run in the driver's 1-GPU test tier: before timing, every `bench.py --gpus N` run checks
    that N shards of one population reproduce the single-GPU filter of the same seed BIT FOR BIT (particles after
    predict / update / resample cycles, resample -> resample, assigned weights) and that the global estimates agree."""
    import torch
    import torch.distributed as dist
    from gpu_se_b200.model.BioreactorModel import X_STEADY
    from gpu_se_b200.sharded import ShardedParticleFilter
    sm, sc = numpy.zeros((2, 5)), numpy.array([numpy.diag([1e-4, 1e-7, 1e-3, 1e-3, 1e-7]),
                                               numpy.diag([1e-3, 1e-6, 1e-2, 1e-2, 1e-6])])
    state = g.MultivariateGaussianSum(sm, sc, [0.75, 0.25])
    meas = g.MultivariateGaussianSum([[1e-1, 0], [0, -1e-1]], [[[6e-2, 0], [0, 8e-2]], [[500, 100], [100, 700]]], [0.85, 0.15])
    x0 = g.MultivariateGaussianSum(sm + numpy.array(X_STEADY)[None, :], sc, [0.75, 0.25])
    f, gg = g.Bioreactor.homeostatic_DEs, g.Bioreactor.static_outputs
    us, zs = trajectory(4, seed=11)
    wts = numpy.random.default_rng(5).random(n) ** 6

    def cycles(pf, assign):
        ests = []
        for k in range(3):
            pf.predict(us[k], DT)
            pf.update(us[k], zs[k])
            pf.resample(r=0.1 + 0.3 * k)
            ests.append(pf.point_estimate())
        pf.resample(r=0.77)                         # resample -> resample, no update in between
        assign(pf, wts)
        pf.resample(r=0.41)
        ests.append(pf.point_estimate())
        pf.predict(us[3], DT)
        return ests

    spf = ShardedParticleFilter(f, gg, n, x0, state, meas, device=dev, seed=4321)
    got = cycles(spf, lambda p, w: p.set_global_weights(w))
    parts = [torch.empty((b - a, 5), dtype=torch.float32, device=dev) for a, b in spf.bounds]
    for s in range(world):
        if s == rank:
            parts[s].copy_(spf.particles)
        dist.broadcast(parts[s], src=s)
    ok = torch.ones(1, dtype=torch.int32, device=dev)
    if rank == 0:
        pf = g.ParticleFilter(f, gg, n, x0, state, meas, device=dev, seed=4321)

        def assign(p, w):
            p.weights = w
        ref = cycles(pf, assign)
        same = bool(torch.equal(torch.cat(parts), pf.particles.as_subclass(torch.Tensor)))
        same = same and all(numpy.allclose(a, b, rtol=1e-9, atol=1e-12) for a, b in zip(got, ref))
        ok[0] = 1 if same else 0
        del pf
    dist.broadcast(ok, src=0)
    spf.close()
    del spf
    torch.cuda.empty_cache()
    return bool(int(ok[0]))


def emit(line):
    """The one JSON line goes to the real stdout; everything else that lands on fd 1 (NCCL's version
    banner, library chatter) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
