# session 2, run 15 (1 GPU): bench with the timed region aligned behind a clock sample; driver-style short run too
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_g1.json 2> gpurun_out/r2_bench_g1.err; tail -c 300 gpurun_out/r2_bench_g1.err
python bench.py --steps 20 --warmup 5 --no-gsf --no-cpu-baseline > gpurun_out/s2_bench_k20.json 2> gpurun_out/s2_bench_k20.err
python bench.py --steps 20 --warmup 5 --no-gsf --no-cpu-baseline > gpurun_out/s2_bench_k20b.json 2> gpurun_out/s2_bench_k20b.err
python - <<'PY'
import json
for t in ("r2_bench_g1", "s2_bench_k20", "s2_bench_k20b"):
    try:
        d=json.load(open("gpurun_out/%s.json" % t))
        print(t, d["steps"], round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4), d["clocks"])
    except Exception as e:
        print(t, "failed", e); print(open("gpurun_out/%s.err" % t).read()[-1500:])
PY
