cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest.log
tail -30 gpurun_out/r2i_pytest.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python bench.py --steps 50 --warmup 10 --no-cpu-baseline --sharded > gpurun_out/r2i_bench_sh1.json 2> gpurun_out/r2i_bench_sh1.err
python - <<'PY'
import json
for n in ("bench","bench_sh1"):
    try:
        d=json.load(open("gpurun_out/r2i_%s.json"%n))
        print(n, round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, round(d["e2e"]["ms_per_step"],4))
    except Exception as e:
        print(n, "failed", e); print(open("gpurun_out/r2i_%s.err"%n).read()[-1500:])
PY
