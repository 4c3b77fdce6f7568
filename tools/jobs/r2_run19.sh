cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graphs.py tests/test_gpu_misc.py tests/test_gpu_gsukf.py tests/test_gpu_closed_loop.py -x -q 2>&1 | tail -5
python tools/host_issue.py
python tools/host_issue.py --sharded
python tools/host_issue.py --estimate
python tools/host_issue.py --sharded --estimate
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gsf --sharded > gpurun_out/r2n_bench_sh.json 2> gpurun_out/r2n_bench_sh.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2n_bench_sh.json"))
print("bench sharded", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4), "host issue", round(d["host_issue_ms_per_step"],4), d["gpu_launches"])
PY
