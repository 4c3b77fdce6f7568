cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
for m in "" "--sharded"; do
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gsf $m > gpurun_out/r2l_bench$m.json 2> gpurun_out/r2l_bench$m.err
python - "$m" <<'PY'
import json,sys
m=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2l_bench%s.json"%m))
    print("bench",m, round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, round(d["e2e"]["ms_per_step"],4))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2l_bench%s.err"%m).read()[-2000:])
PY
done
