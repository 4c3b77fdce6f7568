set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_resample.py -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_pytest.log
tail -4 gpurun_out/r2e_pytest.log
b() { name=$1; shift; env "$@" python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2e_bench_$name.json 2> gpurun_out/r2e_bench_$name.err; }
b default GSE_PREDICT_MINB=4
b fmin3 GSE_FUSED_MINB=3 GSE_PREDICT_MINB=4
python - <<'PY'
import json
for n in ("default","fmin3"):
    try:
        d=json.load(open("gpurun_out/r2e_bench_%s.json"%n))
        print(n, round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, round(d["e2e"]["ms_per_step"],4))
    except Exception as e:
        print(n, "failed", e)
PY
GSE_FUSED_TRACE=1 GSE_FUSED_MINB=3 python tools/fused_trace.py 24 > gpurun_out/r2e_trace3.txt 2>&1
GSE_FUSED_TRACE=1 python tools/fused_trace.py 24 > gpurun_out/r2e_trace4.txt 2>&1
cat gpurun_out/r2e_trace3.txt gpurun_out/r2e_trace4.txt
