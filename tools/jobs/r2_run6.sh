set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest.log
tail -15 gpurun_out/r2f_pytest.log
b() { name=$1; shift; env "$@" python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2f_bench_$name.json 2> gpurun_out/r2f_bench_$name.err; }
b default GSE_PREDICT_MINB=4
b fmin3 GSE_FUSED_MINB=3 GSE_PREDICT_MINB=4
python - <<'PY'
import json
for n in ("default","fmin3"):
    try:
        d=json.load(open("gpurun_out/r2f_bench_%s.json"%n))
        print(n, round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, round(d["e2e"]["ms_per_step"],4))
    except Exception as e:
        print(n, "failed", e)
PY
GSE_FUSED_TRACE=1 GSE_FUSED_MINB=3 python tools/fused_trace.py 24 > gpurun_out/r2f_trace3.txt 2>&1
GSE_FUSED_TRACE=1 python tools/fused_trace.py 24 > gpurun_out/r2f_trace4.txt 2>&1
tail -12 gpurun_out/r2f_trace3.txt; tail -6 gpurun_out/r2f_trace4.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; cat gpurun_out/r2f_bench_reference.json | cut -c1-600
