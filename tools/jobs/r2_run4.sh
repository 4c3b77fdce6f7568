set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_resample.py tests/test_gpu_particle.py tests/test_gpu_lazy.py tests/test_gpu_graphs.py -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
b() { name=$1; shift; env "$@" python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2d_bench_$name.json 2> gpurun_out/r2d_bench_$name.err; }
b default A=1
b fmin3 GSE_FUSED_MINB=3
b both GSE_FUSED_MINB=3 GSE_PREDICT_MINB=4
python - <<'PY'
import json
for n in ("default","fmin3","both"):
    try:
        d=json.load(open("gpurun_out/r2d_bench_%s.json"%n))
        print(n, round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, round(d["e2e"]["ms_per_step"],4))
    except Exception as e:
        print(n, "failed", e)
PY
export GSE_FUSED_MINB=3 GSE_PREDICT_MINB=4
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_resample_fused|k_pf_update|k_pf_predict' -s 9 -c 3 -o gpurun_out/prof_r2d python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
