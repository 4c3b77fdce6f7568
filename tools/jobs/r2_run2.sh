set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resample.py tests/test_gpu_particle.py tests/test_gpu_lazy.py -x -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2b_bench_fused8.json 2> gpurun_out/r2b_bench_fused8.err
GSE_FUSED_ITEMS=16 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2b_bench_fused16.json 2> gpurun_out/r2b_bench_fused16.err
python - <<'PY'
import json
for n in ("fused8","fused16"):
    try:
        d=json.load(open("gpurun_out/r2b_bench_%s.json"%n))
        print(n, d["ms_per_step"], {k:v["ms"] for k,v in d["stages"].items()}, d["e2e"]["ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_resample_fused|k_pf_update|k_pf_predict' -s 9 -c 6 -o gpurun_out/prof_r2b python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
