set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L
./tools/ubench_xu.bin > gpurun_out/r2_ubench.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_resample.py -x -q -m gpu -s > gpurun_out/r2_pytest_resample.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_resample.log
tail -30 gpurun_out/r2_pytest_resample.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_all.log
tail -15 gpurun_out/r2_pytest_all.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_fused8.json 2> gpurun_out/r2_bench_fused8.err
GSE_FUSED_ITEMS=16 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_fused16.json 2> gpurun_out/r2_bench_fused16.err
GSE_RESAMPLE=unfused python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_unfused.json 2> gpurun_out/r2_bench_unfused.err
python - <<'PY'
import json
for n in ("fused8","fused16","unfused"):
    try:
        d=json.load(open("gpurun_out/r2_bench_%s.json"%n))
        print(n, d["ms_per_step"], {k:v["ms"] for k,v in d["stages"].items()}, d["e2e"]["ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
