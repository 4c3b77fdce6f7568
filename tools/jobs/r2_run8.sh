cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_particle.py tests/test_gpu_lazy.py tests/test_philox.py -x -q -m gpu 2>&1 | tail -3
b() { name=$1; shift; env "$@" python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2h_bench_$name.json 2> gpurun_out/r2h_bench_$name.err; }
b p5 GSE_PREDICT_MINB=5
b p4 GSE_PREDICT_MINB=4
b p104 GSE_PREDICT_MINB=104
b p105 GSE_PREDICT_MINB=105
python - <<'PY'
import json
for n in ("p5","p4","p104","p105"):
    try:
        d=json.load(open("gpurun_out/r2h_bench_%s.json"%n))
        print(n, round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, round(d["e2e"]["ms_per_step"],4))
    except Exception as e:
        print(n, "failed", e)
PY
