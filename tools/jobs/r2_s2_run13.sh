# session 2, run 13 (1 GPU): software-pipelined PF update kernel A/B, GS-UKF persistent update default; tests
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_particle.py tests/test_gpu_gsukf.py tests/test_gpu_fused_update.py tests/test_gpu_graphs.py -q -m gpu 2>&1 | tail -3
run() { echo "== $1 $2"; env $1 $2 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-gsf > gpurun_out/s2_t.json 2> gpurun_out/s2_t.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/s2_t.json")); print(round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4))
except Exception as e:
    print("failed", e); print(open("gpurun_out/s2_t.err").read()[-800:])
PY
}
run GSE_UPDATE_PIPE=1 GSE_X=0
run GSE_UPDATE_PIPE=0 GSE_X=0
run GSE_UPDATE_PIPE=1 GSE_UPDATE_CTAS=3
run GSE_UPDATE_PIPE=1 GSE_UPDATE_CTAS=2
python bench.py --workload gsf --log2n 20 --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('gsf 2^20', round(d['ms_per_step'],4), {k:v['ms'] for k,v in d['stages'].items()})"
