# session 2, run 11 (1 GPU): persistent GS-UKF update kernel: tests + stage times for a few grid shapes
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gsukf.py -q -m gpu 2>&1 | tail -3
run() { echo "== $1"; env $1 python bench.py --workload gsf --log2n $2 --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/s2_gsf_t.json 2> gpurun_out/s2_gsf_t.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/s2_gsf_t.json")); print(round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/s2_gsf_t.err").read()[-800:])
PY
}
run GSE_X=0 20
run GSE_GSF_UPDATE_WAVES=2 20
run GSE_GSF_UPDATE_WAVES=4 20
run GSE_GSF_UPDATE_WAVES=64 20
run GSE_GSF_MINB=5 20
run GSE_X=0 16
