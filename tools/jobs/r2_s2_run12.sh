# session 2, run 12 (1 GPU): persistent GS-UKF update kernel, occupancy sweep
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { echo "== $1 $2"; env $1 $2 python bench.py --workload gsf --log2n 20 --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/s2_gsf_t.json 2> gpurun_out/s2_gsf_t.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/s2_gsf_t.json")); print(round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()})
except Exception as e:
    print("failed", e); print(open("gpurun_out/s2_gsf_t.err").read()[-800:])
PY
}
run GSE_GSF_MINB=4 GSE_X=0
run GSE_GSF_MINB=3 GSE_X=0
run GSE_GSF_MINB=5 GSE_GSF_UPDATE_WAVES=2
run GSE_GSF_MINB=4 GSE_GSF_UPDATE_WAVES=2
