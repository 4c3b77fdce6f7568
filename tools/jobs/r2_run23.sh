cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -k "world1" 2>&1 | tail -3
run() { # label, env...
  lbl=$1; shift
  env "$@" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gsf > gpurun_out/r2r_$lbl.json 2> gpurun_out/r2r_$lbl.err
  python - $lbl <<'PY'
import json,sys
l=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2r_%s.json"%l))
    print(l, round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4))
except Exception as e:
    print("failed", l, e)
PY
}
run base A=1
run pf5 GSE_UPDATE_PREFETCH=1
run pf4 GSE_UPDATE_PREFETCH=1 GSE_UPDATE_CTAS=4
run pf6 GSE_UPDATE_PREFETCH=1 GSE_UPDATE_CTAS=6
run pf3 GSE_UPDATE_PREFETCH=1 GSE_UPDATE_CTAS=3
run np6 GSE_UPDATE_CTAS=6
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gsf"
ncu --set full --clock-control none --import-source on -k regex:'k_pf_predict' -s 8 -c 2 -o gpurun_out/r2r_pred_single $B > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_pf_predict' -s 8 -c 2 -o gpurun_out/r2r_pred_sharded $B --sharded > gpurun_out/ncu_b.log 2>&1
tail -1 gpurun_out/ncu_a.log gpurun_out/ncu_b.log
