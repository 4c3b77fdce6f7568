cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { # label, env...
  lbl=$1; shift
  env "$@" python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gsf $EXTRA > gpurun_out/r2t_$lbl.json 2> gpurun_out/r2t_$lbl.err
  python - $lbl <<'PY'
import json,sys
l=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2t_%s.json"%l))
    print(l, round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4))
except Exception as e:
    print("failed", l, e)
PY
}
run base A=1
run table GSE_DEBUG_PREDICT_TABLE=1
EXTRA=--sharded run sharded A=1
