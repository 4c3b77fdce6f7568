cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
GSE_FUSED_TRACE=1 python tools/fused_trace.py 24 --sharded 2>&1 | grep "^step 1[01]:\|Error\|error" 
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gsf --sharded > gpurun_out/r2v_sh.json 2> gpurun_out/r2v_sh.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2v_sh.json"))
print("sharded", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4))
PY
