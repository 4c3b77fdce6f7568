# session 2, run 6 (1 GPU): L2 fetch granularity hint (32 / 64 / 128 B) against the gathers through the ancestor index
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for g in 64 32 128; do
  echo "== GSE_L2_FETCH=$g"
  GSE_L2_FETCH=$g python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-gsf > gpurun_out/s2_l2f_$g.json 2> gpurun_out/s2_l2f_$g.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s2_l2f_$g.json"))
print(round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4))
PY
  GSE_L2_FETCH=$g python tools/e2e_stages.py 2>&1 | tail -3
done
