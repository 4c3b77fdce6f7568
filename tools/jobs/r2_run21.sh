cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for m in "" "--sharded" "--peer-buffers"; do
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gsf $m > gpurun_out/r2p_bench$m.json 2> gpurun_out/r2p_bench$m.err
python - "$m" <<'PY'
import json,sys
m=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2p_bench%s.json"%m))
    print("bench",m, round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4), d["gpu_launches"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2p_bench%s.err"%m).read()[-2000:])
PY
done
GSE_FUSE_UPDATE=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gsf --peer-buffers > gpurun_out/r2p_bench_unfused_peer.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2p_bench_unfused_peer.json"))
print("unfused, peer buffers", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4))
PY
