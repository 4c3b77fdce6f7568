# session 2, run 14 (2 GPUs): lane-parallel moments merge + result block in the sharded estimate path; block order A/B
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
T0=$(date +%s)
run() { n=$1; l2=$2; steps=$3; tag=$4; shift 4;
env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps $steps --warmup 10 --log2n $l2 > gpurun_out/r2_scale_$tag.json 2> gpurun_out/r2_scale_$tag.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_$tag.json"))
    print("$tag", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4), "parity", d["sharded_parity"], "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"])
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2_scale_$tag.err").read()[-2500:])
PY
}
run 2 24 100 g2 GSE_X=0
run 2 24 100 g2_ends GSE_PREDICT_ENDS_FIRST=1
echo "t=$(( $(date +%s) - T0 ))"
timeout 300 python -m pytest tests/test_gpu_sharded.py -q -m gpu -k "2-peer or (estimates and 2-)" > gpurun_out/s2_pytest_w2b.log 2>&1; echo "rc=$?" >> gpurun_out/s2_pytest_w2b.log
tail -4 gpurun_out/s2_pytest_w2b.log
echo "t=$(( $(date +%s) - T0 ))"
