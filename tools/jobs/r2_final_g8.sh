# round-2 multi-GPU evidence on one 8-GPU box: world 4 / 8 parity tests, weak scaling at 2^24 particles per GPU, 2^30 total
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu -k "8-peer or (estimates and 8-) or (fused and 4-)" > gpurun_out/r2_final_pytest8.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final_pytest8.log
tail -4 gpurun_out/r2_final_pytest8.log
run() { n=$1; l2=$2; steps=$3; tag=$4;
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps $steps --warmup 10 --log2n $l2 > gpurun_out/r2_scale_$tag.json 2> gpurun_out/r2_scale_$tag.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_$tag.json"))
    print("$tag", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4), "parity", d["sharded_parity"], "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"])
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2_scale_$tag.err").read()[-2500:])
PY
}
run 8 24 60 g8
run 4 24 60 g4
run 8 27 20 g8_2p30
