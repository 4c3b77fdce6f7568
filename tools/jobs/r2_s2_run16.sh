# session 2, run 16 (2 GPUs): the driver's launch line for N = 2 with the sampler-aligned bench (short and default runs)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { tag=$1; shift;
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > gpurun_out/r2_scale_$tag.json 2> gpurun_out/r2_scale_$tag.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale_$tag.json"))
    print("$tag", d["steps"], round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4), "parity", d["sharded_parity"], "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], d["clocks"])
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2_scale_$tag.err").read()[-2500:])
PY
}
run g2_k20 --steps 20 --warmup 5
run g2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
