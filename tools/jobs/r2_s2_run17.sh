# session 2, run 17 (4 GPUs): the driver's launch line for N = 4 with the final code
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/r2_scale_g4.json 2> gpurun_out/r2_scale_g4.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_scale_g4.json"))
    print("g4", d["steps"], round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4), "parity", d["sharded_parity"], "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], d["clocks"])
except Exception as e:
    print("g4 failed", e); print(open("gpurun_out/r2_scale_g4.err").read()[-2500:])
PY
