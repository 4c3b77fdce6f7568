cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_pytest.log
tail -5 gpurun_out/r2l_pytest.log
for n in 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/r2l_bench_g$n.json 2> gpurun_out/r2l_bench_g$n.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2l_bench_g$n.json"))
    print("g$n", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4), "parity", d["sharded_parity"], "value %.3g"%d["value"])
except Exception as e:
    print("g$n failed", e); print(open("gpurun_out/r2l_bench_g$n.err").read()[-2500:])
PY
done
