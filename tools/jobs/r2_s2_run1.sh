# session 2, run 1 (2 GPUs): post-optimisation sharded numbers at 2^24 / 2^25 per GPU, world-1 sharded code path, world-2 parity tests
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
T0=$(date +%s)
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
run 2 24 100 g2
echo "t=$(( $(date +%s) - T0 ))"
run 2 25 60 g2_2p26
echo "t=$(( $(date +%s) - T0 ))"
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-gsf > gpurun_out/s2_single.json 2> gpurun_out/s2_single.err
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-gsf --sharded > gpurun_out/s2_w1sharded.json 2> gpurun_out/s2_w1sharded.err
python - <<'PY'
import json
for t in ("single", "w1sharded"):
    try:
        d=json.load(open("gpurun_out/s2_%s.json" % t))
        print(t, round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4))
    except Exception as e:
        print(t, "failed", e); print(open("gpurun_out/s2_%s.err" % t).read()[-1500:])
PY
echo "t=$(( $(date +%s) - T0 ))"
timeout 420 python -m pytest tests/test_gpu_sharded.py -q -m gpu -k "2-" --durations=8 > gpurun_out/s2_pytest_w2.log 2>&1; echo "rc=$?" >> gpurun_out/s2_pytest_w2.log
tail -14 gpurun_out/s2_pytest_w2.log
echo "t=$(( $(date +%s) - T0 ))"
