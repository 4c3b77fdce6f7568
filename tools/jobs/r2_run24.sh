cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gsf --sharded > gpurun_out/r2s_bench_sh.json 2> gpurun_out/r2s_bench_sh.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2s_bench_g2.json 2> gpurun_out/r2s_bench_g2.err
python - <<'PY'
import json
for f in ("r2s_bench_sh", "r2s_bench_g2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"],4), "parity", d["sharded_parity"], "%.4g"%d["value"], "%.4g"%d["e2e"]["value"])
    except Exception as e:
        print("failed", f, e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
