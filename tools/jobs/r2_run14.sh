cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2n_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_pytest.log
tail -12 gpurun_out/r2n_pytest.log
python bench.py --workload gsf --log2n 20 --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/r2n_gsf20.json 2> gpurun_out/r2n_gsf20.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2n_gsf20.json")); print("gsf 2^20", round(d["ms_per_step"],4), "%.3g comps/s"%d["value"], {k:v["ms"] for k,v in d["stages"].items()})
PY
python tools/closed_loop.py --log2n 20 --end-time 50 --dt-control 1.0 --out gpurun_out/r2n_closed_loop_pf_2p20_mpc_dt1.json > gpurun_out/r2n_cl1.log 2>&1; tail -1 gpurun_out/r2n_cl1.log | cut -c1-900
python tools/closed_loop.py --log2n 20 --end-time 10 --dt-control 0.1 --out gpurun_out/r2n_closed_loop_pf_2p20_mpc_dt01.json > gpurun_out/r2n_cl01.log 2>&1; tail -1 gpurun_out/r2n_cl01.log | cut -c1-900
python tools/closed_loop.py --log2n 16 --gsf --end-time 50 --dt-control 1.0 --out gpurun_out/r2n_closed_loop_gsf_2p16_mpc_dt1.json > gpurun_out/r2n_clg.log 2>&1; tail -1 gpurun_out/r2n_clg.log | cut -c1-900
