cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2k_bench.json"))
    print("bench", round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, round(d["e2e"]["ms_per_step"],4))
    g=d["gsf"]; print("gsf 2p16 graphs", g["2p16_cuda_graphs"]["ms_per_step"], "%.3g"%g["2p16_cuda_graphs"]["value"], "eager", g["2p16_eager"]["ms_per_step"], g["2p16_eager"].get("stages"), "2p20", g["2p20"]["ms_per_step"], "%.3g"%g["2p20"]["value"], g["2p20"].get("stages"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2k_bench.err").read()[-2000:])
PY
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gsf --sharded > gpurun_out/plain_sh.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_resample_fused|k_pf_predict' -s 20 -c 4 -o gpurun_out/prof_r2k_sh python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gsf --sharded > gpurun_out/ncu_sh.log 2>&1
tail -2 gpurun_out/ncu_sh.log
python bench.py --workload gsf --log2n 20 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain_gsf.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_gsf' -s 6 -c 4 -o gpurun_out/prof_r2k_gsf python bench.py --workload gsf --log2n 20 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gsf.log 2>&1
tail -2 gpurun_out/ncu_gsf.log
