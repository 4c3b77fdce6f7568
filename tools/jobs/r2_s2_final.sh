# session 2 final evidence on one B200: full GPU suite, smoke, default bench line, reference arm, ncu launch list and full
# captures (PF, GS-UKF; each after the same command exited 0 without ncu), run sequences
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_final_pytest.txt
python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2_final_smoke.txt
echo "t=$(( $(date +%s) - T0 ))"
python bench.py > gpurun_out/r2_bench_g1.json 2> gpurun_out/r2_bench_g1.err; tail -c 300 gpurun_out/r2_bench_g1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err
echo "t=$(( $(date +%s) - T0 ))"
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_l.log 2>&1
$B --no-gsf > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_pf_predict|k_pf_update|k_resample_fused' -s 8 -c 6 -o gpurun_out/r2_pf $B --no-gsf > gpurun_out/ncu_pf.log 2>&1
tail -1 gpurun_out/ncu_pf.log
G="python bench.py --workload gsf --log2n 20 --steps 4 --warmup 3 --no-cpu-baseline"
$G > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_gsf' -s 6 -c 4 -o gpurun_out/r2_gsf $G > gpurun_out/ncu_gsf.log 2>&1
tail -1 gpurun_out/ncu_gsf.log
echo "t=$(( $(date +%s) - T0 ))"
python tools/sweep.py --runs 20 --out gpurun_out/r2_sweep_pf_gsf.json > gpurun_out/sweep.log 2>&1; tail -2 gpurun_out/sweep.log
python tools/sweep.py --runs 20 --graphs --out gpurun_out/r2_sweep_pf_gsf_cuda_graphs.json > gpurun_out/sweep_g.log 2>&1; tail -2 gpurun_out/sweep_g.log
echo "t=$(( $(date +%s) - T0 ))"
python tools/e2e_stages.py 2>&1 | tail -3 | tee gpurun_out/r2_e2e_stages.txt
timeout 120 python tools/power.py --t-run 1 --pf-max 24 --gsf-max 20 --out gpurun_out/r2_power_pf_gsf.json > gpurun_out/power.log 2>&1; tail -2 gpurun_out/power.log
echo "t=$(( $(date +%s) - T0 ))"
