# session 2, run 10 (1 GPU): ncu source-level capture of the GS-UKF kernels at 2^20 components
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
G="python bench.py --workload gsf --log2n 20 --steps 4 --warmup 3 --no-cpu-baseline"
$G > gpurun_out/plain10.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_gsf' -s 6 -c 2 -o gpurun_out/s2_gsf $G > gpurun_out/ncu_gsf.log 2>&1
tail -1 gpurun_out/ncu_gsf.log
