# session 2, run 3 (1 GPU): ncu source-level capture of the fused resample and update kernels
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gsf"
$B > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_pf_update|k_resample_fused' -s 8 -c 2 -o gpurun_out/s2_rf $B > gpurun_out/ncu_rf.log 2>&1
tail -2 gpurun_out/ncu_rf.log; ls -la gpurun_out/*.ncu-rep
