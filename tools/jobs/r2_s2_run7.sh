# session 2, run 7 (1 GPU): ncu full + source capture of the MEAN variant of the fused resample (end-to-end loop)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
B="python tools/e2e_stages.py --steps 6"
$B > gpurun_out/plain7.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_resample_fused' -s 8 -c 1 -o gpurun_out/s2_rfm $B > gpurun_out/ncu_rfm.log 2>&1
tail -2 gpurun_out/ncu_rfm.log
