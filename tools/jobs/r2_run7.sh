cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
GSE_FUSED_TRACE=1 python tools/fused_trace.py 24 2>&1 | grep "ESS" > gpurun_out/r2g_ess.txt; cat gpurun_out/r2g_ess.txt
