cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
GSE_FUSED_TRACE=1 python tools/fused_trace.py 24 2>&1 | grep "^step 1[01]:" 
GSE_FUSED_TRACE=1 python tools/fused_trace.py 24 --sharded 2>&1 | grep "^step 1[01]:\|Error\|error" 
