# session 2, run 8 (1 GPU): software-pipelined mean reads in the fused resample (MEAN kernels): tests + e2e stage times
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_resample.py tests/test_gpu_sharded.py -q -m gpu -k "estimate" 2>&1 | tail -3
python tools/e2e_stages.py 2>&1 | tail -3
python tools/e2e_stages.py --no-events 2>&1 | tail -2
