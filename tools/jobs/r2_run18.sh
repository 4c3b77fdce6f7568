cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resample.py tests/test_gpu_sharded.py tests/test_gpu_particle.py tests/test_gpu_lazy.py tests/test_gpu_graphs.py -x -q 2>&1 | tail -15
python tools/host_issue.py
python tools/host_issue.py --sharded
python tools/host_issue.py --estimate
python tools/host_issue.py --sharded --estimate
