# session 2, run 4 (1 GPU): per-stage device and host times of the end-to-end loop
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python tools/e2e_stages.py 2>&1 | tail -4
python tools/e2e_stages.py --no-events 2>&1 | tail -3
