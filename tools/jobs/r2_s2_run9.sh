# session 2, run 9 (1 GPU): estimate written straight to the host-mapped result block; one wait call; full GPU suite
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python tools/e2e_stages.py 2>&1 | tail -3
python tools/e2e_stages.py --no-events 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
