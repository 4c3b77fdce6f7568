# session 2, run 2 (1 GPU): where the end-to-end loop loses time against the device-timed loop
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python tools/host_issue.py --estimate --steps 200 2>&1 | tail -8
python tools/host_issue.py --steps 200 --log2n 10 2>&1 | tail -8
python tools/host_issue.py --estimate --steps 200 --log2n 10 2>&1 | tail -8
