cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_pytest.log
tail -12 gpurun_out/r2m_pytest.log
for mb in 3 4 5 6; do
GSE_GSF_MINB=$mb python bench.py --workload gsf --log2n 20 --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/r2m_gsf20_$mb.json 2> gpurun_out/r2m_gsf20_$mb.err
GSE_GSF_MINB=$mb python bench.py --workload gsf --log2n 16 --steps 300 --warmup 20 --no-cpu-baseline --graphs > gpurun_out/r2m_gsf16_$mb.json 2> gpurun_out/r2m_gsf16_$mb.err
done
python - <<'PY'
import json
for mb in (3,4,5,6):
    for n in (20,16):
        try:
            d=json.load(open("gpurun_out/r2m_gsf%d_%d.json"%(n,mb)))
            print("minb",mb,"2^%d"%n, round(d["ms_per_step"],4), "%.3g comps/s"%d["value"], {k:v["ms"] for k,v in d["stages"].items()})
        except Exception as e:
            print(mb, n, "failed", e)
PY
