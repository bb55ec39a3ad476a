cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_chain_gpu.py -m gpu -x -q > gpurun_out/g15_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g15_pytest.log
tail -3 gpurun_out/g15_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-masters --no-cpu-baseline --no-e2e --no-strong > gpurun_out/g15_bench.json 2> gpurun_out/g15_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g15_bench.json').read().strip().splitlines()[-1])
print('value',d['value'], 'roof', d['roofline']['frac'], d['roofline']['ms_per_launch'])
for s in d['roofline_stages']: print(s['stage'], round(s['ms_per_frame'],4), s.get('frac'))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reduce_apply_scan -s 1 -c 1 -o gpurun_out/g15_fuse -f python tools/one_frame.py --frames 2 > gpurun_out/g15_ncu.log 2>&1; tail -3 gpurun_out/g15_ncu.log
ls -la gpurun_out/g15_fuse.ncu-rep
