cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g22_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g22_pytest.log
tail -4 gpurun_out/g22_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-masters --no-cpu-baseline --no-e2e --no-strong > gpurun_out/g22_bench.json 2> gpurun_out/g22_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g22_bench.json').read().strip().splitlines()[-1])
print('value',d['value'], 'roof', d['roofline']['frac'], d['roofline']['ms_per_launch'])
for s in d['roofline_stages']: print(s['stage'], round(s['ms_per_frame'],4), s.get('frac'))
PY
