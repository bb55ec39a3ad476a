cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_gpu.py tests/test_reference_golden.py tests/test_overscan_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-masters --no-cpu-baseline --no-strong --no-e2e > gpurun_out/g41_bench.json 2> gpurun_out/g41_bench.err; echo "bench rc $?"
python - gpurun_out/g41_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',round(d['value'],1), 'apply alone', round(d['roofline']['ms_per_launch'],4), round(d['roofline']['ms_min'],4), round(d['roofline']['frac'],3))
PY
