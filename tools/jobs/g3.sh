set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_rice_fz.py tests/test_masters.py -m gpu -x -q > gpurun_out/g3_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g3_pytest.log
tail -8 gpurun_out/g3_pytest.log
timeout 300 python tools/rice_bench.py > gpurun_out/g3_rice.txt 2>&1; cat gpurun_out/g3_rice.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-masters --no-cpu-baseline --no-strong > gpurun_out/g3_bench.json 2> gpurun_out/g3_bench.err; echo "bench rc $?"
python -c "
import json
d=json.loads(open('gpurun_out/g3_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'plain',d['e2e_uncompressed']['value'],d['link'])
for s in d['roofline_stages']: print(s['stage'], round(s['ms_per_frame'],4))
"
tail -5 gpurun_out/g3_bench.err
