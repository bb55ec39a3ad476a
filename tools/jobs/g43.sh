cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_zz_rice_fz.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python bench.py --steps 2 --warmup 3 --no-masters --no-strong --no-cpu-baseline > gpurun_out/g43_bench.json 2> gpurun_out/g43_bench.err; echo "bench rc $?"; tail -2 gpurun_out/g43_bench.err | cut -c1-200
python - gpurun_out/g43_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'roof',round(d['roofline']['frac'],3), d['roofline_stages_note'][:60])
PY
