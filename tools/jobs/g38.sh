cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "0,0" "7,5" "6,5" "6,4" "5,4" "4,3"; do
timeout 600 python bench.py --steps 5 --warmup 3 --resident $cfg --no-masters --no-cpu-baseline --no-strong --no-e2e > gpurun_out/g38_bench.json 2> gpurun_out/g38_bench.err; echo "bench resident $cfg rc $?"
python - gpurun_out/g38_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',round(d['value'],1), 'apply alone', round(d['roofline']['ms_per_launch'],4), round(d['roofline']['frac'],3))
PY
done
