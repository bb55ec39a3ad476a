cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_rice_fz.py -m gpu -x -q -k "run_host" 2>&1 | tail -3
timeout 600 python tools/e2e_prof.py --batch 32 --depth 8 > gpurun_out/g29_prof.txt 2>&1; echo "rc $?"
head -16 gpurun_out/g29_prof.txt | cut -c1-180
for cfg in "8 2" "12 4"; do
set -- $cfg
timeout 600 python bench.py --steps 3 --warmup 3 --depth $1 --ahead $2 --no-masters --no-cpu-baseline --no-strong > gpurun_out/g29_bench_d$1_a$2.json 2> gpurun_out/g29_bench.err; echo "bench depth $1 ahead $2 rc $?"
python - gpurun_out/g29_bench_d$1_a$2.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',round(d['value'],1), ' '.join('%s %.1f (d2h %.1f MB)'%(k, d[k]['value'], d[k]['d2h_bytes_per_step']/64e6) for k in ('e2e','e2e_f32_image','e2e_uncompressed')))
PY
done
