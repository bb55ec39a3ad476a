cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-masters --no-cpu-baseline --no-strong > gpurun_out/g24_bench.json 2> gpurun_out/g24_bench.err; echo "bench rc $?"
tail -3 gpurun_out/g24_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g24_bench.json').read().strip().splitlines()[-1])
print('value',d['value'])
for k in ('e2e','e2e_f32_image','e2e_uncompressed'):
    e=d[k]; print(k, round(e['value'],1), e['h2d_bytes_per_step']/64e6, e['d2h_bytes_per_step']/64e6)
print(d['link'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fq_|rice_' --csv --log-file gpurun_out/g24_rice_launches.csv python tools/rice_bench.py --reps 1 > gpurun_out/g24_ncu.log 2>&1
tail -12 gpurun_out/g24_rice_launches.csv | cut -c1-250
