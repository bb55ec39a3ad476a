cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_rice_fz.py -m gpu -x -q -k "fpack or run_host" > gpurun_out/g25_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g25_pytest.log
tail -5 gpurun_out/g25_pytest.log
timeout 600 python tools/rice_bench.py --reps 10 > gpurun_out/g25_rice.txt 2>&1; echo "rice rc $?"
cat gpurun_out/g25_rice.txt | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fq_|rice_' --csv --log-file gpurun_out/g25_rice_launches.csv python tools/rice_bench.py --reps 1 > gpurun_out/g25_ncu.log 2>&1
tail -4 gpurun_out/g25_rice_launches.csv | cut -c1-60,200-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-masters --no-cpu-baseline --no-strong > gpurun_out/g25_bench.json 2> gpurun_out/g25_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g25_bench.json').read().strip().splitlines()[-1])
print('value',d['value'])
for k in ('e2e','e2e_f32_image','e2e_uncompressed'):
    e=d[k]; print(k, round(e['value'],1), e['h2d_bytes_per_step']/64e6, e['d2h_bytes_per_step']/64e6)
PY
