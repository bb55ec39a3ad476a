cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_rice_fz.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/rice_bench.py --reps 10 > gpurun_out/g34_rice.txt 2>&1; echo "rice rc $?"
cat gpurun_out/g34_rice.txt | tail -5
timeout 900 python bench.py --no-masters --no-strong > gpurun_out/g34_bench.json 2> gpurun_out/g34_bench.err; echo "bench rc $?"; tail -3 gpurun_out/g34_bench.err
python - gpurun_out/g34_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',round(d['value'],1), ' '.join('%s %.1f (d2h %.1f MB)'%(k, d[k]['value'], d[k]['d2h_bytes_per_step']/64e6) for k in ('e2e','e2e_f32_image','e2e_uncompressed')))
print(json.dumps(d['roofline'])[:1500]); print(d['gpu_launches'], d['cpu_baseline'])
PY
