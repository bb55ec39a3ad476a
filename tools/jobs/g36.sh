cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g36_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g36_pytest.log
tail -4 gpurun_out/g36_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"; tail -2 gpurun_out/r02_bench_n1.err
python - gpurun_out/r02_bench_n1.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',round(d['value'],1), ' '.join('%s %.1f (d2h %.1f MB)'%(k, d[k]['value'], d[k]['d2h_bytes_per_step']/64e6) for k in ('e2e','e2e_f32_image','e2e_uncompressed')))
print('roof', d['roofline']['frac'], d['roofline']['ms_per_launch'], 'strong', d['strong_scaling']['value'], 'launches', d['gpu_launches'], d['link'])
PY
