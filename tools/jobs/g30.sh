cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "12 8" "16 8" "16 12" "24 12"; do
set -- $cfg
timeout 600 python bench.py --steps 3 --warmup 3 --depth $1 --ahead $2 --no-masters --no-cpu-baseline --no-strong > gpurun_out/g30_bench_d$1_a$2.json 2> gpurun_out/g30_bench.err; echo "bench depth $1 ahead $2 rc $?"; tail -2 gpurun_out/g30_bench.err
python - gpurun_out/g30_bench_d$1_a$2.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',round(d['value'],1), ' '.join('%s %.1f (d2h %.1f MB)'%(k, d[k]['value'], d[k]['d2h_bytes_per_step']/64e6) for k in ('e2e','e2e_f32_image','e2e_uncompressed')))
PY
done
nvidia-smi --query-gpu=memory.used --format=csv
