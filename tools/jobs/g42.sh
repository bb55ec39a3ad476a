cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc $?"; tail -3 gpurun_out/r02_bench_n2.err | cut -c1-300
python - gpurun_out/r02_bench_n2.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',round(d['value'],1), ' '.join('%s %.1f'%(k, d[k]['value']) for k in ('e2e','e2e_f32_image','e2e_uncompressed')))
print('strong', d['strong_scaling']['value'], 'link', d['link']); 
m=d['master_sharded']; print({k:(m[k]['ms_best'], m[k]['speedup_best_vs_1gpu'], m[k]['equal_to_1gpu']) for k in m})
PY
