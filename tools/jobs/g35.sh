cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 560 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench n8 rc $?"; tail -3 gpurun_out/r02_bench_n8.err
python - gpurun_out/r02_bench_n8.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',round(d['value'],1), ' '.join('%s %.1f (d2h %.1f MB)'%(k, d[k]['value'], d[k]['d2h_bytes_per_step']/64e6/8) for k in ('e2e','e2e_f32_image','e2e_uncompressed')))
print('strong', d['strong_scaling']); print(json.dumps(d['master_sharded'])[:2500]); print(json.dumps(d['roofline'])[:400])
PY
