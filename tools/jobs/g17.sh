cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/g17_bench_n8.json 2> gpurun_out/g17_bench_n8.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g17_bench_n8.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'plain',d['e2e_uncompressed']['value'],'strong',d['strong_scaling'])
for k,v in d['master_sharded'].items():
    print(k, 'ms',v['ms'],'1gpu',v['ms_1gpu'],'equal',v['equal_to_1gpu'],'speedup',v['speedup_vs_1gpu'],'best',v['ms_best'],v['speedup_best_vs_1gpu'], v['fused_allgather'])
PY
tail -3 gpurun_out/g17_bench_n8.err
