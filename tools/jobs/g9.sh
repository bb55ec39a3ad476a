cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/g9_bench_n2.json 2> gpurun_out/g9_bench_n2.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g9_bench_n2.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e'],'\nplain',d['e2e_uncompressed'],'\nstrong',d['strong_scaling'],'\nmasters',json.dumps(d['master_sharded'],indent=1))
PY
tail -5 gpurun_out/g9_bench_n2.err
