cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_steps_gpu.py -m gpu -x -q -k "xtalk or stack or master" > gpurun_out/g16_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g16_pytest.log
tail -3 gpurun_out/g16_pytest.log
timeout 200 python tools/xt_bench.py 2>&1 | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-strong > gpurun_out/g16_bench_n2.json 2> gpurun_out/g16_bench_n2.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g16_bench_n2.json').read().strip().splitlines()[-1])
print('value',d['value'])
for s in d['roofline_stages']: print(s['stage'], round(s['ms_per_frame'],4), s.get('frac'))
print(json.dumps(d['master_sharded'],indent=1))
PY
tail -5 gpurun_out/g16_bench_n2.err
