set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/g1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g1_pytest.log
tail -5 gpurun_out/g1_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/g1_bench.json 2> gpurun_out/g1_bench.err; echo "bench rc $?"
tail -c 3000 gpurun_out/g1_bench.json
timeout 300 python tools/kbench.py > gpurun_out/g1_kbench.txt 2>&1
