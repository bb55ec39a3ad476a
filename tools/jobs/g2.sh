set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_rice_fz.py tests/test_masters.py tests/test_steps_gpu.py -m gpu -x -q > gpurun_out/g2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g2_pytest.log
tail -15 gpurun_out/g2_pytest.log
timeout 300 python tools/rice_bench.py > gpurun_out/g2_rice.txt 2>&1; cat gpurun_out/g2_rice.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-masters --no-cpu-baseline > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err; echo "bench rc $?"
tail -c 1500 gpurun_out/g2_bench.json; tail -5 gpurun_out/g2_bench.err
