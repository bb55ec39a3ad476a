set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_steps_gpu.py tests/test_zz_rice_fz.py -m gpu -x -q > gpurun_out/g5_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g5_pytest.log
tail -12 gpurun_out/g5_pytest.log
timeout 200 python tools/xt_bench.py > gpurun_out/g5_xt.txt 2>&1; cat gpurun_out/g5_xt.txt
timeout 300 python tools/rice_bench.py > gpurun_out/g5_rice.txt 2>&1; cat gpurun_out/g5_rice.txt
