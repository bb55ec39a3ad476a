cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_steps_gpu.py -m gpu -x -q -k xtalk > gpurun_out/g8_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g8_pytest.log
tail -4 gpurun_out/g8_pytest.log
timeout 200 python tools/xt_bench.py > gpurun_out/g8_xt.txt 2>&1; cat gpurun_out/g8_xt.txt
