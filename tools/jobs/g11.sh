cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_chain_gpu.py tests/test_steps_gpu.py tests/test_reference_golden.py -m gpu -x -q > gpurun_out/g11_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g11_pytest.log
tail -25 gpurun_out/g11_pytest.log
