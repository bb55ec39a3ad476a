cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_steps_gpu.py tests/test_masters.py tests/test_fullsize_gpu.py -m gpu -x -q -k "stack or master or division or flat" > gpurun_out/g18_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g18_pytest.log
tail -5 gpurun_out/g18_pytest.log
timeout 300 python tools/kbench.py 2>&1 | grep -E "stack|xtalk"
