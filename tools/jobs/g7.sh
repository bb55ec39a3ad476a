cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for args in "128 1056 2" "128 1056 1" "128 1056 0"; do python tools/jobs/xt_dbg.py $args 2>&1 | grep -E "^ok|Error" | head -2; done
timeout 900 python -m pytest tests/test_steps_gpu.py tests/test_zz_rice_fz.py -m gpu -x -q > gpurun_out/g7_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g7_pytest.log
tail -5 gpurun_out/g7_pytest.log
timeout 200 python tools/xt_bench.py > gpurun_out/g7_xt.txt 2>&1; cat gpurun_out/g7_xt.txt
timeout 300 python tools/rice_bench.py > gpurun_out/g7_rice.txt 2>&1; cat gpurun_out/g7_rice.txt
