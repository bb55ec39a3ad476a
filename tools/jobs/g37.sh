cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_rice_fz.py tests/test_chain_gpu.py -m gpu -x -q -k "reduce_night or run_host" > gpurun_out/g37_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g37_pytest.log
tail -25 gpurun_out/g37_pytest.log
