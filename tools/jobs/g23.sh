cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_rice_fz.py -m gpu -x -q -k "fpack or run_host" > gpurun_out/g23_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g23_pytest.log
tail -30 gpurun_out/g23_pytest.log
timeout 600 python tools/rice_bench.py --reps 10 > gpurun_out/g23_rice.txt 2>&1; echo "rice rc $?"
cat gpurun_out/g23_rice.txt | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fq_|rice_' --csv --log-file gpurun_out/g23_rice_launches.csv python tools/rice_bench.py --reps 1 > gpurun_out/g23_ncu.log 2>&1
tail -12 gpurun_out/g23_rice_launches.csv | cut -c1-250
