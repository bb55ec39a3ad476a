cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_rice_fz.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/rice_bench.py --reps 10 > gpurun_out/g33_rice.txt 2>&1; echo "rice rc $?"
cat gpurun_out/g33_rice.txt | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fq_row_stats_kernel|QuantSrc' -c 2 -o gpurun_out/g33_fpack -f python tools/rice_bench.py --reps 1 > gpurun_out/g33_ncu.log 2>&1; tail -2 gpurun_out/g33_ncu.log
ls -la gpurun_out/g33_fpack.ncu-rep
