cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rice_decode_kernel -s 1 -c 1 -o gpurun_out/g20_rice -f python tools/rice_bench.py --reps 2 > gpurun_out/g20_ncu.log 2>&1; tail -3 gpurun_out/g20_ncu.log
ls -la gpurun_out/g20_rice.ncu-rep
