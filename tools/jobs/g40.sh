cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reduce_apply_strip -s 1 -c 1 -o gpurun_out/g40_apply -f python tools/one_frame.py --frames 2 > gpurun_out/g40_ncu.log 2>&1; tail -2 gpurun_out/g40_ncu.log
ls -la gpurun_out/g40_apply.ncu-rep
