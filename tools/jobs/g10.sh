cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --frames 16 --graphs 1 --depth 4 --ahead 2 --out gpurun_out/trace_g1.json > gpurun_out/g10_tl.txt 2>&1; cat gpurun_out/g10_tl.txt | grep -v Warn
rm -f gpurun_out/trace_g1.json
