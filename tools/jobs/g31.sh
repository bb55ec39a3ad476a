cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/timeline.py --host 1 --frames 36 --depth 12 --ahead 4 --graphs 1 --out /tmp/trace_host.json > gpurun_out/g31_tl_host.txt 2>&1; echo "rc $?"
cat gpurun_out/g31_tl_host.txt | cut -c1-150
timeout 600 python tools/timeline.py --frames 36 --depth 12 --ahead 4 --graphs 1 --out /tmp/trace_dev.json > gpurun_out/g31_tl_dev.txt 2>&1; echo "rc $?"
cat gpurun_out/g31_tl_dev.txt | cut -c1-150
