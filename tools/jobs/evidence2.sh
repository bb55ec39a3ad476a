# refresh of the per-kernel evidence after the last kernel changes (launch list, ncu --set full of one frame, isolated calls)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=${1:-r02}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 1 --warmup 3 --batch 4 --no-e2e --no-cpu-baseline --no-masters --no-strong > gpurun_out/${R}_launches_bench.log 2>&1; echo "launch list rc $?"
python tools/launch_summary.py gpurun_out/${R}_launches_bench.csv 60 > gpurun_out/${R}_launches_bench.summary.txt 2>&1
timeout 1200 ncu --set full --clock-control none -s 64 -c 72 -o /tmp/${R}_frame -f python tools/one_frame.py --frames 2 > gpurun_out/${R}_ncu_frame.log 2>&1; echo "ncu full rc $?"
python tools/ncu_summary.py /tmp/${R}_frame.ncu-rep gpurun_out/${R}_ncu_full_summary.txt gpurun_out/${R}_ncu_traffic.json "tools/one_frame.py (second frame), round 2, final code" > /dev/null 2>&1
timeout 300 python tools/kbench.py > gpurun_out/${R}_kbench.txt 2>&1
head -12 gpurun_out/${R}_launches_bench.summary.txt; grep -n "reduce_apply\|vos_std\|xtalk_tile\|sp_scan" gpurun_out/${R}_ncu_full_summary.txt; head -8 gpurun_out/${R}_kbench.txt
