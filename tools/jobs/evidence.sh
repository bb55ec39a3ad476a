# round-2 evidence run (one B200): bench lines, launch list, ncu --set full of one frame, isolated kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
R=${1:-r02}
timeout 900 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err; echo "ref rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 1 --warmup 3 --batch 4 --no-e2e --no-cpu-baseline --no-masters --no-strong > gpurun_out/${R}_launches_bench.log 2>&1; echo "launch list rc $?"
python tools/launch_summary.py gpurun_out/${R}_launches_bench.csv 60 > gpurun_out/${R}_launches_bench.summary.txt 2>&1
# one frame under ncu --set full: skip the first frame's launches (warm-up), take the second frame's
timeout 1200 ncu --set full --clock-control none -s 64 -c 72 -o /tmp/${R}_frame -f python tools/one_frame.py --frames 2 > gpurun_out/${R}_ncu_frame.log 2>&1; echo "ncu full rc $?"
python tools/ncu_summary.py /tmp/${R}_frame.ncu-rep gpurun_out/${R}_ncu_full_summary.txt gpurun_out/${R}_ncu_traffic.json "tools/one_frame.py (second frame), round 2" > /dev/null 2>&1
# the codec kernels (decode, encode x3, row statistics, scan, compact) of tools/rice_bench.py
timeout 900 ncu --set full --clock-control none -k regex:'rice_|fq_' -c 24 -o /tmp/${R}_codec -f python tools/rice_bench.py --reps 1 > gpurun_out/${R}_ncu_codec.log 2>&1; echo "ncu codec rc $?"
python tools/ncu_summary.py /tmp/${R}_codec.ncu-rep gpurun_out/${R}_ncu_codec_summary.txt /tmp/${R}_codec_traffic.json "tools/rice_bench.py --reps 1 (Rice codec + fpack -q 16 kernels), round 2" > /dev/null 2>&1
timeout 300 python tools/kbench.py > gpurun_out/${R}_kbench.txt 2>&1
timeout 300 python tools/rice_bench.py > gpurun_out/${R}_rice_bench.txt 2>&1
timeout 200 python tools/xt_bench.py > gpurun_out/${R}_xt_bench.txt 2>&1
tail -c 600 gpurun_out/${R}_bench_n1.json; echo; head -30 gpurun_out/${R}_launches_bench.summary.txt; head -40 gpurun_out/${R}_ncu_full_summary.txt; cat gpurun_out/${R}_ncu_codec_summary.txt
