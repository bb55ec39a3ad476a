cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/e2e_prof.py --batch 32 --depth 8 > gpurun_out/g28_prof.txt 2>&1; echo "rc $?"
head -75 gpurun_out/g28_prof.txt | cut -c1-180
