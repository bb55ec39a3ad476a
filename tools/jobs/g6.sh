cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for args in "128 1056 2" "128 1056 1" "128 2048 2" "128 4096 2"; do python tools/jobs/xt_dbg.py $args 2>&1 | grep -E "^ok|Error" | head -2; done
