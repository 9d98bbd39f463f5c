cd $GRAFT_REPO_ROOT
timeout 300 python scripts/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1 && XQ_ACT_LANE=1 timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/sanitize_memcheck.log
