cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_explicit.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -5
timeout 300 compute-sanitizer --tool memcheck python tests/tools/sanitize_bst.py > gpurun_out/r02_memcheck_bst.log 2>&1; tail -4 gpurun_out/r02_memcheck_bst.log
timeout 600 compute-sanitizer --tool racecheck python tests/tools/sanitize_bst.py > gpurun_out/r02_racecheck_bst.log 2>&1; tail -4 gpurun_out/r02_racecheck_bst.log
