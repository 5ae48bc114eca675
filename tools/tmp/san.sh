cd $GRAFT_REPO_ROOT
timeout 1000 compute-sanitizer --tool memcheck python tests/tools/sanitize_driver.py > gpurun_out/r02_memcheck_all.log 2>&1; tail -4 gpurun_out/r02_memcheck_all.log
timeout 1300 compute-sanitizer --tool racecheck python tests/tools/sanitize_driver.py > gpurun_out/r02_racecheck_all.log 2>&1; tail -4 gpurun_out/r02_racecheck_all.log
