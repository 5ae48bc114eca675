cd $GRAFT_REPO_ROOT
timeout 300 ncu --set full --import-source on --clock-control none -k regex:ldp_warp_sim -s 2 -c 1 -f -o gpurun_out/r02_k2sim_fast python tests/tools/prof_driver.py sim > /dev/null 2>&1
ls -la gpurun_out/r02_k2sim_fast.ncu-rep
