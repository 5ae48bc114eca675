set -x
cd $GRAFT_REPO_ROOT
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_final_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 200 gpurun_out/r02_bench_n1.json
timeout 400 python bench.py --impl reference > gpurun_out/r02_reference_n1.json 2> gpurun_out/r02_reference_n1.err; tail -c 200 gpurun_out/r02_reference_n1.json
timeout 120 python tests/tools/invpend_on_k2.py > gpurun_out/r02_invpend_on_k2.json 2>&1; cat gpurun_out/r02_invpend_on_k2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1; wc -l gpurun_out/r02_launches_bench.csv
timeout 300 ncu --set full --import-source on --clock-control none -k regex:ldp_box -s 4 -c 2 -f -o gpurun_out/r02_k1c_final python tests/tools/prof_driver.py box > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:ldp_warp_sim -s 1 -c 1 -f -o gpurun_out/r02_k2sim_final python tests/tools/prof_driver.py sim > /dev/null 2>&1
timeout 400 compute-sanitizer --tool memcheck python tests/tools/sanitize_box.py > gpurun_out/r02_memcheck_box.log 2>&1; tail -3 gpurun_out/r02_memcheck_box.log
timeout 600 compute-sanitizer --tool racecheck python tests/tools/sanitize_box.py > gpurun_out/r02_racecheck_box.log 2>&1; tail -3 gpurun_out/r02_racecheck_box.log
ls -la gpurun_out/*final*
