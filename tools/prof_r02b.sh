# r02 (second half): full GPU tests, bench lines, launch list, ncu capture of the explicit-evaluate ring kernel
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_final_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 300 gpurun_out/r02_bench_n1.json
timeout 400 python bench.py --impl reference > gpurun_out/r02_reference_n1.json 2> gpurun_out/r02_reference_n1.err; tail -c 200 gpurun_out/r02_reference_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1; wc -l gpurun_out/r02_launches_bench.csv
timeout 300 ncu --set full --import-source on --clock-control none -k regex:bst_eval_ring -s 2 -c 1 -f -o gpurun_out/r02_k4_ring python tests/tools/prof_driver.py bst > /dev/null 2>&1
ls -la gpurun_out/r02_k4_ring.ncu-rep
