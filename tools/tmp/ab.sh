cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_explicit.py -m gpu -x -q 2>&1 | tail -5
run() { timeout 200 python bench.py --workload explicit_invpend_1e8 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('parity'))"; }
run main; run main
LMPC_K4_NO_RING=1 run noring
for p in 4 6 8 14 18; do LMPC_K4_PARK=$p run park$p; done
for v in r768 rids spv1 r768ids; do LMPC_GPU_LIB=$PWD/build_variants/lib$v.so run $v; done
LMPC_K4_PARK=6 LMPC_GPU_LIB=$PWD/build_variants/libr768.so run r768p6
LMPC_K4_PARK=14 LMPC_GPU_LIB=$PWD/build_variants/libr768.so run r768p14
