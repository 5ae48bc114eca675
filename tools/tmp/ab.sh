cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_explicit.py -m gpu -x -q 2>&1 | tail -3
LMPC_K4_INDEX64=1 timeout 300 python -m pytest tests/test_explicit.py -m gpu -x -q 2>&1 | tail -3
run() { timeout 200 python bench.py --workload explicit_invpend_1e8 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('parity'))"; }
run main; run main
LMPC_K4_INDEX64=1 run idx64
for p in 4 6 10 12; do LMPC_K4_PARK=$p run park$p; done
