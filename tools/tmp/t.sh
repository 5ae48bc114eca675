cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_explicit.py -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --workload explicit_invpend_1e8 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('parity'))"
