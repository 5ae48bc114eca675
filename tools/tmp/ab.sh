cd $GRAFT_REPO_ROOT
for v in "" bcu16 bcu32 il1 il1bcu16; do
  if [ -n "$v" ]; then export LMPC_GPU_LIB=$PWD/build_variants/lib$v.so; fi
  echo "== $v"; timeout 300 python tests/tools/bench_configs.py chain 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d[k] for k in d if k in ('config','theta_scale','N','solves_per_s','gpu_solves_per_s','cold_solves_per_s','mean_iters','bit_identical','flags','ms')} or d)
"
done
