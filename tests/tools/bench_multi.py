"""Single-process multi-GPU host-buffer throughput (lmpc_gpu_solve_batch_multi).

python tests/tools/bench_multi.py [N_per_gpu]     -> one JSON line per device count 1..visible."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lmpc_b200 as lmpc  # noqa: E402
from bench import product_model, theta_range, SEED  # noqa: E402
from oracle import mpqp  # noqa: E402  (theta sampler only)


def main():
    per = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    nd = lmpc.device_count()
    lo, hi = theta_range("invpend", 1.0)
    for G in [g for g in (1, 2, 4, 8) if g <= nd]:
        mg = product_model("invpend").setup(devices=list(range(G)))
        N = per * G
        Th = np.asfortranarray(mpqp.sample_theta(lo, hi, N, seed=SEED))
        best = 1e9
        for it in range(4):
            t0 = time.perf_counter()
            out = mg.solve(Th, want_iters=False, want_active=False)
            dt = time.perf_counter() - t0
            if it:
                best = min(best, dt)
        ok = int((out["exitflag"] >= 1).sum())
        print(json.dumps(dict(tool="bench_multi", gpus=G, N=N, solves_per_s=N / best, s=best, solved=ok)), flush=True)
        mg.close()


if __name__ == "__main__":
    main()
