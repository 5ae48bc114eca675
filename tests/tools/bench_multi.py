"""Single-process multi-GPU host-buffer throughput (lmpc_gpu_solve_batch_multi).

python tests/tools/bench_multi.py [N_per_gpu]   -> one JSON line per device count 1..visible, with three
kinds of caller memory: fresh pageable arrays every call (what a naive caller does; includes the page
faults of the new output arrays), reused pageable arrays, and pinned arrays (lmpc_gpu_host_alloc)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lmpc_b200 as lmpc  # noqa: E402
from bench import product_model, theta_range, SEED  # noqa: E402
from oracle import mpqp  # noqa: E402  (theta sampler only)


def best_of(fn, reps=4):
    best = 1e9
    for it in range(reps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if it:
            best = min(best, dt)
    return best


def main():
    per = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    nd = lmpc.device_count()
    lo, hi = theta_range("invpend", 1.0)
    L = lmpc.lib()
    for G in [g for g in (1, 2, 4, 8) if g <= nd]:
        mg = product_model("invpend").setup(devices=list(range(G)))
        N = per * G
        Th = np.asfortranarray(mpqp.sample_theta(lo, hi, N, seed=SEED))
        res = {}
        res["pageable_fresh"] = best_of(lambda: mg.solve(Th, want_iters=False, want_active=False))
        U = np.zeros((mg.nu, N), order="F"); fl = np.ones(N, np.int32)
        call = lambda th, u, f: L.lmpc_gpu_solve_batch_multi(mg._mh, th.ctypes.data, N, u.ctypes.data, f.ctypes.data,
                                                             None, None, None)
        res["pageable_reused"] = best_of(lambda: call(Th, U, fl))
        Tp = lmpc.pinned_empty(Th.shape, order="F"); Tp[...] = Th
        Up = lmpc.pinned_empty((mg.nu, N), order="F"); fp = lmpc.pinned_empty(N, np.int32)
        res["pinned"] = best_of(lambda: call(Tp, Up, fp))
        assert np.array_equal(Up, U) and np.array_equal(fp, fl) and np.all(fl >= 1)
        print(json.dumps(dict(tool="bench_multi", gpus=G, N=N,
                              solves_per_s={k: N / v for k, v in res.items()})), flush=True)
        mg.close()


if __name__ == "__main__":
    main()
