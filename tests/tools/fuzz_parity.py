"""Randomised parity sweep, larger than the unit tests: every controller x scale x seed, cold and
warm-started batch and closed loops, GPU (through the C ABI) against the CPU oracle, bit for bit.
python tests/tools/fuzz_parity.py [budget_seconds [size_multiplier]]   -> one line per case, exit code 1 on any mismatch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import lmpc_b200, util
from oracle import mpqp

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
mult = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
t_start = time.time()
bad = 0
SIZES = {"invpend": 400_000, "dint_hard": 300_000, "dint_eq": 200_000, "dint_eq_small": 300_000, "genparam": 200_000,
         "simple": 100_000, "simple_q1000": 50_000, "aircraft": 60_000, "aircraft_ksall": 15_000, "chain_small": 20_000,
         "chain": 400}
for name, N in SIZES.items():
    N = int(N * mult)
    mpc = util.controller(name); qp = mpqp.mpc2mpqp(mpc)
    kernels = [0] + ([2, 3] if name in ("invpend", "dint_hard", "dint_eq_small") else []) + ([3] if name in ("aircraft", "dint_eq", "simple") else [])
    for kern in kernels:
        try:
            g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight,
                                    settings=lmpc_b200.default_settings(kernel=kern))
        except lmpc_b200.LmpcError:
            continue
        cst = g.ldp_constants(); cst["senses"] = g.senses
        P = util.oracle_from_constants(qp, mpc.nu, cst, soft_weight=mpc.soft_weight)
        for scale in (1.0, 0.5, 0.1):
            for seed in (101, 202, 303):
                if time.time() - t_start > budget:
                    print("budget reached"); sys.exit(1 if bad else 0)
                n = N if kern == 0 else max(N // 4, 200)
                lo, hi = util.theta_range(name, scale)
                Th = mpqp.sample_theta(lo, hi, n, seed=seed)
                ref = P.solve_batch(Th); out = g.solve(Th)
                ok = ref["flag"] >= 1
                same = (np.array_equal(out["exitflag"], ref["flag"]) and np.array_equal(out["iters"], ref["iters"]) and
                        np.array_equal(out["active_upper"][ok], ref["act_u"][ok]) and
                        np.array_equal(out["active_lower"][ok], ref["act_l"][ok]) and
                        np.array_equal(np.asarray(out["U"])[:, ok], ref["U"][:, ok]))
                # warm start from a perturbed neighbour's working sets
                Th2 = Th + 0.01 * (hi - lo)[:, None] * np.random.default_rng(seed).standard_normal(Th.shape)
                rw = P.solve_batch_warm(Th2, ref["act_u"], ref["act_l"]); ow = g.solve_warm(Th2, ref["act_u"], ref["act_l"])
                okw = rw["flag"] >= 1
                samew = (np.array_equal(ow["exitflag"], rw["flag"]) and np.array_equal(ow["iters"], rw["iters"]) and
                         np.array_equal(np.asarray(ow["U"])[:, okw], rw["U"][:, okw]))
                bad += (not same) + (not samew)
                print(f"{name:15s} kernel {g.kernel} s={scale:4} seed={seed} N={n:7d} cold {'ok' if same else 'MISMATCH'} "
                      f"warm {'ok' if samew else 'MISMATCH'} flags {dict(zip(*np.unique(ref['flag'], return_counts=True)))}", flush=True)
        # closed loop, warm and cold
        nx, nr, nd, nup = mpc.parameter_dims()
        if nx > 0 and name not in ("genparam", "chain"):
            lo, hi = util.theta_range(name, 0.3)
            ntr = max(min(N // 40, 20000), 100)
            X0 = mpqp.sample_theta(lo[:nx], hi[:nx], ntr, seed=7)
            for warm in (True, False):
                o = g.simulate(mpc.F, mpc.G, X0, r=np.zeros(nr) if nr else None, d=np.zeros(nd) if nd else None, N=25,
                               Gd=mpc.Gd if nd else None, warm_start=warm)
                r = P.simulate_batch(mpc.F, mpc.G, mpc.Gd if nd else None, X0, np.zeros((nr, ntr)) if nr else None,
                                     np.zeros((nd, ntr)) if nd else None, None, 25, nr, nd, nup, warm=warm)
                same = np.array_equal(o["Xend"], r["Xend"]) and np.array_equal(o["exitflag"], r["flag"]) and np.array_equal(o["iters"], r["iters"])
                bad += not same
                print(f"{name:15s} kernel {g.kernel} closed loop warm={warm} N={ntr} {'ok' if same else 'MISMATCH'}", flush=True)
        g.close()
print("mismatches:", bad)
sys.exit(1 if bad else 0)
