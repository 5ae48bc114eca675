"""Regenerates tests/golden/*.npz: frozen oracle outputs on seeded theta samples.

The reference itself (Julia + libdaqp) cannot run in this container, so these are NOT outputs
of the reference; they are outputs of the oracle AFTER it was pinned on the reference's known
answers (tests/test_oracle_golden.py).  They freeze the oracle so a later change to it, or to
the mpQP restatement, shows up as a diff.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import mpqp, ldp  # noqa: E402
import util  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("invpend", 1.0, 512), ("invpend", 0.25, 256), ("aircraft", 0.25, 256),
         ("simple", 1.0, 256), ("chain_small", 1.0, 64)]

if __name__ == "__main__":
    for name, scale, N in CASES:
        mpc = util.controller(name)
        qp = mpqp.mpc2mpqp(mpc)
        L = mpqp.qp2ldp(qp, mpc.nu, soft_weight=mpc.soft_weight)
        P = ldp.OracleProblem(L, mpc.nu)
        lo, hi = util.theta_range(name, scale)
        Th = mpqp.sample_theta(lo, hi, N)
        out = P.solve_batch(Th, want_flops=True)
        fn = os.path.join(HERE, f"{name}_s{scale}.npz")
        np.savez_compressed(fn, Theta=Th, U=out["U"], flag=out["flag"], iters=out["iters"],
                            act_u=out["act_u"], act_l=out["act_l"], H=qp.H, f_theta=qp.f_theta,
                            A=qp.A, bu=qp.bu, bl=qp.bl, W=qp.W, senses=qp.senses)
        print(fn, "iters mean %.2f max %d" % (out["iters"].mean(), out["iters"].max()),
              dict(zip(*np.unique(out["flag"], return_counts=True))))
