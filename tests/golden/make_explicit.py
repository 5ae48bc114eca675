"""Regenerates tests/golden/invpend_explicit.npz (takes ~6 minutes: thousands of small LPs).

Explicit solution of the invpend controller over the example's ParameterRange
(/root/reference/src/mpc_examples.jl:128-134) built by oracle/explicit.py, because
ParametricDAQP.jl (the reference's mp-QP solver and tree builder) cannot run here.
Run:  python tests/golden/make_explicit.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mpqp, explicit  # noqa: E402

if __name__ == "__main__":
    mpc = mpqp.invpend()
    qp = mpqp.mpc2mpqp(mpc)
    lo, hi = mpqp.invpend_range(1.0)
    sol = explicit.explicit_solution(mpc, qp, lo, hi)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "invpend_explicit.npz")
    np.savez_compressed(out, halfplanes=sol["halfplanes"], feedbacks=sol["feedbacks"],
                        hp_list=sol["hp_list"], jump_list=sol["jump_list"], lo=lo, hi=hi)
    print(out, "regions", len(sol["regions"]), "laws", len(sol["laws"]), "nodes", len(sol["hp_list"]))
