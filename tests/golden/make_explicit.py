"""Regenerates tests/golden/invpend_explicit.npz (about two minutes: a few thousand small LPs).

Explicit solution of the invpend controller over the example's ParameterRange
(/root/reference/src/mpc_examples.jl:128-134): the search tree the GPU evaluate path walks in the tests and in
bench.py's explicit workload.  ParametricDAQP.jl (the reference's mp-QP solver and tree builder) cannot run
here, so the tree comes from the PRODUCT's builder (linearmpc.jl_b200/explicit.py: exact region polyhedra and
laws, LP-certified tree, clipping to the input bounds as the reference's build_tree!(...; clipping=true) does,
src/explicit.jl:65).  On a GPU box the builder discovers the critical regions with the product's own batched
solver; this container has no GPU, so the CPU oracle stands in for that one step (region DISCOVERY only).
Run:  python tests/golden/make_explicit.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lmpc_b200  # noqa: E402
from oracle import mpqp, ldp  # noqa: E402


class OracleSolver:
    def __init__(self, mpc):
        self.P = ldp.OracleProblem(mpqp.qp2ldp(mpqp.mpc2mpqp(mpc), mpc.nu, soft_weight=mpc.soft_weight), mpc.nu)

    def solve(self, Th, want_iters=False, want_active=True):
        r = self.P.solve_batch(Th)
        return dict(U=r["U"], exitflag=r["flag"], active_upper=r["act_u"], active_lower=r["act_l"])


if __name__ == "__main__":
    ex = importlib.import_module("lmpc_b200.explicit")
    m, rng, _ = lmpc_b200.mpc_examples("invpend")
    lo, hi = np.array(rng[0], float), np.array(rng[1], float)
    sol = ex.build_explicit(m.condense(), OracleSolver(mpqp.invpend()), lo, hi, 1, n_samples=400000,
                            clipping=m.clipping_bounds())
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "invpend_explicit.npz")
    np.savez_compressed(out, halfplanes=sol["halfplanes"], feedbacks=sol["feedbacks"], hp_list=sol["hp_list"],
                        jump_list=sol["jump_list"], lo=lo, hi=hi, clipping=sol["clipping"])
    print(out, "regions", len(sol["regions"]), "laws", len(sol["laws"]), "nodes", len(sol["hp_list"]),
          "depth, leaves", sol["depth"])
