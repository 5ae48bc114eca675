"""compute-sanitizer target for the explicit-evaluate kernels (K4 / K4r): a batch that gives every warp several
tiles of its ring plus a ragged tail, checked against the oracle's walk.
  compute-sanitizer --tool memcheck|racecheck python tests/tools/sanitize_bst.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lmpc_b200
from oracle import mpqp, ldp
z = np.load(os.path.join(ROOT, "tests", "golden", "invpend_explicit.npz"))
ge = lmpc_b200.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"])
B = ldp.OracleBst(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clip=z["clipping"])
for N in (17, 500_017):
    Th = mpqp.sample_theta(z["lo"], z["hi"], N, seed=5)
    U, reg = ge.evaluate(Th)
    Uo, rego = B.evaluate_batch(Th)
    assert np.array_equal(U, Uo) and np.array_equal(reg, rego)
print("bst kernel kind", ge.kernel, "ok")
