"""K1b only (both instantiations: with and without the transposed constant copy) for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import lmpc_b200, util
from oracle import mpqp
for name, N in (("invpend", 5000), ("invpend_rpreview", 2000)):
    mpc = util.controller(name); qp = mpqp.mpc2mpqp(mpc)
    g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight)
    Th = np.asfortranarray(mpqp.sample_theta(*util.theta_range(name, 1.0), N))
    out = g.solve(Th)
    print(name, "kernel", g.kernel, dict(zip(*np.unique(out["exitflag"], return_counts=True))), flush=True)
