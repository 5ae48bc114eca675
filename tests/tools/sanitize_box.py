"""K1b / K1c only for compute-sanitizer: the generic box kernel (both instantiations: with and without the
transposed constant copy), then the add-only fast kernel + REDO launch on the same batches
(LMPC_BOX_FAST_MIN_BATCH = 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import lmpc_b200, util
from oracle import mpqp
for name, N in (("invpend", 5000), ("invpend_rpreview", 2000), ("fast:invpend", 5000), ("fast:invpend_nc8", 3000),
                ("fast:aircraft_box2", 3000)):
    if name.startswith("fast:"):
        os.environ["LMPC_BOX_FAST_MIN_BATCH"] = "0"
        name = name[5:]
    mpc = util.controller(name); qp = mpqp.mpc2mpqp(mpc)
    g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight)
    Th = np.asfortranarray(mpqp.sample_theta(*util.theta_range(name, 1.0), N))
    out = g.solve(Th)
    print(name, "kernel", g.kernel, dict(zip(*np.unique(out["exitflag"], return_counts=True))), flush=True)
