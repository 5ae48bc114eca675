"""Small batches through every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import lmpc_b200, util
from oracle import mpqp
for name, N in (("invpend", 3000), ("dint_hard", 3000), ("aircraft", 600), ("aircraft_ksall", 200), ("chain", 8), ("simple", 600), ("dint_eq", 600), ("dint_eq_small", 600)):
    mpc = util.controller(name); qp = mpqp.mpc2mpqp(mpc)
    g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight)
    Th = mpqp.sample_theta(*util.theta_range(name, 1.0), N)
    out = g.solve(Th)
    print(name, "kernel", g.kernel, dict(zip(*np.unique(out["exitflag"], return_counts=True))), flush=True)
    if name in ("invpend", "aircraft"):
        nx, nr, nd, nup = mpc.parameter_dims()
        r = g.simulate(mpc.F, mpc.G, Th[:nx, :200], r=np.zeros(nr), d=np.zeros(nd) if nd else None, N=10,
                       Gd=mpc.Gd if nd else None, warm_start=True)
        print(" sim", dict(zip(*np.unique(r["exitflag"], return_counts=True))), flush=True)
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "invpend_explicit.npz"))
ge = lmpc_b200.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=np.array([[-2.0, 2.0]]))
U, reg = ge.evaluate(mpqp.sample_theta(z["lo"], z["hi"], 5000))
print("bst", U.shape, reg.max())
