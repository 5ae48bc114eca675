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
    if name in ("invpend", "aircraft", "aircraft_ksall", "dint_eq"):
        w = g.solve_warm(Th, out["active_upper"], out["active_lower"])
        print(" warm", dict(zip(*np.unique(w["exitflag"], return_counts=True))), float(w["iters"].mean()), flush=True)
    if name in ("invpend", "aircraft", "aircraft_ksall"):
        nx, nr, nd, nup = mpc.parameter_dims()
        r = g.simulate(mpc.F, mpc.G, Th[:nx, :200], r=np.zeros(nr), d=np.zeros(nd) if nd else None, N=10,
                       Gd=mpc.Gd if nd else None, warm_start=True)
        print(" sim", dict(zip(*np.unique(r["exitflag"], return_counts=True))), flush=True)
        Cm = mpc.C
        r2 = g.simulate(mpc.F, mpc.G, Th[:nx, :200], r=np.zeros(nr), d=np.zeros(nd) if nd else None, N=10,
                        Gd=mpc.Gd if nd else None, warm_start=True, plant=dict(F=mpc.F * 1.001, G=mpc.G, Gd=mpc.Gd if nd else None),
                        observer=dict(C=Cm, K=0.3 * np.linalg.pinv(Cm), Dd=mpc.Dd if nd else None),
                        noise=1e-3 * np.ones((Cm.shape[0], 4)), r_traj=np.zeros((nr, 3)))
        print(" sim ext", dict(zip(*np.unique(r2["exitflag"], return_counts=True))), flush=True)
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "invpend_explicit.npz"))
ge = lmpc_b200.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=np.array([[-2.0, 2.0]]))
U, reg = ge.evaluate(mpqp.sample_theta(z["lo"], z["hi"], 5000))
print("bst", U.shape, reg.max())
mpc = util.controller("invpend")
ge2 = lmpc_b200.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=np.array([[-2.0, 2.0]]),
                               dims=mpc.parameter_dims())
X0 = mpqp.sample_theta(z["lo"], z["hi"], 3000)[:4] * 0.2
sim = ge2.simulate(mpc.F, mpc.G, X0, N=20, want_traj=True, r_traj=np.zeros((2, 5)))
print("bst sim", sim["Xend"].shape)
