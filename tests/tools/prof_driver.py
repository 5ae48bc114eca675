"""Launches one kernel family a few times so ncu can capture it:  python tests/tools/prof_driver.py box|warp|cta|bst"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import lmpc_b200, util
from oracle import mpqp
dev = torch.device("cuda", 0)
which = sys.argv[1]
def rand_theta(lo, hi, N):
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    T = torch.rand((N, len(lo)), dtype=torch.float64, device=dev, generator=gen)
    return T * torch.tensor(hi - lo, device=dev) + torch.tensor(lo, device=dev)
if which == "sim":
    mpc = util.controller("aircraft"); qp = mpqp.mpc2mpqp(mpc)
    g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight)
    nx, nr, nd, nup = mpc.parameter_dims()
    N, T = 500_000, 40
    lo, hi = util.theta_range("aircraft", 0.25)
    X0 = rand_theta(lo[:nx], hi[:nx], N)
    R = torch.tensor([0.1, 0.0], dtype=torch.float64, device=dev).repeat(N, 1).contiguous()
    D = torch.zeros((N, nd), dtype=torch.float64, device=dev)
    Xe = torch.empty((N, nx), dtype=torch.float64, device=dev); Ul = torch.empty((N, 2), dtype=torch.float64, device=dev)
    fl = torch.empty(N, dtype=torch.int32, device=dev)
    Fm, Gm, Gdm = [np.asfortranarray(a) for a in (mpc.F, mpc.G, mpc.Gd)]
    for _ in range(3):
        rc = lmpc_b200.lib().lmpc_gpu_simulate_batch(g._h, Fm.ctypes.data, Gm.ctypes.data, Gdm.ctypes.data, nx, nr, nd, nup,
                                                     X0.data_ptr(), R.data_ptr(), D.data_ptr(), None, N, T, 1,
                                                     Xe.data_ptr(), Ul.data_ptr(), None, fl.data_ptr(), None)
        assert rc == 0
elif which == "bst":
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "invpend_explicit.npz"))
    ge = lmpc_b200.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"])
    N = 20_000_000
    Th = rand_theta(z["lo"], z["hi"], N); U = torch.empty((N, 1), dtype=torch.float64, device=dev)
    for _ in range(4): ge.evaluate_async(Th, U, N=N)
else:
    name, scale, N = {"box": ("invpend", 1.0, 10_000_000), "warp": ("aircraft", 0.25, 300_000), "cta": ("chain", 0.25, 3000), "big": ("aircraft_ksall", 0.25, 100_000), "thread": ("dint_eq_small", 1.0, 10_000_000), "cta1": ("chain", 1.0, 1500), "rpreview": ("invpend_rpreview", 1.0, 4_000_000), "prestab": ("invpend_prestab", 1.0, 10_000_000)}[which]
    mpc = util.controller(name); qp = mpqp.mpc2mpqp(mpc)
    g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight)
    Th = rand_theta(*util.theta_range(name, scale), N)
    U = torch.empty((N, g.nu), dtype=torch.float64, device=dev); fl = torch.empty(N, dtype=torch.int32, device=dev)
    for _ in range(4): g.solve_async(Th, U, fl, N=N)
torch.cuda.synchronize()
