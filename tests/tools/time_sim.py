"""device-resident timing of the AFTI-16 closed loop (BASELINE config 3 shape): python tests/tools/time_sim.py [N] [T]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import lmpc_b200, util
from oracle import mpqp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
mpc = util.controller("aircraft"); qp = mpqp.mpc2mpqp(mpc)
g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight)
nx, nr, nd, nup = mpc.parameter_dims()
lo, hi = util.theta_range("aircraft", 0.25)
X0 = torch.as_tensor(np.ascontiguousarray(mpqp.sample_theta(lo[:nx], hi[:nx], N).T), device="cuda")
R = torch.tensor([0.1, 0.0], dtype=torch.float64, device="cuda").repeat(N, 1).contiguous()
D = torch.zeros((N, nd), dtype=torch.float64, device="cuda")
Xe = torch.empty((N, nx), dtype=torch.float64, device="cuda"); Ul = torch.empty((N, 2), dtype=torch.float64, device="cuda")
fl = torch.empty(N, dtype=torch.int32, device="cuda"); it = torch.empty(N, dtype=torch.int32, device="cuda")
Fm, Gm, Gdm = [np.asfortranarray(a) for a in (mpc.F, mpc.G, mpc.Gd)]
L = lmpc_b200.lib()
def run():
    rc = L.lmpc_gpu_simulate_batch(g._h, Fm.ctypes.data, Gm.ctypes.data, Gdm.ctypes.data, nx, nr, nd, nup, X0.data_ptr(), R.data_ptr(),
                                   D.data_ptr(), None, N, T, 1, Xe.data_ptr(), Ul.data_ptr(), None, fl.data_ptr(), it.data_ptr())
    assert rc == 0
import time
run(); run(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): run()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(os.environ.get("LMPC_GPU_LIB", "default"), "aircraft closed loop %d x %d: %.4g solves/s" % (N, T, N * T / dt), "iters/solve %.2f" % (it.double().mean().item() / T), flush=True)
