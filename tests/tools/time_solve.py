"""device-resident timing of one controller at several scales: python tests/tools/time_solve.py chain 1.0:100000 0.25:400000"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import lmpc_b200, util
from oracle import mpqp
name = sys.argv[1]
mpc = util.controller(name); qp = mpqp.mpc2mpqp(mpc)
g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight)
for spec in sys.argv[2:]:
    sc, N = spec.split(":"); sc, N = float(sc), int(N)
    lo, hi = util.theta_range(name, sc)
    Th = torch.as_tensor(np.ascontiguousarray(mpqp.sample_theta(lo, hi, N).T), device="cuda")
    U = torch.empty((N, g.nu), dtype=torch.float64, device="cuda"); fl = torch.empty(N, dtype=torch.int32, device="cuda")
    s = torch.cuda.Stream()
    for _ in range(2): g.solve_async(Th, U, fl, N=N, stream=s.cuda_stream)
    s.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    a.record(s)
    for _ in range(reps): g.solve_async(Th, U, fl, N=N, stream=s.cuda_stream)
    b.record(s); s.synchronize()
    print(os.environ.get("LMPC_GPU_LIB", "default"), name, "s=%g" % sc, "%.4g solves/s" % (N * reps / a.elapsed_time(b) * 1e3),
          "ok" if bool((fl >= 1).all()) else "FLAGS<1", flush=True)
