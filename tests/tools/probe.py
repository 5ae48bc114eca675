"""Scratch GPU probe: smoke + kernel-only throughput of the invpend batch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import lmpc_b200
from oracle import mpqp
import __graft_entry__ as ge
ge.smoke()
mpc = mpqp.invpend(); qp = mpqp.mpc2mpqp(mpc)
g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims())
print("fp64 peak TF/s", lmpc_b200.fp64_peak_tflops(0))
for scale in (1.0, 0.25, 0.05):
    lo, hi = mpqp.invpend_range(scale)
    N = 10_000_000
    gen = torch.Generator(device="cuda"); gen.manual_seed(1234)
    Th = torch.rand((N, 7), dtype=torch.float64, device="cuda", generator=gen)
    Th = Th * torch.tensor(hi - lo, device="cuda") + torch.tensor(lo, device="cuda")
    U = torch.empty((N, 1), dtype=torch.float64, device="cuda"); fl = torch.empty(N, dtype=torch.int32, device="cuda")
    it = torch.empty(N, dtype=torch.int32, device="cuda")
    s = torch.cuda.Stream(); 
    with torch.cuda.stream(s):
        for _ in range(3): g.solve_async(Th, U, fl, it, stream=s.cuda_stream)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(5): g.solve_async(Th, U, fl, it, stream=s.cuda_stream)
        b.record(s); s.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"scale {scale}: {ms:.3f} ms per 1e7 -> {N/ms*1e3:.3e} solves/s, mean iters {it.float().mean().item():.2f} max {it.max().item()} flags ok {(fl>0).all().item()}")
