import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, lmpc_b200
from oracle import mpqp
mpc = mpqp.invpend(); qp = mpqp.mpc2mpqp(mpc)
g = lmpc_b200.setup_gpu(qp, nu=1, dims=mpc.parameter_dims())
lo, hi = mpqp.invpend_range(1.0); dev = torch.device("cuda", 0)
N = 10_000_000
gen = torch.Generator(device=dev); gen.manual_seed(1234)
Th = torch.rand((N, 7), dtype=torch.float64, device=dev, generator=gen) * torch.tensor(hi - lo, device=dev) + torch.tensor(lo, device=dev)
U = torch.empty((N, 1), dtype=torch.float64, device=dev); fl = torch.empty(N, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(device=dev)
with torch.cuda.stream(s):
    for _ in range(5): g.solve_async(Th, U, fl, N=N, stream=s.cuda_stream)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(20): g.solve_async(Th, U, fl, N=N, stream=s.cuda_stream)
    b.record(s); s.synchronize()
ms = a.elapsed_time(b) / 20
print(os.environ.get("LMPC_GPU_LIB", "in-tree"), f"{ms:.4f} ms  {N/ms*1e3:.4e} solves/s  checksum {U.sum().item():.10f}")
