"""Evidence for thread-per-QP at n <= 8 (north_star suggests warp-per-QP up to n = 32; SURVEY 7 lets ncu decide):
the headline batch (invpend, 1e7 theta, s = 1) forced onto the warp-per-QP kernel (settings.kernel = 2) and on
the default kernel, same theta, same results.  Prints one JSON line:  python tests/tools/invpend_on_k2.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import lmpc_b200
from oracle import mpqp
mpc = mpqp.invpend(); qp = mpqp.mpc2mpqp(mpc)
lo, hi = mpqp.invpend_range(1.0); dev = torch.device("cuda", 0)
N = 10_000_000
gen = torch.Generator(device=dev); gen.manual_seed(1234)
Th = torch.rand((N, 7), dtype=torch.float64, device=dev, generator=gen) * torch.tensor(hi - lo, device=dev) + torch.tensor(lo, device=dev)
torch.cuda.synchronize()
out = {}
Us = {}
for label, kern in (("thread_per_qp_default", 0), ("warp_per_qp_forced", 2)):
    st = lmpc_b200.default_settings(); st.kernel = kern
    g = lmpc_b200.setup_gpu(qp, nu=1, dims=mpc.parameter_dims(), settings=st)
    U = torch.empty((N, 1), dtype=torch.float64, device=dev); fl = torch.empty(N, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream(device=dev)
    reps = 20 if kern == 0 else 3
    with torch.cuda.stream(s):
        for _ in range(2): g.solve_async(Th, U, fl, N=N, stream=s.cuda_stream)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(reps): g.solve_async(Th, U, fl, N=N, stream=s.cuda_stream)
        b.record(s); s.synchronize()
    ms = a.elapsed_time(b) / reps
    out[label] = {"kernel": int(g.kernel), "ms_per_1e7": ms, "solves_per_s": N / ms * 1e3, "all_flags_ge_1": bool((fl >= 1).all())}
    Us[label] = U
out["bit_identical_u"] = bool(torch.equal(Us["thread_per_qp_default"], Us["warp_per_qp_forced"]))
out["ratio"] = out["thread_per_qp_default"]["solves_per_s"] / out["warp_per_qp_forced"]["solves_per_s"]
print(json.dumps(out))
