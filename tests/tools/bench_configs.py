"""Secondary measurements for the other BASELINE configs (3: AFTI-16 closed loop, 4: chain of
masses, 5: explicit evaluate) plus AFTI-16 single solves.  One JSON line per config on stdout.
Device-resident inputs, CUDA events on the launching stream, CPU oracle port beside each."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import lmpc_b200, util
from oracle import mpqp, ldp

dev = torch.device("cuda", 0)
HBM = 6447.5
which = sys.argv[1:] or ["aircraft", "aircraft_sim", "chain", "explicit"]
fp64 = lmpc_b200.fp64_peak_tflops(0)


def rand_theta(lo, hi, N, seed=1234):
    gen = torch.Generator(device=dev); gen.manual_seed(seed)
    T = torch.rand((N, len(lo)), dtype=torch.float64, device=dev, generator=gen)
    return T * torch.tensor(hi - lo, device=dev) + torch.tensor(lo, device=dev)


def timed(fn, reps=3, warm=1):
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        for _ in range(warm): fn(s)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(reps): fn(s)
        b.record(s); s.synchronize()
    return a.elapsed_time(b) / reps


def solve_cfg(name, scale, N, ncpu):
    mpc = util.controller(name); qp = mpqp.mpc2mpqp(mpc)
    g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight, K=mpc.K)
    lo, hi = util.theta_range(name, scale)
    Th = rand_theta(lo, hi, N)
    U = torch.empty((N, g.nu), dtype=torch.float64, device=dev); fl = torch.empty(N, dtype=torch.int32, device=dev)
    it = torch.empty(N, dtype=torch.int32, device=dev)
    ms = timed(lambda s: g.solve_async(Th, U, fl, it, N=N, stream=s.cuda_stream))
    cst = g.ldp_constants(); cst["senses"] = g.senses
    P = util.oracle_from_constants(qp, mpc.nu, cst, soft_weight=mpc.soft_weight)
    sub = Th[:ncpu].cpu().numpy().T
    cores = ldp.num_threads()
    t0 = time.perf_counter(); ref = P.solve_batch(sub, nthreads=cores, want_flops=True); dt = time.perf_counter() - t0
    same = bool(np.array_equal(U[:ncpu].cpu().numpy().T[:, ref["flag"] >= 1], ref["U"][:, ref["flag"] >= 1]) and
                np.array_equal(fl[:ncpu].cpu().numpy(), ref["flag"]))
    flops = float(ref["flops"].mean())
    flags = {int(k): int(v) for k, v in zip(*np.unique(fl.cpu().numpy(), return_counts=True))}
    bps = 8 * g.nth + 8 * g.nu + 8
    print(json.dumps({"config": name, "kernel": g.kernel, "n": g.n, "m": g.m, "nth": g.nth, "theta_scale": scale, "N": N,
                      "ms": ms, "solves_per_s": N / ms * 1e3, "mean_iters": float(it.float().mean()), "flags": flags,
                      "flops_per_solve": flops, "fp64_tflops": flops * N / ms / 1e9, "fp64_peak": fp64,
                      "fp64_frac": flops * N / ms / 1e9 / fp64, "hbm_gbs": bps * N / ms / 1e6,
                      "cpu_port_solves_per_s": ncpu / dt, "cpu_cores": cores, "cpu_sample": ncpu,
                      "bit_identical_to_oracle_on_sample": same}), flush=True)


if "invpend_scales" in which:
    for sc in (1.0, 0.25, 0.05):
        solve_cfg("invpend", sc, 10_000_000, 2_000_000)
if "condensing_features" in which:
    # controllers condensed with a prestabilising feedback (bounds become general rows) and with a
    # 50-step reference preview (theta is 105 doubles: the solve streams 840 B per instance)
    solve_cfg("invpend_prestab", 1.0, 10_000_000, 200_000)
    solve_cfg("invpend_rpreview", 1.0, 4_000_000, 200_000)
if "small_general" in which:
    # small problems with hard general rows / equalities: the general thread-per-QP kernel K1
    solve_cfg("dint_hard", 1.0, 10_000_000, 1_000_000)
    solve_cfg("dint_eq_small", 1.0, 10_000_000, 1_000_000)
if "aircraft" in which:
    solve_cfg("aircraft", 1.0, 1_000_000, 100_000)
    solve_cfg("aircraft", 0.25, 1_000_000, 100_000)
    solve_cfg("aircraft_ksall", 0.25, 200_000, 50_000)
if "warm" in which:
    # warm-started batch (lmpc_gpu_solve_batch_warm): every instance starts from the optimal working
    # set of a theta 2 % of the range away (a caller-managed closed loop between two samples)
    for name, scale, N in (("aircraft", 0.25, 1_000_000), ("invpend", 1.0, 10_000_000)):
        mpc = util.controller(name); qp = mpqp.mpc2mpqp(mpc)
        g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight)
        lo, hi = util.theta_range(name, scale)
        Th0 = rand_theta(lo, hi, N)
        gen = torch.Generator(device=dev); gen.manual_seed(7)
        Th = Th0 + 0.02 * torch.tensor(hi - lo, device=dev) * torch.randn(Th0.shape, dtype=torch.float64, device=dev, generator=gen)
        prev = g.solve(Th0)
        au0, al0 = prev["active_upper"].clone(), prev["active_lower"].clone()
        cold = g.solve(Th)
        U = torch.empty((N, g.nu), dtype=torch.float64, device=dev); fl = torch.empty(N, dtype=torch.int32, device=dev)
        it = torch.empty(N, dtype=torch.int32, device=dev)
        au, al = au0.clone(), al0.clone()
        def run(s):
            au.copy_(au0); al.copy_(al0)
            lmpc_b200.host._check(lmpc_b200.lib().lmpc_gpu_solve_batch_warm(g._h, Th.data_ptr(), N, U.data_ptr(), fl.data_ptr(),
                                                                       it.data_ptr(), au.data_ptr(), al.data_ptr()))
        t0 = time.perf_counter(); run(None); torch.cuda.synchronize()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps): run(None)
        torch.cuda.synchronize(); ms_w = (time.perf_counter() - t0) / reps * 1e3
        ms_c = timed(lambda s: g.solve_async(Th, U, fl, it, N=N, stream=s.cuda_stream))
        ok = (cold["exitflag"] >= 1)
        print(json.dumps({"config": name + "_warm_batch", "theta_scale": scale, "N": N,
                          "warm_solves_per_s": N / ms_w * 1e3, "cold_solves_per_s": N / ms_c * 1e3,
                          "mean_iters_warm": float(g.solve_warm(Th, au0.clone(), al0.clone())["iters"].float().mean()),
                          "mean_iters_cold": float(cold["iters"].float().mean()),
                          "max_abs_du_vs_cold": float((g.solve_warm(Th, au0.clone(), al0.clone())["U"] - cold["U"])[:, ok].abs().max()),
                          "timer": "wall clock incl. the mask reset copies (warm), CUDA events (cold)"}), flush=True)
if "chain" in which:
    solve_cfg("chain", 1.0, 20_000, 1_000)
    solve_cfg("chain", 0.25, 100_000, 4_000)


def sim_cfg(name, scale, N, T, rvec, label):
    mpc = util.controller(name); qp = mpqp.mpc2mpqp(mpc)
    g = lmpc_b200.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight)
    nx, nr, nd, nup = mpc.parameter_dims()
    lo, hi = util.theta_range(name, scale)
    X0 = rand_theta(lo[:nx], hi[:nx], N)
    R = torch.tensor(rvec, dtype=torch.float64, device=dev).repeat(N, 1).contiguous()
    D = torch.zeros((N, nd), dtype=torch.float64, device=dev)
    Xe = torch.empty((N, nx), dtype=torch.float64, device=dev); Ul = torch.empty((N, mpc.nu), dtype=torch.float64, device=dev)
    fl = torch.empty(N, dtype=torch.int32, device=dev); it = torch.empty(N, dtype=torch.int32, device=dev)
    L = lmpc_b200.lib()
    Fm, Gm = np.asfortranarray(mpc.F), np.asfortranarray(mpc.G)
    Gdm = np.asfortranarray(mpc.Gd) if nd else None
    for warm in (1, 0):
        def run(s=None):
            rc = L.lmpc_gpu_simulate_batch(g._h, Fm.ctypes.data, Gm.ctypes.data, (Gdm.ctypes.data if nd else None), nx, nr, nd, nup,
                                           X0.data_ptr(), R.data_ptr(), (D.data_ptr() if nd else None), None, N, T, warm,
                                           Xe.data_ptr(), Ul.data_ptr(), None, fl.data_ptr(), it.data_ptr())
            assert rc == 0
        run(); torch.cuda.synchronize()
        t0 = time.perf_counter(); run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        cst = g.ldp_constants(); cst["senses"] = g.senses
        P = util.oracle_from_constants(qp, mpc.nu, cst, soft_weight=mpc.soft_weight)
        ncpu = 20000
        t1 = time.perf_counter()
        ref = P.simulate_batch(mpc.F, mpc.G, mpc.Gd if nd else None, X0[:ncpu].cpu().numpy().T, R[:ncpu].cpu().numpy().T,
                               (D[:ncpu].cpu().numpy().T if nd else None), None, T, nr, nd, nup, warm=bool(warm))
        dc = time.perf_counter() - t1
        same = bool(np.array_equal(Xe[:ncpu].cpu().numpy().T, ref["Xend"]) and np.array_equal(fl[:ncpu].cpu().numpy(), ref["flag"]))
        flags = {int(k): int(v) for k, v in zip(*np.unique(fl.cpu().numpy(), return_counts=True))}
        print(json.dumps({"config": label, "warm_start": warm, "trajectories": N, "steps": T, "s": dt,
                          "trajectories_per_s": N / dt, "solves_per_s": N * T / dt,
                          "mean_iters_per_solve": float(it.float().mean()) / T, "worst_flags": flags,
                          "cpu_port_solves_per_s": ncpu * T / dc, "cpu_cores": ldp.num_threads(), "cpu_sample": ncpu,
                          "bit_identical_to_oracle_on_sample": same}), flush=True)

if "aircraft_sim" in which:
    # BASELINE config 3: Monte-Carlo closed loop, r = [0.1, 0], d = 0, T = 40 (mpc_examples.jl:206)
    sim_cfg("aircraft", 0.25, 1_000_000, 40, [0.1, 0.0], "aircraft_closed_loop")
if "invpend_sim" in which:
    # the README's own closed loop (README.md:55-60) as a Monte-Carlo batch: 100 steps = 1 s
    sim_cfg("invpend", 0.25, 2_000_000, 100, [0.0, 0.0], "invpend_closed_loop")
if "explicit_sim" in which:
    # closed loop with the explicit law (Simulation(empc)): 2e6 x0, 100 steps, tree walk + dynamics on chip
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "invpend_explicit.npz"))
    mpc = util.controller("invpend")
    ge = lmpc_b200.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"], dims=mpc.parameter_dims())
    B = ldp.OracleBst(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clip=z["clipping"])
    N, T = 2_000_000, 100
    lo, hi = util.theta_range("invpend", 0.25)
    X0 = rand_theta(lo[:4], hi[:4], N)
    R = torch.zeros((N, 2), dtype=torch.float64, device=dev)
    Xe = torch.empty((N, 4), dtype=torch.float64, device=dev); Ul = torch.empty((N, 1), dtype=torch.float64, device=dev)
    Fm, Gm = np.asfortranarray(mpc.F), np.asfortranarray(mpc.G)
    L = lmpc_b200.lib()
    def run():
        rc = L.lmpc_gpu_simulate_bst_batch(ge._t, Fm.ctypes.data, Gm.ctypes.data, None, 4, 2, 0, 1, X0.data_ptr(),
                                           R.data_ptr(), None, None, None, None, N, T, Xe.data_ptr(), Ul.data_ptr(), None)
        assert rc == 0
    run(); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ncpu = 50000
    t1 = time.perf_counter()
    ref = B.simulate_batch(mpc.F, mpc.G, None, X0[:ncpu].cpu().numpy().T, np.zeros((2, ncpu)), None, None, T, 2, 0, 1)
    dc = time.perf_counter() - t1
    same = bool(np.array_equal(Xe[:ncpu].cpu().numpy().T, ref["Xend"]))
    print(json.dumps({"config": "explicit_invpend_closed_loop", "trajectories": N, "steps": T, "s": dt,
                      "evaluations_per_s": N * T / dt, "cpu_port_evaluations_per_s": ncpu * T / dc,
                      "cpu_cores": ldp.num_threads(), "bit_identical_to_oracle_on_sample": same}), flush=True)
if "explicit" in which:
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "invpend_explicit.npz"))
    ge = lmpc_b200.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"])
    B = ldp.OracleBst(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clip=z["clipping"])
    N = 100_000_000
    Th = rand_theta(z["lo"], z["hi"], N)
    U = torch.empty((N, 1), dtype=torch.float64, device=dev)
    ms = timed(lambda s: ge.evaluate_async(Th, U, N=N, stream=s.cuda_stream))
    ncpu = 4_000_000
    sub = Th[:ncpu].cpu().numpy().T
    t0 = time.perf_counter(); Uo, _ = B.evaluate_batch(sub); dt = time.perf_counter() - t0
    same = bool(np.array_equal(Uo, U[:ncpu].cpu().numpy().T))
    print(json.dumps({"config": "explicit_invpend", "N": N, "nodes": int(len(z["hp_list"])), "ms": ms,
                      "theta_per_s": N / ms * 1e3, "hbm_gbs": 64.0 * N / ms / 1e6, "hbm_frac": 64.0 * N / ms / 1e6 / HBM,
                      "cpu_port_theta_per_s": ncpu / dt, "cpu_cores": ldp.num_threads(), "cpu_sample": ncpu,
                      "bit_identical_to_oracle_on_sample": same}), flush=True)
