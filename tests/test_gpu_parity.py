"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.

Bar (BASELINE.json north_star): exit flag and final active set identical, u within 1e-9
relative.  Because kernels and oracle share the summation order, we additionally assert
bit-identical u wherever both succeed.  Run on the GPU box with `pytest -m gpu`."""
import os

import numpy as np
import pytest

import util
from oracle import mpqp, ldp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_cache = {}


def gpu_controller(lmpc, name, kernel=0):
    key = (name, kernel)
    if key not in _cache:
        mpc = util.controller(name)
        qp = mpqp.mpc2mpqp(mpc)
        st = lmpc.default_settings(kernel=kernel)
        try:
            g = lmpc.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight, settings=st)
        except lmpc.LmpcError as e:
            if e.code == -101 and "no kernel supports" in str(e):
                pytest.skip(f"{name}: kernel {kernel} does not support these dimensions")
            raise
        cst = g.ldp_constants()
        cst["senses"] = g.senses
        P = util.oracle_from_constants(qp, mpc.nu, cst, soft_weight=mpc.soft_weight)
        _cache[key] = (mpc, qp, g, P)
    return _cache[key]


def assert_parity(out, ref, rel=1e-9, bitexact=True):
    assert np.array_equal(out["exitflag"], ref["flag"])
    assert np.array_equal(out["iters"], ref["iters"])
    assert np.array_equal(out["active_upper"], ref["act_u"])
    assert np.array_equal(out["active_lower"], ref["act_l"])
    ok = ref["flag"] >= 1
    U, V = np.asarray(out["U"])[:, ok], ref["U"][:, ok]
    err = np.max(np.abs(U - V) / np.maximum(1.0, np.abs(V))) if ok.any() else 0.0
    assert err <= rel, err
    if bitexact:
        assert np.array_equal(U, V)


@pytest.mark.parametrize("name", ["invpend", "simple", "simple_q1000"])
def test_known_answers_on_gpu(lmpc, name):
    """runtests.jl:62-66, simple.md:107, simple.md:203-205 through compute_control on the GPU."""
    mpc, qp, g, P = gpu_controller(lmpc, name)
    k = util.KNOWN[name]
    u = g.compute_control(np.array(k["x"]), r=k["r"])
    assert abs(u[0] - k["u"]) < k["tol"]


@pytest.mark.parametrize("p,expect", [(0.0, 1.0), (0.75, 0.25)])
def test_known_answer_generalized_parameter_on_gpu(lmpc, p, expect):
    """runtests.jl:1250-1268 through the GPU (f != 0: nonzero v, u_offset)."""
    mpc, qp, g, P = gpu_controller(lmpc, "genparam")
    out = g.solve(mpc.form_parameter([0.0], r=[0.0], p=[p]).reshape(-1, 1))
    assert out["exitflag"][0] >= 1 and abs(out["U"][0, 0] - expect) < 1e-6


CASES = [("invpend", 1.0, 10000), ("invpend", 0.25, 10000), ("invpend", 0.05, 10000),
         ("dint_hard", 1.0, 10000), ("dint_hard", 0.3, 4000), ("genparam", 1.0, 5000),
         ("aircraft", 1.0, 4000), ("aircraft", 0.25, 4000), ("aircraft_ksall", 0.25, 2000),
         ("simple", 1.0, 4000), ("simple_q1000", 0.5, 2000),
         ("chain_small", 1.0, 2000), ("chain", 1.0, 300), ("chain", 0.25, 300),
         ("dint_eq", 1.0, 4000), ("dint_eq_small", 1.0, 4000),   # terminal equality rows
         # soft rows on the whole horizon (the reference's default ks): the refactorisation repair runs
         ("aircraft_ksall", 1.0, 4000), ("aircraft_ksall", 0.5, 4000),
         # every box-kernel instantiation <n, nu>: n = 1..8 with nu = 1, n = 2, 4, 6, 8 with nu = 2
         ("invpend_nc1", 1.0, 4000), ("invpend_nc2", 1.0, 4000), ("invpend_nc3", 1.0, 4000),
         ("invpend_nc4", 0.25, 4000), ("invpend_nc6", 1.0, 4000), ("invpend_nc7", 0.25, 4000),
         ("invpend_nc8", 1.0, 4000), ("invpend_nc8", 0.25, 4000),
         ("aircraft_box1", 1.0, 4000), ("aircraft_box2", 0.25, 4000), ("aircraft_box3", 1.0, 4000),
         ("aircraft_box4", 0.25, 4000), ("aircraft_box4", 1.0, 2000)]


@pytest.mark.parametrize("name,scale,N", CASES)
def test_batch_matches_oracle(lmpc, name, scale, N):
    """BASELINE config 1 (10^4 sampled states) and the other controllers at oracle-sized N."""
    mpc, qp, g, P = gpu_controller(lmpc, name)
    lo, hi = util.theta_range(name, scale)
    Th = mpqp.sample_theta(lo, hi, N)
    assert_parity(g.solve(Th), P.solve_batch(Th))


BOX_CASES = [c for c in CASES if c[0] == "invpend" or c[0].startswith("invpend_nc") or c[0].startswith("aircraft_box")]


@pytest.mark.parametrize("name,scale,N", BOX_CASES)
def test_box_fast_path_matches_oracle(lmpc, monkeypatch, name, scale, N):
    """K1c: the add-only fast kernel + the REDO launch of the generic box kernel, which the library takes from
    2e5 instances on; LMPC_BOX_FAST_MIN_BATCH = 0 pushes these oracle-sized batches through both launches
    (every <n, nu> instantiation): flags, iteration counts, active sets and u identical to the oracle, and to
    the single-launch path the default threshold selects at this size."""
    mpc, qp, g, P = gpu_controller(lmpc, name)
    lo, hi = util.theta_range(name, scale)
    Th = mpqp.sample_theta(lo, hi, N, seed=11)
    single = g.solve(Th)
    launches = lmpc.launch_count()
    monkeypatch.setenv("LMPC_BOX_FAST_MIN_BATCH", "0")
    fast = g.solve(Th)
    assert lmpc.launch_count() - launches == 2          # fast kernel + REDO launch
    assert_parity(fast, P.solve_batch(Th))
    for k in ("U", "exitflag", "iters", "active_upper", "active_lower"):
        assert np.array_equal(fast[k], single[k]), k
    # ragged sizes around the warp / CTA / scan-step boundaries
    for n in (1, 31, 33, 127, 129, 2047, 2049):
        out = g.solve(Th[:, :n])
        assert np.array_equal(out["U"], single["U"][:, :n]) and np.array_equal(out["exitflag"], single["exitflag"][:n])


def test_box_fast_path_with_many_handovers(lmpc, monkeypatch):
    """K1c's hand-over queues under load: invpend far outside its nominal range (s = 3), 1e6 instances, so that every
    CTA of the persistent grid runs several passes and a large share of the instances block a multiplier and travel
    through the queues to the REDO launch.  The two-launch path (default from 2e5 instances on) must equal the
    single-launch generic kernel bit for bit, and a subsample the oracle."""
    import torch
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    lo, hi = util.theta_range("invpend", 3.0)
    N = 1_000_003
    gen = torch.Generator(device="cuda")
    gen.manual_seed(99)
    Th = torch.rand((N, len(lo)), dtype=torch.float64, device="cuda", generator=gen)
    Th = Th * torch.tensor(hi - lo, device="cuda") + torch.tensor(lo, device="cuda")
    launches = lmpc.launch_count()
    fast = g.solve(Th)
    assert lmpc.launch_count() - launches == 2
    monkeypatch.setenv("LMPC_BOX_FAST_MIN_BATCH", str(10 ** 12))
    launches = lmpc.launch_count()
    single = g.solve(Th)
    assert lmpc.launch_count() - launches == 1
    for k in ("U", "exitflag", "iters", "active_upper", "active_lower"):
        assert bool((fast[k] == single[k]).all()), k
    # instances that removed a row needed more iterations than rows they ended with + 1: those were handed over
    k_final = util.bits(single["active_upper"].cpu().numpy().view(np.uint64).reshape(N, -1), 5).sum(1) + \
        util.bits(single["active_lower"].cpu().numpy().view(np.uint64).reshape(N, -1), 5).sum(1)
    handed = (single["iters"].cpu().numpy() > k_final + 1) & (single["exitflag"].cpu().numpy() == 1)
    assert handed.mean() > 0.03, handed.mean()
    idx = torch.randperm(N, device="cuda")[:20_000]
    ref = P.solve_batch(Th[idx].cpu().numpy().T)
    assert np.array_equal(fast["U"][:, idx].cpu().numpy(), ref["U"])
    assert np.array_equal(fast["exitflag"][idx].cpu().numpy(), ref["flag"])
    assert np.array_equal(fast["iters"][idx].cpu().numpy(), ref["iters"])


def test_box_kernel_is_what_runs_on_box_only_problems(lmpc):
    """the cases above really exercise K1b: box-only problems with n <= 8 select the thread kernel and
    the library takes its box fast path (lmpc_gpu_get_info reports the thread kernel, capacity n)."""
    for name in ("invpend_nc1", "invpend_nc8", "aircraft_box1", "aircraft_box4"):
        mpc, qp, g, P = gpu_controller(lmpc, name)
        assert g.kernel == lmpc.KERNEL_THREAD and g.cap == qp.n and not np.any(g.senses)


@pytest.mark.parametrize("scale", [1.0, 0.5, 0.25])
def test_default_output_bounds_never_end_infeasible_at_size(lmpc, scale):
    """AFTI-16 with output bounds on the reference's default ks = 2:Np (18 soft rows) over 1e6 theta:
    flags in {1, 2} only (a soft problem is feasible by construction; round 1 returned -1 on 22 % at
    s = 1), and a 2e4 subsample equals the oracle bit for bit.  One numerically cycling instance per
    3e6 is the known residue (tests/test_oracle_golden.py)."""
    mpc, qp, g, P = gpu_controller(lmpc, "aircraft_ksall")
    lo, hi = util.theta_range("aircraft_ksall", scale)
    N = 1_000_000
    Th = mpqp.sample_theta(lo, hi, N, seed=4321)
    out = g.solve(Th, want_active=False)
    bad = np.count_nonzero((out["exitflag"] != 1) & (out["exitflag"] != 2))
    assert bad <= 1, np.unique(out["exitflag"], return_counts=True)
    sub = np.random.default_rng(1).choice(N, 20000, replace=False)
    ref = P.solve_batch(Th[:, sub])
    assert np.array_equal(out["exitflag"][sub], ref["flag"]) and np.array_equal(out["iters"][sub], ref["iters"])
    assert np.array_equal(np.asarray(out["U"])[:, sub], ref["U"])


def test_closed_loop_at_config_3_size(lmpc):
    """BASELINE config 3 at its stated size: AFTI-16, 10^6 initial states x 40 steps, warm start, the example's
    scenario r = [0.1, 0] (mpc_examples.jl:206).  Size-independent properties on the whole batch (every step of
    every loop ends with flag 1 or 2, finite states, running it twice gives the same bits) and a 1e5-loop
    subsample equal to the oracle's loop bit for bit (states, last controls, flags, iteration totals)."""
    import torch
    mpc, qp, g, P = gpu_controller(lmpc, "aircraft")
    nx, nr, nd, nup = mpc.parameter_dims()
    N, T = 1_000_000, 40
    lo, hi = util.theta_range("aircraft", 0.25)
    X0 = mpqp.sample_theta(lo[:nx], hi[:nx], N, seed=2024)
    a = g.simulate(mpc.F, mpc.G, X0, r=[0.1, 0.0], d=[0.0, 0.0], N=T, Gd=mpc.Gd, warm_start=True)
    b = g.simulate(mpc.F, mpc.G, X0, r=[0.1, 0.0], d=[0.0, 0.0], N=T, Gd=mpc.Gd, warm_start=True)
    assert np.all((a["exitflag"] == 1) | (a["exitflag"] == 2)) and np.all(np.isfinite(a["Xend"]))
    assert np.all(np.abs(a["Ulast"]) <= 0.5 + 1e-9)                        # input bounds, mpc_examples.jl:188
    assert np.array_equal(a["Xend"], b["Xend"]) and np.array_equal(a["iters"], b["iters"])
    sub = np.random.default_rng(3).choice(N, 100_000, replace=False)
    R = np.repeat(np.array([[0.1], [0.0]]), len(sub), axis=1)
    ref = P.simulate_batch(mpc.F, mpc.G, mpc.Gd, X0[:, sub], R, np.zeros((nd, len(sub))), None, T, nr, nd, nup, warm=True)
    assert np.array_equal(a["Xend"][:, sub], ref["Xend"]) and np.array_equal(a["Ulast"][:, sub], ref["Ulast"])
    assert np.array_equal(a["exitflag"][sub], ref["flag"]) and np.array_equal(a["iters"][sub], ref["iters"])


@pytest.mark.parametrize("fn", sorted(f for f in os.listdir(GOLD) if f.endswith(".npz") and "_s" in f))
def test_batch_matches_golden_fixture(lmpc, fn):
    z = np.load(os.path.join(GOLD, fn))
    name = fn.rsplit("_s", 1)[0]
    mpc, qp, g, P = gpu_controller(lmpc, name)
    out = g.solve(z["Theta"])
    assert np.array_equal(out["exitflag"], z["flag"])
    ok = z["flag"] >= 1   # the working set left behind by an infeasible exit is not a result
    assert np.array_equal(out["active_upper"][ok], z["act_u"][ok])
    assert np.array_equal(out["active_lower"][ok], z["act_l"][ok])
    rel = 1e-9 if name in ("invpend", "chain_small") else 1e-5   # setup-step rounding x cond(H)
    assert np.max(np.abs(out["U"][:, ok] - z["U"][:, ok]) / np.maximum(1, np.abs(z["U"][:, ok]))) < rel


@pytest.mark.parametrize("name,kernel", [("invpend", 2), ("dint_hard", 2), ("invpend", 3), ("aircraft", 3),
                                         ("dint_eq_small", 2), ("dint_eq_small", 3), ("dint_eq", 3)])
def test_every_kernel_gives_the_same_answer(lmpc, name, kernel):
    """thread- / warp- / CTA-per-QP kernels are interchangeable on problems they all support."""
    mpc, qp, g, P = gpu_controller(lmpc, name, kernel)
    assert g.kernel == kernel
    lo, hi = util.theta_range(name, 1.0)
    Th = mpqp.sample_theta(lo, hi, 3000, seed=99)
    assert_parity(g.solve(Th), P.solve_batch(Th))


@pytest.mark.parametrize("N", [0, 1, 2, 127, 128, 129, 4097])
def test_ragged_batch_sizes(lmpc, N):
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    lo, hi = mpqp.invpend_range(1.0)
    Th = mpqp.sample_theta(lo, hi, max(N, 1))[:, :N]
    out = g.solve(Th)
    assert out["U"].shape == (1, N)
    if N:
        assert_parity(out, P.solve_batch(Th))


def test_full_size_properties(lmpc):
    """BASELINE config 2 size (10^7 theta): size-independent properties instead of the oracle:
    every instance optimal, bounds respected, idempotent, device path == chunked host path,
    and a 1e5 random subsample equals the oracle bit for bit."""
    import torch
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    lo, hi = mpqp.invpend_range(1.0)
    N = 10_000_000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234)
    Th = torch.rand((N, 7), dtype=torch.float64, device="cuda", generator=gen)
    Th = Th * torch.tensor(hi - lo, device="cuda") + torch.tensor(lo, device="cuda")
    a = g.solve(Th)
    b = g.solve(Th)
    assert bool((a["exitflag"] == 1).all())
    assert bool((a["U"].abs() <= 2.0 + 1e-5).all())          # |u| <= 2 (README.md:47)
    assert bool((a["U"] == b["U"]).all()) and bool((a["active_upper"] == b["active_upper"]).all())
    host = g.solve(np.asfortranarray(Th.cpu().numpy().T), want_active=False)   # 10 pipeline chunks
    assert np.array_equal(host["U"], a["U"].cpu().numpy())
    idx = torch.randperm(N, device="cuda")[:100_000]
    sub = Th[idx].cpu().numpy().T
    ref = P.solve_batch(sub)
    assert np.array_equal(a["U"][:, idx].cpu().numpy(), ref["U"])
    assert np.array_equal(a["exitflag"][idx].cpu().numpy(), ref["flag"])
    # active-set size histogram sanity (at s=1 most instances saturate all 5 bounds)
    k = util.bits(ref["act_u"], 5).sum(1) + util.bits(ref["act_l"], 5).sum(1)
    assert (k == 5).mean() > 0.5


def test_compute_control_c_twin(lmpc):
    """Batched mpc_compute_control(control, state, reference, disturbance): in/out control."""
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    rng = np.random.default_rng(3)
    N = 5000
    X = rng.uniform(-5, 5, (4, N)); R = rng.uniform(-2, 2, (2, N)); Up = rng.uniform(-2, 2, (1, N))
    ctl, flag = g.compute_control_c(Up.copy(), X, R, None)
    Th = np.vstack([X, R, Up])
    ref = P.solve_batch(Th)
    assert np.array_equal(ctl, ref["U"]) and np.array_equal(flag, ref["flag"])
    U2 = g.compute_control(X, r=R, uprev=Up)
    assert np.array_equal(U2, ref["U"])


@pytest.mark.parametrize("name", ["aircraft", "dint_hard", "aircraft_ksall"])
def test_compute_control_c_twin_other_kernels(lmpc, name):
    """theta assembled inside the kernel from state / reference / disturbance / control for the
    general thread kernel, the warp kernel and the large-problem kernel."""
    mpc, qp, g, P = gpu_controller(lmpc, name)
    nx, nr, nd, nup = mpc.parameter_dims()
    lo, hi = util.theta_range(name, 0.25)
    Th = mpqp.sample_theta(lo, hi, 3000, seed=17)
    X, R = Th[:nx], Th[nx:nx + nr]
    D = Th[nx + nr:nx + nr + nd] if nd else None
    ctl0 = np.zeros((g.nu, Th.shape[1]))
    if nup:
        ctl0[:nup] = Th[nx + nr + nd:]
    ctl, flag = g.compute_control_c(ctl0.copy(), X, R, D)
    ref = P.solve_batch(Th)
    ok = ref["flag"] >= 1
    assert np.array_equal(flag, ref["flag"]) and np.array_equal(ctl[:, ok], ref["U"][:, ok])


@pytest.mark.parametrize("name,T", [("invpend", 100), ("aircraft", 40), ("aircraft_ksall", 40), ("chain_small", 15)])
@pytest.mark.parametrize("warm", [False, True])
def test_closed_loop_matches_oracle(lmpc, name, T, warm):
    """BASELINE config 3 shape: N independent Simulation loops (simulation.jl:93-113)."""
    mpc, qp, g, P = gpu_controller(lmpc, name)
    nx, nr, nd, nup = mpc.parameter_dims()
    rng = np.random.default_rng(11)
    N = 512
    lo, hi = util.theta_range(name, 0.25)
    X0 = rng.uniform(lo[:nx], hi[:nx], (N, nx)).T
    r = np.array([0.1, 0.0]) if name.startswith("aircraft") else np.zeros(nr)
    Rm = np.repeat(r.reshape(-1, 1), N, axis=1)
    Dm = np.zeros((nd, N)) if nd else None
    out = g.simulate(mpc.F, mpc.G, X0, r=Rm, d=Dm, N=T, Gd=mpc.Gd if nd else None, warm_start=warm,
                     want_traj=True)
    ref = P.simulate_batch(mpc.F, mpc.G, mpc.Gd if nd else None, X0, Rm, Dm, None, T, nr, nd, nup,
                           warm=warm, want_traj=True)
    assert np.array_equal(out["exitflag"], ref["flag"])
    assert np.array_equal(out["iters"], ref["iters"])
    assert np.array_equal(out["Utraj"], ref["Utraj"])
    assert np.array_equal(out["Xend"], ref["Xend"])


@pytest.mark.parametrize("name,T", [("invpend", 60), ("aircraft", 40), ("aircraft_ksall", 30), ("chain_small", 12)])
def test_closed_loop_with_reference_and_disturbance_trajectories(lmpc, name, T):
    """Scenario trajectories rs[:, k], ds[:, k] shared by every closed loop (simulation.jl:70-110):
    a reference step half way, a disturbance pulse, both shorter than the horizon (last value held)."""
    mpc, qp, g, P = gpu_controller(lmpc, name)
    nx, nr, nd, nup = mpc.parameter_dims()
    rng = np.random.default_rng(5)
    N = 300
    lo, hi = util.theta_range(name, 0.25)
    X0 = rng.uniform(lo[:nx], hi[:nx], (N, nx)).T
    Tr = T - 5
    Rt = None
    if nr:
        Rt = np.zeros((nr, Tr))
        Rt[0, Tr // 2:] = 0.1
        Rt += 0.01 * rng.standard_normal((nr, Tr))
    Dt = None
    if nd:
        Dt = np.zeros((nd, T // 2))
        Dt[:, 3:8] = 0.05
    for warm in (False, True):
        out = g.simulate(mpc.F, mpc.G, X0, N=T, Gd=mpc.Gd if nd else None, warm_start=warm, want_traj=True,
                         r_traj=Rt, d_traj=Dt)
        hold = lambda v: None if v is None else np.hstack([v, np.repeat(v[:, -1:], T - v.shape[1], axis=1)])
        ref = P.simulate_batch(mpc.F, mpc.G, mpc.Gd if nd else None, X0, None, None, None, T, nr, nd, nup,
                               warm=warm, want_traj=True, Rtraj=hold(Rt), Dtraj=hold(Dt))
        assert np.array_equal(out["exitflag"], ref["flag"])
        assert np.array_equal(out["iters"], ref["iters"])
        assert np.array_equal(out["Utraj"], ref["Utraj"])
        assert np.array_equal(out["Xend"], ref["Xend"])
    if nr:   # the trajectory matters: a constant-reference run differs
        const = g.simulate(mpc.F, mpc.G, X0, r=np.zeros(nr), d=np.zeros(nd) if nd else None, N=T,
                           Gd=mpc.Gd if nd else None, want_traj=True)
        assert not np.array_equal(const["Utraj"], out["Utraj"])


def test_warm_equals_cold_on_gpu(lmpc):
    """runtests.jl:85-117 on the GPU: 100 closed-loop steps, |u_warm - u_cold| < 1e-9."""
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    X0 = np.array([[5.0, 5, 0, 0], [1.0, -2, 0.1, 0], [0, 0, 0.15, 0], [10.0, -8, 0.2, 1]]).T
    a = g.simulate(mpc.F, mpc.G, X0, N=100, warm_start=False, want_traj=True)
    b = g.simulate(mpc.F, mpc.G, X0, N=100, warm_start=True, want_traj=True)
    assert np.all(a["exitflag"] >= 1) and np.all(b["exitflag"] >= 1)
    assert np.max(np.abs(a["Utraj"] - b["Utraj"])) < 1e-9
    assert abs(a["Utraj"][0, 0, 0] - 1.7612519326) < 1e-6


def test_unsupported_problems_are_refused(lmpc):
    mpc = util.controller("invpend")
    qp = mpqp.mpc2mpqp(mpc)
    with pytest.raises(lmpc.LmpcError) as e:
        lmpc.GpuMPC(qp.H, qp.f, qp.f_theta, qp.A, qp.bu, qp.bl, qp.W, qp.senses, nu=1, has_binaries=True)
    assert e.value.code == -101
    with pytest.raises(lmpc.LmpcError):
        lmpc.GpuMPC(qp.H, qp.f, qp.f_theta, qp.A, qp.bu, qp.bl, qp.W, qp.senses, nu=1, is_symmetric=False)
    with pytest.raises(lmpc.LmpcError):
        lmpc.GpuMPC(qp.H, qp.f, qp.f_theta, qp.A, qp.bu, qp.bl, qp.W, qp.senses, nu=1, n_break_points=2)


def test_native_library_is_what_ran(lmpc):
    """Guard against a silent fallback: the in-tree .so is loaded and counted launches."""
    before = lmpc.launch_count()
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    g.solve(mpqp.sample_theta(*mpqp.invpend_range(1.0), 256))
    assert lmpc.launch_count() > before
    with open("/proc/self/maps") as f:
        assert "liblmpcgpu.so" in f.read()


@pytest.mark.parametrize("name,N", [("invpend", 100003), ("aircraft", 2001), ("invpend", 1)])
def test_single_process_multi_device_handle(lmpc, name, N):
    """lmpc_gpu_solve_batch_multi: contiguous theta ranges over every visible GPU (one GPU on the
    round-end box: the range logic still runs, with the same device listed twice)."""
    mpc, qp, g, P = gpu_controller(lmpc, name)
    nd = lmpc.device_count()
    devices = list(range(nd)) if nd > 1 else [0, 0]
    mg = lmpc.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight, devices=devices)
    lo, hi = util.theta_range(name, 1.0)
    Th = mpqp.sample_theta(lo, hi, N, seed=5)
    assert_parity(mg.solve(Th), P.solve_batch(Th))
    mg.close()


@pytest.mark.parametrize("N", [5, 300000, 1200007])
def test_pinned_and_pageable_host_buffers_agree(lmpc, N):
    """Pinned caller arrays are DMA'd in place, pageable ones go through the pinned ring
    (several stages at the larger sizes): same bits either way."""
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    lo, hi = mpqp.invpend_range(1.0)
    Th = np.asfortranarray(mpqp.sample_theta(lo, hi, N, seed=21))
    a = g.solve(Th)
    Tp = lmpc.pinned_empty(Th.shape, order="F")
    Tp[...] = Th
    words = (g.m + 63) // 64
    U = lmpc.pinned_empty((g.nu, N), order="F"); fl = lmpc.pinned_empty(N, np.int32)
    it = lmpc.pinned_empty(N, np.int32)
    au = lmpc.pinned_empty((N, words), np.uint64); al = lmpc.pinned_empty((N, words), np.uint64)
    rc = lmpc.lib().lmpc_gpu_solve_batch(g._h, Tp.ctypes.data, N, U.ctypes.data, fl.ctypes.data,
                                         it.ctypes.data, au.ctypes.data, al.ctypes.data)
    assert rc == 0, lmpc.lib().lmpc_gpu_last_error()
    assert np.array_equal(a["U"], U) and np.array_equal(a["exitflag"], fl) and np.array_equal(a["iters"], it)
    assert np.array_equal(a["active_upper"], au) and np.array_equal(a["active_lower"], al)
    if N <= 300000:
        assert_parity(a, P.solve_batch(Th))


WARM_CASES = [("invpend", 0), ("invpend", 2), ("invpend", 3), ("dint_hard", 0), ("dint_eq_small", 0), ("dint_eq", 0),
              ("aircraft", 0), ("aircraft", 3), ("aircraft_ksall", 0), ("simple", 0), ("chain_small", 0), ("chain", 0)]


@pytest.mark.parametrize("name,kernel", WARM_CASES)
def test_warm_started_batch_matches_oracle(lmpc, name, kernel):
    """lmpc_gpu_solve_batch_warm against the oracle's activate_rows + loop: start every instance from
    the optimal working set of a nearby theta (what a caller-managed closed loop does), bit for bit.
    Also from an empty set (== cold start) and from the instance's own optimum (one iteration)."""
    mpc, qp, g, P = gpu_controller(lmpc, name, kernel)
    N = 200 if name == "chain" else 2000
    lo, hi = util.theta_range(name, 0.5)
    Th0 = mpqp.sample_theta(lo, hi, N, seed=3)
    prev = P.solve_batch(Th0)
    Th = Th0 + 0.02 * (hi - lo)[:, None] * np.random.default_rng(1).standard_normal(Th0.shape)
    ref = P.solve_batch_warm(Th, prev["act_u"], prev["act_l"])
    out = g.solve_warm(Th, prev["act_u"], prev["act_l"])
    assert_parity(out, ref)
    cold = P.solve_batch(Th)
    assert ref["iters"].mean() < cold["iters"].mean() or cold["iters"].mean() < 1.5
    # empty working sets: exactly the cold start
    zero = np.zeros_like(prev["act_u"])
    assert_parity(g.solve_warm(Th, zero, zero), cold)
    # the instance's own optimal set: the first iteration confirms it
    own = g.solve_warm(Th, cold["act_u"], cold["act_l"])
    okc = cold["flag"] == 1
    assert np.all(own["iters"][okc] == 1)
    assert np.array_equal(own["active_upper"][okc], cold["act_u"][okc])


def test_warm_start_on_device_buffers(lmpc):
    """Device-resident in/out working sets: a caller-managed closed loop, 30 steps, the masks of
    step t feed step t + 1; same controls as cold solves to 1e-9 (runtests.jl:85-117's bound)."""
    import torch
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    nx, nr, nd, nup = mpc.parameter_dims()
    N = 4096
    lo, hi = util.theta_range("invpend", 0.25)
    rng = np.random.default_rng(2)
    X = rng.uniform(lo[:nx], hi[:nx], (N, nx))
    dev = torch.device("cuda", 0)
    au = torch.zeros((N, 1), dtype=torch.int64, device=dev)
    al = torch.zeros((N, 1), dtype=torch.int64, device=dev)
    uprev = np.zeros((N, nup))
    it_w = it_c = 0
    for t in range(30):
        Th = np.hstack([X, np.zeros((N, nr)), uprev])
        Td = torch.tensor(Th, dtype=torch.float64, device=dev)
        w = g.solve_warm(Td, au, al)
        c = g.solve(Td)
        assert torch.all(w["exitflag"] >= 1)
        assert torch.max(torch.abs(w["U"] - c["U"])).item() < 1e-9
        it_w += int(w["iters"].sum()); it_c += int(c["iters"].sum())
        u = w["U"].T.cpu().numpy()
        X = X @ mpc.F.T + u @ mpc.G.T
        uprev = u[:, :nup]
    assert it_w < it_c


def test_one_handle_many_host_threads(lmpc):
    """Host-buffer calls on one handle from several threads (serialised inside the handle) and
    device-pointer calls on separate streams: every caller gets its own, correct answer."""
    import threading
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    lo, hi = mpqp.invpend_range(1.0)
    Ths = [mpqp.sample_theta(lo, hi, 300000 + 1000 * i, seed=40 + i) for i in range(6)]
    outs = [None] * len(Ths)

    def work(i):
        outs[i] = g.solve(Ths[i])
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(Ths))]
    [t.start() for t in th]
    [t.join() for t in th]
    for i, Th in enumerate(Ths):
        ref = P.solve_batch(Th[:, :20000])
        sub = {k: (v[:, :20000] if k == "U" else v[:20000]) for k, v in outs[i].items()}
        assert_parity(sub, ref)
        assert np.all(outs[i]["exitflag"] >= 1)


@pytest.mark.parametrize("name,kernel", [("invpend", 0), ("invpend", 2), ("dint_hard", 0), ("aircraft", 0),
                                         ("aircraft_ksall", 0), ("chain_small", 0), ("chain", 0)])
@pytest.mark.parametrize("over", [dict(iter_limit=4), dict(iter_limit=7, primal_tol=1e-4),
                                  dict(primal_tol=1e-3, dual_tol=1e-9, cycle_tol=2, progress_tol=1e-3)])
def test_non_default_settings_match_oracle(lmpc, name, kernel, over):
    """lmpc_settings reach the kernels: a tight iteration limit (flag -4 where it bites, on every
    kernel including the box kernel's specialised first iterations), looser tolerances and a short
    cycle guard give the oracle's flags, iteration counts, active sets and controls bit for bit."""
    mpc = util.controller(name)
    qp = mpqp.mpc2mpqp(mpc)
    st = lmpc.default_settings(kernel=kernel, **over)
    g = lmpc.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight, settings=st)
    cst = g.ldp_constants()
    cst["senses"] = g.senses
    P = util.oracle_from_constants(qp, mpc.nu, cst, soft_weight=mpc.soft_weight, **over)
    lo, hi = util.theta_range(name, 1.0)
    Th = mpqp.sample_theta(lo, hi, 100 if name == "chain" else 1500, seed=17)
    ref = P.solve_batch(Th)
    out = g.solve(Th)
    assert np.array_equal(out["exitflag"], ref["flag"])
    assert np.array_equal(out["iters"], ref["iters"])
    done = ref["flag"] >= 1
    assert np.array_equal(out["active_upper"][done], ref["act_u"][done])
    assert np.array_equal(out["active_lower"][done], ref["act_l"][done])
    assert np.array_equal(np.asarray(out["U"])[:, done], ref["U"][:, done])
    if "iter_limit" in over and over["iter_limit"] == 4 and name != "dint_hard":
        assert np.any(ref["flag"] == -4)
    g.close()


@pytest.mark.parametrize("name,T", [("invpend", 50), ("aircraft", 30), ("aircraft_ksall", 20), ("chain_small", 10)])
@pytest.mark.parametrize("mode", ["plant", "observer", "observer+noise+plant"])
def test_closed_loop_with_true_plant_and_observer(lmpc, name, T, mode):
    """lmpc_gpu_simulate_batch_ex: a true plant that differs from the prediction model, the
    steady-state Kalman observer of observer.jl:104-118 in the loop (simulation.jl:92-110) and a
    measurement-noise trajectory; bit-identical to the oracle's restatement on all three closed-loop
    kernels, and the no-mismatch / no-noise observer reproduces the plain closed loop."""
    mpc, qp, g, P = gpu_controller(lmpc, name)
    nx, nr, nd, nup = mpc.parameter_dims()
    rng = np.random.default_rng(9)
    N = 257
    lo, hi = util.theta_range(name, 0.1)
    X0 = rng.uniform(lo[:nx], hi[:nx], (N, nx)).T
    r = np.zeros(nr)
    d = 0.01 * np.ones(nd) if nd else None
    Cm = mpc.C
    ny = Cm.shape[0]
    K = 0.4 * np.linalg.pinv(Cm)
    Dd = mpc.Dd if nd else None
    plant = dict(F=mpc.F * (1 + 1e-3), G=mpc.G * 0.98, Gd=mpc.Gd if nd else None)
    V = 1e-3 * rng.standard_normal((ny, T - 3))
    kw, ext = {}, {}
    if "plant" in mode:
        kw["plant"] = plant
        ext.update(Ft=plant["F"], Gt=plant["G"], Gdt=plant["Gd"])
    if "observer" in mode:
        kw["observer"] = dict(C=Cm, K=K, Dd=Dd)
        ext.update(C=Cm, K=K, Dd=Dd)
    if "noise" in mode:
        kw["noise"] = V
        ext["Vtraj"] = np.hstack([V, np.repeat(V[:, -1:], 3, axis=1)])
    Rm = np.repeat(r.reshape(-1, 1), N, axis=1) if nr else None
    Dm = np.repeat(d.reshape(-1, 1), N, axis=1) if nd else None
    for warm in (True, False):
        out = g.simulate(mpc.F, mpc.G, X0, r=Rm, d=Dm, N=T, Gd=mpc.Gd if nd else None, warm_start=warm,
                         want_traj=True, **kw)
        ref = P.simulate_batch(mpc.F, mpc.G, mpc.Gd if nd else None, X0, Rm, Dm, None, T, nr, nd, nup, warm=warm,
                               want_traj=True, ext=ext)
        assert np.array_equal(out["exitflag"], ref["flag"])
        assert np.array_equal(out["iters"], ref["iters"])
        assert np.array_equal(out["Utraj"], ref["Utraj"])
        assert np.array_equal(out["Xend"], ref["Xend"])
        assert np.array_equal(out["Xhat_end"], ref["Xhat_end"])
    if mode == "observer":
        # exact model, no noise: the estimate is the state and the loop is the plain one
        plain = g.simulate(mpc.F, mpc.G, X0, r=Rm, d=Dm, N=T, Gd=mpc.Gd if nd else None, warm_start=False,
                           want_traj=True)
        assert np.array_equal(plain["Utraj"], out["Utraj"]) and np.array_equal(plain["Xend"], out["Xend"])
        assert np.array_equal(out["Xend"], out["Xhat_end"])
    else:
        assert not np.array_equal(out["Xend"], out["Xhat_end"])


@pytest.mark.parametrize("name", ["aircraft", "aircraft_ksall"])
def test_closed_loop_with_preview_windows(lmpc, name):
    """Sliding preview windows (reference_preview / disturbance_preview slicing, simulation.jl:102-104,
    128-134): the AFTI-16 parameter layout read as a one-output reference and a one-channel
    disturbance previewed over Np = 2 steps.  Step k must see r columns k+1, k+2 and d columns k, k+1
    of the scenario (last column held), and only d_k drives the plant."""
    mpc, qp, g, P = gpu_controller(lmpc, name)
    nx, nr, nd, nup = mpc.parameter_dims()
    assert nr == 2 and nd == 2
    T, N = 25, 200
    rng = np.random.default_rng(12)
    lo, hi = util.theta_range(name, 0.1)
    X0 = rng.uniform(lo[:nx], hi[:nx], (N, nx)).T
    rs = 0.05 * np.cumsum(rng.standard_normal((1, 18)), axis=1) * 0.2     # shorter than the horizon
    ds = 0.02 * rng.standard_normal((1, 30))
    Gd1 = mpc.Gd[:, :1]
    out = g.simulate(mpc.F, mpc.G, X0, N=T, Gd=Gd1, warm_start=True, want_traj=True,
                     r_traj=rs, d_traj=ds, r_preview=2, d_preview=2)
    # the same loop spelled out for the oracle: explicit windows, zero-padded Gd
    hold = lambda v, n: np.hstack([v, np.repeat(v[:, -1:], max(0, n - v.shape[1]), axis=1)])
    # the reference holds column N of its N-column copy (simulation.jl:57-80, 128-134): scenario columns
    # beyond the simulated horizon are never previewed
    rsx, dsx = hold(rs[:, :T], T + 3), hold(ds[:, :T], T + 3)
    Gdp = np.hstack([Gd1, np.zeros((nx, 1))])
    # stepwise restatement with the oracle's single solves (tolerance: BLAS vs fma-chain dynamics)
    for q in range(0, N, 40):
        x, u = X0[:, q].copy(), np.zeros(mpc.nu)
        for k in range(T):
            th = np.concatenate([x, rsx[0, k + 1:k + 3], dsx[0, k:k + 2], u[:nup]])
            u, flag, _ = P.solve(th)
            assert np.allclose(u, out["Utraj"][q, k], rtol=0, atol=2e-5)
            x = mpc.F @ x + mpc.G @ u + Gd1[:, 0] * dsx[0, k]
        assert np.allclose(x, out["Xend"][:, q], rtol=0, atol=2e-5)
    # bit-exact against the oracle's own windowed loop
    Rflat = np.asfortranarray(rsx[:, 1:T + 2]); Dflat = np.asfortranarray(dsx[:, 0:T + 1])
    ref = P.simulate_batch(mpc.F, mpc.G, Gdp, X0, None, None, None, T, nr, nd, nup, warm=True, want_traj=True,
                           Rtraj=Rflat, Dtraj=Dflat, ext=dict(rstride=1, dstride=1))
    assert np.array_equal(out["Utraj"], ref["Utraj"]) and np.array_equal(out["Xend"], ref["Xend"])
    assert np.array_equal(out["exitflag"], ref["flag"])


@pytest.mark.parametrize("name,blocks", [("aircraft", [[1, 2, 7], [3, 7]]), ("invpend", [[2, 3, 5, 40]]),
                                         ("aircraft_ksall", [[1, 1, 1, 1, 6], [1, 1, 1, 1, 6]])])
def test_move_blocked_controllers_on_gpu(lmpc, name, blocks):
    """A controller condensed WITH move blocking by the product's builder (C++), solved on the GPU,
    against the oracle on the numpy restatement of the same blocked problem."""
    import test_condense_cpu as tc
    o = util.controller(name)
    o.Nc = o.Np
    mpqp.move_block(o, blocks)
    qp = mpqp.mpc2mpqp(o)
    m = tc.mirror_of(lmpc, util.controller(name))
    m.Nc = m.Np
    m.move_block(blocks)
    g = m.setup()
    cst = g.ldp_constants()
    cst["senses"] = g.senses
    P = util.oracle_from_constants(qp, o.nu, cst, soft_weight=o.soft_weight)
    lo, hi = util.theta_range(name, 0.25)
    Th = mpqp.sample_theta(lo, hi, 3000, seed=31)
    assert_parity(g.solve(Th), P.solve_batch(Th))
    assert g.n == sum(len(b) for b in o.move_blocks)


@pytest.mark.parametrize("name", ["invpend", "aircraft"])
def test_prestabilising_feedback_in_the_output_map(lmpc, name):
    """lmpc_problem.K: u = z[1:nu] - K x (src/utils.jl:46-50), folded into Uth_offset exactly as qp2ldp
    does (src/codegen.jl:141,269-273).  The reference's condensing with K is out of scope, so K is applied to
    an ordinary mpQP here: the GPU equals the oracle built with the same K bit for bit, and both equal the
    K = 0 result minus K x."""
    mpc = util.controller(name)
    qp = mpqp.mpc2mpqp(mpc)
    nx = mpc.parameter_dims()[0]
    K = np.random.default_rng(4).standard_normal((mpc.nu, nx)) * 0.1
    g = lmpc.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight, K=K)
    g0 = lmpc.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight)
    cst = g.ldp_constants()
    cst["senses"] = g.senses
    L = mpqp.qp2ldp(qp, mpc.nu, K=K, soft_weight=mpc.soft_weight)
    # C++ setup vs numpy: equal to rounding amplified by cond(H) (3.5e8 for AFTI-16)
    assert np.max(np.abs(cst["Uth"] - L.Uth_offset)) <= 1e-6 * max(1.0, np.abs(L.Uth_offset).max())
    P = util.oracle_from_constants(qp, mpc.nu, cst, soft_weight=mpc.soft_weight)
    lo, hi = util.theta_range(name, 0.25)
    Th = mpqp.sample_theta(lo, hi, 3000, seed=8)
    out = g.solve(Th)
    assert_parity(out, P.solve_batch(Th))
    base = g0.solve(Th)
    ok = out["exitflag"] >= 1
    assert np.max(np.abs(out["U"][:, ok] - (base["U"][:, ok] - K @ Th[:nx][:, ok]))) < 1e-10


def test_async_entry_points_on_caller_streams(lmpc):
    """lmpc_gpu_solve_batch_async / lmpc_gpu_evaluate_batch_async: device pointers, the caller's
    stream, no synchronisation inside; two streams in flight at once; the handle's own stream when
    the stream argument is NULL.  Results equal the synchronous host-buffer path bit for bit."""
    import torch
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    dev = torch.device("cuda", 0)
    lo, hi = mpqp.invpend_range(1.0)
    N = 200_000
    Ths = [mpqp.sample_theta(lo, hi, N, seed=60 + i) for i in range(2)]
    Td = [torch.tensor(np.ascontiguousarray(T.T), dtype=torch.float64, device=dev) for T in Ths]
    U = [torch.empty((N, 1), dtype=torch.float64, device=dev) for _ in range(2)]
    fl = [torch.empty(N, dtype=torch.int32, device=dev) for _ in range(2)]
    it = [torch.empty(N, dtype=torch.int32, device=dev) for _ in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    for i in range(2):
        g.solve_async(Td[i], U[i], fl[i], it[i], N=N, stream=streams[i].cuda_stream)
    for s in streams:
        s.synchronize()
    for i in range(2):
        ref = g.solve(Ths[i], want_active=False)
        assert np.array_equal(U[i].cpu().numpy().T, ref["U"]) and np.array_equal(fl[i].cpu().numpy(), ref["exitflag"])
        assert np.array_equal(it[i].cpu().numpy(), ref["iters"])
    # NULL stream = the handle's own stream
    g.solve_async(Td[0], U[1], fl[1], N=N, stream=None)
    torch.cuda.current_stream(dev).synchronize()
    torch.cuda.synchronize()
    assert torch.equal(U[0], U[1]) and g.stream() is not None
    # explicit evaluate, async
    z = np.load(os.path.join(GOLD, "invpend_explicit.npz"))
    ge = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"])
    Te = mpqp.sample_theta(z["lo"], z["hi"], N, seed=3)
    Ted = torch.tensor(np.ascontiguousarray(Te.T), dtype=torch.float64, device=dev)
    Ue = torch.empty((N, 1), dtype=torch.float64, device=dev)
    re = torch.empty(N, dtype=torch.int32, device=dev)
    ge.evaluate_async(Ted, Ue, re, N=N, stream=streams[0].cuda_stream)
    streams[0].synchronize()
    Uh, rh = ge.evaluate(Te)
    assert np.array_equal(Ue.cpu().numpy().T, Uh) and np.array_equal(re.cpu().numpy(), rh)


def test_api_errors_are_codes_not_crashes(lmpc):
    """Error convention of the ABI: bad arguments come back as negative LMPC_E_* return
    codes with a message, never as exceptions across the ABI or device faults."""
    import ctypes as C
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    L = lmpc.lib()
    Th = np.asfortranarray(mpqp.sample_theta(*mpqp.invpend_range(1.0), 16))
    U = np.zeros((1, 16), order="F"); fl = np.zeros(16, np.int32)
    assert L.lmpc_gpu_solve_batch(g._h, None, 16, U.ctypes.data, fl.ctypes.data, None, None, None) == -100
    assert L.lmpc_gpu_solve_batch(g._h, Th.ctypes.data, -1, U.ctypes.data, fl.ctypes.data, None, None, None) == -100
    assert b"bad argument" in L.lmpc_gpu_last_error()
    au = np.zeros((16, 1), np.uint64)
    # active-set words must come in pairs / be present for a warm start
    assert L.lmpc_gpu_solve_batch(g._h, Th.ctypes.data, 16, U.ctypes.data, fl.ctypes.data, None, au.ctypes.data, None) == -100
    assert L.lmpc_gpu_solve_batch_warm(g._h, Th.ctypes.data, 16, U.ctypes.data, fl.ctypes.data, None, None, None) == -100
    # closed loop: the parameter split must add up to nth
    with pytest.raises(lmpc.LmpcError) as e:
        bad = lmpc.setup_gpu(qp, nu=1, dims=(4, 1, 0, 1))
        bad.simulate(mpc.F, mpc.G, Th[:4], N=3)
    assert e.value.code == -100
    # explicit trees are validated on creation: a jump that leaves the node array, a leaf naming a missing law
    hp = np.array([[1.0, 0.0, 0.5]]); fb = np.zeros((1, 1, 3))
    for hl, jl in (([0, 0, 0], [5, 0, 0]), ([0, 0, 7], [1, 0, 0]), ([3, 0, 0], [1, 0, 0])):
        with pytest.raises(lmpc.LmpcError) as e:
            lmpc.GpuExplicitMPC(hp, fb, np.array(hl, np.int32), np.array(jl, np.int32))
        assert e.value.code == -100
    ok = lmpc.GpuExplicitMPC(hp, fb, np.array([0, 0, 0], np.int32), np.array([1, 0, 0], np.int32), dims=(2, 0, 0, 0))
    u, reg = ok.evaluate(np.zeros((2, 5)))
    assert u.shape == (1, 5) and np.all(reg == 0)
    # explicit closed loop limits
    with pytest.raises((lmpc.LmpcError, ValueError)):
        ok.simulate(np.eye(3), np.ones((3, 1)), np.zeros((3, 2)), N=2)      # nx does not match nth
    Fm, Gm, X0 = np.asfortranarray(np.eye(2)), np.asfortranarray(np.ones((2, 1))), np.zeros((2, 3), order="F")
    Xe, Ul = np.zeros((2, 3), order="F"), np.zeros((1, 3), order="F")
    rc = L.lmpc_gpu_simulate_bst_batch(ok._t, Fm.ctypes.data, Gm.ctypes.data, None, 1, 0, 0, 0, X0.ctypes.data, None, None,
                                       None, None, None, 3, 2, Xe.ctypes.data, Ul.ctypes.data, None)
    assert rc == -100 and b"must equal nth" in L.lmpc_gpu_last_error()


# ---- controllers condensed with reference / disturbance previews (types.jl:56-57) ---------------------

def _preview_controller(lmpc, rp, dp, Np=10, Nc=4):
    """AFTI-16 with Np-step previews, built with the product's builder (C++ condenser) and the
    oracle handed the constants the product computed."""
    from test_condense_cpu import mirror_of
    o = mpqp.aircraft(Np=Np, Nc=Nc, ks=(2, 3))
    o.Gd = np.array([[0.1, 0.0], [0.0, 0.02], [0.03, 0.0], [0.0, 0.01]])
    o.reference_preview, o.disturbance_preview = rp, dp
    m = mirror_of(lmpc, o)
    m.reference_preview, m.disturbance_preview = rp, dp
    g = m.setup()
    qp = mpqp.mpc2mpqp(o)
    cst = g.ldp_constants()
    cst["senses"] = g.senses
    P = util.oracle_from_constants(qp, o.nu, cst, soft_weight=o.soft_weight)
    return o, m, g, P


@pytest.mark.gpu
@pytest.mark.parametrize("rp,dp", [(True, False), (False, True), (True, True)])
def test_preview_controller_batch_parity(lmpc, rp, dp):
    """theta = [x; vec(r_1..r_Np); vec(d_0..d_Np-1); uprev] with time-varying previews: the GPU solve
    of the preview-condensed controller against the oracle, instance by instance."""
    o, m, g, P = _preview_controller(lmpc, rp, dp)
    nx, nr, nd, nup = o.parameter_dims()
    assert g.dims == (nx, nr, nd, nup) and nr == (2 * o.Np if rp else 2) and nd == (2 * o.Np if dp else 2)
    N = 3000
    rng = np.random.default_rng(21)
    X = rng.uniform(-1, 1, (nx, N)) * np.array([4, 0.3, 1, 0.3])[:, None]
    Rm = 0.3 * rng.standard_normal((nr, N)) * np.tile([1.0, 0.05], nr // 2)[:, None]
    Dm = 0.2 * rng.standard_normal((nd, N))
    Um = rng.uniform(-0.5, 0.5, (nup, N))
    Th = np.asfortranarray(np.vstack([X, Rm, Dm, Um]))
    out, ref = g.solve(Th), P.solve_batch(Th)
    assert_parity(out, ref)
    assert len(np.unique(ref["act_u"], axis=0)) > 3          # the batch does exercise the constraints
    # the builder's compute_control flattens a per-instance preview the same way
    if rp and dp:
        U = m.compute_control(X[:, :50], r=Rm[:, :50].reshape(2, o.Np, 50, order="F"),
                              d=Dm[:, :50].reshape(2, o.Np, 50, order="F"), uprev=Um[:, :50])
        assert np.array_equal(U, np.asarray(out["U"])[:, :50])


@pytest.mark.gpu
def test_preview_controller_closed_loop_through_builder(lmpc):
    """MPC.simulate on a preview-condensed controller slides the Np-column windows itself
    (simulation.jl:102-104, 128-147): the same loop stepped with compute_control gives the same u."""
    o, m, g, P = _preview_controller(lmpc, True, True)
    T, N, Np = 15, 64, o.Np
    rng = np.random.default_rng(5)
    X0 = rng.uniform(-1, 1, (o.nx, N)) * np.array([4, 0.3, 1, 0.3])[:, None]
    rs = np.vstack([0.2 * np.sin(0.3 * np.arange(12)), np.zeros(12)])     # shorter than T + Np: last held
    ds = 0.2 * rng.standard_normal((2, 40))
    out = m.simulate(X0, N=T, r_traj=rs, d_traj=ds, want_traj=True, warm_start=False)
    hold = lambda v, n: np.hstack([v, np.repeat(v[:, -1:], max(0, n - v.shape[1]), axis=1)])
    rsx, dsx = hold(rs[:, :T], T + Np + 1), hold(ds[:, :T], T + Np + 1)    # clamp at column N, as get_preview does
    for q in range(0, N, 16):
        x, u = X0[:, q].copy(), np.zeros(o.nu)
        for k in range(T):
            th = o.form_parameter(x, rsx[:, k + 1:k + 1 + Np], dsx[:, k:k + Np], u)
            u, flag, _ = P.solve(th)
            assert flag >= 1
            assert np.allclose(u, out["Utraj"][q, k], rtol=0, atol=2e-5)
            x = o.F @ x + o.G @ u + o.Gd @ dsx[:, k]
        assert np.allclose(x, out["Xend"][:, q], rtol=0, atol=2e-5)


# ---- prestabilising feedback condensed by the builder (setup.jl:186-199) ------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("name", ["invpend", "aircraft"])
def test_prestabilised_controller_parity_and_equivalence(lmpc, name):
    """u = v - K x with the QP over v.  (i) GPU == oracle on the prestabilised controller (bounds are
    general rows there, ms = 0); (ii) with Nc = Np it is a change of variables of the plain controller,
    so both give the same control (the reference asserts this kind of agreement at 1e-6 .. 1e-10)."""
    from test_condense_cpu import mirror_of
    if name == "invpend":
        o = mpqp.invpend(Np=10, Nc=10)
        lo, hi = util.theta_range("invpend", 0.25)
    else:
        o = mpqp.aircraft(Np=6, Nc=6, ks=(2, 3))
        lo, hi = util.theta_range("aircraft", 0.1)
    plain = mirror_of(lmpc, o)
    gp = plain.setup()
    o.set_prestabilizing_feedback()
    m = mirror_of(lmpc, o)
    m.set_prestabilizing_feedback()
    g = m.setup()
    qp = mpqp.mpc2mpqp(o)
    assert qp.A.shape[0] == len(qp.bu) and g.dims == o.parameter_dims()
    cst = g.ldp_constants()
    cst["senses"] = g.senses
    L = mpqp.qp2ldp(qp, o.nu, K=o.K, soft_weight=o.soft_weight)
    P = util.oracle_from_constants(qp, o.nu, cst, soft_weight=o.soft_weight)
    assert np.allclose(L.Uth_offset, cst["Uth"], rtol=1e-9, atol=1e-11)     # K is in the output map
    N = 4000
    Th = np.asfortranarray(mpqp.sample_theta(lo, hi, N, seed=33))
    out, ref = g.solve(Th), P.solve_batch(Th)
    assert_parity(out, ref)
    base = gp.solve(Th)
    ok = (out["exitflag"] == 1) & (base["exitflag"] == 1)
    assert ok.sum() > 500
    Ua, Ub = np.asarray(out["U"])[:, ok], np.asarray(base["U"])[:, ok]
    assert np.max(np.abs(Ua - Ub)) <= 1e-6 * max(1.0, np.abs(Ub).max())
    assert len(np.unique(base["active_upper"][ok], axis=0)) > 3


@pytest.mark.gpu
def test_reference_case_disturbance_preview_simulation(lmpc):
    """test/runtests.jl:382-408 through the builder on the GPU: a 20-step closed loop with d stepping
    0 -> 1 after 8 steps.  With the preview the controls differ (norm > 1e-2) and the output energy
    drops below 0.9 of the no-preview loop; the GPU loop equals the oracle's stepwise loop."""
    from test_condense_cpu import mirror_of
    from test_preview_cpu import _dist_preview_case
    o = _dist_preview_case()
    d_traj = np.hstack([np.zeros((1, 8)), np.ones((1, 12))])
    T = 20
    def ys_of(U):
        x, ys = np.zeros(2), []
        for k in range(T):
            ys.append(o.C @ x)
            x = o.F @ x + o.G @ U[k] + o.Gd @ d_traj[:, k]
        return np.array(ys)
    m0 = mirror_of(lmpc, o)
    U0 = m0.simulate(np.zeros((2, 1)), N=T, d_traj=d_traj, want_traj=True, warm_start=False)["Utraj"][0]
    o.disturbance_preview = True
    m1 = mirror_of(lmpc, o)
    m1.disturbance_preview = True
    out = m1.simulate(np.zeros((2, 1)), N=T, d_traj=d_traj, want_traj=True, warm_start=False)
    U1 = out["Utraj"][0]
    assert m1.gpu.dims == (2, 1, 5, 0) and np.all(out["exitflag"] >= 1)
    assert np.linalg.norm(U1 - U0) > 1e-2
    assert np.linalg.norm(ys_of(U1)) / np.linalg.norm(ys_of(U0)) < 0.9
    # the oracle's loop, step by step, on the same controller
    qp = mpqp.mpc2mpqp(o)
    P = ldp.OracleProblem(mpqp.qp2ldp(qp, o.nu, soft_weight=o.soft_weight), o.nu)
    dsx = np.hstack([d_traj, np.repeat(d_traj[:, -1:], o.Np + 1, axis=1)])
    x = np.zeros(2)
    for k in range(T):
        u, flag, _ = P.solve(o.form_parameter(x, None, dsx[:, k:k + o.Np], None))
        assert flag >= 1 and np.allclose(u, U1[k], rtol=0, atol=1e-9)
        x = o.F @ x + o.G @ u + o.Gd @ dsx[:, k]
