"""Round-2 entry points through the C ABI on a GPU: broadcast (Monte-Carlo) compute_control with a
device-side failure count, the oriented tree import, the single-process multi-device closed loop and
explicit evaluate, and the result window (the final gather fused into the solve kernel's stores)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from oracle import mpqp, ldp
from test_gpu_parity import gpu_controller

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "invpend_explicit.npz")


@pytest.mark.parametrize("name", ["invpend", "aircraft"])
def test_broadcast_blocks_equal_the_full_theta_solve(lmpc, name):
    """compute_control(mpc, x; r, d, uprev) over many x with the other arguments fixed: the shared blocks
    travel once (stride-0 segments); results are the same bits as the solve on the assembled theta, for
    host and device x, with per-instance and with shared r; flags may stay on the device (count only)."""
    import torch
    mpc, qp, g, P = gpu_controller(lmpc, name)
    nx, nr, nd, nup = mpc.parameter_dims()
    lo, hi = util.theta_range(name, 0.5)
    N = 50_000
    Th = mpqp.sample_theta(lo, hi, N, seed=7)
    r, d, up = Th[nx:nx + nr, 0].copy(), Th[nx + nr:nx + nr + nd, 0].copy(), Th[nx + nr + nd:, 0].copy()
    Th[nx:] = Th[nx:, :1]
    ref = g.solve(Th)
    U, flag, nfail = g.compute_control_bcast(Th[:nx], r=r, d=d if nd else None, uprev=up, want_flags=True)
    assert np.array_equal(U, ref["U"]) and np.array_equal(flag, ref["exitflag"])
    assert nfail == int(np.count_nonzero(ref["exitflag"] < 1))
    U2, flag2, nfail2 = g.compute_control_bcast(Th[:nx], r=r, d=d if nd else None, uprev=up)     # no flag array
    assert flag2 is None and np.array_equal(U2, ref["U"]) and nfail2 == nfail
    Xd = torch.as_tensor(np.ascontiguousarray(Th[:nx].T), device="cuda")
    U3, _, nfail3 = g.compute_control_bcast(Xd, r=r, d=d if nd else None, uprev=up)
    assert np.array_equal(U3.cpu().numpy(), ref["U"]) and nfail3 == nfail
    # per-instance reference, shared d / uprev
    Th2 = mpqp.sample_theta(lo, hi, N, seed=8)
    Th2[nx + nr:] = Th2[nx + nr:, :1]
    ref2 = g.solve(Th2)
    U4, _, _ = g.compute_control_bcast(Th2[:nx], r=Th2[nx:nx + nr], d=Th2[nx + nr:nx + nr + nd, 0] if nd else None,
                                       uprev=Th2[nx + nr + nd:, 0])
    assert np.array_equal(U4, ref2["U"])


def test_failure_count_sees_infeasible_instances(lmpc):
    mpc, qp, g, P = gpu_controller(lmpc, "simple")
    nx, nr, nd, nup = mpc.parameter_dims()
    lo, hi = util.theta_range("simple", 1.0)
    Th = mpqp.sample_theta(lo, hi, 20000, seed=5)
    Th[nx:] = Th[nx:, :1]
    ref = g.solve(Th)
    bad = int(np.count_nonzero(ref["exitflag"] < 1))
    assert bad > 100
    _, _, nfail = g.compute_control_bcast(Th[:nx], r=Th[nx:nx + nr, 0], uprev=Th[nx + nr:, 0])
    assert nfail == bad


def _swap_children(hl, jl):
    """the same tree in the other orientation: children of every internal node exchanged"""
    hl2, jl2 = hl.copy(), jl.copy()
    for i in range(len(jl)):
        if jl[i] == 0:
            continue
        c0, c1 = i + jl[i], i + jl[i] + 1
        hl2[c0], jl2[c0] = hl[c1], (jl[c1] + 1 if jl[c1] else 0)
        hl2[c1], jl2[c1] = hl[c0], (jl[c0] - 1 if jl[c0] else 0)
    return hl2, jl2


def test_oriented_tree_import(lmpc):
    """A tree whose producer takes next + 1 on a'theta <= b (believed to be ParametricDAQP's order) loads
    through le_child = 1 and evaluates exactly like the library-order tree; loaded with the WRONG
    orientation it does not (so the flag is not a no-op)."""
    z = np.load(GOLD)
    hl2, jl2 = _swap_children(z["hp_list"], z["jump_list"])
    # the swapped tree walked by a literal transcription of `id = next_id + (val <= b)` gives the same leaves
    Th = mpqp.sample_theta(z["lo"], z["hi"], 4000, seed=11)
    B = ldp.OracleBst(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clip=z["clipping"])
    Uo, rego = B.evaluate_batch(Th)
    for q in range(0, 4000, 97):
        i = 0
        while jl2[i] != 0:
            hp = z["halfplanes"][hl2[i]]
            i = i + jl2[i] + int(hp[:-1] @ Th[:, q] <= hp[-1])
        assert hl2[i] == rego[q]
    a = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"], dims=(4, 2, 0, 1))
    b = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], hl2, jl2, clipping=z["clipping"], dims=(4, 2, 0, 1), le_child=1)
    c = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], hl2, jl2, clipping=z["clipping"], dims=(4, 2, 0, 1), le_child=0)
    Ua, ra = a.evaluate(Th)
    Ub, rb = b.evaluate(Th)
    Uc, rc = c.evaluate(Th)
    assert np.array_equal(Ua, Uo) and np.array_equal(Ub, Uo) and np.array_equal(rb, rego)
    assert np.count_nonzero(rc != rego) > 1000


def test_single_process_multi_device_closed_loop_and_explicit(lmpc):
    """lmpc_gpu_simulate_batch_multi / lmpc_gpu_evaluate_batch_multi: same results as the one-device calls
    (device 0 listed twice on a 1-GPU box; real pairs are covered by the test below)."""
    mpc, qp, g, P = gpu_controller(lmpc, "aircraft")
    nx, nr, nd, nup = mpc.parameter_dims()
    ndev = lmpc.device_count()
    devs = [0, 1 % ndev] if ndev >= 2 else [0, 0]
    mg = lmpc.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), soft_weight=mpc.soft_weight, devices=devs)
    lo, hi = util.theta_range("aircraft", 0.1)
    X0 = mpqp.sample_theta(lo[:nx], hi[:nx], 3001, seed=2)
    one = g.simulate(mpc.F, mpc.G, X0, r=[0.1, 0.0], d=[0.0, 0.0], N=20, Gd=mpc.Gd)
    two = mg.simulate(mpc.F, mpc.G, X0, r=[0.1, 0.0], d=[0.0, 0.0], N=20, Gd=mpc.Gd)
    assert np.array_equal(one["Xend"], two["Xend"]) and np.array_equal(one["exitflag"], two["exitflag"])
    assert np.array_equal(one["Ulast"], two["Ulast"])
    z = np.load(GOLD)
    e1 = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"], dims=(4, 2, 0, 1))
    e2 = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"], dims=(4, 2, 0, 1),
                             devices=devs)
    Th = mpqp.sample_theta(z["lo"], z["hi"], 70001, seed=3)
    U1, r1 = e1.evaluate(Th)
    U2, r2 = e2.evaluate(Th)
    assert np.array_equal(U1, U2) and np.array_equal(r1, r2)


def test_solve_batch_multi_on_two_real_devices(lmpc):
    if lmpc.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    mg = lmpc.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), devices=[0, 1])
    lo, hi = util.theta_range("invpend", 1.0)
    Th = mpqp.sample_theta(lo, hi, 300_001, seed=9)
    a, b = g.solve(Th), mg.solve(Th)
    for k in ("U", "exitflag", "iters", "active_upper", "active_lower"):
        assert np.array_equal(a[k], b[k]), k


_CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import lmpc_b200, util
from oracle import mpqp
dev, off, N = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
win = lmpc_b200.ResultWindow(dev, handle=bytes.fromhex(sys.argv[1]))
mpc = util.controller("invpend"); qp = mpqp.mpc2mpqp(mpc)
g = lmpc_b200.setup_gpu(qp, nu=1, dims=mpc.parameter_dims(), device=dev)
lo, hi = util.theta_range("invpend", 1.0)
Th = torch.as_tensor(np.ascontiguousarray(mpqp.sample_theta(lo, hi, N, seed=100 + off).T), device=f"cuda:{{dev}}")
ntot = int(sys.argv[5])
# window layout: U (8 B x ntot) then flags (4 B x ntot); this process owns columns [off, off + N)
lmpc_b200.lib().lmpc_gpu_solve_batch_async(g._h, Th.data_ptr(), N, win.ptr + 8 * off, win.ptr + 8 * ntot + 4 * off,
                                           None, None, None, None)
torch.cuda.synchronize(dev)
win.close()
"""


def test_result_window_gathers_across_processes(lmpc, tmp_path):
    """the gather fused into the solve: a second PROCESS maps rank 0's window through CUDA IPC and its
    kernel stores u / flags straight into the gathered array (its own GPU when the box has two)."""
    import torch
    ndev = lmpc.device_count()
    N, ntot = 40_000, 80_000
    win = lmpc.ResultWindow(0, nbytes=12 * ntot)
    mpc, qp, g, P = gpu_controller(lmpc, "invpend")
    lo, hi = util.theta_range("invpend", 1.0)
    Th0 = mpqp.sample_theta(lo, hi, N, seed=100)
    Thd = torch.as_tensor(np.ascontiguousarray(Th0.T), device="cuda:0")
    lmpc.lib().lmpc_gpu_solve_batch_async(g._h, Thd.data_ptr(), N, win.ptr, win.ptr + 8 * ntot, None, None, None, None)
    torch.cuda.synchronize()
    script = tmp_path / "child.py"
    script.write_text(_CHILD.format(root=ROOT))
    subprocess.check_call([sys.executable, str(script), win.handle_bytes().hex(), str(1 if ndev >= 2 else 0),
                           str(N), str(N), str(ntot)])
    Uw, Fw = win.read(0, ntot, np.float64), win.read(8 * ntot, ntot, np.int32)
    Th1 = mpqp.sample_theta(lo, hi, N, seed=100 + N)
    ref0, ref1 = P.solve_batch(Th0), P.solve_batch(Th1)
    assert np.array_equal(Uw[:N], ref0["U"][0]) and np.array_equal(Uw[N:], ref1["U"][0])
    assert np.array_equal(Fw[:N], ref0["flag"]) and np.array_equal(Fw[N:], ref1["flag"])
    win.close()


def test_chain_two_pass_cta_path(lmpc):
    """chain of masses (n = 180): first pass with two CTAs per SM at a reduced working-set capacity, second
    pass at full capacity for the instances that outgrew it.  Results are those of the full capacity, bit
    for bit the oracle's (z cache + cooperative backward sweep included), and the second pass really ran."""
    mpc, qp, g, P = gpu_controller(lmpc, "chain")
    fast_cap, before = g.tiering()
    assert g.kernel == lmpc.KERNEL_CTA and 0.8 * g.cap <= fast_cap < g.cap
    lo, hi = util.theta_range("chain", 1.4)            # beyond the nominal range: working sets up to ~175 rows
    Th = mpqp.sample_theta(lo, hi, 800, seed=5)
    out, ref = g.solve(Th), P.solve_batch(Th)
    for a, b in (("exitflag", "flag"), ("iters", "iters"), ("active_upper", "act_u"), ("active_lower", "act_l")):
        assert np.array_equal(out[a], ref[b]), a
    assert np.array_equal(out["U"], ref["U"]) and np.all(ref["flag"] == 1)
    k = (util.bits(ref["act_u"], P.m) | util.bits(ref["act_l"], P.m)).sum(1)
    over = int(np.count_nonzero(k > fast_cap))          # adds are never undone on this problem class
    assert over >= 1 and g.tiering()[1] - before >= over
    # small theta: tiny working sets, everything ends in the first pass
    lo, hi = util.theta_range("chain", 0.25)
    Th = mpqp.sample_theta(lo, hi, 1500, seed=6)
    mid = g.tiering()[1]
    out, ref = g.solve(Th), P.solve_batch(Th)
    assert np.array_equal(out["U"], ref["U"]) and np.array_equal(out["iters"], ref["iters"])
    assert g.tiering()[1] == mid


def test_certified_iteration_bound_sizes_the_kernel(lmpc):
    """(f-4) certify(mpc; range).max_iterations (src/certify.jl:18-29) as `certified_iterations`: a hard
    iteration cap and a smaller on-chip capacity.  The chain of masses restricted to a quarter of its range
    never needs more than ~56 iterations, so with a bound of 60 its factor shrinks from 180 to 60 rows: eight
    QPs per CTA in warp mode instead of one per CTA.  Same bits as the oracle given the same cap / limit; an
    instance outside the certified range ends with the iteration-limit flag instead of a wrong answer."""
    mpc = util.controller("chain")
    qp = mpqp.mpc2mpqp(mpc)
    st = lmpc.default_settings(certified_iterations=60)
    g = lmpc.setup_gpu(qp, nu=mpc.nu, dims=mpc.parameter_dims(), settings=st)
    assert g.cap == 60 and g.tiering()[0] == 0
    cst = g.ldp_constants()
    P = util.oracle_from_constants(qp, mpc.nu, cst, iter_limit=61)
    lo, hi = util.theta_range("chain", 0.25)
    Th = mpqp.sample_theta(lo, hi, 3000, seed=8)
    out, ref = g.solve(Th), P.solve_batch(Th)
    assert np.all(ref["flag"] == 1) and ref["iters"].max() <= 60
    for a, b in (("exitflag", "flag"), ("iters", "iters"), ("active_upper", "act_u"), ("active_lower", "act_l")):
        assert np.array_equal(out[a], ref[b]), a
    assert np.array_equal(out["U"], ref["U"])
    lo, hi = util.theta_range("chain", 1.0)                  # outside the certificate's range
    Th = mpqp.sample_theta(lo, hi, 200, seed=8)
    out, ref = g.solve(Th), P.solve_batch(Th)
    assert np.array_equal(out["exitflag"], ref["flag"]) and np.all(out["exitflag"] == -4)


def test_chain_full_step_size_properties(lmpc):
    """BASELINE config 4 at the size of one bench step (1e5 chain-of-masses QPs, n = 180, s = 1; the config's 1e6 are ten
    such steps): every instance optimal, the applied control within its bounds, both passes of the two-pass CTA scheme
    used, idempotent, device path == chunked host path, and a 1000-instance subsample equal to the oracle bit for bit
    (exit flag, iteration count, active set, u)."""
    import torch
    mpc, qp, g, P = gpu_controller(lmpc, "chain")
    lo, hi = util.theta_range("chain", 1.0)
    N = 100_000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4321)
    Th = torch.rand((N, len(lo)), dtype=torch.float64, device="cuda", generator=gen)
    Th = Th * torch.tensor(hi - lo, device="cuda") + torch.tensor(lo, device="cuda")
    before = g.tiering()[1]
    a = g.solve(Th)
    assert bool((a["exitflag"] == 1).all())
    assert g.tiering()[1] > before                              # some working sets outgrow the reduced capacity
    umin, umax = np.asarray(mpc.umin, float), np.asarray(mpc.umax, float)
    U = a["U"].cpu().numpy()
    assert np.all(U >= umin[:, None] - 1e-6) and np.all(U <= umax[:, None] + 1e-6)
    b = g.solve(Th)
    assert bool((a["U"] == b["U"]).all()) and bool((a["iters"] == b["iters"]).all())
    sub = Th[:20_000].cpu().numpy().T
    host = g.solve(np.asfortranarray(sub), want_active=False)
    assert np.array_equal(host["U"], U[:, :20_000])
    idx = torch.randperm(N, device="cuda")[:1000]
    ref = P.solve_batch(Th[idx].cpu().numpy().T)
    assert np.array_equal(U[:, idx.cpu().numpy()], ref["U"])
    for k, r in (("exitflag", "flag"), ("iters", "iters")):
        assert np.array_equal(a[k][idx].cpu().numpy(), ref[r]), k
    assert np.array_equal(a["active_upper"][idx].cpu().numpy().view(np.uint64), ref["act_u"])   # int64 tensor: same bits
    assert np.array_equal(a["active_lower"][idx].cpu().numpy().view(np.uint64), ref["act_l"])
