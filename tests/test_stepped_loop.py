"""Stepped closed loop: user-defined true plant and observers with their own state
(OffsetFreeObserver) around the batched solve.  CPU part: the oracle's general loop pinned on the plain loop
and on a numpy transcription of the reference's Simulation loop; GPU part: lmpc_gpu_simulate_batch_stepped
against the oracle, and the reference's own closed-loop test cases (test/runtests.jl:989-1010, 1071-1082)."""
import numpy as np
import pytest

import util
from oracle import mpqp, ldp


def _dint(offset_free=False, lmpc=None, **kw):
    """the double integrator of test/runtests.jl:990-1003 with the product builder (lmpc) or as the oracle's MPC"""
    F, G, Cm = np.array([[1, 0.1], [0, 1.0]]), np.array([[0.005], [0.1]]), np.array([[1.0, 0]])
    if lmpc is None:
        m = mpqp.MPC(F, G, C=Cm, Np=20, **kw)
        m.set_objective(Q=[1.0], R=[0.0], Rr=[0.1])
        m.set_bounds(umin=[-1.0], umax=[1.0])
        return m
    m = lmpc.MPC(F, G, C=Cm, Np=20, **kw)
    m.set_objective(Q=[1.0], R=[0.0], Rr=[0.1])
    m.set_bounds(umin=[-1.0], umax=[1.0])
    if offset_free:
        m.set_offset_free_observer(method="velocity", Q=[1e-3, 1e-3], R=[1e-4])
    return m


def _oracle_problem(o):
    qp = mpqp.mpc2mpqp(o)
    return ldp.OracleProblem(mpqp.qp2ldp(qp, o.nu), o.nu), qp


def test_oracle_general_loop_equals_the_plain_loop():
    """no observer, the model as the plant: orc_simulate_batch_stepped == orc_simulate_batch bit for bit (cold
    starts; warm, the stepped loop rebuilds the factorisation from the carried working set, the fused loop keeps
    it: same optimum, different iteration counts)"""
    o = util.controller("aircraft")
    P, _ = _oracle_problem(o)
    nx, nr, nd, nup = o.parameter_dims()
    lo, hi = util.theta_range("aircraft", 0.25)
    N, T = 40, 25
    X0 = mpqp.sample_theta(lo[:nx], hi[:nx], N, seed=3)
    R = np.repeat(np.array([[0.1], [0.0]]), N, axis=1)
    D = 0.01 * np.ones((nd, N))
    for warm in (False, True):
        a = P.simulate_batch(o.F, o.G, o.Gd, X0, R, D, None, T, nr, nd, nup, warm=warm, want_traj=True)
        b = ldp.simulate_batch_stepped(P, nx, nr, nup, X0, R, D, None, T, dict(Ft=o.F, Gt=o.G, Gdt=o.Gd, ndm=nd),
                                       warm=warm, want_traj=True)
        if not warm:
            for k in ("Xend", "Ulast", "Utraj", "flag", "iters"):
                assert np.array_equal(a[k], b[k]), k
        else:
            assert np.array_equal(a["flag"], b["flag"]) and np.max(np.abs(a["Utraj"] - b["Utraj"])) < 1e-4


def _numpy_offset_free_loop(o, P, obs, x0, r, T, f_true, ndm=0):
    """simulation.jl:93-113 + observer.jl:104-124 + get_control_disturbance (:203-222), written out with numpy"""
    nx = o.nx
    x, xa, u = x0.copy(), np.concatenate([x0, np.zeros(obs["F"].shape[0] - nx)]), np.zeros(o.nu)
    xs, us = [], []
    for _ in range(T):
        y = obs["C"][:, :nx] @ x
        xa = xa + obs["K"] @ (y - obs["C"] @ xa)                       # correct!
        th = np.concatenate([xa[:nx], r, xa[nx:], u])                 # theta = [xhat; r; dhat; uprev]
        u, flag, _ = P.solve(th)
        assert flag >= 1
        xa = obs["F"] @ xa + obs["G"] @ u                              # predict!
        x = f_true(x, u)
        xs.append(x.copy()); us.append(u.copy())
    return np.array(xs), np.array(us), xa


def _velocity_observer(o):
    """set_offset_free_observer!(mpc; method=:velocity, Q, R) restated (setup.jl:393-487) on the oracle's MPC"""
    from scipy.linalg import solve_discrete_are
    F, Cm = o.F, o.C
    nx, ny = o.nx, o.ny
    Pm = solve_discrete_are(F.T, Cm.T, np.diag([1e-3, 1e-3]), np.diag([1e-4]))
    K = Pm @ Cm.T @ np.linalg.inv(Cm @ Pm @ Cm.T + np.diag([1e-4]))
    Bd, Cd = K, np.eye(ny) - Cm @ K
    obs = dict(F=np.block([[F, Bd], [np.zeros((ny, nx)), np.eye(ny)]]), G=np.vstack([o.G, np.zeros((ny, o.nu))]),
               C=np.hstack([Cm, Cd]), K=np.vstack([K, np.eye(ny)]))
    aug = _dint(Gd=Bd, Dd=Cd)
    return obs, aug


def test_oracle_offset_free_loop_matches_the_written_out_reference_loop():
    """The reference's own case (runtests.jl:989-1010): an unmodelled constant push [0.01, 0] leaves the nominal
    controller 5e-2 off the reference, the velocity-form offset-free observer removes the offset."""
    f_true = lambda x, u: np.array([[1, 0.1], [0, 1.0]]) @ x + np.array([0.005, 0.1]) * u[0] + np.array([0.01, 0.0])
    nominal = _dint()
    Pn, _ = _oracle_problem(nominal)
    x, u = np.zeros(2), np.zeros(1)
    for _ in range(100):
        u, flag, _ = Pn.solve(np.concatenate([x, [0.5], u]))
        x = f_true(x, u)
    assert abs(x[0] - 0.5) > 5e-2
    obs, aug = _velocity_observer(nominal)
    Pa, _ = _oracle_problem(aug)
    assert aug.parameter_dims() == (2, 1, 1, 1)
    xs, us, xa = _numpy_offset_free_loop(aug, Pa, obs, np.zeros(2), np.array([0.5]), 100, f_true)
    assert abs(xs[-1, 0] - 0.5) < 1e-3
    cfg = dict(Ft=aug.F, Gt=aug.G, ft=[0.01, 0.0], ny=1, Cm=aug.C, nxo=3, ndm=0, Fo=obs["F"], Go=obs["G"],
               Co=obs["C"], Ko=obs["K"])
    out = ldp.simulate_batch_stepped(Pa, 2, 1, 1, np.zeros((2, 1)), np.array([[0.5]]), None, None, 100, cfg,
                                     warm=False, want_traj=True)
    assert out["flag"][0] >= 1
    assert np.max(np.abs(out["Utraj"][0, :, 0] - us[:, 0])) < 1e-9 and np.max(np.abs(out["Xend"][:, 0] - xs[-1])) < 1e-9
    assert np.max(np.abs(out["Xhat_end"][:, 0] - xa)) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("warm", [True, False])
def test_gpu_offset_free_loop_matches_oracle(lmpc, warm):
    """lmpc_gpu_simulate_batch_stepped with the OffsetFreeObserver built by the product's own builder: bit-identical
    to the oracle's general loop on a batch of initial states, and the reference's pass criteria at x0 = 0."""
    nominal = _dint(lmpc=lmpc)
    sim_n = nominal.simulate(np.zeros((2, 1)), r=[0.5], N=100, plant=dict(F=nominal.F, G=nominal.G, f_offset=[0.01, 0.0]))
    assert abs(sim_n["Xend"][0, 0] - 0.5) > 5e-2
    m = _dint(offset_free=True, lmpc=lmpc)
    assert m.Gd.shape == (2, 1) and m.Dd.shape == (1, 1) and m.nd == 1 and m.nd_measured == 0
    rng = np.random.default_rng(0)
    N, T = 256, 100
    X0 = np.hstack([np.zeros((2, 1)), 0.2 * rng.normal(size=(2, N - 1))])
    plant = dict(F=nominal.F, G=nominal.G, f_offset=[0.01, 0.0])
    out = m.simulate(X0, r=[0.5], N=T, plant=plant, warm_start=warm, want_traj=True)
    assert np.all(out["exitflag"] >= 1) and abs(out["Xend"][0, 0] - 0.5) < 1e-3
    g = m.gpu
    cst = g.ldp_constants()
    o = _dint(Gd=m.Gd, Dd=m.Dd)
    L = mpqp.qp2ldp(mpqp.mpc2mpqp(o), 1)
    L.Mbar, L.Dth, L.du, L.dl = cst["Mbar"], cst["Dth"], cst["du"], cst["dl"]
    L.Rinv = np.vstack([cst["Rinv"], np.zeros((L.Rinv.shape[0] - 1, L.Rinv.shape[1]))])
    L.Uth_offset, L.u_offset = cst["Uth"], cst["uoff"]
    P = ldp.OracleProblem(L, 1, cap=cst["cap"])
    P.Gram[:] = cst["Gram"]
    ob = m.observer
    cfg = dict(Ft=nominal.F, Gt=nominal.G, ft=[0.01, 0.0], ny=1, Cm=nominal.C, nxo=3, ndm=0, Fo=ob["F"], Go=ob["G"],
               Co=ob["C"], Ko=ob["K"])
    ref = ldp.simulate_batch_stepped(P, 2, 1, 1, X0, np.full((1, N), 0.5), None, None, T, cfg, warm=warm, want_traj=True)
    assert np.array_equal(out["exitflag"], ref["flag"]) and np.array_equal(out["iters"], ref["iters"])
    assert np.array_equal(out["Utraj"], ref["Utraj"]) and np.array_equal(out["Xend"], ref["Xend"])
    assert np.array_equal(out["Xhat_end"], ref["Xhat_end"])


@pytest.mark.gpu
def test_gpu_nonlinear_plant_callback(lmpc):
    """A user-defined NONLINEAR true plant (scenario.dynamics, simulation.jl:39,110): the README pendulum's
    controller drives Euler steps of the nonlinear cart-pole-like model written with torch ops on the library's
    stream; against the oracle's loop with the same function in numpy (1e-9: sin / cos differ in the last bits)."""
    import torch
    o = mpqp.invpend()
    from test_condense_cpu import mirror_of
    m = mirror_of(lmpc, o)
    Ts = 0.01
    # the README pendulum before linearisation: gravity acts through sin(theta) (Euler steps of 1 ms x 10)
    def f_np(k, x, u, d):
        x = x.copy()
        for _ in range(10):
            s_ = np.sin(x[2])
            x = x + (Ts / 10) * np.array([x[1], -10 * x[1] + 9.81 * s_ + 100 * u[0], x[3], -20 * x[1] + 39.24 * s_ + 200 * u[0]])
        return x
    calls = []
    def f_torch(k, X, U, D, Xn, stream):
        calls.append(k)
        assert D is None and X.shape[1] == 4 and U.shape[1] == 1
        x, u = X.clone(), U[:, 0]
        for _ in range(10):
            s_ = torch.sin(x[:, 2])
            dx = torch.stack([x[:, 1], -10 * x[:, 1] + 9.81 * s_ + 100 * u, x[:, 3], -20 * x[:, 1] + 39.24 * s_ + 200 * u], dim=1)
            x = x + (Ts / 10) * dx
        Xn.copy_(x)
    rng = np.random.default_rng(1)
    N, T = 64, 300
    X0 = np.vstack([0.1 * rng.normal(size=(2, N)), 0.15 * rng.normal(size=(1, N)), 0.1 * rng.normal(size=(1, N))])
    out = m.simulate(X0, r=[0.0, 0.0], N=T, dynamics=f_torch, want_traj=True)
    assert calls == list(range(T)) and np.all(out["exitflag"] >= 1)
    P, _ = _oracle_problem(o)
    nx, nr, nd, nup = o.parameter_dims()
    ref = ldp.simulate_batch_stepped(P, nx, nr, nup, X0, np.zeros((nr, N)), None, None, T, dict(dyn=f_np, ndm=0),
                                     warm=True, want_traj=True)
    assert np.array_equal(out["exitflag"], ref["flag"])
    assert np.max(np.abs(out["Utraj"] - ref["Utraj"])) < 1e-8 and np.max(np.abs(out["Xend"] - ref["Xend"])) < 1e-8
    # the nonlinear pendulum is brought upright from these initial angles
    assert np.all(np.abs(out["Xend"][2]) < 0.02)


@pytest.mark.gpu
def test_gpu_constant_offset_with_state_observer(lmpc):
    """runtests.jl:1071-1082: f_offset in the model, a Kalman state observer, the affine plant as `dynamics`:
    x1 settles at the reference 0.5 (output bounds +-0.5 are soft)."""
    F, G = np.array([[1, 0.1], [0, 1.0]]), np.array([[0.005], [0.1]])
    m = lmpc.MPC(F, G, C=[[1.0, 0]], Np=25, f_offset=[0.1, 0.1])
    m.set_objective(R=[0.0], Rr=[1.0], Q=[1.0])
    m.set_bounds(umin=[-2], umax=[2], ymin=[-0.5], ymax=[0.5])
    m.set_state_observer(Q=1e-3)
    rng = np.random.default_rng(2)
    noise = 0.01 * rng.normal(size=(1, 200))
    out = m.simulate(np.zeros((2, 4)), r=[0.5], N=200, plant=dict(F=F, G=G, f_offset=[0.1, 0.1]), noise=noise, want_traj=True)
    assert np.all(out["exitflag"] >= 1)
    assert np.all(np.abs(out["Xend"][0] - 0.5) < 0.1)
