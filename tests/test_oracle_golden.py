"""CPU tests that pin the oracle (oracle/) before anything is compared against it.

Anchors: the reference's own known answers, brute-force KKT enumeration,
and a KKT certificate check of every oracle solution.  No GPU needed."""
import os

import numpy as np
import pytest
from scipy.optimize import linprog

import util
from oracle import mpqp, ldp, kkt_enum

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make(name):
    mpc = util.controller(name)
    qp = mpqp.mpc2mpqp(mpc)
    L = mpqp.qp2ldp(qp, mpc.nu, soft_weight=mpc.soft_weight)
    return mpc, qp, L, ldp.OracleProblem(L, mpc.nu)


@pytest.mark.parametrize("name", ["invpend", "simple", "simple_q1000"])
def test_known_answers(name):
    """runtests.jl:62-66 (1.7612519326), simple.md:107 (-1), simple.md:203-205 (1.0)."""
    mpc, qp, L, P = make(name)
    k = util.KNOWN[name]
    u, flag, _ = P.solve(mpc.form_parameter(k["x"], r=k["r"]))
    assert flag >= 1
    assert abs(u[0] - k["u"]) < k["tol"]


@pytest.mark.parametrize("p,expect", [(0.0, 1.0), (0.75, 0.25)])
def test_known_answer_generalized_parameter(p, expect):
    """runtests.jl:1250-1268: u = 1.0 (p = 0) and 0.25 (p = 0.75), tol 1e-6.  The only known
    answer with f != 0: exercises v = R^-T f, du/dl += Mext v and the output offset u_offset."""
    mpc, qp, L, P = make("genparam")
    assert (mpc.parameter_dims(), mpc.affine_parameter_dim()) == ((1, 1, 0, 0), 1)
    assert np.any(qp.f != 0) and np.any(L.u_offset != 0)
    u, flag, _ = P.solve(mpc.form_parameter([0.0], r=[0.0], p=[p]))
    assert flag >= 1 and abs(u[0] - expect) < 1e-6


def test_problem_dimensions():
    """n, m, ms, nth of the BASELINE controllers."""
    _, qp, _, _ = make("invpend")
    assert (qp.n, qp.m, qp.ms, qp.nth) == (5, 5, 5, 7)
    assert 1.4e3 < np.linalg.cond(qp.H) < 1.8e3
    _, qp, _, _ = make("aircraft")
    assert (qp.n, qp.m, qp.ms, qp.nth) == (20, 22, 20, 10)
    assert np.count_nonzero(qp.senses & mpqp.SOFT) == 2
    _, qp, _, _ = make("aircraft_ksall")
    assert (qp.n, qp.m, qp.ms, qp.nth) == (20, 38, 20, 10)
    _, qp, _, _ = make("chain")
    assert (qp.n, qp.m, qp.ms, qp.nth) == (180, 180, 180, 20)


def test_preprocessing_matches_reference():
    """runtests.jl:1306-1318: general rows on u alone fold into the simple bounds."""
    F, G = mpqp.zoh(np.array([[0, 1.0], [10, 0]]), np.array([[0], [1.0]]), 0.1)
    mpc = mpqp.MPC(F, G, Np=10)
    mpc.set_bounds(umin=[-1], umax=[1])
    mpc.add_constraint(Au=np.array([[-1.0]]), lb=[-0.9], ub=[1.5], ks=range(1, 11))
    mpc.add_constraint(Au=np.array([[1.0]]), lb=[-0.5], ub=[2.0], ks=range(1, 11))
    qp = mpqp.mpc2mpqp(mpc)
    assert qp.A.shape[0] == 0
    assert np.all(qp.bu == 0.9 * np.ones(10))
    assert np.all(qp.bl == -0.5 * np.ones(10))


@pytest.mark.parametrize("scale", [1.0, 0.25, 0.05])
def test_invpend_matches_kkt_enumeration(scale):
    """3^5 = 243 candidate working sets per theta; unique optimum and active set."""
    mpc, qp, L, P = make("invpend")
    lo, hi = mpqp.invpend_range(scale)
    Th = mpqp.sample_theta(lo, hi, 120, seed=1234 + int(scale * 100))
    out = P.solve_batch(Th)
    au, al = util.bits(out["act_u"], 5), util.bits(out["act_l"], 5)
    for q in range(Th.shape[1]):
        sh = L.Dth @ Th[:, q]
        v, eu, el, margin = kkt_enum.enumerate_ldp(L.Mbar, L.du + sh, L.dl + sh)
        x = L.Rinv[:1] @ v + L.u_offset + L.Uth_offset @ Th[:, q]
        assert abs(x[0] - out["U"][0, q]) <= 1e-9 * max(1.0, abs(x[0]))
        assert out["flag"][q] == 1
        if margin > 1e-5:  # non-degenerate: the active set is unique
            assert set(np.where(au[q])[0]) == eu and set(np.where(al[q])[0]) == el


def kkt_certificate(L, P, theta, act_u, act_l, rho, tol=1e-6):
    """Independent optimality check of one oracle answer from its final active set."""
    sh = L.Dth @ theta
    du, dl = L.du + sh, L.dl + sh
    idx = np.where(act_u | act_l)[0]
    soft = (L.senses & mpqp.SOFT) != 0
    if len(idx):
        Mw = L.Mbar[idx]
        dw = np.where(act_u[idx], du[idx], dl[idx])
        G = Mw @ Mw.T + np.diag(np.where(soft[idx], rho, 0.0))
        lam = np.linalg.solve(G, -dw)
        v = -Mw.T @ lam
        sgn = np.where(act_u[idx], 1.0, -1.0)
        free = (L.senses[idx] & mpqp.ACTIVE) != 0      # equality rows: multiplier of either sign
        assert np.all((lam * sgn >= -1e-9 * (1 + np.abs(lam).max())) | free), "dual infeasible"
    else:
        v = np.zeros(L.Mbar.shape[1])
    mv = L.Mbar @ v
    inact = ~(act_u | act_l) & ((L.senses & mpqp.IMMUTABLE) == 0)
    assert np.all(du[inact] - mv[inact] >= -tol * 1.001) and np.all(mv[inact] - dl[inact] >= -tol * 1.001)
    return v


# tolerance of the independent re-solve: the aircraft Hessian has cond 3.5e8 and its active
# Gram matrices reach cond ~1e10, so two correct fp64 factorisations differ by ~1e-6 in u (an
# 80-bit re-solve differs from BOTH by the same amount); the others are well conditioned.
@pytest.mark.parametrize("name,scale,rtol", [("aircraft", 0.25, 2e-5), ("simple", 1.0, 2e-6),
                                             ("chain_small", 1.0, 1e-8), ("invpend", 1.0, 1e-9),
                                             ("genparam", 1.0, 1e-8), ("dint_eq", 1.0, 1e-8),
                                             ("dint_eq_small", 1.0, 1e-8)])
def test_oracle_solutions_satisfy_kkt(name, scale, rtol):
    mpc, qp, L, P = make(name)
    lo, hi = util.theta_range(name, scale)
    Th = mpqp.sample_theta(lo, hi, 200, seed=7)
    out = P.solve_batch(Th)
    au, al = util.bits(out["act_u"], P.m), util.bits(out["act_l"], P.m)
    eq = (L.senses & mpqp.ACTIVE) != 0
    hard = ((L.senses & mpqp.SOFT) == 0) & (((L.senses & mpqp.IMMUTABLE) == 0) | eq)
    for q in range(Th.shape[1]):
        if out["flag"][q] >= 1:
            assert np.all((au[q] | al[q])[eq]), "an equality row left the working set"
            v = kkt_certificate(L, P, Th[:, q], au[q], al[q], L.rho_soft)
            x = L.Rinv[:mpc.nu] @ v + L.u_offset + L.Uth_offset @ Th[:, q]
            assert np.allclose(x, out["U"][:, q], rtol=rtol, atol=rtol)
        else:
            assert out["flag"][q] == -1
            sh = L.Dth @ Th[:, q]
            A = np.vstack([L.Mbar[hard], -L.Mbar[hard]])
            b = np.concatenate([(L.du + sh)[hard], -(L.dl + sh)[hard]])
            res = linprog(np.zeros(A.shape[1]), A_ub=A, b_ub=b, bounds=[(None, None)] * A.shape[1])
            assert res.status == 2, "oracle said infeasible but the hard rows admit a point"


@pytest.mark.parametrize("fn", sorted(f for f in os.listdir(GOLD) if f.endswith(".npz") and "_s" in f))
def test_oracle_reproduces_golden(fn):
    """Frozen oracle outputs (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLD, fn))
    name = fn.rsplit("_s", 1)[0]
    mpc, qp, L, P = make(name)
    assert np.allclose(qp.H, z["H"], rtol=1e-12, atol=1e-12) and np.allclose(qp.W, z["W"], rtol=1e-12, atol=1e-12)
    out = P.solve_batch(z["Theta"])
    assert np.array_equal(out["flag"], z["flag"])
    assert np.array_equal(out["act_u"], z["act_u"]) and np.array_equal(out["act_l"], z["act_l"])
    ok = z["flag"] >= 1
    assert np.allclose(out["U"][:, ok], z["U"][:, ok], rtol=1e-10, atol=1e-12)


def test_warm_equals_cold_closed_loop():
    """runtests.jl:85-117: 100 closed-loop steps, |u_warm - u_cold| < 1e-9."""
    mpc, qp, L, P = make("invpend")
    nx, nr, nd, nup = mpc.parameter_dims()
    X0 = np.array([[5.0, 5, 0, 0], [1.0, -2, 0.1, 0], [0, 0, 0.15, 0]]).T
    kw = dict(Fx=mpc.F, Gx=mpc.G, Gdx=None, X0=X0, R=np.zeros((nr, 3)), D=None, U0=None, T=100,
              nr=nr, nd=nd, nuprev=nup, want_traj=True)
    cold = P.simulate_batch(warm=False, **kw)
    warm = P.simulate_batch(warm=True, **kw)
    assert np.all(cold["flag"] >= 1) and np.all(warm["flag"] >= 1)
    assert np.max(np.abs(cold["Utraj"] - warm["Utraj"])) < 1e-9
    # the pendulum angle is being regulated (cart position takes longer than 1 s)
    assert np.all(np.abs(cold["Xend"][2]) < 0.5)


def test_thread_count_does_not_change_results():
    mpc, qp, L, P = make("invpend")
    lo, hi = mpqp.invpend_range(1.0)
    Th = mpqp.sample_theta(lo, hi, 1001)
    a = P.solve_batch(Th, nthreads=1)
    b = P.solve_batch(Th, nthreads=5)
    assert np.array_equal(a["U"], b["U"]) and np.array_equal(a["act_u"], b["act_u"])


def test_soft_rows_equal_the_reference_explicit_slack_form():
    """The oracle handles soft rows as rho = 1/soft_weight on the NORMALISED rows' LDL diagonal.
    The reference states the same problem with explicit slack variables weighted by soft_weight
    (make_singlesided, src/utils.jl:341-364; rho_soft at src/setup.jl:26).  Solve that explicit
    form with an unrelated method (SLSQP) and compare: 50 theta, 13 with active soft rows."""
    from scipy.optimize import minimize
    mpc, qp, L, P = make("simple")
    ss = mpqp.make_singlesided(qp, soft_weight=mpc.soft_weight, single_soft=False)
    H, f, fth, A, b, W = ss["H"], ss["f"], ss["f_theta"], ss["A"], ss["b"], ss["W"]
    lo, hi = util.theta_range("simple", 0.6)
    Th = mpqp.sample_theta(lo, hi, 60, seed=3)
    out = P.solve_batch(Th)
    n = qp.n
    sc = np.ones(H.shape[0])
    sc[n:] = 1 / np.sqrt(mpc.soft_weight)          # rescale the slack variables for conditioning
    Hs, As = H * sc[:, None] * sc[None, :], A * sc[None, :]
    checked = soft_active = 0
    for q in range(Th.shape[1]):
        if out["flag"][q] < 1:
            continue
        g, rhs = (f + fth @ Th[:, q]) * sc, b + W @ Th[:, q]
        res = minimize(lambda z: 0.5 * z @ Hs @ z + g @ z, np.zeros(len(sc)), jac=lambda z: Hs @ z + g,
                       method="SLSQP", constraints=[{"type": "ineq", "fun": lambda z: rhs - As @ z,
                                                     "jac": lambda z: -As}],
                       options={"ftol": 1e-15, "maxiter": 500})
        assert abs(res.x[0] - out["U"][0, q]) < 1e-6
        checked += 1
        soft_active += int(out["flag"][q] == 2)
    assert checked >= 40 and soft_active >= 5


def test_closed_loop_with_trajectories_equals_the_stepwise_loop():
    """orc_simulate_batch_traj against the Simulation loop written out in numpy
    (simulation.jl:93-110: theta_k = [x_k; rs[:, k]; ds[:, k]; u_{k-1}], x+ = F x + G u + Gd ds[:, k]),
    one cold solve per step.  Tolerance: the stepwise loop rounds the dynamics differently (BLAS
    vs the oracle's fma chains) and the AFTI-16 problem amplifies 1 ulp of x to ~1e-7 of u (cond(H)
    3.5e8, see test_oracle_solutions_satisfy_kkt); a wrong trajectory column would show at 1e-2."""
    mpc, qp, L, P = make("aircraft")
    nx, nr, nd, nup = mpc.parameter_dims()
    T, N = 25, 6
    rng = np.random.default_rng(3)
    lo, hi = util.theta_range("aircraft", 0.25)
    X0 = rng.uniform(lo[:nx], hi[:nx], (N, nx)).T
    Rt = np.zeros((nr, T)); Rt[0, 10:] = 0.1
    Dt = np.zeros((nd, T)); Dt[:, 4:9] = 0.05
    out = P.simulate_batch(mpc.F, mpc.G, mpc.Gd, X0, None, None, None, T, nr, nd, nup, warm=False,
                           want_traj=True, Rtraj=Rt, Dtraj=Dt)
    for q in range(N):
        x, u = X0[:, q].copy(), np.zeros(mpc.nu)
        for k in range(T):
            th = np.concatenate([x, Rt[:, k], Dt[:, k], u[:nup]])
            u, flag, _ = P.solve(th)
            assert flag >= 1
            assert np.allclose(u, out["Utraj"][q, k], rtol=0, atol=2e-5)
            x = mpc.F @ x + mpc.G @ u + mpc.Gd @ Dt[:, k]
        assert np.allclose(x, out["Xend"][:, q], rtol=0, atol=2e-5)


@pytest.mark.parametrize("name,tol", [("invpend", 1e-9), ("dint_hard", 1e-9), ("dint_eq", 1e-9), ("chain_small", 1e-9),
                                      ("simple", 1e-5)])
def test_warm_start_from_masks_reaches_the_cold_solution(name, tol):
    """activate_rows: starting from the optimal working set of a nearby theta changes the iteration
    count, not the answer (runtests.jl:85-117 asserts warm == cold to 1e-9); from an empty set it IS
    the cold start; from the instance's own optimal set the first iteration confirms it."""
    mpc, qp, L, P = make(name)
    lo, hi = util.theta_range(name, 0.5)
    Th0 = mpqp.sample_theta(lo, hi, 1500, seed=3)
    prev = P.solve_batch(Th0)
    Th = Th0 + 0.02 * (hi - lo)[:, None] * np.random.default_rng(1).standard_normal(Th0.shape)
    cold = P.solve_batch(Th)
    warm = P.solve_batch_warm(Th, prev["act_u"], prev["act_l"])
    assert np.array_equal(cold["flag"], warm["flag"])
    ok = cold["flag"] >= 1
    assert np.max(np.abs(cold["U"][:, ok] - warm["U"][:, ok])) <= tol
    assert warm["iters"].sum() < cold["iters"].sum()
    zero = np.zeros_like(prev["act_u"])
    z = P.solve_batch_warm(Th, zero, zero)
    assert np.array_equal(z["U"], cold["U"]) and np.array_equal(z["iters"], cold["iters"])
    own = P.solve_batch_warm(Th, cold["act_u"], cold["act_l"])
    assert np.all(own["iters"][cold["flag"] == 1] == 1)


def test_closed_loop_with_observer_equals_the_stepwise_loop():
    """orc_simulate_batch_ex against simulation.jl:92-110 + observer.jl:104-118 written out in numpy:
    set_state!(x0); per step y = C x + v, correct!, compute_control on the estimate, predict! with the
    model, true plant step with a mismatched F, G."""
    mpc, qp, L, P = make("invpend")
    nx, nr, nd, nup = mpc.parameter_dims()
    T, N = 40, 5
    rng = np.random.default_rng(4)
    lo, hi = util.theta_range("invpend", 0.1)
    X0 = rng.uniform(lo[:nx], hi[:nx], (N, nx)).T
    Cm = mpc.C
    ny = Cm.shape[0]
    K = 0.4 * np.linalg.pinv(Cm)
    Ft, Gt = mpc.F * (1 + 1e-3), mpc.G * 0.98
    V = 1e-3 * rng.standard_normal((ny, T))
    out = P.simulate_batch(mpc.F, mpc.G, None, X0, np.zeros((nr, N)), None, None, T, nr, nd, nup, warm=False,
                           want_traj=True, ext=dict(Ft=Ft, Gt=Gt, C=Cm, K=K, Vtraj=V))
    for q in range(N):
        x, xh, u = X0[:, q].copy(), X0[:, q].copy(), np.zeros(mpc.nu)
        for k in range(T):
            y = Cm @ x + V[:, k]
            xh = xh + K @ (y - Cm @ xh)
            u, flag, _ = P.solve(np.concatenate([xh, np.zeros(nr), u[:nup]]))
            assert flag >= 1
            assert np.allclose(u, out["Utraj"][q, k], rtol=0, atol=1e-8)
            xh = mpc.F @ xh + mpc.G @ u
            x = Ft @ x + Gt @ u
        assert np.allclose(x, out["Xend"][:, q], rtol=0, atol=1e-8)
        assert np.allclose(xh, out["Xhat_end"][:, q], rtol=0, atol=1e-8)


def _exact_kkt(L, P, th, au, al, nu):
    """Re-solve the equality-constrained problem on a given working set in 50-digit arithmetic
    (Gram matrix included, so the conditioning of the soft regularisation does not touch the check).
    Returns (lam, sides, worst primal violation over inactive rows, control)."""
    import mpmath
    mpmath.mp.dps = 50
    soft = (L.senses & mpqp.SOFT) != 0
    sh = L.Dth @ th
    dup, dlo = L.du + sh, L.dl + sh
    W = np.where(au | al)[0]
    sd = np.where(au[W], 1.0, -1.0)
    Mw = mpmath.matrix(L.Mbar[W].tolist())
    G = Mw * Mw.T
    for i in range(len(W)):
        if soft[W[i]]:
            G[i, i] += mpmath.mpf(L.rho_soft)
    dw = np.where(sd > 0, dup[W], dlo[W])
    lam = mpmath.lu_solve(G, mpmath.matrix((-dw).tolist()))
    u = np.array([float(x) for x in -(Mw.T * lam)])
    lam = np.array([float(x) for x in lam])
    mu = L.Mbar @ u
    pv = np.maximum(mu - dup, dlo - mu)
    pv[W] = -1.0
    pv[(L.senses & mpqp.IMMUTABLE) != 0] = -1.0
    return lam, sd, pv.max(), L.Rinv[:nu] @ u + L.u_offset + L.Uth_offset @ th


def test_default_output_bounds_never_end_infeasible():
    """AFTI-16 with output bounds on ks = 2:Np, the reference's DEFAULT ks (setup.jl:64,94): 18 soft rows.
    A soft-constrained problem is feasible by construction and the reference asserts exitflag >= 1
    (utils.jl:46).  Round 1 returned INFEASIBLE on 22 % of the example's own parameter range because
    pivots of hard rows factorised after soft rows (order rho |c|^2 ~ 1e-10) drowned in rounding; the
    repair (oracle: rebuild) refactorises with the hard block first.  Anchors:
      * flags in {1, 2} on 1e5 theta at each of s = 1, 0.5, 0.25 (3e6 checked offline: one CYCLE);
      * with the repair switched off the old failure is back (so the test is red on round 1's code);
      * on theta that failed before, the final working set is the exact optimum: re-solved in 50-digit
        arithmetic its multipliers have the right signs and no inactive row is violated beyond the
        noise of u = -M'lam with lam ~ 1e8 (bounds below; the same limit applies to any fp64 dual method)."""
    mpc, qp, L, P = make("aircraft_ksall")
    P0 = ldp.OracleProblem(L, mpc.nu, no_repair=1)
    for s, nbad_min, utol in [(1.0, 15000, 5e-3), (0.5, 1000, 5e-4), (0.25, 0, 1e-4)]:
        lo, hi = util.theta_range("aircraft_ksall", s)
        Th = mpqp.sample_theta(lo, hi, 100000)
        out = P.solve_batch(Th)
        assert set(np.unique(out["flag"]).tolist()) <= {1, 2}, np.unique(out["flag"], return_counts=True)
        old = P0.solve_batch(Th)
        bad = np.where(old["flag"] < 0)[0]
        assert len(bad) >= nbad_min
        au, al = util.bits(out["act_u"], P.m), util.bits(out["act_l"], P.m)
        pick = bad[:12] if len(bad) else np.arange(12)
        for q in pick:
            lam, sd, pviol, ctl = _exact_kkt(L, P, Th[:, q], au[q], al[q], mpc.nu)
            assert np.all(lam * sd >= -1e-6 * max(1.0, np.abs(lam).max())), q   # dual feasible (upper: lam >= 0)
            assert pviol <= 1e-3 * s, (q, pviol)
            assert np.max(np.abs(ctl - out["U"][:, q])) <= utol, (q, ctl, out["U"][:, q])


def test_repair_leaves_the_baseline_controllers_alone():
    """BASELINE AFTI-16 (2 soft rows) never needed the repair: identical results with it switched off."""
    mpc, qp, L, P = make("aircraft")
    P0 = ldp.OracleProblem(L, mpc.nu, no_repair=1)
    for s in (1.0, 0.25):
        lo, hi = util.theta_range("aircraft", s)
        Th = mpqp.sample_theta(lo, hi, 50000)
        a, b = P.solve_batch(Th), P0.solve_batch(Th)
        assert set(np.unique(a["flag"]).tolist()) <= {1, 2}
        assert np.array_equal(a["U"], b["U"]) and np.array_equal(a["iters"], b["iters"])
