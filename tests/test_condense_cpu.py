"""Host-side condensing behind the C ABI against the numpy restatement of
mpc2mpqp in oracle/mpqp.py, through the Python builder mirror (no GPU needed)."""
import numpy as np
import pytest

import util
from oracle import mpqp


def mirror_of(lmpc, o):
    """Rebuild an oracle-side controller description with the product's builder API."""
    m = lmpc.MPC(o.F, o.G, C=o.C, Gd=o.Gd if o.nd_model else None, Dd=o.Dd if o.nd_model else None,
                 Np=o.Np, Nc=o.Nc, f_offset=o.f_offset, h_offset=o.h_offset, xo=o.xo, uo=o.uo)
    m.set_objective(Q=o.Q, R=o.R, Rr=o.Rr, Qf=o.Qf, Ex=o.Ex, ex=o.ex, Eu=o.Eu, eu=o.eu)
    if len(o.umax):
        m.set_input_bounds(o.umin, o.umax)
    for c in o.constraints:
        m.add_constraint(Ax=c.Ax, Au=c.Au, Ar=c.Ar if c.Ar.size else None, Ad=c.Ad if c.Ad.size else None,
                         Aup=c.Aup if c.Aup.size else None, Ap=c.Ap if c.Ap.size else None,
                         ub=c.ub, lb=c.lb, ks=c.ks, soft=c.soft, prio=c.prio)
    m.reference_tracking = o.reference_tracking
    m.soft_weight = o.soft_weight
    return m


@pytest.mark.parametrize("name", ["invpend", "aircraft", "aircraft_ksall", "simple", "simple_q1000",
                                  "dint_hard", "chain_small", "chain"])
def test_condense_matches_numpy_restatement(lmpc, name):
    o = util.controller(name)
    ref = mpqp.mpc2mpqp(o)
    q = mirror_of(lmpc, o).condense()
    assert q["dims"] == o.parameter_dims()
    assert q["H"].shape == ref.H.shape and q["A"].shape == ref.A.shape and q["W"].shape == ref.W.shape
    sc = max(1.0, np.abs(ref.H).max())
    assert np.max(np.abs(q["H"] - ref.H)) <= 1e-12 * sc
    assert np.max(np.abs(q["f_theta"] - ref.f_theta)) <= 1e-12 * max(1.0, np.abs(ref.f_theta).max())
    assert np.allclose(q["A"], ref.A, rtol=1e-12, atol=1e-13) and np.allclose(q["W"], ref.W, rtol=1e-12, atol=1e-13)
    assert np.allclose(q["bu"], ref.bu, rtol=1e-13) and np.allclose(q["bl"], ref.bl, rtol=1e-13)
    assert np.array_equal(q["senses"], ref.senses) and np.all(q["f"] == 0)


def _genparam_rich():
    """README pendulum with every generalized-parameter term in use: Ex (np = 2), ex, Eu (np = 1, padded),
    eu, a parametric state constraint (Ap, np = 2) and a parametric input constraint."""
    o = mpqp.invpend(Np=12, Nc=4)
    o.set_objective(Ex=np.array([[0.3, 0.0], [0.0, -0.2], [0.1, 0.1], [0.0, 0.0]]), ex=[0.05, 0.0, -0.02, 0.0],
                    Eu=np.array([[-0.4, 0.25]]), eu=[0.07])
    o.add_constraint(Ax=np.array([[0.0, 0.0, 1.0, 0.0]]), Ap=np.array([[0.5, -1.0]]), ub=[0.4], lb=[-0.4],
                     ks=range(2, 8), soft=True)
    o.add_constraint(Au=np.array([[1.0]]), Ap=np.array([[1.0, 0.0]]), ub=[1.5], ks=range(1, 4))
    return o


@pytest.mark.parametrize("preview", [False, True])
@pytest.mark.parametrize("blocks", [None, [[1, 2, 9]]])
def test_generalized_parameters_match_numpy_restatement(lmpc, blocks, preview):
    """Ex / ex / Eu / eu (setup.jl:118-132, mpc2mpqp.jl:478-508) and the constraints' Ap (setup.jl:61-84):
    theta = [x; r; d; uprev; p], f != 0, the p columns of f_theta and W; also under move blocking."""
    for o in (mpqp.generalized_parameter_example(), _genparam_rich()):
        m = mirror_of(lmpc, o)
        o.parameter_preview = m.parameter_preview = preview
        if blocks is not None:
            if o.Np != 12:
                continue
            o.Nc = m.Nc = o.Np
            mpqp.move_block(o, blocks)
            m.move_block(blocks)
        ref = mpqp.mpc2mpqp(o)
        q = m.condense()
        npar = o.affine_parameter_dim() * (o.Np if preview else 1)
        assert m.affine_parameter_dim() * (o.Np if preview else 1) == npar
        assert q["f_theta"].shape[1] == sum(o.parameter_dims()) + npar
        assert q["H"].shape == ref.H.shape and q["A"].shape == ref.A.shape and q["W"].shape == ref.W.shape
        assert np.any(ref.f != 0) and np.any(ref.f_theta[:, -npar:] != 0) and np.any(ref.W[:, -npar:] != 0)
        assert np.max(np.abs(q["H"] - ref.H)) <= 1e-12 * max(1.0, np.abs(ref.H).max())
        assert np.max(np.abs(q["f"] - ref.f)) <= 1e-12 * max(1.0, np.abs(ref.f).max())
        assert np.max(np.abs(q["f_theta"] - ref.f_theta)) <= 1e-12 * max(1.0, np.abs(ref.f_theta).max())
        assert np.allclose(q["A"], ref.A, rtol=1e-12, atol=1e-13) and np.allclose(q["W"], ref.W, rtol=1e-12, atol=1e-13)
        assert np.allclose(q["bu"], ref.bu, rtol=1e-13) and np.allclose(q["bl"], ref.bl, rtol=1e-13)
        assert np.array_equal(q["senses"], ref.senses)


@pytest.mark.gpu
@pytest.mark.parametrize("p,expect", [(0.0, 1.0), (0.75, 0.25)])
def test_generalized_parameter_known_answer_through_the_builder(lmpc, p, expect):
    """runtests.jl:1250-1268 with the product's own builder: model -> C++ condensing -> GPU."""
    m = lmpc.MPC([[1.0]], [[1.0]], C=[[1.0]], Np=4, Nc=4)
    m.set_bounds(umin=[0.0], umax=[2.0])
    m.set_objective(Q=[0.0], R=[1e-6], Eu=[[-1.0]], eu=[-0.1])
    m.add_constraint(Au=[[1.0]], Ap=[[1.0]], ub=[1.0], ks=range(1, 5))
    u = m.compute_control(np.array([0.0]), r=[0.0], p=[p])
    assert abs(u[0] - expect) < 1e-6


def _assert_same_mpqp(q, ref):
    assert q["H"].shape == ref.H.shape and q["A"].shape == ref.A.shape and q["W"].shape == ref.W.shape
    assert np.max(np.abs(q["H"] - ref.H)) <= 1e-12 * max(1.0, np.abs(ref.H).max())
    assert np.max(np.abs(q["f"] - ref.f)) <= 1e-11 * max(1.0, np.abs(ref.f).max())
    assert np.max(np.abs(q["f_theta"] - ref.f_theta)) <= 1e-12 * max(1.0, np.abs(ref.f_theta).max())
    for k in ("A", "W"):       # relative to the largest entry (feedback rows go through powers of F - G K)
        r = getattr(ref, k)
        assert r.size == 0 or np.max(np.abs(q[k] - r)) <= 1e-11 * max(1.0, np.abs(r).max())
    assert np.allclose(q["bu"], ref.bu, rtol=1e-12, atol=1e-13) and np.allclose(q["bl"], ref.bl, rtol=1e-12, atol=1e-13)
    assert np.array_equal(q["senses"], ref.senses)


@pytest.mark.parametrize("name,rp,dp,pre,blocks", [
    ("invpend", False, False, False, None), ("invpend", True, False, False, None),
    ("aircraft", False, False, False, None), ("aircraft_ksall", False, True, False, None),
    ("aircraft_ksall", True, True, False, None), ("aircraft", False, False, True, None),
    ("invpend", False, False, True, [[1, 2, 47]]), ("aircraft_ksall", False, False, False, [[2, 2, 6], [3, 7]])])
def test_model_offsets_match_numpy_restatement(lmpc, name, rp, dp, pre, blocks):
    """f_offset / h_offset (model.jl:45-72; mpc2mpqp.jl:395-399, 518-530, 684-689): the C++ condenser carries the
    constant as a trailing extended state, the restatement as an extra disturbance channel (itself checked
    against simulations of the affine model in test_preview_cpu.py); with previews, a feedback, move blocking."""
    rng = np.random.default_rng(2)
    o = util.controller(name)
    if name.startswith("aircraft"):
        o.Gd = np.array([[0.1, 0.0], [0.0, 0.02], [0.03, 0.0], [0.0, 0.01]])
    o.f_offset, o.h_offset = 0.1 * rng.normal(size=o.nx), 0.2 * rng.normal(size=o.ny)
    o.reference_preview, o.disturbance_preview = rp, dp
    if pre:
        o.set_prestabilizing_feedback()
    m = mirror_of(lmpc, o)
    m.reference_preview, m.disturbance_preview = rp, dp
    if pre:
        m.set_prestabilizing_feedback(o.K)
    if blocks is not None:
        o.Nc = m.Nc = o.Np
        mpqp.move_block(o, blocks)
        m.move_block(blocks)
    ref = mpqp.mpc2mpqp(o)
    assert np.any(ref.f != 0)
    _assert_same_mpqp(m.condense(), ref)


@pytest.mark.parametrize("pre", [False, True])
def test_regulator_operating_point_matches_numpy_restatement(lmpc, pre):
    """reference_tracking off with (xo, uo) (mpc2mpqp.jl:433-443, 457-459), also with the feedback terms."""
    o = mpqp.invpend(Np=12, Nc=4)
    o.reference_tracking = False
    o.Rr = np.zeros((1, 1))
    o.xo, o.uo = np.array([0.3, -0.1, 0.05, 0.2]), np.array([0.4])
    o.f_offset = np.array([0.01, 0.0, -0.02, 0.03])
    if pre:
        o.set_prestabilizing_feedback()
    m = mirror_of(lmpc, o)
    if pre:
        m.set_prestabilizing_feedback(o.K)
    ref = mpqp.mpc2mpqp(o)
    assert ref.f_theta.shape[1] == 4
    _assert_same_mpqp(m.condense(), ref)


@pytest.mark.gpu
def test_affine_model_controller_on_gpu_matches_oracle(lmpc):
    """A linearised (affine) pendulum model -> C++ condensing -> GPU, against the oracle's solve of the
    restated mpQP on sampled theta (f != 0 through the collapsed offset)."""
    from oracle import ldp
    o = mpqp.invpend()
    o.f_offset, o.h_offset = np.array([0.02, -0.01, 0.03, 0.0]), np.array([0.05, -0.02])
    m = mirror_of(lmpc, o)
    g = m.setup()
    qp = mpqp.mpc2mpqp(o)
    P = ldp.OracleProblem(mpqp.qp2ldp(qp, o.nu), o.nu)
    lo, hi = mpqp.invpend_range(1.0)
    Th = mpqp.sample_theta(lo, hi, 2000)
    out, ref = g.solve(Th), P.solve_batch(Th)
    assert np.array_equal(out["exitflag"], ref["flag"])
    assert np.max(np.abs(out["U"] - ref["U"])) < 1e-9
    # the offset matters: the plain model gives a different control on the same theta
    plain = mirror_of(lmpc, mpqp.invpend()).setup().solve(Th)
    assert np.max(np.abs(plain["U"] - out["U"])) > 1e-3


def test_offsets_with_generalized_parameters_are_refused(lmpc):
    m = lmpc.MPC([[1.0]], [[1.0]], Np=3, f_offset=[0.1])
    m.set_objective(Eu=[[1.0]])
    with pytest.raises(lmpc.LmpcError):
        m.condense()


def test_condense_preprocessing_known_answer(lmpc):
    """runtests.jl:1306-1318 through the product's own builder."""
    F, G = mpqp.zoh(np.array([[0, 1.0], [10, 0]]), np.array([[0], [1.0]]), 0.1)
    m = lmpc.MPC(F, G, Np=10)
    m.set_bounds(umin=[-1], umax=[1])
    m.add_constraint(Au=[[-1.0]], lb=[-0.9], ub=[1.5], ks=range(1, 11))
    m.add_constraint(Au=[[1.0]], lb=[-0.5], ub=[2.0], ks=range(1, 11))
    q = m.condense()
    assert q["A"].shape[0] == 0
    assert np.all(q["bu"] == 0.9) and np.all(q["bl"] == -0.5)


def test_condense_rejects_bad_model(lmpc):
    m = lmpc.MPC(np.eye(2), np.ones((2, 1)), Np=3, Nc=5)
    with pytest.raises(lmpc.LmpcError):
        m.condense()


@pytest.mark.gpu
def test_builder_end_to_end_known_answers(lmpc):
    """model -> C++ condensing -> GPU, no oracle in the loop: runtests.jl:62-66 and simple.md:107,203."""
    o = util.controller("invpend")
    m = mirror_of(lmpc, o)
    assert abs(m.compute_control(np.array([5.0, 5, 0, 0]))[0] - 1.7612519326) < 1e-6
    s = lmpc.MPC([[1, 0.5], [0, 1.0]], [[0], [1.0]], C=[[1.0, 0], [1, 1]], Np=10)
    s.set_objective(Q=[1, 1], R=0, Rr=[1])
    s.set_bounds(umin=[-3], umax=[3], ymin=[0, 0], ymax=[1, 2])
    s.add_constraint(Ax=[[2.0, -1]], lb=[-1], ub=[2])
    assert abs(s.compute_control(np.array([0.5, 1.0]), r=[0.0, 0.0])[0] + 1.0) < 1e-6
    s.set_objective(Q=[1000, 1], Rr=[1])
    assert abs(s.compute_control(np.array([0.0, 0.0]), r=[1.0, 0.0], uprev=[0.0])[0] - 1.0) < 1e-6
    sim = m.simulate(np.array([[0.0, 0, 0.15, 0]]).T, N=200)
    assert sim["exitflag"][0] >= 1 and abs(sim["Xend"][2, 0]) < 0.05


@pytest.mark.parametrize("name,blocks", [("invpend", [[1, 1, 48]]), ("invpend", [[2, 3, 5, 40]]),
                                         ("aircraft", [[1, 2, 7], [3, 7]]), ("dint_hard", [[1, 3]]),
                                         ("aircraft_ksall", [[2, 2, 6], [2, 2, 6]])])
def test_move_blocking_matches_numpy_restatement(lmpc, name, blocks):
    """move_block! (setup.jl:196-251) + apply_move_block (mpc2mpqp.jl:830-857): the C++ condensing
    against the numpy restatement, including per-input blocks of different lengths."""
    o = util.controller(name)
    o.Nc = o.Np
    mpqp.move_block(o, blocks)
    ref = mpqp.mpc2mpqp(o)
    m = mirror_of(lmpc, util.controller(name))
    m.Nc = m.Np
    m.move_block(blocks)
    assert m.Nc == o.Nc and m.move_blocks == o.move_blocks
    q = m.condense()
    assert q["H"].shape == ref.H.shape == (sum(len(b) for b in o.move_blocks),) * 2
    assert q["A"].shape == ref.A.shape and q["bu"].shape == ref.bu.shape
    assert np.max(np.abs(q["H"] - ref.H)) <= 1e-12 * max(1.0, np.abs(ref.H).max())
    assert np.max(np.abs(q["f_theta"] - ref.f_theta)) <= 1e-12 * max(1.0, np.abs(ref.f_theta).max())
    assert np.allclose(q["A"], ref.A, rtol=1e-12, atol=1e-13) and np.allclose(q["W"], ref.W, rtol=1e-12, atol=1e-13)
    assert np.allclose(q["bu"], ref.bu, rtol=1e-13) and np.allclose(q["bl"], ref.bl, rtol=1e-13)
    assert np.array_equal(q["senses"], ref.senses)


def test_move_blocked_problem_is_the_full_problem_with_tied_moves():
    """Independent check of the blocking algebra: the optimum of the blocked mpQP equals the
    optimum of the UNBLOCKED one under the equality constraints 'moves of a block are equal' (and
    held after the last block), solved by SLSQP."""
    from scipy.optimize import minimize
    from oracle import ldp
    blocks = [[2, 3, 5]]
    full = mpqp.invpend(Np=10, Nc=10)
    blk = mpqp.invpend(Np=10, Nc=10)
    mpqp.move_block(blk, blocks)
    qf, qb = mpqp.mpc2mpqp(full), mpqp.mpc2mpqp(blk)
    assert qb.n == 3 and blk.Nc == 6
    P = ldp.OracleProblem(mpqp.qp2ldp(qb, 1), 1)
    rng = np.random.default_rng(0)
    lo, hi = mpqp.invpend_range(0.2)
    # U (10 moves) tied: [z1 z1 z2 z2 z2 z3 z3 z3 z3 z3]
    Tm = np.zeros((10, 3)); Tm[0:2, 0] = 1; Tm[2:5, 1] = 1; Tm[5:, 2] = 1
    for _ in range(12):
        th = lo + (hi - lo) * rng.random(len(lo))
        u_blk, flag, _ = P.solve(th)
        assert flag >= 1
        fvec = qf.f + qf.f_theta @ th
        obj = lambda z: 0.5 * (Tm @ z) @ qf.H @ (Tm @ z) + fvec @ (Tm @ z)
        res = minimize(obj, np.zeros(3), jac=lambda z: Tm.T @ (qf.H @ (Tm @ z) + fvec), method="SLSQP",
                       bounds=[(-2.0, 2.0)] * 3, options=dict(ftol=1e-14, maxiter=200))
        assert res.success and abs(res.x[0] - u_blk[0]) < 1e-6
