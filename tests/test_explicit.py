"""Explicit-MPC path: the search-tree fixture, the oracle's walk, and the
batched evaluate kernel."""
import os

import numpy as np
import pytest

import util
from oracle import mpqp, ldp, explicit

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "invpend_explicit.npz")


def load_tree():
    z = np.load(GOLD)
    return z, ldp.OracleBst(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clip=z["clipping"])


def test_explicit_known_answer_and_matches_implicit():
    """runtests.jl:178-183 (1.7612519326 through the tree) and :381,:490 (explicit == implicit)."""
    z, B = load_tree()
    mpc = mpqp.invpend()
    qp = mpqp.mpc2mpqp(mpc)
    P = ldp.OracleProblem(mpqp.qp2ldp(qp, mpc.nu), mpc.nu)
    u, _ = B.evaluate_batch(mpc.form_parameter([5.0, 5, 0, 0]).reshape(-1, 1))
    assert abs(u[0, 0] - 1.7612519326) < 1e-6
    Th = mpqp.sample_theta(z["lo"], z["hi"], 50000, seed=21)
    U, reg = B.evaluate_batch(Th)
    ref = P.solve_batch(Th)
    assert np.max(np.abs(U - ref["U"])) < 1e-10
    assert reg.min() >= 0 and reg.max() < z["feedbacks"].shape[0]


def test_tree_arrays_are_well_formed():
    z, _ = load_tree()
    hl, jl = z["hp_list"], z["jump_list"]
    n = len(hl)
    inner = jl != 0
    assert np.all(jl[inner] > 0) and np.all(np.arange(n)[inner] + jl[inner] + 1 < n)
    assert np.all(hl[inner] < z["halfplanes"].shape[0]) and np.all(hl[~inner] < z["feedbacks"].shape[0])
    assert np.allclose(np.linalg.norm(z["halfplanes"][:, :-1], axis=1), 1.0)


def test_region_enumeration_small_problem():
    """The builder itself on a 2-variable problem: regions partition the box and the tree
    reproduces the implicit solution."""
    mpc = mpqp.double_integrator_hard(Np=2, Nc=2)
    mpc.constraints.clear()
    qp = mpqp.mpc2mpqp(mpc)
    lo, hi = np.array([-20.0, -10.0, -20.0]), np.array([20.0, 10.0, 20.0])
    sol = explicit.explicit_solution(mpc, qp, lo, hi, n_samples=2000)
    assert 3 <= len(sol["regions"]) <= 9
    B = ldp.OracleBst(sol["halfplanes"], sol["feedbacks"], sol["hp_list"], sol["jump_list"])
    P = ldp.OracleProblem(mpqp.qp2ldp(qp, mpc.nu), mpc.nu)
    Th = mpqp.sample_theta(lo, hi, 5000, seed=4)
    U, _ = B.evaluate_batch(Th)
    assert np.max(np.abs(U - P.solve_batch(Th)["U"])) < 1e-10


@pytest.mark.gpu
def test_gpu_evaluate_matches_oracle_walk(lmpc):
    z, B = load_tree()
    g = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"],
                            dims=(4, 2, 0, 1))
    for N in (0, 1, 255, 256, 257, 100000):
        Th = mpqp.sample_theta(z["lo"], z["hi"], max(N, 1), seed=3)[:, :N]
        U, reg = g.evaluate(Th)
        Uo, rego = B.evaluate_batch(Th) if N else (np.zeros((1, 0)), np.zeros(0, np.int32))
        assert np.array_equal(U, Uo) and np.array_equal(reg, rego)
    u = g.compute_control(np.array([5.0, 5, 0, 0]))
    assert abs(u[0] - 1.7612519326) < 1e-6           # runtests.jl:178-183


@pytest.mark.gpu
def test_gpu_explicit_equals_gpu_implicit(lmpc):
    """runtests.jl:381: norm(u_explicit - u_implicit) < 1e-10, here on 2e5 theta, plus clipping."""
    z, B = load_tree()
    mpc = mpqp.invpend()
    qp = mpqp.mpc2mpqp(mpc)
    gi = lmpc.setup_gpu(qp, nu=1, dims=mpc.parameter_dims())
    ge = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"],
                             dims=mpc.parameter_dims())
    Th = mpqp.sample_theta(z["lo"], z["hi"], 200000, seed=8)
    Ue, _ = ge.evaluate(Th)
    Ui = gi.solve(Th, want_active=False)["U"]
    assert np.max(np.abs(Ue - Ui)) < 1e-10
    assert np.all(np.abs(Ue) <= 2.0)


@pytest.mark.gpu
def test_gpu_evaluate_device_buffers_and_large_batch(lmpc):
    """BASELINE config 5 at its stated size, 1e8 theta (5.6 GB of parameters resident in HBM): every result
    within the input bounds, idempotent, device-resident evaluate equals the chunked host path and a
    subsample equals the oracle's walk bit for bit."""
    import torch
    z, B = load_tree()
    g = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"])
    N = 100_000_000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234)
    lo, hi = torch.tensor(z["lo"], device="cuda"), torch.tensor(z["hi"], device="cuda")
    Th = torch.rand((N, 7), dtype=torch.float64, device="cuda", generator=gen)
    Th.mul_(hi - lo).add_(lo)
    U, reg = g.evaluate(Th)
    assert bool((U.abs() <= 2.0 + 1e-9).all())
    U2, reg2 = g.evaluate(Th)
    assert bool((U == U2).all()) and bool((reg == reg2).all())
    del U2, reg2
    sub = Th[:300000].cpu().numpy().T
    Uh, regh = g.evaluate(np.asfortranarray(sub))
    assert np.array_equal(Uh, U[:, :300000].cpu().numpy())
    Uo, _ = B.evaluate_batch(sub[:, :50000])
    assert np.array_equal(Uo, Uh[:, :50000])


def test_product_builder_logic_with_a_cpu_stand_in():
    """The product's explicit builder (linearmpc.jl_b200/explicit.py) takes any object with a
    GpuMPC-like solve(); here the CPU oracle stands in for the GPU so that region discovery, the
    adjacency closure (two regions of this problem have a Chebyshev radius of 5e-3 and are never
    sampled), the laws and the tree are covered without a device."""
    import lmpc_b200
    from lmpc_b200 import explicit as pex
    mpc = mpqp.invpend(Np=30, Nc=3)
    qp = mpqp.mpc2mpqp(mpc)
    P = ldp.OracleProblem(mpqp.qp2ldp(qp, 1), 1)

    class StandIn:
        def solve(self, Th, want_iters=False):
            o = P.solve_batch(Th)
            return dict(exitflag=o["flag"], active_upper=o["act_u"], active_lower=o["act_l"])

    q = dict(H=qp.H, f=qp.f, f_theta=qp.f_theta, A=qp.A, bu=qp.bu, bl=qp.bl, W=qp.W, senses=qp.senses)
    lo, hi = mpqp.invpend_range(0.3)
    sol = pex.build_explicit(q, StandIn(), lo, hi, 1, n_samples=30000, tree_samples=4000)
    regs, laws = explicit.critical_regions(mpqp.make_singlesided(qp), lo, hi, 1)
    assert len(sol["regions"]) == len(regs) == 15 and len(sol["laws"]) == len(laws)
    B = ldp.OracleBst(sol["halfplanes"], sol["feedbacks"], sol["hp_list"], sol["jump_list"])
    Th = mpqp.sample_theta(lo, hi, 30000, seed=9)
    U, _ = B.evaluate_batch(Th)
    assert np.max(np.abs(U - P.solve_batch(Th)["U"])) < 1e-10
    # soft rows are made explicit with one slack column each (utils.jl:345-380), no longer refused
    q2 = dict(q)
    q2["senses"] = np.asarray(q["senses"]) | 8
    H2, f2, fth2, A2, b2, W2, src2 = pex.make_singlesided(q2, soft_weight=1e6)
    n, m = qp.H.shape[0], len(qp.bu)
    assert H2.shape == (n + m, n + m) and np.allclose(np.diag(H2)[n:], 1e6) and A2.shape[1] == n + m
    assert np.all(A2[:, n:] <= 0) and np.count_nonzero(A2[:, n:]) == A2.shape[0]   # a slack relaxes both sides of its row


@pytest.mark.gpu
def test_product_explicit_builder_matches_implicit_and_enumeration(lmpc):
    """lmpc_b200.MPC.explicit: regions discovered by batched implicit solves, exact laws
    and half-planes, LP-certified tree.  On a 3-move pendulum controller: the same full-dimensional
    regions as the test-side combinatorial enumeration, and explicit == implicit to 1e-10 on fresh
    parameters (runtests.jl:381,490)."""
    from lmpc_b200 import explicit as pex
    ref_mpc = mpqp.invpend(Np=30, Nc=3)
    from scipy.linalg import expm
    A = np.array([[0, 1, 0, 0], [0, -10, 9.81, 0], [0, 0, 0, 1], [0, -20, 39.24, 0.0]])
    Bm = 100 * np.array([[0], [1.0], [0], [2.0]])
    E = expm(np.block([[A * 0.01, Bm * 0.01], [np.zeros((1, 5))]]))
    mpc = lmpc.MPC(E[:4, :4], E[:4, 4:], C=[[1.0, 0, 0, 0], [0, 0, 1.0, 0]], Np=30, Nc=3)
    mpc.set_objective(Q=[1.2 ** 2, 1], R=[0.0], Rr=[1.0])
    mpc.set_bounds(umin=[-2.0], umax=[2.0])
    g = mpc.setup()
    lo, hi = mpqp.invpend_range(0.3)
    empc, sol = mpc.explicit(lo, hi, n_samples=100_000, seed=1)
    # test-side enumeration of the same problem (3^3 candidate sets)
    qp = mpqp.mpc2mpqp(ref_mpc)
    regs, laws = explicit.critical_regions(mpqp.make_singlesided(qp), lo, hi, 1)
    assert len(sol["regions"]) == len(regs) and len(sol["laws"]) == len(laws)
    rep = pex.verify(empc, g, lo, hi, n=300_000)
    assert rep["max_err"] < 1e-10 and rep["frac_within_tol"] == 1.0
    # and against the CPU oracle's implicit solve
    P = ldp.OracleProblem(mpqp.qp2ldp(qp, 1), 1)
    Th = mpqp.sample_theta(lo, hi, 20000, seed=6)
    U, _ = empc.evaluate(Th)
    assert np.max(np.abs(U - P.solve_batch(Th)["U"])) < 1e-10


def test_oracle_explicit_closed_loop_equals_implicit_closed_loop():
    """Simulation(empc) vs Simulation(mpc): the reference asserts 1e-10 between explicit and implicit
    (runtests.jl:381,490); here over 100 closed-loop steps from the README's initial state."""
    z, B = load_tree()
    mpc = mpqp.invpend()
    qp = mpqp.mpc2mpqp(mpc)
    P = ldp.OracleProblem(mpqp.qp2ldp(qp, 1), 1)
    X0 = np.array([[5.0, 5, 0, 0], [1.0, -2, 0.1, 0], [0, 0, 0.15, 0], [10.0, -8, 0.2, 1]]).T
    R = np.zeros((2, X0.shape[1]))
    a = B.simulate_batch(mpc.F, mpc.G, None, X0, R, None, None, 100, 2, 0, 1, want_traj=True)
    b = P.simulate_batch(mpc.F, mpc.G, None, X0, R, None, None, 100, 2, 0, 1, want_traj=True)
    assert np.max(np.abs(a["Utraj"] - b["Utraj"])) < 1e-10
    assert abs(a["Utraj"][0, 0, 0] - 1.7612519326) < 1e-6


@pytest.mark.gpu
def test_gpu_explicit_closed_loop_matches_oracle(lmpc):
    """lmpc_gpu_simulate_bst_batch: tree walk + dynamics per step on chip, bit-identical to the
    oracle's loop; with clipping, a reference trajectory, host and device buffers."""
    import torch
    z, _ = load_tree()
    clip = np.array([[-2.0, 2.0]])
    B = ldp.OracleBst(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clip=clip)
    mpc = mpqp.invpend()
    g = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=clip,
                            dims=mpc.parameter_dims())
    rng = np.random.default_rng(3)
    N, T = 5000, 60
    lo, hi = mpqp.invpend_range(0.2)
    X0 = rng.uniform(lo[:4], hi[:4], (N, 4)).T
    R = np.zeros((2, N))
    out = g.simulate(mpc.F, mpc.G, X0, r=R, N=T, want_traj=True)
    ref = B.simulate_batch(mpc.F, mpc.G, None, X0, R, None, None, T, 2, 0, 1, want_traj=True)
    assert np.array_equal(out["Utraj"], ref["Utraj"]) and np.array_equal(out["Xend"], ref["Xend"])
    assert np.array_equal(out["Ulast"], ref["Ulast"])
    Rt = np.zeros((2, 40)); Rt[0, 20:] = 0.05
    out2 = g.simulate(mpc.F, mpc.G, X0, N=T, want_traj=True, r_traj=Rt)
    Rfull = np.hstack([Rt, np.repeat(Rt[:, -1:], T - 40, axis=1)])
    ref2 = B.simulate_batch(mpc.F, mpc.G, None, X0, None, None, None, T, 2, 0, 1, want_traj=True, Rtraj=Rfull)
    assert np.array_equal(out2["Utraj"], ref2["Utraj"]) and np.array_equal(out2["Xend"], ref2["Xend"])
    assert not np.array_equal(out2["Utraj"], out["Utraj"])
    # the implicit closed loop on the GPU gives the same controls to 1e-10
    gi = lmpc.setup_gpu(mpqp.mpc2mpqp(mpc), nu=1, dims=mpc.parameter_dims())
    imp = gi.simulate(mpc.F, mpc.G, X0, r=R, N=T, want_traj=True)
    assert np.max(np.abs(imp["Utraj"] - out["Utraj"])) < 1e-10


@pytest.mark.gpu
def test_product_explicit_with_feedback_and_with_disturbance_preview(lmpc):
    """(i) prestabilised controller: the explicit laws give v and are corrected by -K x
    (build_tree!, src/explicit.jl:65-71); (ii) disturbance preview: theta carries d_0..d_{Np-1}
    and `control_explicit ~ control_preview` (test/runtests.jl:378-380, atol 1e-10)."""
    from lmpc_b200 import explicit as pex
    # (i) double integrator, regulation, LQR prestabilisation
    m = lmpc.MPC(np.array([[1.0, 1.0], [0.0, 1.0]]), np.array([[0.5], [1.0]]), Np=5, Nc=5)
    m.reference_tracking = False
    m.set_objective(Q=[1.0, 1.0], R=[0.5])
    m.set_bounds(umin=[-1.0], umax=[1.0])
    plain = m.setup()
    m.set_prestabilizing_feedback()
    g = m.setup()
    lo, hi = np.array([-3.0, -1.5]), np.array([3.0, 1.5])
    empc, sol = m.explicit(lo, hi, n_samples=100_000, seed=2)
    assert len(sol["regions"]) > 3
    Th = mpqp.sample_theta(lo, hi, 50_000, seed=9)
    U, _ = empc.evaluate(Th)
    imp = g.solve(Th)
    ok = imp["exitflag"] == 1
    assert ok.mean() > 0.5 and np.max(np.abs(U[:, ok] - np.asarray(imp["U"])[:, ok])) < 1e-9
    base = plain.solve(Th)
    both = ok & (base["exitflag"] == 1)
    assert np.max(np.abs(U[:, both] - np.asarray(base["U"])[:, both])) < 1e-8   # same control as without K
    # (ii) the reference's disturbance-preview case, horizon cut to 3 (the LP-certified tree over the
    # 8-dimensional parameter of Np = 5 takes minutes of host time)
    m = lmpc.MPC(np.array([[1.0, 1.0], [0.0, 1.0]]), np.array([[0.0], [1.0]]), Gd=np.array([[0.0], [1.0]]),
                 C=np.array([[1.0, 0.0]]), Np=3, Nc=3)
    m.set_bounds(umin=[-0.5], umax=[0.5])
    m.set_objective(Q=[10.0], R=[0.1])
    m.disturbance_preview = True
    g = m.setup()
    assert g.dims == (2, 1, 3, 0)
    hi = np.concatenate([[0.5, 0.25], [0.2], np.ones(3)])
    empc, sol = m.explicit(-hi, hi, n_samples=50_000, seed=3)
    u_imp = m.compute_control(np.zeros(2), d=np.array([[0.0, 1.0, 1.0]]))
    u_exp = empc.compute_control(np.zeros(2), d=np.array([0.0, 1.0, 1.0]))
    assert abs(u_imp[0] - u_exp[0]) < 1e-10
    rep = pex.verify(empc, g, -hi, hi, n=50_000)
    # the implicit solve leaves rows violated by less than primal_tol = 1e-6 inactive, the explicit law is
    # exact: next to region boundaries the two differ by that band, nowhere by more
    assert rep["max_err"] < 2e-6 and rep["frac_within_tol"] > 0.999


# ---- product-side explicit builder (linearmpc.jl_b200/explicit.py): soft rows, clipping, codegen arrays ------
class _OracleSolver:
    """stands in for a GpuMPC in the explicit builder on a box without a GPU: region DISCOVERY only (the
    builder computes polyhedra and laws itself); on a GPU box the product's own batched solver does this"""

    def __init__(self, mpc):
        self.qp = mpqp.mpc2mpqp(mpc)
        self.P = ldp.OracleProblem(mpqp.qp2ldp(self.qp, mpc.nu, soft_weight=mpc.soft_weight), mpc.nu)

    def solve(self, Th, want_iters=False, want_active=True):
        r = self.P.solve_batch(Th)
        return dict(U=r["U"], exitflag=r["flag"], active_upper=r["act_u"], active_lower=r["act_l"])


def _product_builder(lmpc):
    import importlib
    return importlib.import_module(lmpc.__name__ + ".explicit")


def _walk(sol, Th, clip=None):
    B = ldp.OracleBst(sol["halfplanes"], sol["feedbacks"], sol["hp_list"], sol["jump_list"], clip=clip)
    return B.evaluate_batch(Th)[0]


def test_explicit_builder_handles_soft_rows(lmpc):
    """(f-2) ExplicitMPC on a controller WITH soft constraints (the reference makes them explicit through the
    slack columns of make_singlesided, src/utils.jl:345-380): double integrator with a soft velocity bound.
    explicit == implicit to 1e-8 on fresh parameters, including those where the soft row is active."""
    from test_condense_cpu import mirror_of
    ex = _product_builder(lmpc)
    o = mpqp.double_integrator_hard(Np=3, Nc=3)
    for c in o.constraints:
        c.soft = True
    m = mirror_of(lmpc, o)
    q = m.condense()
    assert np.count_nonzero(np.asarray(q["senses"]) & 8) >= 2
    lo, hi = np.array([-3.0, -1.5, -2.0]), np.array([3.0, 1.5, 2.0])
    S = _OracleSolver(o)
    sol = ex.build_explicit(q, S, lo, hi, o.nu, n_samples=40000, soft_weight=o.soft_weight)
    Th = mpqp.sample_theta(lo, hi, 20000, seed=17)
    ref = S.P.solve_batch(Th)
    assert np.count_nonzero(ref["flag"] == 2) > 500          # soft rows really are active in the sample
    err = np.abs(_walk(sol, Th) - ref["U"])
    # soft_weight = 1e6 puts cond ~1e6 into both computations: they agree to ~6e-9 where soft rows are active
    assert np.max(err) < 1e-7 and np.mean(err < 1e-8) > 0.999


def test_clipping_aware_tree_is_smaller_and_right_after_clipping(lmpc):
    """build_tree!(...; clipping=true) is the reference default (src/explicit.jl:65): with the input bounds
    as clipping values the leaves only have to be right AFTER clipping, so saturated regions are absorbed."""
    ex = _product_builder(lmpc)
    m, rng, _ = lmpc.mpc_examples("invpend", Np=20, Nc=3)
    o = mpqp.invpend(Np=20, Nc=3)
    q = m.condense()
    lo, hi = np.array(rng[0]) * 0.5, np.array(rng[1]) * 0.5
    lo[-1], hi[-1] = -2.0, 2.0
    S = _OracleSolver(o)
    clip = m.clipping_bounds()
    assert np.array_equal(clip, [[-2.0, 2.0]])
    exact = ex.build_explicit(q, S, lo, hi, 1, n_samples=60000)
    clipped = ex.build_explicit(q, S, lo, hi, 1, n_samples=60000, clipping=clip)
    Th = mpqp.sample_theta(lo, hi, 30000, seed=23)
    ref = S.P.solve_batch(Th)["U"]
    assert np.max(np.abs(_walk(exact, Th) - ref)) < 1e-9
    assert np.max(np.abs(_walk(clipped, Th, clip=clip) - ref)) < 1e-9
    assert len(clipped["hp_list"]) < len(exact["hp_list"]) and clipped["depth"][0] <= exact["depth"][0]


def test_parametricdaqp_codegen_arrays_round_trip(lmpc, tmp_path):
    """The reader of ParametricDAQP.codegen's emitted arrays (src/codegen.jl:62-66; same array style as
    write_float_array, src/codegen.jl:220-236): a tree written in the package's layout and branch order is read
    back bit for bit, its orientation is recognised from the evaluate function in the file, and the committed
    sample (tests/golden/pdaqp_sample.c/.h) parses.  The format follows the package's public generator as
    recollected; it is NOT pinned on a file produced by the real package (none is available here)."""
    ex = _product_builder(lmpc)
    z = np.load(GOLD)
    path = str(tmp_path / "empc.c")
    ex.write_pdaqp_codegen(path, z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], fname="empc")
    t = ex.load_pdaqp_codegen(path)
    assert t["le_child"] == 1 and t["nth"] == 7 and t["nu"] == 1 and t["fname"] == "empc"
    # "%.20f" (the generator's format) keeps 20 decimals: exact to 5e-21 absolute, which is every bit of the
    # entries that matter and a truncation only of numerically-zero ones (5e-19 and below)
    assert np.max(np.abs(t["halfplanes"] - z["halfplanes"])) <= 5e-21 + 1e-16 * 0
    assert np.max(np.abs(t["feedbacks"] - z["feedbacks"])) <= 5e-21
    big = np.abs(z["halfplanes"]) > 1e-3
    assert np.array_equal(t["halfplanes"][big], z["halfplanes"][big])
    # the file holds the tree in the package's branch order: walking it with `id = next + (val <= b)` gives
    # the leaves of the library-order walk
    Th = mpqp.sample_theta(z["lo"], z["hi"], 2000, seed=31)
    _, rego = ldp.OracleBst(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clip=z["clipping"]).evaluate_batch(Th)
    for q in range(0, 2000, 41):
        i = 0
        while t["jump_list"][i] != 0:
            hp = t["halfplanes"][t["hp_list"][i]]
            i = i + t["jump_list"][i] + int(hp[:-1] @ Th[:, q] <= hp[-1])
        assert t["hp_list"][i] == rego[q]
    # library order written as such is recognised as such
    ex.write_pdaqp_codegen(str(tmp_path / "lib.c"), z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"],
                           fname="lib", le_child=0)
    t0 = ex.load_pdaqp_codegen(str(tmp_path / "lib.c"))
    assert t0["le_child"] == 0 and np.array_equal(t0["hp_list"], z["hp_list"]) and np.array_equal(t0["jump_list"], z["jump_list"])
    # the committed sample
    samp = os.path.join(os.path.dirname(GOLD), "pdaqp_sample.c")
    ts = ex.load_pdaqp_codegen(samp)
    assert ts["le_child"] == 1 and ts["halfplanes"].shape[1] == ts["nth"] + 1 and len(ts["hp_list"]) == len(ts["jump_list"])


def _random_tree(rng, nth, nu, n_hp, depth, p_leaf=0.25):
    """Random binary search tree in the ParametricDAQP layout (children of node i at i + jump, i + jump + 1;
    jump == 0 marks a leaf, whose index names the affine law)."""
    hp = rng.normal(size=(n_hp, nth + 1))
    hp[:, -1] *= 0.3
    nodes = []                                  # [jump, index] in breadth-first order

    def grow():
        queue = [(0, 0)]
        nodes.append([0, 0])
        n_fb = 0
        while queue:
            i, d = queue.pop(0)
            if d >= depth or (d > 0 and rng.random() < p_leaf):
                nodes[i] = [0, n_fb]
                n_fb += 1
                continue
            nodes[i] = [len(nodes) - i, int(rng.integers(n_hp))]
            nodes.append([0, 0]); nodes.append([0, 0])
            queue.append((len(nodes) - 2, d + 1)); queue.append((len(nodes) - 1, d + 1))
        return n_fb
    n_fb = grow()
    fb = rng.normal(size=(n_fb, nu, nth + 1))
    nd = np.array(nodes, np.int32)
    return hp, fb, nd[:, 1].copy(), nd[:, 0].copy()


@pytest.mark.gpu
@pytest.mark.parametrize("nth,nu,depth", [(1, 1, 3), (3, 2, 9), (7, 1, 14), (8, 3, 6), (10, 2, 5), (12, 2, 10), (16, 1, 7), (5, 1, 0)])
def test_gpu_evaluate_random_trees_ragged_batches_and_band_hits(lmpc, nth, nu, depth):
    """The lane-refill ring kernel (bst_eval.cu, K4r) on trees it was not tuned on: every nth up to 16, one / two /
    run-time output counts, a leaf root, unbalanced depths, batch sizes around the 32-theta tile, theta that sit ON
    half-planes (inside the fp32 screen's error band: decided by the fp64 re-walk) and non-finite theta (NaN
    compares false everywhere: the walk must still end and agree with the oracle's).  Bit-exact against the oracle."""
    rng = np.random.default_rng(100 * nth + nu)
    hp, fb, hl, jl = _random_tree(rng, nth, nu, n_hp=37, depth=depth, p_leaf=0.3 if depth > 10 else 0.25)
    clip = np.vstack([-0.7 * np.ones(nu), 0.9 * np.ones(nu)]).T.copy() if nu < 3 else None
    B = ldp.OracleBst(hp, fb, hl, jl, clip=clip)
    g = lmpc.GpuExplicitMPC(hp, fb, hl, jl, clipping=clip)
    # the ring needs 8 table replicas and two theta tiles per warp next to the tables: big nth / many laws fall back
    if (nth, nu, depth) != (12, 2, 10):
        assert g.kernel == 2, g.kernel                  # nth = 16 runs the ring with 512-thread CTAs
    for N in (1, 31, 32, 33, 1000, 65 * 1024 + 17, 1_000_003):
        Th = rng.normal(size=(nth, N))
        if N >= 1000:
            # project some theta onto a half-plane of the tree: a'theta - b == 0 up to rounding
            for q in range(0, N, 7):
                a, b = hp[q % len(hp), :-1], hp[q % len(hp), -1]
                Th[:, q] -= a * ((a @ Th[:, q] - b) / (a @ a))
            Th[0, 5] = np.nan
            Th[nth - 1, 6] = np.inf
            Th[0, 8] = -np.inf
            Th[:, 9] = 1e300
        U, reg = g.evaluate(np.asfortranarray(Th))
        Uo, rego = B.evaluate_batch(Th)
        assert np.array_equal(reg, rego), (N, int((reg != rego).sum()))
        assert np.array_equal(U, Uo, equal_nan=True), N
