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
    return z, ldp.OracleBst(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"])


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
    g = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"],
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
    ge = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"],
                             clipping=np.array([[-2.0, 2.0]]), dims=mpc.parameter_dims())
    Th = mpqp.sample_theta(z["lo"], z["hi"], 200000, seed=8)
    Ue, _ = ge.evaluate(Th)
    Ui = gi.solve(Th, want_active=False)["U"]
    assert np.max(np.abs(Ue - Ui)) < 1e-10
    assert np.all(np.abs(Ue) <= 2.0)


@pytest.mark.gpu
def test_gpu_evaluate_device_buffers_and_large_batch(lmpc):
    """BASELINE config 5 shape at 1e7 (1e8 needs 6.4 GB; same code path): device-resident
    evaluate equals the chunked host path; every result within the input bounds."""
    import torch
    z, B = load_tree()
    g = lmpc.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"])
    N = 10_000_000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234)
    lo, hi = torch.tensor(z["lo"], device="cuda"), torch.tensor(z["hi"], device="cuda")
    Th = torch.rand((N, 7), dtype=torch.float64, device="cuda", generator=gen) * (hi - lo) + lo
    U, reg = g.evaluate(Th)
    assert bool((U.abs() <= 2.0 + 1e-9).all())
    sub = Th[:300000].cpu().numpy().T
    Uh, regh = g.evaluate(np.asfortranarray(sub))
    assert np.array_equal(Uh, U[:, :300000].cpu().numpy())
    Uo, _ = B.evaluate_batch(sub[:, :50000])
    assert np.array_equal(Uo, Uh[:, :50000])
