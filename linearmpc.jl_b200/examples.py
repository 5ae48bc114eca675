"""mpc_examples(name, Np, Nc): the reference's example controllers (src/mpc_examples.jl) for the product's
own builder (`lmpc_b200.MPC`, C++ condensing behind lmpc_gpu_condense), with their parameter ranges and
scenarios.  Only the examples on the batched path's BASELINE configs are mirrored:

  "invpend"   src/mpc_examples.jl:104-141 (linearised at the origin == README.md:36-53)
  "aircraft"  src/mpc_examples.jl:174-208 (AFTI-16; `ks` chooses the steps of the soft output bounds)
  "chain"     not in the reference: BASELINE config 4, built through the public constructor MPC(F, G; ...)
              as SURVEY 8 prescribes (nm unit masses / springs, nm - 1 actuators, Ts = 0.5, Q = R = I)

Returns (mpc, range, scenario): range = (lo, hi) over theta = [x; r; d; uprev] (ParameterRange,
src/types.jl:202-224), scenario = dict(x0, r, d, T) (the example's first Scenario)."""
import numpy as np

from .host import MPC


def zoh(A, B, Ts):
    """src/utils.jl:291-295: exp([A B; 0 0] Ts) by scaling and squaring of a Taylor series (dense, tiny)."""
    A, B = np.atleast_2d(np.asarray(A, float)), np.atleast_2d(np.asarray(B, float))
    nx, nu = B.shape
    M = np.zeros((nx + nu, nx + nu))
    M[:nx, :nx], M[:nx, nx:] = A * Ts, B * Ts
    s = max(0, int(np.ceil(np.log2(max(np.linalg.norm(M, 1), 1e-16)))) + 4)
    Ms = M / 2.0 ** s
    E, term = np.eye(nx + nu), np.eye(nx + nu)
    for k in range(1, 20):
        term = term @ Ms / k
        E = E + term
    for _ in range(s):
        E = E @ E
    return E[:nx, :nx], E[:nx, nx:]


def mpc_examples(name, Np=None, Nc=None, **kw):
    if name == "invpend":
        Np, Nc = Np or 50, Nc or 5
        A = np.array([[0, 1, 0, 0], [0, -10, 9.81, 0], [0, 0, 0, 1], [0, -20, 39.24, 0.0]])
        B = 100 * np.array([[0], [1.0], [0], [2.0]])
        F, G = zoh(A, B, 0.01)
        mpc = MPC(F, G, C=[[1.0, 0, 0, 0], [0, 0, 1.0, 0]], Np=Np, Nc=Nc, Ts=0.01)
        mpc.set_objective(Q=[1.2 ** 2, 1], R=[0.0], Rr=[1.0])
        mpc.set_bounds(umin=[-2.0], umax=[2.0])
        hi = np.concatenate([20 * np.ones(4), 20 * np.ones(2), [2.0]])
        return mpc, (-hi, hi), dict(x0=np.array([0.0, 0.0, 0.15, 0.0]), r=np.zeros(2), d=None, T=2.0)
    if name == "aircraft":
        Np, Nc = Np or 10, Nc or 10
        A = np.array([[-0.0151, -60.5651, 0, -32.174], [-0.0001, -1.3411, 0.9929, 0],
                      [0.00018, 43.2541, -0.86939, 0], [0, 0, 1, 0.0]])
        B = np.array([[-2.516, -13.136], [-0.1689, -0.2514], [-17.251, -1.5766], [0, 0.0]])
        F, G = zoh(A, B, 0.05)
        C = np.array([[0, 1.0, 0, 0], [0, 0, 0, 1.0]]) / np.array([[1.0], [200.0]])
        Dd = np.array([[1.0, 0], [0, 200.0]]) / np.array([[1.0], [200.0]])
        mpc = MPC(F, 50 * G, C=C, Gd=np.zeros((4, 2)), Dd=Dd, Np=Np, Nc=Nc, Ts=0.05)
        mpc.set_objective(Q=np.array([10.0, 10.0]) ** 2, R=np.zeros(2), Rr=np.array([0.1, 0.1]) ** 2)
        mpc.set_input_bounds(umin=[-0.5, -0.5], umax=[0.5, 0.5])
        ks = kw.get("ks", (2,))
        if len(ks):
            mpc.set_output_bounds(ymin=[-0.5, -0.5], ymax=[0.5, 0.5], ks=list(ks))
        hi = np.concatenate([20 * np.ones(4), [1, 0.05], 20 * np.ones(2), [0.5, 0.5]])
        return mpc, (-hi, hi), dict(x0=np.zeros(4), r=np.array([0.1, 0.0]), d=np.zeros(2), T=2.0)
    if name == "chain":
        nm = int(kw.get("nm", 10))
        Np, Nc = Np or 20, Nc or 20
        Fx = -2 * np.eye(nm) + np.eye(nm, k=1) + np.eye(nm, k=-1)
        A = np.block([[np.zeros((nm, nm)), np.eye(nm)], [Fx, np.zeros((nm, nm))]])
        Bf = np.zeros((nm, nm - 1))
        for j in range(nm - 1):
            Bf[j, j], Bf[j + 1, j] = 1.0, -1.0
        F, G = zoh(A, np.vstack([np.zeros((nm, nm - 1)), Bf]), 0.5)
        mpc = MPC(F, G, Np=Np, Nc=Nc, Ts=0.5)
        mpc.reference_tracking = False
        mpc.set_objective(Q=np.ones(2 * nm), R=np.ones(nm - 1), Rr=np.zeros(nm - 1))
        mpc.set_input_bounds(umin=-0.5 * np.ones(nm - 1), umax=0.5 * np.ones(nm - 1))
        hi = 4.0 * np.ones(2 * nm)
        return mpc, (-hi, hi), dict(x0=np.zeros(2 * nm), r=None, d=None, T=10.0)
    raise KeyError(name)
