"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's mpQP condensing.

This module restates, for the subset of features used by the BASELINE configs,
what LinearMPC.jl does once per controller:

* ``zoh``                      -> /root/reference/src/utils.jl:291-295
* extended system / weights    -> /root/reference/src/mpc2mpqp.jl:649-731
* state predictor (Phi, Gamma) -> /root/reference/src/mpc2mpqp.jl:20-46
* dense objective              -> /root/reference/src/mpc2mpqp.jl:407-533
* control bounds / general rows-> /root/reference/src/mpc2mpqp.jl:206-245, 249-354
* preprocessing                -> /root/reference/src/mpc2mpqp.jl:733-828
* senses / clamping            -> /root/reference/src/mpc2mpqp.jl:868-899
* qp2ldp (theta-absorbed LDP)  -> /root/reference/src/codegen.jl:239-280
* make_singlesided             -> /root/reference/src/utils.jl:329-391
* form_parameter               -> /root/reference/src/explicit.jl:54-63

It is the *fixture generator and checker* for the CUDA path.  Nothing in the
product package imports it; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may.  Unsupported reference features (prestabilising feedback,
previews, move blocking, offsets, binaries, multiple objectives) raise.

Pinned against the reference's own known answers in tests/test_oracle_golden.py
(1.7612519326 at runtests.jl:62-66, -1 / 1.0 at docs/src/manual/simple.md:107,203,
and the preprocessing result at runtests.jl:1306-1318).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
from scipy.linalg import expm, block_diag, cholesky, solve_triangular

# DAQP sense bits (DAQPBase constants used at mpc2mpqp.jl:868-885)
ACTIVE, LOWER, IMMUTABLE, SOFT, BINARY = 1, 2, 4, 8, 16
EQUALITY = ACTIVE + IMMUTABLE
BIG = 1e30


def zoh(A, B, Ts):
    """Zero-order hold discretisation (utils.jl:291-295)."""
    A = np.atleast_2d(np.asarray(A, float))
    B = np.asarray(B, float).reshape(A.shape[0], -1)
    nx, nu = B.shape
    blk = np.zeros((nx + nu, nx + nu))
    blk[:nx, :nx] = A * Ts
    blk[:nx, nx:] = B * Ts
    E = expm(blk)
    return E[:nx, :nx].copy(), E[:nx, nx:].copy()


def _diagify(v, n):
    """matrixify (utils.jl:297-299): scalar / vector -> diagonal matrix."""
    v = np.asarray(v, float)
    if v.ndim == 0:
        return np.eye(n) * float(v)
    if v.ndim == 1:
        return np.diag(v)
    return v.copy()


@dataclass
class Constraint:
    """lb <= Ax x_k + Au u_k + Ar r + Ad d + Aup u^- <= ub for k in ks (types.jl:4-18)."""
    Au: np.ndarray
    Ax: np.ndarray
    Ar: np.ndarray
    Ad: np.ndarray
    Aup: np.ndarray
    Ap: np.ndarray
    ub: np.ndarray
    lb: np.ndarray
    ks: Sequence[int]  # 1-based, k=1 <-> x0 (as in the reference)
    soft: bool = False
    prio: int = 0


@dataclass
class MPQP:
    """Dense mpQP, same field meaning as types.jl:75-98 (row-major numpy arrays)."""
    H: np.ndarray
    f: np.ndarray
    f_theta: np.ndarray
    A: np.ndarray
    bu: np.ndarray
    bl: np.ndarray
    W: np.ndarray
    senses: np.ndarray
    prio: np.ndarray

    @property
    def n(self):
        return self.H.shape[0]

    @property
    def m(self):
        return self.bu.shape[0]

    @property
    def ms(self):
        return self.bu.shape[0] - self.A.shape[0]

    @property
    def nth(self):
        return self.f_theta.shape[1]


@dataclass
class MPC:
    """Controller description: the slice of types.jl:108-160 the hot path needs."""
    F: np.ndarray
    G: np.ndarray
    C: Optional[np.ndarray] = None
    Gd: Optional[np.ndarray] = None
    Dd: Optional[np.ndarray] = None
    Ts: float = -1.0
    Np: int = 10
    Nc: Optional[int] = None
    Q: Optional[np.ndarray] = None
    R: Optional[np.ndarray] = None
    Rr: Optional[np.ndarray] = None
    Qf: Optional[np.ndarray] = None
    umin: np.ndarray = field(default_factory=lambda: np.zeros(0))
    umax: np.ndarray = field(default_factory=lambda: np.zeros(0))
    constraints: List[Constraint] = field(default_factory=list)
    reference_tracking: bool = True
    # time-varying previews over the horizon (types.jl:56-57, 67-68): theta carries ny*Np / nd*Np values
    reference_preview: bool = False
    disturbance_preview: bool = False
    parameter_preview: bool = False   # types.jl:58,69: Np blocks of generalized parameters in theta
    # types.jl:54,65: with reference_preview, collapse the previewed trajectory to ONE setpoint (theta carries ny
    # values again); traj2setpoint (ny x ny*Np) maps a trajectory to it, set by mpc2mpqp (mpc2mpqp.jl:550-568)
    reference_condensation: bool = False
    traj2setpoint: Optional[np.ndarray] = None
    preprocess_mpqp: bool = True
    soft_weight: float = 1e6
    # generalized-parameter (affine) objective terms (Ex p + ex)'x + (Eu p + eu)'u, setup.jl:118-132
    Ex: Optional[np.ndarray] = None
    ex: Optional[np.ndarray] = None
    Eu: Optional[np.ndarray] = None
    eu: Optional[np.ndarray] = None
    # affine model x+ = F x + G u + Gd d + f_offset, y = C x + Dd d + h_offset, operating point (xo, uo)
    # (model.jl:17-41,61-64); xo / uo only enter the cost of a regulator (reference_tracking off)
    f_offset: Optional[np.ndarray] = None
    h_offset: Optional[np.ndarray] = None
    xo: Optional[np.ndarray] = None
    uo: Optional[np.ndarray] = None

    def __post_init__(self):
        self.F = np.atleast_2d(np.asarray(self.F, float))
        nx = self.F.shape[0]
        self.G = np.asarray(self.G, float).reshape(nx, -1)
        nu = self.G.shape[1]
        self.C = np.eye(nx) if self.C is None else np.atleast_2d(np.asarray(self.C, float))
        ny = self.C.shape[0]
        Gd = np.zeros((nx, 0)) if self.Gd is None else np.asarray(self.Gd, float).reshape(nx, -1)
        Dd = np.zeros((ny, 0)) if self.Dd is None else np.asarray(self.Dd, float).reshape(ny, -1)
        nd = max(Gd.shape[1], Dd.shape[1])  # model.jl:55-57
        self.Gd = np.hstack([Gd, np.zeros((nx, nd - Gd.shape[1]))])
        self.Dd = np.hstack([Dd, np.zeros((ny, nd - Dd.shape[1]))])
        self.nx, self.nu, self.ny, self.nd_model = nx, nu, ny, nd
        if self.Nc is None:
            self.Nc = self.Np
        # default weights: MPCWeights(nu,nx,nr), types.jl:35-38
        self.Q = np.eye(ny) if self.Q is None else _diagify(self.Q, ny)
        self.R = np.eye(nu) if self.R is None else _diagify(self.R, nu)
        self.Rr = np.zeros((nu, nu)) if self.Rr is None else _diagify(self.Rr, nu)
        self.Qf = np.zeros((ny, ny)) if self.Qf is None else _diagify(self.Qf, ny)
        self.umin = np.atleast_1d(np.asarray(self.umin, float))
        self.umax = np.atleast_1d(np.asarray(self.umax, float))
        self.K = np.zeros((nu, nx))
        vec = lambda v, n: np.zeros(n) if v is None or np.size(v) == 0 else np.asarray(v, float).reshape(n)
        self.f_offset, self.h_offset = vec(self.f_offset, nx), vec(self.h_offset, ny)
        self.xo, self.uo = vec(self.xo, nx), vec(self.uo, nu)

    # -- builder API mirroring setup.jl ------------------------------------
    def set_objective(self, Q=None, R=None, Rr=None, Qf=None, Ex=None, ex=None, Eu=None, eu=None):
        """setup.jl:118-132 (subset)."""
        if Ex is not None:
            self.Ex = np.atleast_2d(np.asarray(Ex, float))
        if ex is not None:
            self.ex = np.atleast_1d(np.asarray(ex, float))
        if Eu is not None:
            self.Eu = np.atleast_2d(np.asarray(Eu, float))
        if eu is not None:
            self.eu = np.atleast_1d(np.asarray(eu, float))
        if Q is not None:
            self.Q = _diagify(Q, self.ny)
        if R is not None:
            self.R = _diagify(R, self.nu)
        if Rr is not None:
            self.Rr = _diagify(Rr, self.nu)
        if Qf is not None:
            self.Qf = _diagify(Qf, self.ny)

    def set_input_bounds(self, umin=(), umax=()):
        """setup.jl:36-46."""
        umin = np.atleast_1d(np.asarray(umin, float))
        umax = np.atleast_1d(np.asarray(umax, float))
        nb = max(len(umin), len(umax))
        if nb == 0:
            return
        assert nb == self.nu
        self.umin = np.concatenate([umin, -BIG * np.ones(nb - len(umin))])
        self.umax = np.concatenate([umax, BIG * np.ones(nb - len(umax))])

    def add_constraint(self, Ax=None, Au=None, Ar=None, Ad=None, Aup=None, Ap=None,
                       ub=(), lb=(), ks=None, soft=False, prio=0):
        """setup.jl:61-84."""
        if Ax is None and Au is None:
            return
        ub = np.atleast_1d(np.asarray(ub, float))
        lb = np.atleast_1d(np.asarray(lb, float))
        m = max(len(ub), len(lb))
        if m == 0:
            return
        ub = np.concatenate([ub, BIG * np.ones(m - len(ub))])
        lb = np.concatenate([lb, -BIG * np.ones(m - len(lb))])
        Ax = np.zeros((m, self.nx)) if Ax is None else np.atleast_2d(np.asarray(Ax, float))
        Au = np.zeros((m, self.nu)) if Au is None else np.atleast_2d(np.asarray(Au, float))
        e = np.zeros((0, 0))
        Ar = e if Ar is None else np.atleast_2d(np.asarray(Ar, float))
        Ad = e if Ad is None else np.atleast_2d(np.asarray(Ad, float))
        Aup = e if Aup is None else np.atleast_2d(np.asarray(Aup, float))
        Ap = e if Ap is None else np.atleast_2d(np.asarray(Ap, float))
        ks = list(range(2, self.Np + 1)) if ks is None else list(ks)
        self.constraints.append(Constraint(Au, Ax, Ar, Ad, Aup, Ap, ub, lb, ks, soft, prio))

    def set_output_bounds(self, ymin=(), ymax=(), ks=None, soft=True, prio=0):
        """setup.jl:94-98 (soft by default)."""
        self.add_constraint(Ax=self.C, Ad=self.Dd if self.nd_model > 0 else None,
                            lb=ymin, ub=ymax, ks=ks, soft=soft, prio=prio)

    def set_bounds(self, umin=(), umax=(), ymin=(), ymax=()):
        """setup.jl:105-108."""
        if len(np.atleast_1d(umin)) or len(np.atleast_1d(umax)):
            self.set_input_bounds(umin, umax)
        if len(np.atleast_1d(ymin)) or len(np.atleast_1d(ymax)):
            self.set_output_bounds(ymin, ymax)

    def set_prestabilizing_feedback(self, K=None):
        """setup.jl:186-199: u = v - K x, the QP is over v.  K = None: infinite-horizon LQR gain of
        (F, G, R + Rr, C'QC) (the reference calls `ared`)."""
        if K is None:
            from scipy.linalg import solve_discrete_are
            Rt = self.R + self.Rr
            Qx = self.C.T @ self.Q @ self.C
            P = solve_discrete_are(self.F, self.G, Qx, Rt)
            K = np.linalg.solve(Rt + self.G.T @ P @ self.G, self.G.T @ P @ self.F)
        self.K = np.asarray(K, float).reshape(self.nu, self.nx)

    # -- parameter layout ---------------------------------------------------
    def affine_parameter_dim(self):
        """np: widest generalized-parameter block in use (utils.jl:207-217)."""
        dims = [0]
        for E in (self.Ex, self.Eu):
            if E is not None:
                dims.append(E.shape[1])
        dims += [c.Ap.shape[1] for c in self.constraints if c.Ap.size]
        return max(dims)

    def parameter_dims(self):
        """(nx, nr, nd, nuprev) -- mpc2mpqp.jl:147-164 (previews: Np blocks each); np via affine_parameter_dim."""
        nr = self.ny if self.reference_tracking else 0
        if self.reference_preview and not self.reference_condensation and nr > 0:
            nr *= self.Np
        nd = self.nd_model
        if self.disturbance_preview and nd > 0:
            nd *= self.Np
        aup = any(c.Aup.size and np.any(c.Aup != 0) for c in self.constraints)
        nuprev = self.nu if (np.any(self.Rr != 0) or aup) else 0
        return self.nx, nr, nd, nuprev

    def form_parameter(self, x, r=None, d=None, uprev=None, p=None):
        """theta = [x; r; d; uprev; p] (explicit.jl:54-63); previews flattened column by column with
        the last column held when the trajectory is short (utils.jl:78-134, 155-205)."""
        nx, nr, nd, nup = self.parameter_dims()
        x = np.asarray(x, float).reshape(nx)
        if self.reference_preview and nr > 0:
            r = _flatten_preview(r, self.ny, self.Np)
            if self.reference_condensation:          # condense_reference, utils.jl:140-147
                r = self.traj2setpoint @ r
        if self.disturbance_preview and nd > 0:
            d = _flatten_preview(d, self.nd_model, self.Np)
        r = np.zeros(nr) if r is None else np.asarray(r, float).reshape(-1)[:nr]
        d = np.zeros(nd) if d is None else np.asarray(d, float).reshape(-1)[:nd]
        up = np.zeros(nup) if uprev is None else np.asarray(uprev, float).reshape(-1)[:nup]
        return np.concatenate([x, r, d, up, self.format_affine_parameters(p)])

    def format_affine_parameters(self, p):
        """utils.jl:225-259: a vector of np values is held over the horizon, an np x cols matrix is cut to Np
        columns or extended by holding its last column (first column only without parameter_preview)."""
        nb = self.affine_parameter_dim()
        tot = nb * self.Np if self.parameter_preview else nb
        if tot == 0:
            return np.zeros(0)
        if p is None:
            return np.zeros(tot)
        p = np.asarray(p, float)
        if p.ndim == 1 and p.size == nb:
            return np.tile(p, self.Np) if self.parameter_preview else p
        if p.ndim == 1 and p.size == tot:
            return p
        if p.ndim == 2:
            if p.shape[0] != nb:
                raise ValueError(f"Generalized parameter matrix must have {nb} rows")
            return _flatten_preview(p, nb, self.Np) if self.parameter_preview else p[:, 0].copy()
        raise ValueError("Generalized parameters must be a vector or matrix")


def _flatten_preview(v, rows, Np):
    """vec of a rows x Np preview: a vector is repeated, a short matrix holds its last column."""
    if v is None:
        return np.zeros(rows * Np)
    v = np.asarray(v, float)
    if v.ndim == 1:
        if v.size != rows:
            raise ValueError(f"preview vector length ({v.size}) must match ({rows})")
        return np.tile(v, Np)
    if v.shape[0] != rows:
        raise ValueError(f"preview matrix must have {rows} rows")
    if v.shape[1] >= Np:
        return v[:, :Np].reshape(-1, order="F")
    ext = np.hstack([v, np.repeat(v[:, -1:], Np - v.shape[1], axis=1)])
    return ext.reshape(-1, order="F")


# ---------------------------------------------------------------------------
# condensing
# ---------------------------------------------------------------------------

def extended_system(mpc: MPC):
    """x~ = [x; r; d; u^-]  (mpc2mpqp.jl:649-690 with K = 0, no offset); previewed r / d are not states."""
    nx, nr, nd, nup = mpc.parameter_dims()
    nu = mpc.nu
    K = mpc.K
    F, G, C = mpc.F - mpc.G @ K, mpc.G.copy(), mpc.C.copy()
    if mpc.reference_preview:
        nr = 0
    if mpc.disturbance_preview:
        nd = 0
    if nr > 0:
        F = block_diag(F, np.eye(mpc.ny))
        G = np.vstack([G, np.zeros((mpc.ny, nu))])
        C = np.hstack([C, -np.eye(mpc.ny)])
    if nd > 0:
        F = block_diag(F, np.eye(nd))
        F[:nx, -nd:] = mpc.Gd
        G = np.vstack([G, np.zeros((nd, nu))])
        C = np.hstack([C, mpc.Dd])
    if nup > 0:
        F = block_diag(F, np.zeros((nu, nu)))
        F[-nu:, :nx] = -K                      # u^- of the next step is v - K x
        G = np.vstack([G, np.eye(nu)])
        nye, nxe = C.shape
        C = np.vstack([np.hstack([C, np.zeros((nye, nu))]),
                       np.hstack([K, np.zeros((nu, nxe - nx)), np.eye(nu)])])
    if np.any(mpc.R != 0) and np.any(K != 0):  # u'Ru with u = v - K x: output K x weighted by R
        C = np.vstack([C, np.hstack([K, np.zeros((nu, C.shape[1] - nx))])])
    return F, G, C


def extended_cost(mpc: MPC):
    """Q~, R~, S~, Qf~  (mpc2mpqp.jl:692-731 with K = 0)."""
    nx, nr, nd, nup = mpc.parameter_dims()
    nu = mpc.nu
    Q, R, Rr = mpc.Q.copy(), mpc.R.copy(), mpc.Rr.copy()
    Qf = Q.copy() if not np.any(mpc.Qf != 0) else mpc.Qf.copy()
    S = np.zeros((nx, nu))
    if mpc.reference_preview:
        nr = 0
    if mpc.disturbance_preview:
        nd = 0
    if nr > 0:
        S = np.vstack([S, np.zeros((mpc.ny, nu))])
    if nd > 0:
        S = np.vstack([S, np.zeros((nd, nu))])
    if nup > 0:
        Q = block_diag(Q, Rr)
        Qf = block_diag(Qf, np.zeros((nu, nu)))
        S = np.vstack([S, -Rr])
        S[:nx] -= mpc.K.T @ Rr
        R = R + Rr
    if np.any(mpc.R != 0) and np.any(mpc.K != 0):
        Q = block_diag(Q, mpc.R)
        Qf = block_diag(Qf, np.zeros((nu, nu)))
        S[:nx] -= mpc.K.T @ mpc.R
    return Q, R, S, Qf


def state_predictor(F, G, Np, Nc):
    """X = Phi x0 + Gamma U with u_k = u_{Nc-1} for k >= Nc (mpc2mpqp.jl:20-46)."""
    nx, nu = G.shape
    Phi = np.zeros(((Np + 1) * nx, nx))
    Gam = np.zeros(((Np + 1) * nx, Nc * nu))
    Phi[:nx] = np.eye(nx)
    for k in range(Np):
        prev = slice(k * nx, (k + 1) * nx)
        cur = slice((k + 1) * nx, (k + 2) * nx)
        Phi[cur] = F @ Phi[prev]
        Gam[cur] = F @ Gam[prev]
        j = min(k, Nc - 1)
        Gam[cur, j * nu:(j + 1) * nu] += G
    return Phi, Gam


def disturbance_predictor(F, E, Np):
    """Psi with X = ... + Psi [d_0; ...; d_{Np-1}] (mpc2mpqp.jl:48-58)."""
    nx, nd = E.shape
    Psi = np.zeros(((Np + 1) * nx, Np * nd))
    for k in range(1, Np + 1):
        Psi[k * nx:(k + 1) * nx] = F @ Psi[(k - 1) * nx:k * nx]
        Psi[k * nx:(k + 1) * nx, (k - 1) * nd:k * nd] += E
    return Psi


def dense_objective(mpc: MPC, F, G, C, Phi, Gam):
    """H, f, f_theta (mpc2mpqp.jl:407-533, single objective, no previews/offsets)."""
    Q, R, S, Qf = extended_cost(mpc)
    N, Nc = mpc.Np, mpc.Nc
    nxe, nu = G.shape
    posQ = np.where(np.diag(Q) > 0)[0]
    Qp, Cp = Q[np.ix_(posQ, posQ)], C[posQ, :]
    posF = np.where(np.diag(Qf) > 0)[0]
    Qfp, Cf = Qf[np.ix_(posF, posF)], C[posF, :]

    H = np.kron(np.eye(Nc), R)
    H[-nu:, -nu:] += (N - Nc) * R
    f0 = np.zeros(Nc * nu)
    regulator = not mpc.reference_tracking
    if regulator and np.any(mpc.uo != 0):          # nonzero operating point (mpc2mpqp.jl:433-443)
        Uo = np.tile(mpc.uo, Nc)
        f0 -= H @ Uo
        if np.any(mpc.K != 0) and np.any(R != 0):
            KR = np.vstack([-mpc.K.T @ R, np.zeros((nxe - mpc.nx, nu))])
            KRtot = np.vstack([np.kron(np.eye(Nc), KR), np.zeros(((N - Nc + 1) * nxe, Nc * nu))])
            if N > Nc:
                KRtot[Nc * nxe:N * nxe, -nu:] = np.tile(KR, (N - Nc, 1))
            GKR = Gam.T @ KRtot
            f0 -= (GKR + GKR.T) @ Uo
    CQC = Cp.T @ Qp @ Cp
    CQCtot = np.kron(np.eye(N + 1), CQC)
    CQCtot[-nxe:, -nxe:] = Cf.T @ Qfp @ Cf
    H = H + Gam.T @ CQCtot @ Gam
    f_theta = Gam.T @ CQCtot @ Phi
    if regulator and np.any(mpc.xo != 0):          # mpc2mpqp.jl:457-459
        f0 -= Gam.T @ CQCtot @ np.tile(np.concatenate([mpc.xo, np.zeros(nxe - mpc.nx)]), N + 1)
    if np.any(S != 0):
        Stot = np.vstack([np.kron(np.eye(Nc), S), np.zeros(((N - Nc + 1) * nxe, Nc * nu))])
        if N > Nc:
            Stot[Nc * nxe:N * nxe, -nu:] = np.tile(S, (N - Nc, 1))
        GS = Gam.T @ Stot
        H = H + GS + GS.T
        f_theta = f_theta + Stot.T @ Phi
    ny = mpc.ny
    nxp, nrp, ndp, _ = mpc.parameter_dims()
    Cfull, Qfull, Qffull = C[:ny, :], Q[:ny, :ny], Qf[:ny, :ny]
    if nrp > 0 and mpc.reference_preview:
        # ref_preview_cost (mpc2mpqp.jl:535-577, no condensation): block k of the preview is the
        # reference of x_{k+1}; the reference paired with x_0 does not touch U and is dropped
        CQ = np.kron(np.eye(N + 1), Cfull.T @ Qfull)
        CQ[-nxe:, -ny:] = Cfull.T @ Qffull
        Fr = (-Gam.T @ CQ)[:, ny:]
        if mpc.reference_condensation:
            # mpc2mpqp.jl:550-568: the setpoint whose (weighted) effect on the unconstrained optimum -H^-1 Fr r is
            # closest to the trajectory's; the first control move is weighted 1e6
            Is = np.tile(np.eye(ny), (N, 1))
            Wc = np.eye(nu * Nc)
            Wc[np.arange(nu), np.arange(nu)] = 1e6
            WinvHFr = (Wc @ np.linalg.inv(H)) @ Fr
            mpc.traj2setpoint = np.linalg.lstsq(WinvHFr @ Is, WinvHFr, rcond=None)[0]
            Fr = Fr @ Is
        f_theta = np.hstack([f_theta[:, :nxp], Fr, f_theta[:, nxp:]])
    if ndp > 0 and mpc.disturbance_preview:
        # disturbance_preview_cost (mpc2mpqp.jl:579-610): y_k = C x_k + Dd d_{k-1}, x through Psi
        Psi = disturbance_predictor(F, np.vstack([mpc.Gd, np.zeros((nxe - mpc.nx, mpc.nd_model))]), N)
        CY = np.kron(np.eye(N), Cfull)
        Gy = CY @ Gam[nxe:, :]
        Yd = CY @ Psi[nxe:, :] + np.kron(np.eye(N), mpc.Dd[:ny, :])
        Qy = np.kron(np.eye(N), Qfull)
        Qy[-ny:, -ny:] = Qffull
        Fd = Gy.T @ Qy @ Yd
        split = nxp + nrp
        f_theta = np.hstack([f_theta[:, :split], Fd, f_theta[:, split:]])
    H = (H + H.T) / 2
    f = f0
    # generalized-parameter terms (mpc2mpqp.jl:478-508), constant parameter across the horizon
    npar = mpc.affine_parameter_dim()
    nxm = mpc.nx
    Ex = np.zeros((nxm, npar)) if mpc.Ex is None else mpc.Ex
    ex = np.zeros(nxm) if mpc.ex is None else mpc.ex
    Eu = np.zeros((nu, npar)) if mpc.Eu is None else mpc.Eu
    eu = np.zeros(nu) if mpc.eu is None else mpc.eu
    if Ex.shape != (nxm, npar) or Eu.shape != (nu, npar):    # mpc2mpqp.jl:486-490
        raise ValueError(f"Affine objective matrices Ex / Eu must have {npar} columns")
    # stage_parameter_matrix (mpc2mpqp.jl:148): one shared p, or stage k <-> p block k with parameter_preview
    stage = (lambda E: np.kron(np.eye(N), E)) if mpc.parameter_preview else (lambda E: np.tile(E, (N, 1)))
    Umap = np.kron(np.vstack([np.eye(Nc), np.zeros((N - Nc, Nc))]), np.eye(nu))
    f = f + Umap.T @ np.tile(eu, N)
    xsel = np.hstack([np.eye(nxm), np.zeros((nxm, nxe - nxm))])
    Gx = np.kron(np.eye(N), xsel) @ Gam[nxe:, :]
    f = f + Gx.T @ np.tile(ex, N)
    if npar > 0:
        Fp = Umap.T @ stage(Eu) + Gx.T @ stage(Ex)
        f_theta = np.hstack([f_theta, Fp])
    return H, f, f_theta


def control_bounds(mpc: MPC, nth, Phi=None, Gam=None):
    """Leading simple rows, or with a prestabilising feedback the general rows
    lb <= v_k - K x_k <= ub, i.e. A = I - K~ Gamma, W = K~ Phi (mpc2mpqp.jl:206-245)."""
    Nc, nu, nx = mpc.Nc, mpc.nu, mpc.nx
    bu = np.tile(mpc.umax, Nc)
    bl = np.tile(mpc.umin, Nc)
    W = np.zeros((Nc * nu, nth))
    A = np.zeros((0, Nc * nu))
    if np.any(mpc.K != 0):
        nxe = Phi.shape[1]
        Kfull = np.kron(np.eye(Nc), np.hstack([mpc.K, np.zeros((nu, nxe - nx))]))
        A = np.eye(Nc * nu) - Kfull @ Gam[:Nc * nxe, :]
        We = Kfull @ Phi[:Nc * nxe, :]
        _, nrp, ndp, _ = mpc.parameter_dims()
        nrx = 0 if mpc.reference_preview else nrp
        ndx = 0 if mpc.disturbance_preview else ndp
        parts = [We[:, :nx], np.zeros((Nc * nu, nrp)) if mpc.reference_preview else We[:, nx:nx + nrx]]
        if mpc.disturbance_preview and ndp > 0:
            Psi = disturbance_predictor(Phi[nxe:2 * nxe, :],
                                        np.vstack([mpc.Gd, np.zeros((nxe - nx, mpc.nd_model))]), mpc.Np)
            parts.append(Kfull @ Psi[:Nc * nxe, :])
        else:
            parts.append(We[:, nx + nrx:nx + nrx + ndx])
        parts.append(We[:, nx + nrx + ndx:])
        W = np.hstack(parts + [np.zeros((Nc * nu, nth - sum(p.shape[1] for p in parts)))])
    return bu, bl, W, A


def general_constraints(mpc: MPC, Phi, Gam):
    """Rows lb <= [Ax Ar Ad Aup] x~_{k-1} + Au u_{k-1} <= ub (mpc2mpqp.jl:249-354)."""
    Np, Nc, nu = mpc.Np, mpc.Nc, mpc.nu
    nx, nr, nd, nup = mpc.parameter_dims()
    nrp, ndp = nr, nd
    if mpc.reference_preview:
        nr = 0
    if mpc.disturbance_preview:
        nd = 0
    nxe = nx + nr + nd + nup
    ndm = mpc.nd_model
    Wd_direct = []
    eyeX = np.eye(Np + 1)
    eyeU = np.vstack([np.eye(Nc), np.zeros((1 + Np - Nc, Nc))])
    Axt, Aut = np.zeros((0, nxe * (Np + 1))), np.zeros((0, nu * Nc))
    ub, lb, soft, prio = [], [], [], []
    for c in mpc.constraints:
        mi = c.Au.shape[0]
        kmax = Np + 1 if not np.any(c.Au != 0) else Np
        ks = [k for k in c.ks if k <= kmax]
        Ni = len(ks)
        Ar = c.Ar if (c.Ar.size and nr > 0) else np.zeros((mi, nr))
        Ad = c.Ad if (c.Ad.size and nd > 0) else np.zeros((mi, nd))
        Aup = c.Aup if c.Aup.size else np.zeros((mi, nup))
        row = np.hstack([c.Ax - c.Au @ mpc.K, Ar, Ad, Aup])
        sel = [k - 1 for k in ks]
        Aut = np.vstack([Aut, np.kron(eyeU[sel, :], c.Au)])
        Axt = np.vstack([Axt, np.kron(eyeX[sel, :], row)])
        ub.append(np.tile(c.ub, Ni))
        lb.append(np.tile(c.lb, Ni))
        soft += [c.soft] * (mi * Ni)
        prio += [c.prio] * (mi * Ni)
        if mpc.reference_preview and nrp > 0 and c.Ar.size and np.any(c.Ar != 0):
            raise ValueError("constraints on the reference are not supported with reference_preview")
        if mpc.disturbance_preview and ndp > 0:
            # disturbance_preview_direct (mpc2mpqp.jl:94-109): row block of x_{k-1} sees -Ad d_{k-2}
            Wdd = np.zeros((mi * Ni, ndm * Np))
            if c.Ad.size:
                for i, k in enumerate(ks):
                    if 2 <= k <= Np + 1:
                        Wdd[i * mi:(i + 1) * mi, (k - 2) * ndm:(k - 1) * ndm] = -c.Ad
            Wd_direct.append(Wdd)
    A = Axt @ Gam + Aut
    W = -Axt @ Phi
    # insert_preview_parameter_blocks (mpc2mpqp.jl:70-92)
    if mpc.reference_preview and nrp > 0:
        W = np.hstack([W[:, :nx], np.zeros((W.shape[0], nrp)), W[:, nx:]])
    if mpc.disturbance_preview and ndp > 0:
        Fx = Phi[nxe:2 * nxe, :]   # one-step map of the extended state
        Psi = disturbance_predictor(Fx, np.vstack([mpc.Gd, np.zeros((nxe - nx, ndm))]), Np)
        Wd = -Axt @ Psi + np.vstack(Wd_direct)
        split = nx + nrp + nd
        W = np.hstack([W[:, :split], Wd, W[:, split:]])
    npar = mpc.affine_parameter_dim()
    if npar > 0:   # parameter_preview_direct without preview (mpc2mpqp.jl:125-143): W_p = -Ap per row
        Wp = []
        for c in mpc.constraints:
            mi = c.Au.shape[0]
            kmax = Np + 1 if not np.any(c.Au != 0) else Np
            kk = [k for k in c.ks if 1 <= k <= kmax]
            Apm = c.Ap if c.Ap.size else np.zeros((mi, npar))
            if Apm.shape[1] != npar:
                raise ValueError(f"Ap must have {npar} columns")
            if mpc.parameter_preview:      # stage k sees p block min(k, Np)
                blk = np.zeros((mi * len(kk), npar * Np))
                for i, k in enumerate(kk):
                    cid = min(k, Np) - 1
                    blk[i * mi:(i + 1) * mi, cid * npar:(cid + 1) * npar] = -Apm
                Wp.append(blk)
            else:
                Wp.append(np.tile(-Apm, (len(kk), 1)))
        W = np.hstack([W, np.vstack(Wp)])
    return (A, np.concatenate(ub), np.concatenate(lb), W,
            np.array(soft, bool), np.array(prio, int))


def _remove_redundant(A, bu, bl, W, soft, prio):
    """mpc2mpqp.jl:733-773 -- zero rows dropped, sign-canonical rows, single-entry
    rows folded into simple bounds; rows are normalised ONLY when something was
    removed (the reference's behaviour)."""
    A, bu, bl, W = A.copy(), bu.copy(), bl.copy(), W.copy()
    ns = len(bu) - A.shape[0]
    keep = list(range(ns))
    scale = [1.0] * ns
    for i in range(A.shape[0]):
        a = A[i]
        idx = ns + i
        nf = np.linalg.norm(a)
        if nf > 1e-10:
            nz = int(np.argmax(np.abs(a) > 1e-12))
            if a[nz] < 0:
                a[:] = -a + 0.0
                bu[idx], bl[idx] = -bl[idx], -bu[idx]
                W[idx] = -W[idx] + 0.0
            if np.count_nonzero(np.abs(a) > 1e-12) == 1:
                if nz < ns and prio[nz] == prio[idx] and soft[nz] == soft[idx] \
                        and not np.any(W[idx] - W[nz]):
                    bu[nz] = min(bu[nz], bu[idx] / nf)
                    bl[nz] = max(bl[nz], bl[idx] / nf)
                    continue
            keep.append(idx)
            scale.append(1.0 / nf)
    if len(keep) < len(bu):
        keep = np.array(keep, int)
        scale = np.array(scale)
        keepA = keep[ns:] - ns
        A = A[keepA] * scale[ns:, None]
        bu, bl, W = bu[keep] * scale, bl[keep] * scale, W[keep] * scale[:, None]
        soft, prio = soft[keep], prio[keep]
    return A, bu, bl, W, soft, prio


def _remove_duplicate(A, bu, bl, W, soft, prio):
    """mpc2mpqp.jl:775-828 -- merge general rows equal to 6 digits."""
    ns = len(bu) - A.shape[0]
    ext = np.hstack([A, W[ns:], soft[ns:, None].astype(float), prio[ns:, None].astype(float)])
    groups, order = {}, []
    for i in range(ext.shape[0]):
        key = tuple(np.round(ext[i], 6) + 0.0)
        if key not in groups:
            groups[key] = []
            order.append(key)
        groups[key].append(i)
    if len(order) == A.shape[0]:
        return A, bu, bl, W, soft, prio
    first = np.array([groups[k][0] for k in order], int)
    A2 = A[first]
    bu2 = np.concatenate([bu[:ns], [min(bu[ns + j] for j in groups[k]) for k in order]])
    bl2 = np.concatenate([bl[:ns], [max(bl[ns + j] for j in groups[k]) for k in order]])
    W2 = np.vstack([W[:ns], W[ns + first]])
    soft2 = np.concatenate([soft[:ns], soft[ns + first]])
    prio2 = np.concatenate([prio[:ns], prio[ns + first]])
    return A2, bu2, bl2, W2, soft2, prio2


def format_move_block(block, Np):
    """setup.jl:235-251: pad / clip a block list to the prediction horizon."""
    b = [int(v) for v in block]
    tot = sum(b)
    if tot < Np:
        b[-1] += Np - tot
    elif tot > Np:
        acc, i = 0, 0
        while True:
            acc += b[i]
            if acc >= Np:
                break
            i += 1
        b = b[:i + 1]
        b[-1] += Np - acc
    return b


def move_block(mpc: "MPC", blocks):
    """move_block!(mpc, blocks) (setup.jl:196-233): blocks = one list per input."""
    mpc.move_blocks = [format_move_block(b, mpc.Np) for b in blocks]
    mpc.Nc = max(sum(b[:-1]) for b in mpc.move_blocks) + 1


def apply_move_block(mpc, H, f, f_theta, A, bu, bl, W, soft, prio):
    """apply_move_block (mpc2mpqp.jl:830-857): U_old = T z, blocks interleaved pass by pass, the
    last block of every input clipped to one step, only the first bound of each block kept."""
    nu, Nc = mpc.nu, mpc.Nc
    nub = len(mpc.umax)
    n_old = nu * Nc
    n_new = sum(len(mb) for mb in mpc.move_blocks)
    T = np.zeros((n_old, n_new))
    counter = list(range(1, nu + 1))           # 1-based, as in the reference
    keep, new_id = [], 1
    for p in range(1, max(len(mb) for mb in mpc.move_blocks) + 1):
        for iu, mb in enumerate(mpc.move_blocks):
            if len(mb) < p:
                continue
            block = mb[p - 1] if len(mb) != p else 1
            rows = np.arange(counter[iu], counter[iu] + nu * (block - 1) + 1, nu)
            T[rows - 1, new_id - 1] = 1.0
            if counter[iu] <= nub * Nc:
                keep.append(counter[iu])
            counter[iu] += nu * block
            new_id += 1
    nb = nub * Nc                               # control-bound rows: simple, or the top of A with a feedback
    keep = np.array(keep, int) - 1
    full = np.concatenate([keep, np.arange(nb, len(bu))]).astype(int)
    if np.any(mpc.K != 0) and nb:               # mpc2mpqp.jl:854: the bounds are rows of A then
        A = A[np.concatenate([keep, np.arange(nb, A.shape[0])]).astype(int)]
    return (T.T @ H @ T, T.T @ f, T.T @ f_theta, A @ T, bu[full], bl[full], W[full], soft[full], prio[full])


def _with_offset_channel(mpc: MPC):
    """A constant dynamics offset f_offset is one more state that stays at 1 (mpc2mpqp.jl:684-689): here it
    rides as an extra measured-disturbance channel d_off = 1 with Gd column f_offset and Dd column 0, which
    gives the same columns of f_theta and W (only their position in the extended state differs); the
    caller collapses them into f, bu and bl (mpc2mpqp.jl:395-399, 518-522)."""
    import copy
    aug = copy.copy(mpc)
    ndm = mpc.nd_model
    aug.Gd = np.hstack([mpc.Gd, mpc.f_offset.reshape(-1, 1)])
    aug.Dd = np.hstack([mpc.Dd, np.zeros((mpc.ny, 1))])
    aug.nd_model = ndm + 1
    aug.f_offset = np.zeros(mpc.nx)
    aug.constraints = []
    for c in mpc.constraints:
        c2 = copy.copy(c)
        if c.Ad.size:
            c2.Ad = np.hstack([c.Ad, np.zeros((c.Ad.shape[0], 1))])
        aug.constraints.append(c2)
    nx, nr, _, _ = mpc.parameter_dims()
    if mpc.disturbance_preview:
        cols = [nx + nr + k * (ndm + 1) + ndm for k in range(mpc.Np)]
    else:
        cols = [nx + nr + ndm]
    return aug, cols


def mpc2mpqp(mpc: MPC) -> MPQP:
    """The dense mpQP of `mpc` (mpc2mpqp.jl:612-647)."""
    offset_cols = None
    if np.any(mpc.f_offset != 0):
        if mpc.affine_parameter_dim() > 0:
            # the reference collapses the LAST column of f_theta / W, which is a p column then
            raise ValueError("f_offset together with generalized parameters is not supported")
        orig = mpc
        mpc, offset_cols = _with_offset_channel(mpc)
    F, G, C = extended_system(mpc)
    Phi, Gam = state_predictor(F, G, mpc.Np, mpc.Nc)
    H, f, f_theta = dense_objective(mpc, F, G, C, Phi, Gam)
    n = H.shape[0]
    nth = sum(mpc.parameter_dims()) + mpc.affine_parameter_dim() * (mpc.Np if mpc.parameter_preview else 1)
    A = np.zeros((0, n))
    if len(mpc.umax):
        bu, bl, W, A = control_bounds(mpc, nth, Phi, Gam)
        soft = np.zeros(n, bool)
        prio = np.zeros(n, int)
    else:
        bu, bl, W = np.zeros(0), np.zeros(0), np.zeros((0, nth))
        soft, prio = np.zeros(0, bool), np.zeros(0, int)
    if mpc.constraints:
        Ag, bug, blg, Wg, sg, pg = general_constraints(mpc, Phi, Gam)
        A = np.vstack([A, Ag])
        bu, bl, W = np.concatenate([bu, bug]), np.concatenate([bl, blg]), np.vstack([W, Wg])
        soft, prio = np.concatenate([soft, sg]), np.concatenate([prio, pg])
    if offset_cols is not None:                     # collapse the constant channel (value 1)
        f = f + f_theta[:, offset_cols].sum(axis=1)
        bu, bl = bu + W[:, offset_cols].sum(axis=1), bl + W[:, offset_cols].sum(axis=1)
        f_theta, W = np.delete(f_theta, offset_cols, axis=1), np.delete(W, offset_cols, axis=1)
        mpc = orig
    nxp, nrp, _, _ = mpc.parameter_dims()
    if nrp > 0 and np.any(mpc.h_offset != 0):       # r - h_offset is what the outputs track (mpc2mpqp.jl:523-530)
        ho = np.tile(mpc.h_offset, mpc.Np) if (mpc.reference_preview and not mpc.reference_condensation) else mpc.h_offset
        f = f - f_theta[:, nxp:nxp + nrp] @ ho
    if getattr(mpc, "move_blocks", None):
        H, f, f_theta, A, bu, bl, W, soft, prio = apply_move_block(mpc, H, f, f_theta, A, bu, bl, W, soft, prio)
    # sort general rows by priority, stable (mpc2mpqp.jl:859-866)
    ns = len(bu) - A.shape[0]
    order = np.argsort(prio[ns:], kind="stable")
    A = A[order]
    full = np.concatenate([np.arange(ns), order + ns])
    bu, bl, W, soft, prio = bu[full], bl[full], W[full], soft[full], prio[full]
    if mpc.preprocess_mpqp:
        A, bu, bl, W, soft, prio = _remove_redundant(A, bu, bl, W, soft, prio)
        A, bu, bl, W, soft, prio = _remove_duplicate(A, bu, bl, W, soft, prio)
    senses = np.zeros(len(bu), np.int32)
    for i in range(len(bu)):
        if bu[i] > 1e20 and bl[i] < -1e20:
            senses[i] = IMMUTABLE
        elif abs(bu[i] - bl[i]) < 1e-12:
            senses[i] = EQUALITY
    senses[soft] += SOFT
    bu = np.clip(bu, -BIG, BIG)
    bl = np.clip(bl, -BIG, BIG)
    return MPQP(H, f, f_theta, A, bu, bl, W, senses, prio.astype(np.int32))


# ---------------------------------------------------------------------------
# LDP constants (the device layout)
# ---------------------------------------------------------------------------

@dataclass
class LDP:
    """theta-absorbed least-distance problem (codegen.jl:239-280).

    min 1/2 ||v||^2  s.t.  dl + Dth theta <= Mbar v <= du + Dth theta
    control = uscaling * (Rinv_n v)[0:nu]/... see ``control_from_v``.
    """
    Mbar: np.ndarray      # m x n, unit rows; first ms rows = normalised rows of R^-1
    Dth: np.ndarray       # m x nth (row-major)
    du: np.ndarray
    dl: np.ndarray
    Rinv: np.ndarray      # n x n upper triangular, un-normalised
    Uth_offset: np.ndarray  # nu x nth
    u_offset: np.ndarray  # nu
    norm: np.ndarray      # m row norms of [I_ms; A] R^-1
    senses: np.ndarray
    ms: int
    rho_soft: float


def qp2ldp(qp: MPQP, nu: int, K: Optional[np.ndarray] = None, soft_weight: float = 1e6) -> LDP:
    n, ms = qp.n, qp.ms
    R = cholesky((qp.H + qp.H.T) / 2, lower=False)  # H = R'R, R upper
    Rinv = solve_triangular(R, np.eye(n), lower=False)
    Mext = np.vstack([np.eye(n)[:ms], qp.A]) @ Rinv
    Vth = solve_triangular(R.T, qp.f_theta, lower=True)
    v = solve_triangular(R.T, qp.f, lower=True)
    Dth = qp.W + Mext @ Vth
    dd = Mext @ v
    du, dl = qp.bu + dd, qp.bl + dd
    nrm = np.linalg.norm(Mext, axis=1)
    s = np.where(nrm > 0, nrm, 1.0)
    Mbar = Mext / s[:, None]
    Dth = Dth / s[:, None]
    du, dl = du / s, dl / s
    Uth = -np.linalg.solve(qp.H, qp.f_theta)[:nu]
    if K is not None:
        Uth[:, :K.shape[1]] -= K
    uo = -np.linalg.solve(qp.H, qp.f)[:nu]
    return LDP(Mbar, Dth, du, dl, Rinv, Uth, uo, nrm, qp.senses.copy(), ms, 1.0 / soft_weight)


def make_singlesided(qp: MPQP, soft_weight=1e6, single_soft=True):
    """A U <= b + W theta with explicit soft slack (utils.jl:329-391)."""
    m, n, ms = qp.m, qp.n, qp.ms
    A0 = np.vstack([np.hstack([np.eye(ms), np.zeros((ms, n - ms))]), qp.A])
    A = np.vstack([A0, -A0])
    senses = np.tile(qp.senses, 2)
    H, f, f_theta = qp.H, qp.f, qp.f_theta
    softm = (qp.senses & SOFT) == SOFT
    if softm.any():
        ids = np.where(softm)[0]
        R = cholesky((H + H.T) / 2, lower=False)
        Ms = solve_triangular(R.T, A0[softm].T, lower=True).T
        nf = np.linalg.norm(Ms, axis=1)
        if single_soft:
            col = np.zeros((2 * m, 1))
            col[ids, 0] = -nf
            col[ids + m, 0] = -nf
            A = np.hstack([A, col])
            ns = 1
        else:
            ns = len(ids)
            blk = np.zeros((2 * m, ns))
            blk[ids, np.arange(ns)] = -nf
            blk[ids + m, np.arange(ns)] = -nf
            A = np.hstack([A, blk])
        H = block_diag(H, soft_weight * np.eye(ns))
        f = np.concatenate([f, np.zeros(ns)])
        f_theta = np.vstack([f_theta, np.zeros((ns, f_theta.shape[1]))])
    b = np.concatenate([qp.bu, -qp.bl])
    W = np.vstack([qp.W, -qp.W])
    keep = b < 1e20
    return dict(H=H, f=f, f_theta=f_theta, A=A[keep], b=b[keep], W=W[keep], senses=senses[keep])


# ---------------------------------------------------------------------------
# BASELINE controllers
# ---------------------------------------------------------------------------

def invpend(Np=50, Nc=5) -> MPC:
    """README.md:36-53 == mpc_examples.jl:104-141 linearised at the origin."""
    A = np.array([[0, 1, 0, 0], [0, -10, 9.81, 0], [0, 0, 0, 1], [0, -20, 39.24, 0.0]])
    B = 100 * np.array([[0], [1.0], [0], [2.0]])
    C = np.array([[1.0, 0, 0, 0], [0, 0, 1.0, 0]])
    Ts = 0.01
    F, G = zoh(A, B, Ts)
    mpc = MPC(F, G, C=C, Ts=Ts, Np=Np, Nc=Nc)
    mpc.set_objective(Q=[1.2 ** 2, 1], R=[0.0], Rr=[1.0])
    mpc.set_bounds(umin=[-2.0], umax=[2.0])
    return mpc


def invpend_range(scale=1.0):
    """mpc_examples.jl:128-134 + types.jl:202-224 (uprev range = input bounds)."""
    lo = np.concatenate([-20 * scale * np.ones(4), -20 * scale * np.ones(2), [-2.0]])
    return lo, -lo


def aircraft(Np=10, Nc=10, ks=(2,)) -> MPC:
    """mpc_examples.jl:174-208."""
    A = np.array([[-0.0151, -60.5651, 0, -32.174],
                  [-0.0001, -1.3411, 0.9929, 0],
                  [0.00018, 43.2541, -0.86939, 0],
                  [0, 0, 1, 0.0]])
    B = np.array([[-2.516, -13.136], [-0.1689, -0.2514], [-17.251, -1.5766], [0, 0.0]])
    C = np.array([[0, 1.0, 0, 0], [0, 0, 0, 1.0]])
    Ts = 0.05
    F, G = zoh(A, B, Ts)
    C = C / np.array([[1.0], [200.0]])
    Dd = np.array([[1.0, 0], [0, 200.0]]) / np.array([[1.0], [200.0]])
    mpc = MPC(F, 50 * G, C=C, Dd=Dd, Ts=Ts, Np=Np, Nc=Nc)
    mpc.set_objective(Q=np.array([10.0, 10.0]) ** 2, R=np.zeros(2), Rr=np.array([0.1, 0.1]) ** 2)
    mpc.set_input_bounds(umin=[-0.5, -0.5], umax=[0.5, 0.5])
    mpc.set_output_bounds(ymin=[-0.5, -0.5], ymax=[0.5, 0.5], ks=list(ks))
    return mpc


def aircraft_range(scale=1.0):
    """mpc_examples.jl:198-204; theta = [x(4); r(2); d(2); uprev(2)]."""
    hi = np.concatenate([20 * scale * np.ones(4), scale * np.array([1, 0.05]),
                         20 * scale * np.ones(2), [0.5, 0.5]])
    return -hi, hi


def chain_of_masses(nm=10, Np=20, Nc=20) -> MPC:
    """BASELINE config 4 (synthesised; closest reference example is mass_spring,
    mpc_examples.jl:241-286): nm unit masses joined by unit springs, nm-1 actuators
    acting between neighbours, Ts = 0.5 zoh, Q = I, R = I, |u| <= 0.5, regulation."""
    Fx = -2 * np.eye(nm) + np.eye(nm, k=1) + np.eye(nm, k=-1)
    A = np.block([[np.zeros((nm, nm)), np.eye(nm)], [Fx, np.zeros((nm, nm))]])
    Bf = np.zeros((nm, nm - 1))
    for j in range(nm - 1):
        Bf[j, j], Bf[j + 1, j] = 1.0, -1.0
    B = np.vstack([np.zeros((nm, nm - 1)), Bf])
    F, G = zoh(A, B, 0.5)
    mpc = MPC(F, G, Ts=0.5, Np=Np, Nc=Nc, reference_tracking=False)
    mpc.set_objective(Q=np.ones(2 * nm), R=np.ones(nm - 1), Rr=np.zeros(nm - 1))
    mpc.set_input_bounds(umin=-0.5 * np.ones(nm - 1), umax=0.5 * np.ones(nm - 1))
    return mpc


def chain_range(nm=10, scale=1.0):
    hi = 4 * scale * np.ones(2 * nm)
    return -hi, hi


def simple_md(q1=1.0) -> MPC:
    """docs/src/manual/simple.md:53-71 (q1 = 1000 for the second known answer)."""
    mpc = MPC(np.array([[1, 0.5], [0, 1.0]]), np.array([[0], [1.0]]), C=np.array([[1.0, 0], [1, 1]]),
              Np=10, Nc=10)
    mpc.set_objective(Q=[q1, 1], R=0, Rr=[1])
    mpc.set_bounds(umin=[-3], umax=[3], ymin=[0, 0], ymax=[1, 2])
    mpc.add_constraint(Ax=np.array([[2.0, -1]]), lb=[-1], ub=[2])
    return mpc


def double_integrator_hard(Np=4, Nc=4) -> MPC:
    """Small controller with HARD general rows (velocity limits) for the thread-per-QP kernel's
    general / singular branch; same plant as runtests.jl:1067-1074."""
    F, G = np.array([[1, 0.1], [0, 1.0]]), np.array([[0.005], [0.1]])
    mpc = MPC(F, G, C=np.array([[1.0, 0]]), Ts=0.1, Np=Np, Nc=Nc)
    mpc.set_objective(Q=[1.0], R=[0.1], Rr=[0.0])
    mpc.set_input_bounds(umin=[-1.0], umax=[1.0])
    mpc.add_constraint(Ax=np.array([[0.0, 1.0]]), lb=[-0.5], ub=[0.5], ks=range(2, Np + 2))
    return mpc


def double_integrator_terminal_eq(Np=10, Nc=10) -> MPC:
    """Terminal EQUALITY constraint (position at the end of the horizon pinned to 0.2: lb == ub,
    hard) on the double integrator: the rows get sense EQUALITY = ACTIVE + IMMUTABLE
    (mpc2mpqp.jl:868-885) and enter the working set before the first iteration."""
    F, G = np.array([[1, 0.1], [0, 1.0]]), np.array([[0.005], [0.1]])
    mpc = MPC(F, G, C=np.array([[1.0, 0]]), Ts=0.1, Np=Np, Nc=Nc)
    mpc.set_objective(Q=[1.0], R=[0.1], Rr=[0.0])
    mpc.set_input_bounds(umin=[-1.0], umax=[1.0])
    mpc.add_constraint(Ax=np.array([[1.0, 0.0]]), lb=[0.2], ub=[0.2], ks=[Np + 1], soft=False)
    return mpc


def generalized_parameter_example() -> MPC:
    """runtests.jl:1250-1256: scalar integrator, linear cost (Eu p + eu)'u, parametric constraint
    u + p <= 1; known answers u = 1.0 (p = 0) and 0.25 (p = 0.75) at runtests.jl:1264-1268."""
    mpc = MPC(np.array([[1.0]]), np.array([[1.0]]), C=np.array([[1.0]]), Np=4, Nc=4)
    mpc.set_bounds(umin=[0.0], umax=[2.0])
    mpc.set_objective(Q=[0.0], R=[1e-6], Eu=np.array([[-1.0]]), eu=[-0.1])
    mpc.add_constraint(Au=np.array([[1.0]]), Ap=np.array([[1.0]]), ub=[1.0], ks=range(1, 5))
    return mpc


def sample_theta(lo, hi, N, seed=1234):
    """theta ~ U(lo, hi), seed 1234 (runtests.jl:8); returns nth x N column-major."""
    rng = np.random.Generator(np.random.PCG64(seed))
    T = lo[:, None] + (hi - lo)[:, None] * rng.random((len(lo), N))
    return np.asfortranarray(T)
