"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle.so (ldp_oracle.c).

See ldp_oracle.h for what the C code restates and its parity status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_up = C.POINTER(C.c_uint64)


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("ms", C.c_int), ("nth", C.c_int), ("nu", C.c_int),
                ("Mbar", c_dp), ("Gram", c_dp), ("Dth", c_dp), ("du", c_dp), ("dl", c_dp),
                ("Rinv", c_dp), ("Uth", c_dp), ("uoff", c_dp), ("sense", c_ip),
                ("rho_soft", C.c_double), ("primal_tol", C.c_double), ("dual_tol", C.c_double),
                ("zero_tol", C.c_double), ("progress_tol", C.c_double), ("fval_bound", C.c_double),
                ("cycle_tol", C.c_int), ("iter_limit", C.c_int), ("no_repair", C.c_int), ("cap", C.c_int)]


class _Bst(C.Structure):
    _fields_ = [("nth", C.c_int), ("nu", C.c_int), ("n_hp", C.c_int), ("n_fb", C.c_int),
                ("n_nodes", C.c_int), ("halfplanes", c_dp), ("feedbacks", c_dp), ("clip", c_dp),
                ("hp_list", c_ip), ("jump_list", c_ip)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "ldp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


def _up(a):
    return a.ctypes.data_as(c_up) if a is not None else None


DEFAULT_SETTINGS = dict(primal_tol=1e-6, dual_tol=1e-12, zero_tol=1e-11, progress_tol=1e-6,
                        fval_bound=1e30, cycle_tol=10, iter_limit=10000, no_repair=0, cap=0)


def gram(Mbar):
    """Gram matrix with the fixed summation order shared with the product's setup code."""
    m, n = Mbar.shape
    G = np.zeros((m, m))
    for i in range(m):
        for j in range(m):
            acc = 0.0
            for c in range(n):
                acc += Mbar[i, c] * Mbar[j, c]
            G[i, j] = acc
    return G


class OracleProblem:
    """Holds the arrays alive and exposes the C struct."""

    def __init__(self, ldp, nu, **settings):
        s = dict(DEFAULT_SETTINGS)
        s.update(settings)
        self.settings = s
        c = np.ascontiguousarray
        self.Mbar = c(ldp.Mbar, float)
        self.Gram = c(ldp.Mbar @ ldp.Mbar.T if ldp.Mbar.shape[0] > 64 else gram(ldp.Mbar))
        self.Dth = c(ldp.Dth, float)
        self.du, self.dl = c(ldp.du, float), c(ldp.dl, float)
        self.Rinv = c(ldp.Rinv[:nu], float)
        self.Uth = c(ldp.Uth_offset, float)
        self.uoff = c(ldp.u_offset, float)
        self.sense = c(ldp.senses, np.int32)
        self.m, self.n = self.Mbar.shape
        self.nth, self.nu = self.Dth.shape[1], nu
        self.c = _Problem(self.n, self.m, ldp.ms, self.nth, nu, _dp(self.Mbar), _dp(self.Gram),
                          _dp(self.Dth), _dp(self.du), _dp(self.dl), _dp(self.Rinv), _dp(self.Uth),
                          _dp(self.uoff), _ip(self.sense), ldp.rho_soft, s["primal_tol"],
                          s["dual_tol"], s["zero_tol"], s["progress_tol"], s["fval_bound"],
                          s["cycle_tol"], s["iter_limit"], int(s["no_repair"]), int(s["cap"]))

    def solve_batch(self, Theta, nthreads=0, want_flops=False):
        """Theta: nth x N (any layout).  Returns dict(U nu x N, flag, iters, act_u, act_l[, flops])."""
        Th = np.ascontiguousarray(np.asarray(Theta, float).T)  # N x nth == nth x N column-major
        N = Th.shape[0]
        words = (self.m + 63) // 64
        U = np.zeros((N, self.nu))
        flag = np.zeros(N, np.int32)
        iters = np.zeros(N, np.int32)
        au = np.zeros((N, words), np.uint64)
        al = np.zeros((N, words), np.uint64)
        fl = np.zeros(N) if want_flops else None
        lib().orc_solve_batch(C.byref(self.c), _dp(Th), C.c_int64(N), _dp(U), _ip(flag), _ip(iters),
                              _up(au), _up(al), _dp(fl), C.c_int(nthreads))
        out = dict(U=U.T, flag=flag, iters=iters, act_u=au, act_l=al)
        if want_flops:
            out["flops"] = fl
        return out

    def solve_batch_warm(self, Theta, act_u, act_l, nthreads=0):
        """Warm start from the working sets act_u / act_l (N x words uint64; not modified)."""
        Th = np.ascontiguousarray(np.asarray(Theta, float).T)
        N = Th.shape[0]
        U = np.zeros((N, self.nu))
        flag = np.zeros(N, np.int32)
        iters = np.zeros(N, np.int32)
        au = np.ascontiguousarray(act_u, np.uint64).copy()
        al = np.ascontiguousarray(act_l, np.uint64).copy()
        lib().orc_solve_batch_warm(C.byref(self.c), _dp(Th), C.c_int64(N), _dp(U), _ip(flag), _ip(iters),
                                   _up(au), _up(al), C.c_int(nthreads))
        return dict(U=U.T, flag=flag, iters=iters, act_u=au, act_l=al)

    def solve(self, theta):
        r = self.solve_batch(np.asarray(theta, float).reshape(-1, 1), nthreads=1)
        return r["U"][:, 0], int(r["flag"][0]), r

    def simulate_batch(self, Fx, Gx, Gdx, X0, R, D, U0, T, nr, nd, nuprev, warm=False,
                       nthreads=0, want_traj=False, Rtraj=None, Dtraj=None, ext=None, want_flops=False):
        """Rtraj (nr x T) / Dtraj (nd x T): trajectories shared by all closed loops (override R / D).
        ext: dict with any of Ft, Gt, Gdt (true plant), C, Dd, K (observer), Vtraj (ny x T noise)."""
        nx = Fx.shape[0]
        c = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, float).T)
        X0c, Rc, Dc, U0c = c(X0), c(R), c(D), c(U0)
        N = X0c.shape[0]
        Fx, Gx = np.ascontiguousarray(Fx, float), np.ascontiguousarray(Gx, float)
        Gdx = np.ascontiguousarray(Gdx, float) if nd > 0 else np.zeros((nx, 0))
        Xend = np.zeros((N, nx))
        Ulast = np.zeros((N, self.nu))
        Utraj = np.zeros((N, T, self.nu)) if want_traj else None
        flag = np.zeros(N, np.int32)
        iters = np.zeros(N, np.int32)
        Rt, Dt = c(Rtraj), c(Dtraj)      # T x nr, T x nd row-major == column-major nr x T
        ex, Xhat, keep = None, None, []
        fl = np.zeros(N) if want_flops else None
        if ext is None and want_flops:
            ex = _SimExt(None, None, None, None, None, None, None, 0, None, 0, 0, fl.ctypes.data)
        if ext is not None:
            rm = lambda k: None if ext.get(k) is None else np.ascontiguousarray(np.asarray(ext[k], float))
            m = {k: rm(k) for k in ("Ft", "Gt", "Gdt", "C", "Dd", "K")}
            V = None if ext.get("Vtraj") is None else np.ascontiguousarray(np.asarray(ext["Vtraj"], float).T)
            ny = 0 if m["C"] is None else m["C"].shape[0]
            Xhat = np.zeros((N, nx))
            keep = [m, V]
            vp = lambda a: None if a is None else a.ctypes.data
            ex = _SimExt(vp(m["Ft"]), vp(m["Gt"]), vp(m["Gdt"]), vp(m["C"]), vp(m["Dd"]), vp(m["K"]), vp(V),
                         ny, vp(Xhat), int(ext.get("rstride", 0)), int(ext.get("dstride", 0)),
                         fl.ctypes.data if fl is not None else None)
        lib().orc_simulate_batch_ex(C.byref(self.c), nx, nr, nd, nuprev, _dp(Fx), _dp(Gx), _dp(Gdx),
                                    _dp(X0c), _dp(Rc), _dp(Dc), _dp(U0c), _dp(Rt), _dp(Dt),
                                    C.byref(ex) if ex is not None else None,
                                    C.c_int64(N), T, int(warm),
                                    _dp(Xend), _dp(Ulast), _dp(Utraj), _ip(flag), _ip(iters),
                                    C.c_int(nthreads))
        out = dict(Xend=Xend.T, Ulast=Ulast.T, Utraj=Utraj, flag=flag, iters=iters)
        if fl is not None:
            out["flops"] = fl
        if Xhat is not None:
            out["Xhat_end"] = Xhat.T
        return out


DYN_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, c_dp, c_dp, c_dp, c_dp)


class _SimStepped(C.Structure):
    _fields_ = [("dyn", DYN_FN), ("user", C.c_void_p)] + \
               [(k, C.c_void_p) for k in ("Ft", "Gt", "Gdt", "ft")] + [("ny", C.c_int)] + \
               [(k, C.c_void_p) for k in ("Cm", "Ddm", "hm", "Vtraj")] + [("nxo", C.c_int), ("ndm", C.c_int)] + \
               [(k, C.c_void_p) for k in ("Fo", "Go", "Gdo", "fo", "Co", "Ddo", "ho", "Ko", "Rtraj", "Dtraj", "Xhat_end")]


def simulate_batch_stepped(P, nx, nr, nuprev, X0, R, D, U0, T, cfg, warm=True, nthreads=0, want_traj=False):
    """orc_simulate_batch_stepped: cfg is a dict of row-major matrices / vectors (ldp_oracle.h: orc_sim_stepped),
    'dyn' a Python function (k, x, u, d) -> xnext for a nonlinear plant; X0 nx x N, R nr x N, D ndm x N."""
    c = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, float).T)
    X0c, Rc, Dc, U0c = c(X0), c(R), c(D), c(U0)
    N = X0c.shape[0]
    nu = P.nu
    keep = {}
    def vp(k, transpose=False):
        a = cfg.get(k)
        if a is None:
            return None
        a = np.ascontiguousarray(np.asarray(a, float).T if transpose else np.asarray(a, float))
        keep[k] = a
        return a.ctypes.data
    ndm = int(cfg.get("ndm", 0))
    nxo = int(cfg.get("nxo", nx))
    ny = int(cfg.get("ny", 0))
    Xhat = np.zeros((N, nxo))
    dyn = cfg.get("dyn")
    if dyn is not None:
        def tramp(user, k, x, u, d, xn):
            xv = np.ctypeslib.as_array(x, (nx,)); uv = np.ctypeslib.as_array(u, (nu,))
            dv = np.ctypeslib.as_array(d, (max(ndm, 1),))[:ndm]
            np.ctypeslib.as_array(xn, (nx,))[:] = dyn(k, xv, uv, dv)
        cb = DYN_FN(tramp)
    else:
        cb = DYN_FN(0)
    st = _SimStepped(cb, None, vp("Ft"), vp("Gt"), vp("Gdt"), vp("ft"), ny, vp("Cm"), vp("Ddm"), vp("hm"),
                     vp("Vtraj", True), nxo, ndm, vp("Fo"), vp("Go"), vp("Gdo"), vp("fo"), vp("Co"), vp("Ddo"),
                     vp("ho"), vp("Ko"), vp("Rtraj", True), vp("Dtraj", True), Xhat.ctypes.data)
    Xend, Ulast = np.zeros((N, nx)), np.zeros((N, nu))
    Utraj = np.zeros((N, T, nu)) if want_traj else None
    flag, iters = np.zeros(N, np.int32), np.zeros(N, np.int32)
    lib().orc_simulate_batch_stepped(C.byref(P.c), nx, nr, nuprev, _dp(X0c), _dp(Rc), _dp(Dc), _dp(U0c),
                                     C.byref(st), C.c_int64(N), T, int(warm), _dp(Xend), _dp(Ulast), _dp(Utraj),
                                     _ip(flag), _ip(iters), C.c_int(nthreads))
    return dict(Xend=Xend.T, Ulast=Ulast.T, Utraj=Utraj, flag=flag, iters=iters, Xhat_end=Xhat.T)


class _SimExt(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("Ft", "Gt", "Gdt", "C", "Dd", "K", "Vtraj")] + [("ny", C.c_int),
                                                                                      ("Xhat_end", C.c_void_p),
                                                                                      ("rstride", C.c_int),
                                                                                      ("dstride", C.c_int),
                                                                                      ("flops", C.c_void_p)]


class OracleBst:
    def __init__(self, halfplanes, feedbacks, hp_list, jump_list, clip=None):
        c = np.ascontiguousarray
        self.hp = c(halfplanes, float)          # n_hp x (nth+1)
        self.fb = c(feedbacks, float)           # n_fb x nu x (nth+1)
        self.hl = c(hp_list, np.int32)
        self.jl = c(jump_list, np.int32)
        self.clip = None if clip is None else c(clip, float)
        self.nth = self.hp.shape[1] - 1
        self.nu = self.fb.shape[1]
        self.c = _Bst(self.nth, self.nu, self.hp.shape[0], self.fb.shape[0], len(self.hl),
                      _dp(self.hp), _dp(self.fb), _dp(self.clip), _ip(self.hl), _ip(self.jl))

    def simulate_batch(self, Fx, Gx, Gdx, X0, R, D, U0, T, nr, nd, nuprev, nthreads=0, want_traj=False,
                       Rtraj=None, Dtraj=None):
        """Closed loop with the explicit law; arguments as OracleProblem.simulate_batch."""
        nx = np.asarray(Fx).shape[0]
        c = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, float).T)
        X0c, Rc, Dc, U0c, Rt, Dt = c(X0), c(R), c(D), c(U0), c(Rtraj), c(Dtraj)
        N = X0c.shape[0]
        Fx, Gx = np.ascontiguousarray(Fx, float), np.ascontiguousarray(np.asarray(Gx, float).reshape(nx, -1))
        Gdx = np.ascontiguousarray(Gdx, float) if nd > 0 else np.zeros((nx, 0))
        Xend, Ulast = np.zeros((N, nx)), np.zeros((N, self.nu))
        Utraj = np.zeros((N, T, self.nu)) if want_traj else None
        lib().orc_bst_simulate_batch(C.byref(self.c), nx, nr, nd, nuprev, _dp(Fx), _dp(Gx), _dp(Gdx),
                                     _dp(X0c), _dp(Rc), _dp(Dc), _dp(U0c), _dp(Rt), _dp(Dt), C.c_int64(N), T,
                                     _dp(Xend), _dp(Ulast), _dp(Utraj), C.c_int(nthreads))
        return dict(Xend=Xend.T, Ulast=Ulast.T, Utraj=Utraj)

    def evaluate_batch(self, Theta, nthreads=0):
        Th = np.ascontiguousarray(np.asarray(Theta, float).T)
        N = Th.shape[0]
        U = np.zeros((N, self.nu))
        reg = np.zeros(N, np.int32)
        lib().orc_bst_evaluate_batch(C.byref(self.c), _dp(Th), C.c_int64(N), _dp(U), _ip(reg),
                                     C.c_int(nthreads))
        return U.T, reg


def num_threads():
    return lib().orc_num_threads()
