"""Shared helpers for the tests: controllers, oracle problems built from the handle's own
constants (so the online path is compared bit for bit), bit-mask decoding."""
import numpy as np

from oracle import mpqp, ldp

KNOWN = {
    # /root/reference/test/runtests.jl:62-66 (also :69-83 via generated C, :178-197 via BST)
    "invpend": dict(x=[5.0, 5, 0, 0], r=None, u=1.7612519326, tol=1e-6),
    # /root/reference/docs/src/manual/simple.md:107
    "simple": dict(x=[0.5, 1.0], r=[0.0, 0.0], u=-1.0, tol=1e-6),
    # /root/reference/docs/src/manual/simple.md:203-205
    "simple_q1000": dict(x=[0.0, 0.0], r=[1.0, 0.0], u=1.0, tol=1e-6),
}


def controller(name):
    if name == "invpend":
        return mpqp.invpend()
    if name == "aircraft":
        return mpqp.aircraft()
    if name == "aircraft_ksall":
        return mpqp.aircraft(ks=range(2, 11))
    if name == "chain":
        return mpqp.chain_of_masses()
    if name == "chain_small":
        return mpqp.chain_of_masses(nm=4, Np=10, Nc=10)
    if name == "dint_hard":
        return mpqp.double_integrator_hard()
    if name == "dint_eq":
        return mpqp.double_integrator_terminal_eq()
    if name == "dint_eq_small":
        return mpqp.double_integrator_terminal_eq(Np=6, Nc=6)
    if name == "genparam":
        return mpqp.generalized_parameter_example()
    if name == "invpend_prestab":      # README pendulum with the LQR prestabilising feedback (setup.jl:196-199)
        mpc = mpqp.invpend()
        mpc.set_prestabilizing_feedback()
        return mpc
    if name == "invpend_rpreview":     # README pendulum with a 50-step reference preview (nth = 4 + 100 + 1)
        mpc = mpqp.invpend()
        mpc.reference_preview = True
        return mpc
    if name.startswith("invpend_nc"):  # README pendulum with Nc = 1..8 moves: one box-kernel instantiation <n, 1> each
        return mpqp.invpend(Nc=int(name[len("invpend_nc"):]))
    if name.startswith("aircraft_box"):  # AFTI-16 with input bounds only, Nc = 1..4: box-only, nu = 2, n = 2 Nc
        Nc = int(name[len("aircraft_box"):])
        return mpqp.aircraft(Np=10, Nc=Nc, ks=())
    if name == "simple":
        return mpqp.simple_md(1.0)
    if name == "simple_q1000":
        return mpqp.simple_md(1000.0)
    raise KeyError(name)


def theta_range(name, scale=1.0):
    if name in ("invpend", "invpend_prestab") or name.startswith("invpend_nc"):
        return mpqp.invpend_range(scale)
    if name == "invpend_rpreview":     # [x(4); r_1..r_50 (2 each); uprev]
        lo, hi = mpqp.invpend_range(scale)
        return (np.concatenate([lo[:4], np.tile(lo[4:6], 50), lo[6:]]),
                np.concatenate([hi[:4], np.tile(hi[4:6], 50), hi[6:]]))
    if name.startswith("aircraft"):
        return mpqp.aircraft_range(scale)
    if name == "chain":
        return mpqp.chain_range(10, scale)
    if name == "chain_small":
        return mpqp.chain_range(4, scale)
    if name == "genparam":      # theta = [x; r; p]
        return np.array([-1.0, -1.0, -0.5]) * scale, np.array([1.0, 1.0, 1.5]) * scale
    if name == "dint_hard":
        hi = np.array([2.0, 1.0, 2.0]) * scale
        return -hi, hi
    if name.startswith("dint_eq"):
        hi = np.array([0.5, 0.5, 1.0]) * scale
        return -hi, hi
    if name.startswith("simple"):
        hi = np.array([1.0, 2.0, 1.0, 1.0, 3.0]) * scale
        return -hi, hi
    raise KeyError(name)


def oracle_from_constants(qp, nu, cst, soft_weight=1e6, **settings):
    """OracleProblem that uses the constants the product computed (Mbar, Gram, Dth, ...), so any
    difference seen afterwards comes from the online path alone."""
    L = mpqp.qp2ldp(qp, nu, soft_weight=soft_weight)
    L.Mbar, L.Dth, L.du, L.dl = cst["Mbar"], cst["Dth"], cst["du"], cst["dl"]
    Rfull = np.zeros_like(L.Rinv)
    Rfull[:nu] = cst["Rinv"]
    L.Rinv = Rfull
    L.Uth_offset, L.u_offset = cst["Uth"], cst["uoff"]
    if "senses" in cst:
        L.senses = np.asarray(cst["senses"], np.int32)
    if "cap" in cst and "cap" not in settings:      # the working-set capacity the kernels provision
        settings["cap"] = int(cst["cap"])
    P = ldp.OracleProblem(L, nu, **settings)
    P.Gram[:] = cst["Gram"]
    return P


def bits(words, m):
    """(N x words) uint64 -> (N x m) bool."""
    w = np.asarray(words).astype(np.uint64)
    out = np.zeros((w.shape[0], m), bool)
    for j in range(m):
        out[:, j] = (w[:, j // 64] >> np.uint64(j % 64)) & np.uint64(1)
    return out
