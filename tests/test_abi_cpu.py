"""CPU tests of the C-ABI boundary: the library loads, exports every declared symbol, refuses
to run without a GPU, validates arguments, and its host-side setup step matches qp2ldp."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import util
from oracle import mpqp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "lmpc_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lmpc_gpu_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lmpc):
    L = lmpc.lib()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/lmpc_gpu.h but not exported"


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "linearmpc.jl_b200")
    for dp, _, fns in os.walk(pkg):
        if os.path.basename(dp) == "build":
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, fn


def has_gpu(lmpc):
    return lmpc.device_count() > 0


def test_no_gpu_fails_loudly(lmpc):
    if has_gpu(lmpc):
        pytest.skip("GPU present")
    mpc = util.controller("invpend")
    qp = mpqp.mpc2mpqp(mpc)
    with pytest.raises(lmpc.LmpcError) as e:
        lmpc.setup_gpu(qp, nu=1, dims=mpc.parameter_dims())
    assert e.value.code == -104 and "no CPU fallback" in str(e.value)
    with pytest.raises(lmpc.LmpcError):
        lmpc.fp64_peak_tflops(0)
    with pytest.raises(lmpc.LmpcError):
        lmpc.GpuExplicitMPC(np.zeros((1, 3)), np.zeros((2, 1, 3)), [0, 0, 1], [1, 0, 0])


@pytest.mark.parametrize("name", ["invpend", "aircraft", "aircraft_ksall", "simple", "chain_small", "chain"])
def test_host_setup_matches_qp2ldp(lmpc, name):
    """lmpc_gpu_create's setup step == qp2ldp (codegen.jl:239-280) to rounding (scaled by cond)."""
    mpc = util.controller(name)
    qp = mpqp.mpc2mpqp(mpc)
    L = mpqp.qp2ldp(qp, mpc.nu, soft_weight=mpc.soft_weight)
    c = lmpc.host_ldp(qp.H, qp.f, qp.f_theta, qp.A, qp.bu, qp.bl, qp.W, qp.senses, mpc.nu)
    tol = 1e-13 * max(1.0, np.linalg.cond(qp.H))
    finite = np.abs(L.du) < 1e20
    for k, ref in [("Mbar", L.Mbar), ("Dth", L.Dth), ("Rinv", L.Rinv[:mpc.nu]), ("Uth", L.Uth_offset),
                   ("uoff", L.u_offset)]:
        assert np.max(np.abs(c[k] - ref)) <= tol * max(1.0, np.abs(ref).max()), k
    assert np.allclose(c["du"][finite], L.du[finite], rtol=tol, atol=tol)
    assert np.allclose(np.diag(c["Gram"]), 1.0, atol=1e-12)
    assert np.allclose(c["Gram"], c["Mbar"] @ c["Mbar"].T, atol=1e-13)
    nsoft = int(np.count_nonzero(qp.senses & mpqp.SOFT))
    general = qp.m > qp.ms or nsoft > 0
    assert c["cap"] == min(qp.m, qp.n + (1 if general else 0) + nsoft)


def test_host_setup_rejects_bad_problems(lmpc):
    mpc = util.controller("invpend")
    qp = mpqp.mpc2mpqp(mpc)
    Hbad = qp.H.copy()
    Hbad[0, 0] = -1.0
    with pytest.raises(lmpc.LmpcError) as e:
        lmpc.host_ldp(Hbad, qp.f, qp.f_theta, qp.A, qp.bu, qp.bl, qp.W, qp.senses, 1)
    assert e.value.code == -102
    s = qp.senses.copy()
    s[0] |= mpqp.BINARY
    with pytest.raises(lmpc.LmpcError) as e:
        lmpc.host_ldp(qp.H, qp.f, qp.f_theta, qp.A, qp.bu, qp.bl, qp.W, s, 1)
    assert e.value.code == -101
    with pytest.raises(lmpc.LmpcError) as e:
        lmpc.host_ldp(qp.H, qp.f, qp.f_theta, qp.A, qp.bu, qp.bl, qp.W, qp.senses, 9)
    assert e.value.code == -100


def test_default_settings_follow_the_doc_table(lmpc):
    """docs/src/manual/solver.md:49-56."""
    s = lmpc.default_settings()
    assert (s.primal_tol, s.dual_tol, s.progress_tol, s.cycle_tol, s.iter_limit) == (1e-6, 1e-12, 1e-6, 10, 10000)


def test_form_parameter_layout(lmpc):
    """theta = [x; r; d; uprev] (explicit.jl:54-63), same as the oracle's restatement."""
    mpc = util.controller("aircraft")
    cls = lmpc.GpuMPC.__new__(lmpc.GpuMPC)
    cls.dims, cls.uprev, cls.nu = mpc.parameter_dims(), np.array([0.1, -0.2]), 2
    x = np.arange(4.0)
    th = cls.form_parameter(x, r=[1.0, 2.0], d=[3.0, 4.0])
    assert np.array_equal(th[:, 0], np.concatenate([x, [1, 2], [3, 4], [0.1, -0.2]]))
    th = cls.form_parameter(x, uprev=[9.0, 8.0])
    assert np.array_equal(th[:, 0], mpc.form_parameter(x, uprev=[9.0, 8.0]))
    X = np.arange(12.0).reshape(4, 3)
    th = cls.form_parameter(X, r=[1.0, 2.0])
    assert th.shape == (10, 3) and np.array_equal(th[4:6, 2], [1, 2]) and th.flags.f_contiguous
    with pytest.raises(ValueError):
        cls.form_parameter(np.zeros(3))


# ---- struct layouts: the library's own sizeof / offsetof vs the Julia and ctypes mirrors ----------------
_JL_TYPES = {"Int32": (4, 4), "Cint": (4, 4), "Int64": (8, 8), "Float64": (8, 8), "UInt64": (8, 8)}


def _julia_struct_layout(name):
    """Parse `struct name ... end` in LinearMPCGPU.jl and lay the fields out by the C rules Julia uses
    for isbits structs (natural alignment, size rounded up to the largest alignment)."""
    src = open(os.path.join(ROOT, "linearmpc.jl_b200", "julia", "LinearMPCGPU.jl")).read()
    body = re.search(r"^(?:mutable )?struct\s+" + name + r"\b(.*?)^end", src, flags=re.S | re.M).group(1)
    body = re.sub(r"^\s*\w+\(\)\s*=.*$", "", body, flags=re.M)           # inner constructors
    body = re.sub(r"#.*", "", body)
    fields = re.findall(r"(\w+)::([\w{}]+)", body)
    off, offs, amax = 0, [], 1
    for _, ty in fields:
        size, align = (8, 8) if ty.startswith("Ptr{") else _JL_TYPES[ty]
        off = (off + align - 1) // align * align
        offs.append(off)
        off += size
        amax = max(amax, align)
    return [f for f, _ in fields], offs, (off + amax - 1) // amax * amax


def _abi_layout(lmpc):
    L = lmpc.lib()
    L.lmpc_gpu_abi_layout.restype = C.c_int32
    n = L.lmpc_gpu_abi_layout(None, 0)
    buf = (C.c_int32 * n)()
    assert L.lmpc_gpu_abi_layout(buf, n) == n
    v, i, out = list(buf), 0, {}
    while i < n:
        sid, size, nf = v[i], v[i + 1], v[i + 2]
        out[sid] = (size, v[i + 3:i + 3 + nf])
        i += 3 + nf
    return out


def test_struct_layouts_match_the_julia_and_ctypes_mirrors(lmpc):
    """include/lmpc_gpu.h structs as compiled (lmpc_gpu_abi_layout) == the hand-written Julia structs
    (field count, offsets, size) == the ctypes mirrors of host.py.  Round 1 shipped a Julia LmpcSimExt
    without the trailing rstride / dstride: C then read 8 bytes past the Julia object."""
    lay = _abi_layout(lmpc)
    assert sorted(lay) == [1, 2, 3, 4, 5, 6]
    for sid, jl in [(1, "LmpcProblem"), (2, "LmpcSettings"), (3, "LmpcSimExt"), (6, "LmpcSimStepped")]:
        names, offs, size = _julia_struct_layout(jl)
        assert (size, offs) == (lay[sid][0], lay[sid][1]), (jl, names, offs, size, lay[sid])
    import sys
    host = sys.modules[lmpc.__name__ + ".host"]
    _Problem, _Settings, _SimExt, _Constraint, _Model = (host._Problem, host._Settings, host._SimExt,
                                                          host._Constraint, host._Model)
    for sid, T in [(1, _Problem), (2, _Settings), (3, _SimExt), (4, _Constraint), (5, _Model), (6, host._SimStepped)]:
        offs = [getattr(T, f[0]).offset for f in T._fields_]
        assert (C.sizeof(T), offs) == (lay[sid][0], lay[sid][1]), T.__name__


def test_julia_struct_field_names_follow_the_header():
    """same field names, same order, as the C declarations (a renamed or reordered field is a bug even
    when the sizes happen to agree)."""
    hdr = open(os.path.join(ROOT, "include", "lmpc_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for cname, jl in [("lmpc_problem", "LmpcProblem"), ("lmpc_settings", "LmpcSettings"), ("lmpc_sim_ext", "LmpcSimExt"),
                      ("lmpc_sim_stepped", "LmpcSimStepped")]:
        body = re.search(r"typedef struct \{([^{}]*)\}\s*" + cname + r"\s*;", hdr).group(1)
        cfields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(double|int32_t|int64_t|uint64_t|void|lmpc_plant_fn)\s*", "", decl)
            cfields += [x.strip().lstrip("*").strip() for x in decl.split(",")]
        assert _julia_struct_layout(jl)[0] == cfields, (cname, cfields)


def test_setup_flags_infeasible_and_overdetermined(lmpc):
    """DAQP.setup's other two flags (src/setup.jl:14-23) are raised once, at create / host setup, instead of
    showing up per instance at solve time: -1 (a row whose bounds cross) and -6 (equality rows that are
    dependent or outnumber the variables)."""
    mpc = util.controller("invpend")
    qp = mpqp.mpc2mpqp(mpc)
    bl = qp.bl.copy()
    bl[2] = qp.bu[2] + 1.0                                   # lower bound above the upper bound
    with pytest.raises(lmpc.LmpcError) as e:
        lmpc.host_ldp(qp.H, qp.f, qp.f_theta, qp.A, qp.bu, bl, qp.W, qp.senses, 1)
    assert e.value.code == -105 and "setup flag -1" in str(e.value)
    # six equality rows on five variables: the same row twice makes the set dependent
    n, nth = qp.n, qp.f_theta.shape[1]
    A = np.vstack([np.ones((1, n)), np.ones((1, n))])
    bu = np.concatenate([qp.bu, [0.3, 0.3]]); bl2 = np.concatenate([qp.bl, [0.3, 0.3]])
    W = np.vstack([qp.W, np.zeros((2, nth))])
    s = np.concatenate([qp.senses, [5, 5]]).astype(np.int32)  # EQUALITY = ACTIVE + IMMUTABLE (mpc2mpqp.jl:868-885)
    with pytest.raises(lmpc.LmpcError) as e:
        lmpc.host_ldp(qp.H, qp.f, qp.f_theta, A, bu, bl2, W, s, 1)
    assert e.value.code == -106 and "setup flag -6" in str(e.value)
    # one such row is fine
    c = lmpc.host_ldp(qp.H, qp.f, qp.f_theta, A[:1], bu[:-1], bl2[:-1], W[:-1], s[:-1], 1)
    assert c["cap"] >= 1
