"""ctypes binding of liblmpcgpu.so and the host-side mirror of the reference interface.

Reference interface mirrored (argument meaning and error behaviour kept):
  * ``compute_control(mpc, x; r, d, uprev, check)``   /root/reference/src/utils.jl:43-51
  * ``solve(mpc, theta)``                              /root/reference/src/utils.jl:268-283
  * ``form_parameter``                                 /root/reference/src/explicit.jl:54-63
  * ``Simulation(mpc; x0, r, d, N)``                   /root/reference/src/simulation.jl:37-116
  * ``compute_control(empc, x; ...)``                  /root/reference/src/utils.jl:53-60
Batch semantics: where the reference takes a vector ``x`` (nx), these take a matrix
(nx x N, one instance per column) and return nu x N.

numpy arrays are accepted as host buffers; objects exposing ``data_ptr()`` (torch CUDA
tensors) or ``__cuda_array_interface__`` are passed through as device pointers.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

KERNEL_AUTO, KERNEL_THREAD, KERNEL_WARP, KERNEL_CTA = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LmpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liblmpcgpu error {code}: {msg}")
        self.code = code


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("ms", C.c_int32), ("nth", C.c_int32),
                ("nu", C.c_int32), ("nx", C.c_int32),
                ("H", C.c_void_p), ("f", C.c_void_p), ("f_theta", C.c_void_p), ("A", C.c_void_p),
                ("bu", C.c_void_p), ("bl", C.c_void_p), ("W", C.c_void_p), ("senses", C.c_void_p),
                ("K", C.c_void_p), ("soft_weight", C.c_double),
                ("has_binaries", C.c_int32), ("is_symmetric", C.c_int32),
                ("n_break_points", C.c_int32)]


class _Settings(C.Structure):
    _fields_ = [("primal_tol", C.c_double), ("dual_tol", C.c_double), ("zero_tol", C.c_double),
                ("progress_tol", C.c_double), ("fval_bound", C.c_double),
                ("cycle_tol", C.c_int32), ("iter_limit", C.c_int32), ("kernel", C.c_int32),
                ("certified_iterations", C.c_int32)]


def lib_path():
    """In-tree library; LMPC_GPU_LIB names another build of the same sources (A/B measurements)."""
    return os.environ.get("LMPC_GPU_LIB") or os.path.join(_HERE, "liblmpcgpu.so")


def load_library():
    """Load liblmpcgpu.so.  Fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise LmpcError(-1, f"{path} not found: build it with `python __graft_entry__.py` "
                            f"or `make -C linearmpc.jl_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(path)
    L.lmpc_gpu_last_error.restype = C.c_char_p
    L.lmpc_gpu_launch_count.restype = C.c_int64
    L.lmpc_gpu_stream.restype = C.c_void_p
    L.lmpc_gpu_stream.argtypes = [C.c_void_p]
    L.lmpc_gpu_host_alloc.restype = C.c_void_p
    L.lmpc_gpu_host_alloc.argtypes = [C.c_size_t]
    L.lmpc_gpu_host_free.argtypes = [C.c_void_p]
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.lmpc_gpu_create.argtypes = [C.POINTER(_Problem), C.POINTER(_Settings), C.c_int, C.POINTER(vp)]
    L.lmpc_gpu_destroy.argtypes = [vp]
    L.lmpc_gpu_get_ldp.argtypes = [vp] * 9
    L.lmpc_gpu_get_info.argtypes = [vp, vp, vp, vp]
    L.lmpc_gpu_get_tiering.argtypes = [vp, vp, vp]
    L.lmpc_gpu_solve_batch.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    L.lmpc_gpu_create_multi.argtypes = [C.POINTER(_Problem), C.POINTER(_Settings), vp, i32, C.POINTER(vp)]
    L.lmpc_gpu_destroy_multi.argtypes = [vp]
    L.lmpc_gpu_solve_batch_multi.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    L.lmpc_gpu_solve_batch_warm.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    L.lmpc_gpu_solve_batch_async.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp]
    L.lmpc_gpu_compute_control_batch.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i64, vp]
    L.lmpc_gpu_simulate_batch.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, i64,
                                          i32, i32, vp, vp, vp, vp, vp]
    L.lmpc_gpu_simulate_batch_ex.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp,
                                             i64, i32, i32, vp, vp, vp, vp, vp]
    L.lmpc_gpu_simulate_batch_stepped.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, i64, i32, i32,
                                                  vp, vp, vp, vp, vp]
    L.lmpc_gpu_simulate_batch_traj.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp,
                                               i64, i32, i32, vp, vp, vp, vp, vp]
    L.lmpc_gpu_create_bst.argtypes = [vp, i32, vp, i32, vp, vp, i32, vp, i32, i32, C.c_int,
                                      C.POINTER(vp)]
    L.lmpc_gpu_destroy_bst.argtypes = [vp]
    L.lmpc_gpu_bst_kernel.argtypes = [vp]
    L.lmpc_gpu_simulate_bst_batch.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i64, i32,
                                              vp, vp, vp]
    L.lmpc_gpu_evaluate_batch.argtypes = [vp, vp, i64, vp, vp]
    L.lmpc_gpu_evaluate_batch_async.argtypes = [vp, vp, i64, vp, vp, vp]
    L.lmpc_gpu_compute_control_batch_bcast.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i64, i32, vp, vp]
    L.lmpc_gpu_simulate_batch_multi.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp,
                                                i64, i32, i32, vp, vp, vp, vp, vp]
    L.lmpc_gpu_create_bst_oriented.argtypes = [vp, i32, vp, i32, vp, vp, i32, vp, i32, i32, i32, C.c_int,
                                               C.POINTER(vp)]
    L.lmpc_gpu_create_bst_multi.argtypes = [vp, i32, vp, i32, vp, vp, i32, vp, i32, i32, vp, i32, C.POINTER(vp)]
    L.lmpc_gpu_destroy_bst_multi.argtypes = [vp]
    L.lmpc_gpu_evaluate_batch_multi.argtypes = [vp, vp, i64, vp, vp]
    L.lmpc_gpu_window_create.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp), vp]
    L.lmpc_gpu_window_open.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.lmpc_gpu_window_close.argtypes = [vp, C.c_int]
    L.lmpc_gpu_window_read.argtypes = [vp, vp, C.c_size_t]
    L.lmpc_gpu_abi_layout.argtypes = [vp, i32]
    L.lmpc_gpu_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.lmpc_gpu_default_settings.argtypes = [C.POINTER(_Settings)]
    _LIB = L
    return L


def lib():
    return load_library()


def _check(rc):
    if rc != 0:
        raise LmpcError(rc, lib().lmpc_gpu_last_error().decode())


class _Pinned:
    def __init__(self, nbytes):
        self.p = lib().lmpc_gpu_host_alloc(max(nbytes, 8))
        if not self.p:
            raise LmpcError(-102, lib().lmpc_gpu_last_error().decode())

    def __del__(self):
        if getattr(self, "p", None):
            lib().lmpc_gpu_host_free(self.p)
            self.p = None


def pinned_empty(shape, dtype=np.float64, order="C"):
    """numpy array in page-locked host memory (lmpc_gpu_host_alloc): the library DMAs such arrays
    in place; ordinary (pageable) arrays are staged through an internal pinned ring."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    own = _Pinned(n * dtype.itemsize)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(own.p)
    a = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape, order=order)
    _PIN_OWNERS[id(buf)] = own
    import weakref
    weakref.finalize(buf, _PIN_OWNERS.pop, id(buf), None)
    return a


_PIN_OWNERS = {}


def device_count():
    return lib().lmpc_gpu_device_count()


def launch_count():
    return int(lib().lmpc_gpu_launch_count())


def fp64_peak_tflops(device=0):
    out = C.c_double(0.0)
    _check(lib().lmpc_gpu_fp64_peak(device, C.byref(out)))
    return out.value


def default_settings(**over):
    s = _Settings()
    lib().lmpc_gpu_default_settings(C.byref(s))
    for k, v in over.items():
        setattr(s, k, v)
    return s


def _ptr(a):
    """host numpy array / device tensor -> raw pointer (None -> NULL)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    if hasattr(a, "__cuda_array_interface__"):
        return a.__cuda_array_interface__["data"][0]
    raise TypeError(f"unsupported buffer type {type(a)}")


def _is_device(a):
    return a is not None and not isinstance(a, np.ndarray)


def _wait_producer(*tensors):
    """The library's streams are non-blocking (include/lmpc_gpu.h, top): whatever torch queued on its
    current stream to produce a device argument must have finished before a synchronous entry point reads it."""
    for t in tensors:
        if t is not None and getattr(t, "is_cuda", False):
            import torch
            torch.cuda.current_stream(t.device).synchronize()
            return


def _colmajor(a, rows):
    """Batch argument -> (rows x N) column-major float64 host array (instance per column).
    A vector of length `rows` is one instance; with rows == 1 a vector is 1 x N."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(rows, 1) if a.size == rows else (a.reshape(1, -1) if rows == 1 else a)
    if a.ndim != 2 or a.shape[0] != rows:
        raise ValueError(f"expected {rows} rows, got shape {a.shape}")
    return np.asfortranarray(a)


def host_ldp(H, f, f_theta, A, bu, bl, W, senses, nu, nx=0, K=None, soft_weight=1e6):
    """Run ONLY the host-side setup (Cholesky, R^-1, normalisation, theta-absorbed bounds) and
    return its constants; needs no GPU.  Used to check the setup step against qp2ldp."""
    H = np.asfortranarray(np.asarray(H, np.float64)); n = H.shape[0]
    bu = np.ascontiguousarray(bu, np.float64); bl = np.ascontiguousarray(bl, np.float64); m = bu.shape[0]
    f_theta = np.asfortranarray(np.asarray(f_theta, np.float64)); nth = f_theta.shape[1]
    A = np.asfortranarray(np.asarray(A, np.float64).reshape(-1, n)); ms = m - A.shape[0]
    W = np.asfortranarray(np.asarray(W, np.float64).reshape(m, nth))
    f = np.ascontiguousarray(f, np.float64) if f is not None and np.size(f) else np.zeros(n)
    senses = np.ascontiguousarray(senses, np.int32)
    Kc = None if K is None else np.asfortranarray(np.asarray(K, np.float64).reshape(nu, nx))
    p = _Problem(n, m, ms, nth, nu, nx, _ptr(H), _ptr(f), _ptr(f_theta), _ptr(A) if A.size else None,
                 _ptr(bu), _ptr(bl), _ptr(W) if W.size else None, _ptr(senses), _ptr(Kc),
                 float(soft_weight), 0, 1, 0)
    out = dict(Mbar=np.zeros((m, n)), Gram=np.zeros((m, m)), Dth=np.zeros((m, nth)), du=np.zeros(m),
               dl=np.zeros(m), Rinv=np.zeros((nu, n)), Uth=np.zeros((nu, nth)), uoff=np.zeros(nu),
               senses=np.zeros(m, np.int32))
    cap = C.c_int32()
    L = lib()
    L.lmpc_gpu_host_ldp.argtypes = [C.POINTER(_Problem)] + [C.c_void_p] * 10
    _check(L.lmpc_gpu_host_ldp(C.byref(p), *[_ptr(out[k]) for k in
                               ("Mbar", "Gram", "Dth", "du", "dl", "Rinv", "Uth", "uoff", "senses")],
                               C.cast(C.byref(cap), C.c_void_p)))
    out["cap"] = cap.value
    return out


class _SimExt(C.Structure):
    """lmpc_sim_ext (include/lmpc_gpu.h)."""
    _fields_ = [(k, C.c_void_p) for k in ("Rtraj", "Dtraj", "Ft", "Gt", "Gdt", "C", "Dd", "K", "Vtraj")] + \
               [("ny", C.c_int32), ("Xhat_end", C.c_void_p), ("rstride", C.c_int32), ("dstride", C.c_int32)]


def hold_trajectory(v, rows, steps):
    """Scenario trajectory -> rows x steps, last value held: the reference copies scenario.r / scenario.d
    into an N-column array before the loop starts (simulation.jl:57-80)."""
    if v is None or rows == 0:
        return None
    v = np.asarray(v, float).reshape(rows, -1)
    if v.shape[1] < steps:
        v = np.hstack([v, np.repeat(v[:, -1:], steps - v.shape[1], axis=1)])
    return np.asfortranarray(v[:, :steps])


def preview_window(v, rows_total, Np, first, steps):
    """Flat trajectory for a sliding window of Np columns starting at column k + first of the held
    `steps`-column trajectory (get_preview, simulation.jl:102-104,128-134: rs[:, min(k + i, end)] with
    `end` = steps, so scenario columns beyond `steps` are never seen).  Returns (flat, stride)."""
    base = rows_total // Np
    v = hold_trajectory(v, base, steps)
    need = first + steps - 1 + Np                      # last column index read + 1
    v = np.hstack([v, np.repeat(v[:, -1:], need - v.shape[1], axis=1)])
    return np.asfortranarray(v[:, first:need]), base


class GpuMPC:
    """Device-resident implicit controller: the GPU backend `setup_gpu!` would attach to an MPC.

    Built from the dense mpQP that `mpc2mpqp` produces (types.jl:75-98) plus K and
    soft_weight.  numpy inputs are ordinary (row-major) arrays; they are handed to the C ABI
    column-major, as Julia would.
    """

    def _pack(self, H, f, f_theta, A, bu, bl, W, senses, nu, nx, K, soft_weight, has_binaries,
              is_symmetric, n_break_points, dims):
        """numpy mpQP -> lmpc_problem (column-major views kept alive on self)."""
        F = lambda x, shape=None: np.asfortranarray(np.asarray(x, np.float64).reshape(shape) if shape else np.asarray(x, np.float64))
        H = F(H)
        n = H.shape[0]
        bu, bl = F(bu).ravel(), F(bl).ravel()
        m = bu.shape[0]
        f_theta = F(f_theta)
        nth = f_theta.shape[1] if f_theta.ndim == 2 else 0
        A = F(A).reshape(-1, n) if np.size(A) else np.zeros((0, n), order="F")
        A = np.asfortranarray(A)
        ms = m - A.shape[0]
        W = np.asfortranarray(F(W).reshape(m, nth))
        f = F(f).ravel() if f is not None and np.size(f) else np.zeros(n)
        senses = np.ascontiguousarray(senses, np.int32)
        self.n, self.m, self.ms, self.nth, self.nu = n, m, ms, nth, int(nu)
        self.nx = int(nx) if nx is not None else 0
        self.dims = dims  # optional (nx, nr, nd, nuprev) for form_parameter / compute_control
        Kc = None
        if K is not None and np.any(np.asarray(K) != 0):
            Kc = np.asfortranarray(np.asarray(K, np.float64).reshape(self.nu, self.nx))
        self._keep = (H, f, f_theta, A, bu, bl, W, senses, Kc)
        p = _Problem(n, m, ms, nth, self.nu, self.nx, _ptr(H), _ptr(f), _ptr(f_theta),
                     _ptr(A) if A.size else None, _ptr(bu), _ptr(bl), _ptr(W) if W.size else None,
                     _ptr(senses), _ptr(Kc), float(soft_weight), int(has_binaries),
                     int(is_symmetric), int(n_break_points))
        return p

    def __init__(self, H, f, f_theta, A, bu, bl, W, senses, nu, nx=None, K=None,
                 soft_weight=1e6, has_binaries=False, is_symmetric=True, n_break_points=0,
                 device=0, settings=None, dims=None):
        L = lib()
        p = self._pack(H, f, f_theta, A, bu, bl, W, senses, nu, nx, K, soft_weight, has_binaries,
                       is_symmetric, n_break_points, dims)
        m = self.m
        st = settings if settings is not None else default_settings()
        self._h = C.c_void_p()
        _check(L.lmpc_gpu_create(C.byref(p), C.byref(st), device, C.byref(self._h)))
        self.device = device
        self.uprev = np.zeros(self.nu)   # mirrors mpc.uprev (utils.jl:48)
        k, cap = C.c_int32(), C.c_int32()
        se = np.zeros(m, np.int32)
        _check(L.lmpc_gpu_get_info(self._h, C.byref(k), C.byref(cap), _ptr(se)))
        self.kernel, self.cap, self.senses = k.value, cap.value, se

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().lmpc_gpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup constants (for parity checks against qp2ldp) ------------------------------
    def ldp_constants(self):
        n, m, nth, nu = self.n, self.m, self.nth, self.nu
        out = dict(Mbar=np.zeros((m, n)), Gram=np.zeros((m, m)), Dth=np.zeros((m, nth)),
                   du=np.zeros(m), dl=np.zeros(m), Rinv=np.zeros((nu, n)),
                   Uth=np.zeros((nu, nth)), uoff=np.zeros(nu))
        _check(lib().lmpc_gpu_get_ldp(self._h, *[_ptr(out[k]) for k in
                                                 ("Mbar", "Gram", "Dth", "du", "dl", "Rinv", "Uth", "uoff")]))
        out["senses"], out["cap"] = np.array(self.senses, np.int32), int(self.cap)   # lmpc_gpu_get_info
        return out

    def tiering(self):
        """(fast_cap, redone): see lmpc_gpu_get_tiering"""
        fc, rd = C.c_int32(0), C.c_int64(0)
        _check(lib().lmpc_gpu_get_tiering(self._h, C.byref(fc), C.byref(rd)))
        return int(fc.value), int(rd.value)

    # -- solve(mpc, theta), batched ---------------------------------------------------------
    def solve(self, Theta, want_iters=True, want_active=True):
        """Theta: nth x N host array (or device tensor of N x nth contiguous = nth x N column-major).

        Returns dict(U nu x N, exitflag N, iters N, active_upper, active_lower (N x words))."""
        words = (self.m + 63) // 64
        if _is_device(Theta):
            import torch
            N = Theta.numel() // self.nth
            dev = Theta.device
            U = torch.empty((N, self.nu), dtype=torch.float64, device=dev)
            flag = torch.empty(N, dtype=torch.int32, device=dev)
            iters = torch.empty(N, dtype=torch.int32, device=dev) if want_iters else None
            au = torch.empty((N, words), dtype=torch.int64, device=dev) if want_active else None
            al = torch.empty((N, words), dtype=torch.int64, device=dev) if want_active else None
            _wait_producer(Theta)
            _check(lib().lmpc_gpu_solve_batch(self._h, _ptr(Theta), N, _ptr(U), _ptr(flag),
                                              _ptr(iters), _ptr(au), _ptr(al)))
            return dict(U=U.T, exitflag=flag, iters=iters, active_upper=au, active_lower=al)
        Th = _colmajor(Theta, self.nth)
        N = Th.shape[1]
        U = np.zeros((self.nu, N), order="F")
        flag = np.zeros(N, np.int32)
        iters = np.zeros(N, np.int32) if want_iters else None
        au = np.zeros((N, words), np.uint64) if want_active else None
        al = np.zeros((N, words), np.uint64) if want_active else None
        _check(lib().lmpc_gpu_solve_batch(self._h, _ptr(Th), N, _ptr(U), _ptr(flag), _ptr(iters),
                                          _ptr(au), _ptr(al)))
        return dict(U=U, exitflag=flag, iters=iters, active_upper=au, active_lower=al)

    def solve_warm(self, Theta, active_upper, active_lower, want_iters=True):
        """Warm-started solve (lmpc_gpu_solve_batch_warm): start every instance from the working set
        in active_upper / active_lower (N x words uint64, e.g. the previous solve's result).
        Host arrays (not modified) or device tensors (updated in place)."""
        words = (self.m + 63) // 64
        if _is_device(Theta):
            import torch
            N = Theta.numel() // self.nth
            dev = Theta.device
            U = torch.empty((N, self.nu), dtype=torch.float64, device=dev)
            flag = torch.empty(N, dtype=torch.int32, device=dev)
            iters = torch.empty(N, dtype=torch.int32, device=dev) if want_iters else None
            _wait_producer(Theta)
            _check(lib().lmpc_gpu_solve_batch_warm(self._h, _ptr(Theta), N, _ptr(U), _ptr(flag), _ptr(iters),
                                                   _ptr(active_upper), _ptr(active_lower)))
            return dict(U=U.T, exitflag=flag, iters=iters, active_upper=active_upper, active_lower=active_lower)
        Th = _colmajor(Theta, self.nth)
        N = Th.shape[1]
        U = np.zeros((self.nu, N), order="F")
        flag = np.zeros(N, np.int32)
        iters = np.zeros(N, np.int32) if want_iters else None
        au = np.ascontiguousarray(active_upper, np.uint64).reshape(N, words).copy()
        al = np.ascontiguousarray(active_lower, np.uint64).reshape(N, words).copy()
        _check(lib().lmpc_gpu_solve_batch_warm(self._h, _ptr(Th), N, _ptr(U), _ptr(flag), _ptr(iters),
                                               _ptr(au), _ptr(al)))
        return dict(U=U, exitflag=flag, iters=iters, active_upper=au, active_lower=al)

    def solve_async(self, Theta_dev, U_dev, flag_dev, iters_dev=None, au_dev=None, al_dev=None,
                    N=None, stream=None):
        """Device pointers only, no synchronisation (bench's device-resident leg)."""
        N = Theta_dev.numel() // self.nth if N is None else N
        _check(lib().lmpc_gpu_solve_batch_async(self._h, _ptr(Theta_dev), N, _ptr(U_dev),
                                                _ptr(flag_dev), _ptr(iters_dev), _ptr(au_dev),
                                                _ptr(al_dev), stream))

    def stream(self):
        return lib().lmpc_gpu_stream(self._h)

    # -- form_parameter / compute_control -----------------------------------------------------
    def form_parameter(self, x, r=None, d=None, uprev=None, p=None):
        """theta = [x; r; d; uprev; p] per column (explicit.jl:54-63); missing r/d/p default to 0 and
        missing uprev to the controller's stored previous control (utils.jl:48, explicit.jl:60).
        p: the generalized parameters (format_affine_parameters, utils.jl:225-240), nth - the rest rows."""
        if self.dims is None:
            raise ValueError("GpuMPC was built without dims=(nx, nr, nd, nuprev)")
        nx, nr, nd, nup = self.dims
        X = _colmajor(x, nx)
        N = X.shape[1]
        def blk(v, rows, default=None):
            if rows == 0:
                return np.zeros((0, N))
            if v is None:
                base = np.zeros(rows) if default is None else np.asarray(default, float)[:rows]
                return np.repeat(base.reshape(rows, 1), N, axis=1)
            v = np.asarray(v, float)
            # any other size is an error, as in format_reference / format_disturbance (utils.jl:78-205)
            if v.ndim == 1:
                if v.size != rows:
                    raise ValueError(f"vector length ({v.size}) must match ({rows})")
                return np.repeat(v.reshape(rows, 1), N, axis=1)
            if v.shape[0] != rows:
                raise ValueError(f"matrix must have {rows} rows, got {v.shape[0]}")
            if v.shape[1] == 1 and N > 1:
                return np.repeat(v, N, axis=1)
            if v.shape[1] != N:
                raise ValueError(f"matrix must have {N} columns (one per instance), got {v.shape[1]}")
            return v
        Rm, Dm, Um = blk(r, nr), blk(d, nd), blk(uprev, nup, self.uprev)
        if nd > 0 and Dm.shape[0] != nd:
            raise ValueError(f"Disturbance vector must have length {nd}")
        npar = max(0, getattr(self, "nth", nx + nr + nd + nup) - (nx + nr + nd + nup))
        if npar == 0 and p is not None and np.size(p):
            raise ValueError("this controller has no generalized parameters")
        return np.asfortranarray(np.vstack([X, Rm, Dm, Um, blk(p, npar)]))

    def compute_control(self, x, r=None, d=None, uprev=None, p=None, check=True):
        """Batched compute_control (utils.jl:43-51): returns nu x N (a vector for a single x)."""
        single = np.ndim(x) == 1
        Th = self.form_parameter(x, r, d, uprev, p)
        out = self.solve(Th, want_iters=False, want_active=False)
        if check:
            assert np.all(out["exitflag"] >= 1), "exitflag < 1 (utils.jl:46)"
        U = out["U"]
        self.uprev = U[:, -1].copy()
        return U[:, 0].copy() if single else U

    def compute_control_bcast(self, x, r=None, d=None, uprev=None, want_flags=False):
        """Monte-Carlo form (lmpc_gpu_compute_control_batch_bcast): x is nx x N, every one of r / d / uprev
        that is a VECTOR is shared by all instances and uploaded once; matrices are per instance.
        Returns (U nu x N, exitflag or None, number of instances with exitflag < 1)."""
        nx, nr, nd, nup = self.dims
        dev = hasattr(x, "data_ptr")
        if dev:
            import torch
            N = x.numel() // nx
            U = torch.empty((N, self.nu), dtype=torch.float64, device=x.device)
            flag = torch.empty(N, dtype=torch.int32, device=x.device) if want_flags else None
            X = x
        else:
            X = _colmajor(x, nx)
            N = X.shape[1]
            U = np.zeros((self.nu, N), order="F")
            flag = np.zeros(N, np.int32) if want_flags else None
        mask, blocks = 0, []
        for bit, v, rows in ((1, r, nr), (2, d, nd), (4, uprev if uprev is not None else self.uprev, nup)):
            if rows == 0:
                blocks.append(None)
                continue
            if v is None:
                v = np.zeros(rows)
            if hasattr(v, "data_ptr"):
                if v.numel() == rows:
                    mask |= bit
                blocks.append(v)
                continue
            v = np.asarray(v, float)
            if v.ndim == 1 or v.shape[1] == 1:
                if v.size != rows:
                    raise ValueError(f"vector length ({v.size}) must match ({rows})")
                mask |= bit
                v = np.ascontiguousarray(v.reshape(rows))
                if dev:
                    import torch
                    v = torch.as_tensor(v, device=x.device)
                blocks.append(v)
            else:
                if v.shape != (rows, N):
                    raise ValueError(f"matrix must be {rows} x {N}")
                if dev:
                    raise ValueError("per-instance blocks must be device tensors when x is one")
                blocks.append(np.asfortranarray(v))
        nfail = C.c_int64(0)
        _wait_producer(X)
        _check(lib().lmpc_gpu_compute_control_batch_bcast(self._h, _ptr(U), _ptr(X), _ptr(blocks[0]), _ptr(blocks[1]),
                                                          _ptr(blocks[2]), nr, nd, nup, N, mask, _ptr(flag),
                                                          C.addressof(nfail)))
        return (U.T if dev else U), flag, int(nfail.value)

    def compute_control_c(self, control, state, reference=None, disturbance=None):
        """Batched twin of the generated mpc_compute_control(control, state, reference,
        disturbance) (codegen/mpc_update_qp.h:3-7): `control` (nu x N) is in/out."""
        nx, nr, nd, nup = self.dims
        ctl = _colmajor(control, self.nu)
        st = _colmajor(state, nx)
        N = st.shape[1]
        ref = _colmajor(reference, nr) if nr > 0 else None
        dis = _colmajor(disturbance, nd) if (nd > 0 and disturbance is not None) else None
        flag = np.zeros(N, np.int32)
        _check(lib().lmpc_gpu_compute_control_batch(self._h, _ptr(ctl), _ptr(st), _ptr(ref), _ptr(dis),
                                                    nr, nd, nup, N, _ptr(flag)))
        return ctl, flag

    # -- Simulation, batched ----------------------------------------------------------------
    def simulate(self, F, G, x0, r=None, d=None, u0=None, N=None, T=None, Gd=None, warm_start=True,
                 want_traj=False, r_traj=None, d_traj=None, plant=None, observer=None, noise=None,
                 r_preview=0, d_preview=0):
        """N independent closed loops of `steps` steps (simulation.jl:93-113) with the default
        linear dynamics (model.jl:70).  `N` = number of steps, as in the reference's keyword.

        r, d: constant per closed loop (vector = the same for all, matrix = one column each).
        r_traj (nr x Tr), d_traj (nd x Td): the scenario's trajectories, shared by all closed loops;
        shorter than the horizon -> the last column is held (simulation.jl:70-80).
        r_preview = Np > 0: the controller was condensed with reference_preview, nr = ny_ref * Np; r_traj
        is then the ny_ref x Tr scenario reference and step k sees columns k+1 .. k+Np (last held), as
        get_preview does (simulation.jl:102,128-134); d_preview likewise with columns k .. k+Np-1.
        plant = dict(F=, G=, Gd=): the true plant when it differs from the prediction model
        (scenario.dynamics); observer = dict(C=, K=, Dd=): steady-state Kalman observer run in the
        loop (observer.jl:104-118), the controller then acts on the estimate; noise (ny x T):
        additive measurement noise shared by all closed loops.  With any of them the result also
        has Xhat_end (final estimates) and Xend is the TRUE final state."""
        nx, nr, nd, nup = self.dims
        steps = int(N if N is not None else T)
        X0 = _colmajor(x0, nx)
        n = X0.shape[1]
        Fm, Gm = np.asfortranarray(np.asarray(F, float)), np.asfortranarray(np.asarray(G, float).reshape(nx, -1))
        Gdm = np.asfortranarray(np.asarray(Gd, float).reshape(nx, -1)) if (nd > 0 and Gd is not None) else None
        if Gdm is not None and Gdm.shape[1] < nd:      # disturbance preview: only d_k itself drives the plant
            Gdm = np.asfortranarray(np.hstack([Gdm, np.zeros((nx, nd - Gdm.shape[1]))]))
        def bc(v, rows):
            if rows == 0:
                return None
            if v is None:
                return np.zeros((rows, n), order="F")
            v = np.asarray(v, float)
            if v.ndim == 1:
                v = np.repeat(v.reshape(rows, 1), n, axis=1)
            return np.asfortranarray(v)
        Rm, Dm, Um = bc(r, nr), bc(d, nd), bc(u0, nup) if u0 is not None else None
        hold = lambda v, rows: hold_trajectory(v, rows, steps)
        rstride = dstride = 0
        window = lambda v, rows_total, Np, first: preview_window(v, rows_total, Np, first, steps)
        if r_preview and r_traj is not None and nr:
            Rt, rstride = window(r_traj, nr, int(r_preview), 1)
            Dt = hold(d_traj, nd)
        else:
            Rt, Dt = hold(r_traj, nr), hold(d_traj, nd)
        if d_preview and d_traj is not None and nd:
            Dt, dstride = window(d_traj, nd, int(d_preview), 0)
        Xend = np.zeros((nx, n), order="F")
        Ulast = np.zeros((self.nu, n), order="F")
        Utraj = np.zeros((n, steps, self.nu)) if want_traj else None
        flag = np.zeros(n, np.int32)
        iters = np.zeros(n, np.int32)
        cm = lambda a, rows: None if a is None else np.asfortranarray(np.asarray(a, float).reshape(rows, -1))
        ext = _SimExt()
        keep = [Rt, Dt]
        ext.Rtraj, ext.Dtraj = _ptr(Rt), _ptr(Dt)
        extended = plant is not None or observer is not None
        Xhat = np.zeros((nx, n), order="F") if extended else None
        if plant is not None:
            pm = [cm(plant.get("F"), nx), cm(plant.get("G"), nx), cm(plant.get("Gd"), nx) if nd > 0 else None]
            keep += pm
            ext.Ft, ext.Gt, ext.Gdt = _ptr(pm[0]), _ptr(pm[1]), _ptr(pm[2])
        if observer is not None:
            Cm = np.asfortranarray(np.atleast_2d(np.asarray(observer["C"], float)))
            ny = Cm.shape[0]
            Km = cm(observer["K"], nx)
            Ddm = cm(observer.get("Dd"), ny) if (nd > 0 and observer.get("Dd") is not None) else None
            Vm = None
            if noise is not None:
                Vm = hold(noise, ny)
            keep += [Cm, Km, Ddm, Vm]
            ext.C, ext.K, ext.Dd, ext.Vtraj, ext.ny = _ptr(Cm), _ptr(Km), _ptr(Ddm), _ptr(Vm), ny
        ext.Xhat_end = _ptr(Xhat)
        ext.rstride, ext.dstride = rstride, dstride
        _check(lib().lmpc_gpu_simulate_batch_ex(self._h, _ptr(Fm), _ptr(Gm), _ptr(Gdm), nx, nr, nd, nup,
                                                _ptr(X0), _ptr(Rm), _ptr(Dm), _ptr(Um), C.byref(ext),
                                                n, steps, int(bool(warm_start)), _ptr(Xend), _ptr(Ulast),
                                                _ptr(Utraj), _ptr(flag), _ptr(iters)))
        out = dict(Xend=Xend, Ulast=Ulast, Utraj=Utraj, exitflag=flag, iters=iters)
        if extended:
            out["Xhat_end"] = Xhat
        return out


    def simulate_stepped(self, F, G, x0, r=None, d=None, u0=None, N=100, Gd=None, warm_start=True, want_traj=False,
                         r_traj=None, d_traj=None, dynamics=None, plant=None, measurement=None, observer=None,
                         noise=None):
        """lmpc_gpu_simulate_batch_stepped: the Simulation loop stepped from the host (one kernel sequence per
        step, loop state in HBM) for a user-defined true plant and / or an observer with its own state.

        dynamics(k, X, U, D, Xnext, stream): the true plant (scenario.dynamics, simulation.jl:39,110); X, U, D,
          Xnext are torch views (N x nx, N x nu, N x ndm or None, N x nx) of the library's device buffers and
          the function must queue its work on `stream` (a torch.cuda.ExternalStream) writing Xnext in place.
        plant = dict(F=, G=, Gd=, f_offset=): an affine true plant instead (default: the model F, G, Gd).
        observer = dict(F=, G=, Gd=, f_offset=, C=, Dd=, h_offset=, K=): KalmanFilter (observer.jl:1-124) with
          nxo = F.shape[0] >= nx states; the extra nxo - nx states are the OffsetFreeObserver's disturbance
          estimates and take the LAST nxo - nx disturbance slots of theta.
        measurement = dict(C=, Dd=, h_offset=): the true measurement map (default: the observer's C / Dd / h_offset
          restricted to the plant state)."""
        nx, nr, nd, nup = self.dims
        steps = int(N)
        X0 = _colmajor(x0, nx)
        n = X0.shape[1]
        nxo = nx if observer is None else np.atleast_2d(np.asarray(observer["F"], float)).shape[0]
        ndo = nxo - nx
        ndm = nd - ndo
        if ndm < 0:
            raise ValueError("the observer estimates more disturbances than the controller has")
        cm = lambda a, rows: None if a is None else np.asfortranarray(np.asarray(a, float).reshape(rows, -1))
        vec = lambda a, rows: None if a is None else np.ascontiguousarray(np.asarray(a, float).reshape(rows))
        Fm, Gm = cm(F, nx), cm(G, nx)
        Gdm = cm(np.asarray(Gd, float).reshape(nx, -1)[:, :ndm], nx) if (ndm > 0 and Gd is not None) else None
        def bc(v, rows):
            if rows == 0:
                return None
            if v is None:
                return np.zeros((rows, n), order="F")
            v = np.asarray(v, float)
            if v.ndim == 1:
                v = np.repeat(v.reshape(rows, 1), n, axis=1)
            return np.asfortranarray(v)
        Rm, Dm, Um = bc(r, nr), bc(d, ndm), (bc(u0, nup) if u0 is not None else None)
        Rt, Dt = hold_trajectory(r_traj, nr, steps), hold_trajectory(d_traj, ndm, steps)
        cfg = _SimStepped()
        keep = [Rt, Dt]
        cfg.Rtraj, cfg.Dtraj, cfg.ndm, cfg.nxo = _ptr(Rt), _ptr(Dt), ndm, nxo
        if plant is not None:
            pm = [cm(plant.get("F"), nx), cm(plant.get("G"), nx), cm(plant.get("Gd"), nx) if ndm > 0 else None,
                  vec(plant.get("f_offset"), nx)]
            keep += pm
            cfg.Ft, cfg.Gt, cfg.Gdt, cfg.ft = [_ptr(a) for a in pm]
        cb = None
        if dynamics is not None:
            import torch
            def tramp(user, stream, Nn, k, X, U, D, Xn):
                try:
                    es = torch.cuda.ExternalStream(stream, device=self.device)
                    with torch.cuda.stream(es):
                        dynamics(int(k), device_view(X, nx, Nn, self.device), device_view(U, self.nu, Nn, self.device),
                                 device_view(D, ndm, Nn, self.device) if D else None,
                                 device_view(Xn, nx, Nn, self.device), es)
                    return 0
                except Exception:            # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return 1
            cb = PLANT_FN(tramp)
            cfg.dynamics = cb
        Xhat = np.zeros((nxo, n), order="F")
        cfg.Xhat_end = _ptr(Xhat)
        if observer is not None:
            Co = cm(observer["C"], -1) if np.ndim(observer["C"]) == 1 else np.asfortranarray(np.atleast_2d(np.asarray(observer["C"], float)))
            ny = Co.shape[0]
            om = dict(Fo=cm(observer["F"], nxo), Go=cm(observer["G"], nxo),
                      Gdo=cm(observer.get("Gd"), nxo) if ndm > 0 and observer.get("Gd") is not None else None,
                      fo=vec(observer.get("f_offset"), nxo), Co=Co,
                      Ddo=cm(observer.get("Dd"), ny) if ndm > 0 and observer.get("Dd") is not None else None,
                      ho=vec(observer.get("h_offset"), ny), Ko=cm(observer["K"], nxo))
            meas = measurement or {}
            mm = dict(Cm=cm(meas.get("C", Co[:, :nx]), ny),
                      Ddm=cm(meas.get("Dd", observer.get("Dd")), ny) if ndm > 0 and meas.get("Dd", observer.get("Dd")) is not None else None,
                      hm=vec(meas.get("h_offset", observer.get("h_offset")), ny),
                      Vtraj=hold_trajectory(noise, ny, steps) if noise is not None else None)
            keep += list(om.values()) + list(mm.values())
            cfg.ny = ny
            for k, v in {**om, **mm}.items():
                setattr(cfg, k, _ptr(v))
        Xend = np.zeros((nx, n), order="F")
        Ulast = np.zeros((self.nu, n), order="F")
        Utraj = np.zeros((n, steps, self.nu)) if want_traj else None
        flag, iters = np.zeros(n, np.int32), np.zeros(n, np.int32)
        _check(lib().lmpc_gpu_simulate_batch_stepped(self._h, _ptr(Fm), _ptr(Gm), _ptr(Gdm), nx, nr, nup, _ptr(X0),
                                                     _ptr(Rm), _ptr(Dm), _ptr(Um), C.byref(cfg), C.c_int64(n), steps,
                                                     int(bool(warm_start)), _ptr(Xend), _ptr(Ulast), _ptr(Utraj),
                                                     _ptr(flag), _ptr(iters)))
        return dict(Xend=Xend, Ulast=Ulast, Utraj=Utraj, exitflag=flag, iters=iters, Xhat_end=Xhat)


PLANT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                       C.c_void_p)


class _SimStepped(C.Structure):
    """ctypes mirror of lmpc_sim_stepped (include/lmpc_gpu.h)"""
    _fields_ = [("dynamics", PLANT_FN), ("user", C.c_void_p)] + \
               [(k, C.c_void_p) for k in ("Ft", "Gt", "Gdt", "ft")] + [("ny", C.c_int32)] + \
               [(k, C.c_void_p) for k in ("Cm", "Ddm", "hm", "Vtraj")] + [("nxo", C.c_int32), ("ndm", C.c_int32)] + \
               [(k, C.c_void_p) for k in ("Fo", "Go", "Gdo", "fo", "Co", "Ddo", "ho", "Ko", "Rtraj", "Dtraj", "Xhat_end")]


class _DevArray:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = dict(shape=shape, typestr="<f8", data=(int(ptr), False), version=2, strides=None)


def device_view(ptr, rows, n, device=0):
    """(n x rows) torch view of a column-major rows x n device array owned by the library (no copy)."""
    import torch
    return torch.as_tensor(_DevArray(ptr, (int(n), int(rows))), device=torch.device("cuda", device))


def setup_gpu(mpqp, nu, dims=None, K=None, soft_weight=1e6, device=0, settings=None, devices=None,
              **kw):
    """`setup_gpu!(mpc)`: build the device-resident backend from an MPQP-like object with
    fields H, f, f_theta, A, bu, bl, W, senses (types.jl:75-98).  `devices=[...]` builds the
    single-process multi-GPU variant (host batches only)."""
    g = lambda k, default=None: getattr(mpqp, k, default) if not isinstance(mpqp, dict) else mpqp.get(k, default)
    nx = dims[0] if dims is not None else kw.pop("nx", None)
    common = dict(nu=nu, nx=nx, K=K, soft_weight=soft_weight,
                  has_binaries=bool(g("has_binaries", False)),
                  is_symmetric=bool(g("is_symmetric", True)),
                  n_break_points=len(g("break_points", ()) or ()), settings=settings, dims=dims)
    args = (g("H"), g("f"), g("f_theta"), g("A"), g("bu"), g("bl"), g("W"), g("senses"))
    if devices is not None:
        return MultiGpuMPC(*args, devices=devices, **common)
    return GpuMPC(*args, device=device, **common)


class MultiGpuMPC(GpuMPC):
    """One process, several GPUs: a replica of the controller on every listed device; `solve`
    splits a HOST batch into contiguous theta ranges, one per GPU, with no inter-GPU traffic
    (include/lmpc_gpu.h: lmpc_gpu_create_multi / lmpc_gpu_solve_batch_multi)."""

    def __init__(self, H, f, f_theta, A, bu, bl, W, senses, nu, nx=None, K=None,
                 soft_weight=1e6, has_binaries=False, is_symmetric=True, n_break_points=0,
                 devices=None, settings=None, dims=None):
        L = lib()
        p = self._pack(H, f, f_theta, A, bu, bl, W, senses, nu, nx, K, soft_weight, has_binaries,
                       is_symmetric, n_break_points, dims)
        st = settings if settings is not None else default_settings()
        devices = list(range(device_count())) if devices is None else list(devices)
        dv = np.ascontiguousarray(devices, np.int32)
        self._h = None
        self._mh = C.c_void_p()
        _check(L.lmpc_gpu_create_multi(C.byref(p), C.byref(st), _ptr(dv), len(devices),
                                       C.byref(self._mh)))
        self.devices = devices
        self.uprev = np.zeros(self.nu)

    def solve(self, Theta, want_iters=True, want_active=True):
        words = (self.m + 63) // 64
        Th = _colmajor(Theta, self.nth)
        N = Th.shape[1]
        U = np.zeros((self.nu, N), order="F")
        flag = np.zeros(N, np.int32)
        iters = np.zeros(N, np.int32) if want_iters else None
        au = np.zeros((N, words), np.uint64) if want_active else None
        al = np.zeros((N, words), np.uint64) if want_active else None
        _check(lib().lmpc_gpu_solve_batch_multi(self._mh, _ptr(Th), N, _ptr(U), _ptr(flag),
                                                _ptr(iters), _ptr(au), _ptr(al)))
        return dict(U=U, exitflag=flag, iters=iters, active_upper=au, active_lower=al)

    def simulate(self, F, G, x0, r=None, d=None, u0=None, N=100, Gd=None, warm_start=True):
        """N_traj closed loops split over the devices (lmpc_gpu_simulate_batch_multi): constant r / d per loop."""
        nx, nr, nd, nup = self.dims
        X0 = _colmajor(x0, nx)
        n = X0.shape[1]
        def bc(v, rows):
            if rows == 0:
                return None
            if v is None:
                return np.zeros((rows, n), order="F")
            v = np.asarray(v, float)
            return np.asfortranarray(np.repeat(v.reshape(rows, 1), n, axis=1) if v.ndim == 1 else v)
        Rm, Dm, Um = bc(r, nr), bc(d, nd), (bc(u0, nup) if u0 is not None else None)
        Fm, Gm = np.asfortranarray(np.asarray(F, float)), np.asfortranarray(np.asarray(G, float).reshape(nx, -1))
        Gdm = np.asfortranarray(np.asarray(Gd, float).reshape(nx, -1)) if (nd > 0 and Gd is not None) else None
        Xend, Ulast = np.zeros((nx, n), order="F"), np.zeros((self.nu, n), order="F")
        flag, iters = np.zeros(n, np.int32), np.zeros(n, np.int32)
        _check(lib().lmpc_gpu_simulate_batch_multi(self._mh, _ptr(Fm), _ptr(Gm), _ptr(Gdm), nx, nr, nd, nup,
                                                   _ptr(X0), _ptr(Rm), _ptr(Dm), _ptr(Um), None, n, int(N),
                                                   int(bool(warm_start)), _ptr(Xend), _ptr(Ulast), None,
                                                   _ptr(flag), _ptr(iters)))
        return dict(Xend=Xend, Ulast=Ulast, exitflag=flag, iters=iters)

    def close(self):
        if getattr(self, "_mh", None):
            lib().lmpc_gpu_destroy_multi(self._mh)
            self._mh = None

    __del__ = close


class ResultWindow:
    """The final result gather fused into the solve (lmpc_gpu_window_*): a device buffer on one GPU that the
    other processes of the box map through CUDA IPC; a rank passes `ptr + its byte offset` as the U / flag
    output of solve_async and its kernel's stores land in the gathered array over NVLink."""

    def __init__(self, device, nbytes=None, handle=None):
        self.owner = handle is None
        self.device = device
        p = C.c_void_p()
        if self.owner:
            self.handle = (C.c_ubyte * 64)()
            _check(lib().lmpc_gpu_window_create(device, int(nbytes), C.byref(p), self.handle))
        else:
            self.handle = (C.c_ubyte * 64)(*bytes(handle))
            _check(lib().lmpc_gpu_window_open(device, self.handle, C.byref(p)))
        self.ptr = p.value

    def handle_bytes(self):
        return bytes(self.handle)

    def read(self, offset, shape, dtype):
        """host copy of part of the window"""
        out = np.empty(shape, dtype)
        _check(lib().lmpc_gpu_window_read(out.ctypes.data, self.ptr + int(offset), out.nbytes))
        return out

    def close(self):
        if getattr(self, "ptr", None):
            lib().lmpc_gpu_window_close(self.ptr, int(self.owner))
            self.ptr = None

    __del__ = close


class GpuExplicitMPC:
    """Device-resident explicit controller: what `compute_control(empc, ...)` evaluates.

    halfplanes: n_hp x (nth+1) rows [a, b]; feedbacks: n_fb x nu x (nth+1) rows [F, g];
    hp_list / jump_list: node arrays (0-based; jump 0 = leaf); clipping: nu x 2 or None.
    """

    def __init__(self, halfplanes, feedbacks, hp_list, jump_list, clipping=None, dims=None, device=0,
                 le_child=0, devices=None):
        """le_child: the child taken when a'theta <= b (0: `next`, this library's and its builder's
        convention; 1: `next + 1`, see lmpc_gpu_create_bst_oriented).  devices: a list builds one replica
        per GPU (lmpc_gpu_create_bst_multi; host batches only, le_child must be 0)."""
        hp = np.ascontiguousarray(halfplanes, np.float64)          # row-major n_hp x ld == col-major ld x n_hp
        fb = np.ascontiguousarray(feedbacks, np.float64)           # n_fb x nu x ld: block r col-major ld x nu
        self.nth = hp.shape[1] - 1 if hp.size else fb.shape[2] - 1
        self.nu = fb.shape[1]
        hl = np.ascontiguousarray(hp_list, np.int32)
        jl = np.ascontiguousarray(jump_list, np.int32)
        clip = None if clipping is None else np.asfortranarray(np.asarray(clipping, np.float64).reshape(self.nu, 2))
        self._keep = (hp, fb, hl, jl, clip)
        self.dims = dims
        self.uprev = np.zeros(self.nu)
        self._t = C.c_void_p()
        self._mt = None
        if devices is not None:
            if le_child != 0:
                raise ValueError("multi-device trees take the library's own orientation (le_child = 0)")
            dv = np.ascontiguousarray(list(devices), np.int32)
            self._mt = C.c_void_p()
            _check(lib().lmpc_gpu_create_bst_multi(_ptr(hp), hp.shape[0], _ptr(fb), fb.shape[0], _ptr(hl), _ptr(jl),
                                                   len(hl), _ptr(clip), self.nth, self.nu, _ptr(dv), len(dv),
                                                   C.byref(self._mt)))
            self._t = None
            return
        _check(lib().lmpc_gpu_create_bst_oriented(_ptr(hp), hp.shape[0], _ptr(fb), fb.shape[0], _ptr(hl), _ptr(jl),
                                                  len(hl), _ptr(clip), self.nth, self.nu, int(le_child), device,
                                                  C.byref(self._t)))

    @classmethod
    def from_codegen(cls, path_c, path_h=None, clipping=None, dims=None, device=0, le_child=None):
        """A tree computed by the real package, read from the C arrays ParametricDAQP.codegen emitted
        (src/codegen.jl:62-66).  The branch orientation is taken from the generated evaluate function in the
        file; le_child overrides it."""
        from . import explicit as ex
        t = ex.load_pdaqp_codegen(path_c, path_h)
        lc = t["le_child"] if le_child is None else le_child
        if lc is None:
            raise ValueError("branch orientation not found in the file: pass le_child = 0 or 1")
        return cls(t["halfplanes"], t["feedbacks"], t["hp_list"], t["jump_list"], clipping=clipping, dims=dims,
                   device=device, le_child=lc)

    @property
    def kernel(self):
        """Evaluate kernel of this tree: 0 generic fp64 rows, 1 fp32-screened walk, 2 fp32-screened lane-refill ring."""
        return int(lib().lmpc_gpu_bst_kernel(self._t)) if getattr(self, "_t", None) else -1

    def close(self):
        if getattr(self, "_t", None) is not None and self._t:
            lib().lmpc_gpu_destroy_bst(self._t)
            self._t = None
        if getattr(self, "_mt", None) is not None and self._mt:
            lib().lmpc_gpu_destroy_bst_multi(self._mt)
            self._mt = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def evaluate(self, Theta, want_region=True):
        if _is_device(Theta):
            import torch
            N = Theta.numel() // self.nth
            U = torch.empty((N, self.nu), dtype=torch.float64, device=Theta.device)
            reg = torch.empty(N, dtype=torch.int32, device=Theta.device) if want_region else None
            _wait_producer(Theta)
            _check(lib().lmpc_gpu_evaluate_batch(self._t, _ptr(Theta), N, _ptr(U), _ptr(reg)))
            return U.T, reg
        Th = _colmajor(Theta, self.nth)
        N = Th.shape[1]
        U = np.zeros((self.nu, N), order="F")
        reg = np.zeros(N, np.int32) if want_region else None
        if self._mt is not None:
            _check(lib().lmpc_gpu_evaluate_batch_multi(self._mt, _ptr(Th), N, _ptr(U), _ptr(reg)))
        else:
            _check(lib().lmpc_gpu_evaluate_batch(self._t, _ptr(Th), N, _ptr(U), _ptr(reg)))
        return U, reg

    def evaluate_async(self, Theta_dev, U_dev, region_dev=None, N=None, stream=None):
        N = Theta_dev.numel() // self.nth if N is None else N
        _check(lib().lmpc_gpu_evaluate_batch_async(self._t, _ptr(Theta_dev), N, _ptr(U_dev),
                                                   _ptr(region_dev), stream))

    form_parameter = GpuMPC.form_parameter

    def compute_control(self, x, r=None, d=None, uprev=None, check=True):
        single = np.ndim(x) == 1
        U, _ = self.evaluate(self.form_parameter(x, r, d, uprev), want_region=False)
        self.uprev = U[:, -1].copy()
        return U[:, 0].copy() if single else U

    def simulate(self, F, G, x0, r=None, d=None, u0=None, N=100, Gd=None, want_traj=False,
                 r_traj=None, d_traj=None):
        """Simulation(empc, scenario) (simulation.jl:37-116) for every column of x0: N steps of the
        explicit law + linear dynamics on the GPU.  Arguments as GpuMPC.simulate."""
        if self.dims is None:
            raise ValueError("GpuExplicitMPC was built without dims=(nx, nr, nd, nuprev)")
        nx, nr, nd, nup = self.dims
        steps = int(N)
        X0 = _colmajor(x0, nx)
        n = X0.shape[1]
        Fm, Gm = np.asfortranarray(np.asarray(F, float)), np.asfortranarray(np.asarray(G, float).reshape(nx, -1))
        Gdm = np.asfortranarray(np.asarray(Gd, float).reshape(nx, -1)) if (nd > 0 and Gd is not None) else None
        def bc(v, rows):
            if rows == 0:
                return None
            if v is None:
                return np.zeros((rows, n), order="F")
            v = np.asarray(v, float)
            if v.ndim == 1:
                v = np.repeat(v.reshape(rows, 1), n, axis=1)
            return np.asfortranarray(v)
        def hold(v, rows):
            if v is None or rows == 0:
                return None
            v = np.asarray(v, float).reshape(rows, -1)
            if v.shape[1] < steps:
                v = np.hstack([v, np.repeat(v[:, -1:], steps - v.shape[1], axis=1)])
            return np.asfortranarray(v[:, :steps])
        Rm, Dm, Um = bc(r, nr), bc(d, nd), bc(u0, nup) if u0 is not None else None
        Rt, Dt = hold(r_traj, nr), hold(d_traj, nd)
        Xend = np.zeros((nx, n), order="F")
        Ulast = np.zeros((self.nu, n), order="F")
        Utraj = np.zeros((n, steps, self.nu)) if want_traj else None
        _check(lib().lmpc_gpu_simulate_bst_batch(self._t, _ptr(Fm), _ptr(Gm), _ptr(Gdm), nx, nr, nd, nup,
                                                 _ptr(X0), _ptr(Rm), _ptr(Dm), _ptr(Um), _ptr(Rt), _ptr(Dt),
                                                 n, steps, _ptr(Xend), _ptr(Ulast), _ptr(Utraj)))
        return dict(Xend=Xend, Ulast=Ulast, Utraj=Utraj)


# ---------------------------------------------------------------------------------------------
# Builder mirror: MPC(F, G; ...) / set_objective! / set_bounds! / add_constraint! / setup!
# (src/types.jl:162-172, src/setup.jl:36-132) on top of lmpc_gpu_condense.  No oracle involved.
# ---------------------------------------------------------------------------------------------
class _Constraint(C.Structure):
    _fields_ = [("rows", C.c_int32), ("Au", C.c_void_p), ("Ax", C.c_void_p), ("Ar", C.c_void_p),
                ("Ad", C.c_void_p), ("Aup", C.c_void_p), ("ub", C.c_void_p), ("lb", C.c_void_p),
                ("ks", C.c_void_p), ("nks", C.c_int32), ("soft", C.c_int32), ("prio", C.c_int32),
                ("Ap", C.c_void_p)]


class _Model(C.Structure):
    _fields_ = [("nx", C.c_int32), ("nu", C.c_int32), ("ny", C.c_int32), ("nd", C.c_int32),
                ("Np", C.c_int32), ("Nc", C.c_int32),
                ("F", C.c_void_p), ("G", C.c_void_p), ("C", C.c_void_p), ("Gd", C.c_void_p), ("Dd", C.c_void_p),
                ("Q", C.c_void_p), ("R", C.c_void_p), ("Rr", C.c_void_p), ("Qf", C.c_void_p),
                ("umin", C.c_void_p), ("umax", C.c_void_p),
                ("constraints", C.c_void_p), ("n_constraints", C.c_int32),
                ("reference_tracking", C.c_int32), ("preprocess", C.c_int32),
                ("move_blocks", C.c_void_p), ("n_move_blocks", C.c_void_p),
                ("reference_preview", C.c_int32), ("disturbance_preview", C.c_int32), ("K", C.c_void_p),
                ("np", C.c_int32), ("Ex", C.c_void_p), ("ex", C.c_void_p), ("Eu", C.c_void_p), ("eu", C.c_void_p),
                ("parameter_preview", C.c_int32),
                ("f_offset", C.c_void_p), ("h_offset", C.c_void_p), ("xo", C.c_void_p), ("uo", C.c_void_p),
                ("reference_condensation", C.c_int32)]


def _diag(v, n):
    v = np.asarray(v, np.float64)
    if v.ndim == 0:
        return np.eye(n) * float(v)
    return np.diag(v) if v.ndim == 1 else v


class MPC:
    """`MPC(F, G; C, Gd, Dd, Np, Nc)` with the reference's builder calls; `setup()` condenses on the
    host (C++) and uploads to the GPU, after which `compute_control` / `solve` / `simulate` work as
    on GpuMPC.  Discrete-time models only (use your own zoh for continuous ones, utils.jl:291-295)."""

    def __init__(self, F, G, C=None, Gd=None, Dd=None, Np=10, Nc=None, Ts=-1.0, f_offset=None, h_offset=None,
                 xo=None, uo=None):
        f64 = lambda a: np.asarray(a, np.float64)
        self.F = np.atleast_2d(f64(F)); self.nx = self.F.shape[0]
        self.G = f64(G).reshape(self.nx, -1); self.nu = self.G.shape[1]
        self.C = None if C is None else np.atleast_2d(f64(C))
        self.ny = self.nx if self.C is None else self.C.shape[0]
        gd = None if Gd is None else f64(Gd).reshape(self.nx, -1)
        dd = None if Dd is None else f64(Dd).reshape(self.ny, -1)
        self.nd = max(0 if gd is None else gd.shape[1], 0 if dd is None else dd.shape[1])
        pad = lambda a, r: None if self.nd == 0 else (np.zeros((r, self.nd)) if a is None else np.hstack([a, np.zeros((r, self.nd - a.shape[1]))]))
        self.Gd, self.Dd = pad(gd, self.nx), pad(dd, self.ny)
        self.Np, self.Nc, self.Ts = int(Np), int(Nc if Nc is not None else Np), Ts
        self.Q = self.R = self.Rr = self.Qf = None
        self.umin = self.umax = None
        self.constraints = []
        self.reference_tracking, self.preprocess_mpqp, self.soft_weight = True, True, 1e6
        self.reference_preview = self.disturbance_preview = False   # types.jl:56-57
        self.reference_condensation, self.traj2setpoint = False, None   # types.jl:54,65; set by condense()
        self.K = None                                               # prestabilising feedback, setup.jl:186-199
        self.Ex = self.ex = self.Eu = self.eu = None                # generalized-parameter cost, setup.jl:118-132
        self.parameter_preview = False                              # types.jl:58,69
        # affine model / operating point (Model(...; f_offset, h_offset, xo, uo), model.jl:45-72)
        vec = lambda v, n: None if v is None or np.size(v) == 0 else f64(v).reshape(n)
        self.f_offset, self.h_offset = vec(f_offset, self.nx), vec(h_offset, self.ny)
        self.xo, self.uo = vec(xo, self.nx), vec(uo, self.nu)
        self.move_blocks = []
        self.gpu = None
        self.mpqp = None
        self.observer, self.nd_measured, self._ndo = None, None, 0   # set_state_observer / set_offset_free_observer

    def set_prestabilizing_feedback(self, K=None):
        """set_prestabilizing_feedback!(mpc[, K]) (setup.jl:186-199): u = v - K x with the QP over v;
        without K the infinite-horizon LQR gain of (F, G, R + Rr, C'QC), as the reference's `ared` call."""
        if K is None:
            from scipy.linalg import solve_discrete_are
            Cm = np.eye(self.nx) if self.C is None else self.C
            Q = np.eye(self.ny) if self.Q is None else self.Q
            R = (np.eye(self.nu) if self.R is None else self.R) + (0 if self.Rr is None else self.Rr)
            P = solve_discrete_are(self.F, self.G, Cm.T @ Q @ Cm, R)
            K = np.linalg.solve(R + self.G.T @ P @ self.G, self.G.T @ P @ self.F)
        self.K = np.asarray(K, np.float64).reshape(self.nu, self.nx)
        self.gpu = self.mpqp = None

    # setup.jl:328-340 (KalmanFilter, observer.jl:50-68)
    @staticmethod
    def kalman_gain(F, Cm, Q=None, R=None):
        """Steady-state gain of KalmanFilter(F, G, C; Q, R): P = ared(F', C', R, Q), K = P C' (C P C' + R)^-1."""
        from scipy.linalg import solve_discrete_are
        F, Cm = np.atleast_2d(np.asarray(F, float)), np.atleast_2d(np.asarray(Cm, float))
        nx, ny = F.shape[0], Cm.shape[0]
        Q = np.eye(nx) if Q is None else _diag(Q, nx)
        R = np.eye(ny) if R is None else _diag(R, ny)
        P = solve_discrete_are(F.T, Cm.T, Q, R)
        return P @ Cm.T @ np.linalg.inv(Cm @ P @ Cm.T + R)

    def set_state_observer(self, Q=None, R=None, K=None):
        """set_state_observer!(mpc; Q, R): steady-state Kalman filter on the prediction model, run inside
        `simulate` (correct! -> compute_control on the estimate -> predict!, simulation.jl:92-108)."""
        Cm = np.eye(self.nx) if self.C is None else self.C
        self.observer = dict(F=self.F, G=self.G, Gd=self.Gd, f_offset=self.f_offset, C=Cm, Dd=self.Dd,
                             h_offset=self.h_offset, K=self.kalman_gain(self.F, Cm, Q, R) if K is None else np.asarray(K, float))
        self.nd_measured = self.nd
        return self.observer

    def set_offset_free_observer(self, method="state_disturbance", Q=None, R=None, K=None, Bd=None, Cd=None,
                                 Kx=None, Kd=None):
        """set_offset_free_observer!(mpc; method, Q, R, K, Bd, Cd, Kx, Kd) (setup.jl:393-487): the controller model
        gets ndo constant disturbance channels [Gd Bd], [Dd Cd] and the observer estimates [x; d]; in the closed loop
        the estimated disturbances are appended to the measured ones (observer.jl:203-222).  Methods as in the
        reference: state_disturbance / velocity (Bd = K, Cd = I - C K, Kx = K, Kd = I), output_disturbance
        (Bd = 0, Cd = I), general (Bd, Cd given)."""
        nx, ny = self.nx, self.ny
        Cm = np.eye(nx) if self.C is None else self.C
        ndm = getattr(self, "nd_measured", None)
        ndm = self.nd if ndm is None or getattr(self, "_ndo", 0) == 0 else ndm
        Gd = np.zeros((nx, 0)) if self.Gd is None else self.Gd[:, :ndm]          # strip_offset_free_model
        Dd = np.zeros((ny, 0)) if self.Dd is None else self.Dd[:, :ndm]
        if method in ("state_disturbance", "velocity"):
            K = self.kalman_gain(self.F, Cm, Q, R) if K is None else np.asarray(K, float).reshape(nx, ny)
            Bd, Cd, Kx, Kd = K, np.eye(ny) - Cm @ K, K, np.eye(ny)
        elif method == "output_disturbance":
            Bd, Cd = np.zeros((nx, ny)), np.eye(ny)
        elif method == "general":
            if Bd is None or Cd is None:
                raise ValueError("method 'general' requires Bd and Cd")
            Bd, Cd = np.asarray(Bd, float).reshape(nx, -1), np.asarray(Cd, float).reshape(ny, -1)
        else:
            raise ValueError(f"unknown offset-free method {method}")
        ndo = Bd.shape[1]
        if np.linalg.matrix_rank(np.block([[self.F - np.eye(nx), Bd], [Cm, Cd]])) != nx + ndo:
            raise ValueError("Offset-free disturbance model violates rank([F-I Bd; C Cd]) = nx + nd")
        Fa = np.block([[self.F, Bd], [np.zeros((ndo, nx)), np.eye(ndo)]])
        Ga = np.vstack([self.G, np.zeros((ndo, self.nu))])
        Gda = np.vstack([Gd, np.zeros((ndo, ndm))])
        Ca = np.hstack([Cm, Cd])
        fa = None if self.f_offset is None else np.concatenate([self.f_offset, np.zeros(ndo)])
        if Kx is not None or Kd is not None:
            Ka = np.vstack([np.zeros((nx, ny)) if Kx is None else np.asarray(Kx, float).reshape(nx, ny),
                            np.zeros((ndo, ny)) if Kd is None else np.asarray(Kd, float).reshape(ndo, ny)])
        else:
            Ka = self.kalman_gain(Fa, Ca, Q, R)
        self.observer = dict(F=Fa, G=Ga, Gd=Gda if ndm else None, f_offset=fa, C=Ca, Dd=Dd if ndm else None,
                             h_offset=self.h_offset, K=Ka)
        self.Gd, self.Dd = np.hstack([Gd, Bd]), np.hstack([Dd, Cd])                # append_offset_free_model
        self.nd, self.nd_measured, self._ndo = ndm + ndo, ndm, ndo
        self.gpu = self.mpqp = None
        return self.observer

    # setup.jl:196-251
    def move_block(self, block):
        """move_block!(mpc, block): keep the control constant in blocks, e.g. [2, 1, 3].  An int gives
        equal blocks, a list applies to every input, a list of lists is per input; blocks are padded /
        clipped to the prediction horizon and Nc becomes max(sum(block[:-1])) + 1, as in the reference."""
        self.gpu = None
        scalar = isinstance(block, (int, float, np.integer, np.floating))
        if block is None or (scalar and int(block) <= 0) or (not scalar and len(block) == 0):
            self.move_blocks, self.Nc = [], self.Np
            return
        if scalar:
            block = [int(block)] * (self.Np // int(block) + 1)
        if isinstance(block[0], (int, float, np.integer, np.floating)):
            block = [list(block) for _ in range(self.nu)]
        if len(block) != self.nu:
            raise ValueError("move_block: need blocks for every control input")

        def fmt(b):                                  # format_move_block, setup.jl:235-251
            b = [int(v) for v in b]
            tot = sum(b)
            if tot < self.Np:
                b[-1] += self.Np - tot
            elif tot > self.Np:
                acc, i = 0, 0
                while True:
                    acc += b[i]
                    if acc >= self.Np:
                        break
                    i += 1
                b = b[:i + 1]
                b[-1] += self.Np - acc
            return b
        self.move_blocks = [fmt(b) for b in block]
        self.Nc = max(sum(b[:-1]) for b in self.move_blocks) + 1

    # setup.jl:118-132
    def set_objective(self, Q=None, R=None, Rr=None, Qf=None, Ex=None, ex=None, Eu=None, eu=None):
        if Q is not None: self.Q = _diag(Q, self.ny)
        if R is not None: self.R = _diag(R, self.nu)
        if Rr is not None: self.Rr = _diag(Rr, self.nu)
        if Qf is not None: self.Qf = _diag(Qf, self.ny)
        # linear cost (Ex p + ex)'x_k + (Eu p + eu)'u_k in the generalized parameter p
        if Ex is not None: self.Ex = np.asarray(Ex, np.float64).reshape(self.nx, -1)
        if ex is not None: self.ex = np.asarray(ex, np.float64).reshape(self.nx)
        if Eu is not None: self.Eu = np.asarray(Eu, np.float64).reshape(self.nu, -1)
        if eu is not None: self.eu = np.asarray(eu, np.float64).reshape(self.nu)
        self.gpu = None

    def affine_parameter_dim(self):
        """np: the widest generalized-parameter block in use (get_affine_parameter_base_dim, utils.jl:207-217)."""
        dims = [0] + [E.shape[1] for E in (self.Ex, self.Eu) if E is not None]
        dims += [c["Ap"].shape[1] for c in self.constraints if c["Ap"] is not None]
        return max(dims)

    # setup.jl:36-46
    def set_input_bounds(self, umin=None, umax=None):
        big = 1e30
        lo = np.atleast_1d(np.asarray([] if umin is None else umin, np.float64))
        hi = np.atleast_1d(np.asarray([] if umax is None else umax, np.float64))
        if max(len(lo), len(hi)) == 0:
            return
        self.umin = np.concatenate([lo, -big * np.ones(self.nu - len(lo))])
        self.umax = np.concatenate([hi, big * np.ones(self.nu - len(hi))])
        self.gpu = None

    # setup.jl:61-84
    def add_constraint(self, Ax=None, Au=None, Ar=None, Ad=None, Aup=None, Ap=None, ub=(), lb=(), ks=None,
                       soft=False, prio=0):
        if Ax is None and Au is None:
            return
        ub = np.atleast_1d(np.asarray(ub, np.float64)); lb = np.atleast_1d(np.asarray(lb, np.float64))
        m = max(len(ub), len(lb))
        if m == 0:
            return
        ub = np.concatenate([ub, 1e30 * np.ones(m - len(ub))]); lb = np.concatenate([lb, -1e30 * np.ones(m - len(lb))])
        mat = lambda a, c: None if a is None else np.asfortranarray(np.asarray(a, np.float64).reshape(m, c))
        ks = np.ascontiguousarray(list(range(2, self.Np + 1)) if ks is None else list(ks), np.int32)
        self.constraints.append(dict(rows=m, Au=mat(Au, self.nu), Ax=mat(Ax, self.nx), Ar=mat(Ar, self.ny),
                                     Ad=mat(Ad, self.nd), Aup=mat(Aup, self.nu), ub=ub, lb=lb, ks=ks,
                                     Ap=None if Ap is None else np.asfortranarray(np.asarray(Ap, np.float64).reshape(m, -1)),
                                     soft=int(bool(soft)), prio=int(prio)))
        self.gpu = None

    # setup.jl:94-98 (soft by default)
    def set_output_bounds(self, ymin=(), ymax=(), ks=None, soft=True, prio=0):
        Cm = np.eye(self.nx) if self.C is None else self.C
        self.add_constraint(Ax=Cm, Ad=self.Dd, lb=ymin, ub=ymax, ks=ks, soft=soft, prio=prio)

    # setup.jl:105-108
    def set_bounds(self, umin=None, umax=None, ymin=(), ymax=()):
        if umin is not None or umax is not None:
            self.set_input_bounds(umin, umax)
        if len(np.atleast_1d(ymin)) or len(np.atleast_1d(ymax)):
            self.set_output_bounds(ymin, ymax)

    def condense(self):
        """mpc2mpqp on the host (C++): returns the dense mpQP as numpy arrays (row-major views)."""
        L = lib()
        L.lmpc_gpu_condense.argtypes = [C.POINTER(_Model), C.POINTER(C.c_void_p)]
        L.lmpc_gpu_mpqp_dims.argtypes = [C.c_void_p, C.c_void_p]
        L.lmpc_gpu_mpqp_get.argtypes = [C.c_void_p] * 9
        L.lmpc_gpu_mpqp_free.argtypes = [C.c_void_p]
        L.lmpc_gpu_condense_error.restype = C.c_char_p
        keep = []
        def cm(a):
            if a is None:
                return None
            a = np.asfortranarray(np.asarray(a, np.float64)); keep.append(a)
            return a.ctypes.data
        npar = self.affine_parameter_dim()
        def wide(a, rows):                       # every block in use must have np columns (mpc2mpqp.jl:486-490)
            if a is not None and a.shape[1] != npar:
                raise ValueError(f"generalized-parameter blocks (Ex, Eu, Ap) must all have {npar} columns")
            return a
        cons = (_Constraint * max(len(self.constraints), 1))()
        for i, c in enumerate(self.constraints):
            cons[i] = _Constraint(c["rows"], cm(c["Au"]), cm(c["Ax"]), cm(c["Ar"]), cm(c["Ad"]), cm(c["Aup"]),
                                  cm(c["ub"]), cm(c["lb"]), c["ks"].ctypes.data, len(c["ks"]), c["soft"], c["prio"],
                                  cm(wide(c["Ap"], c["rows"])))
        md = _Model(self.nx, self.nu, self.ny, self.nd, self.Np, self.Nc, cm(self.F), cm(self.G), cm(self.C),
                    cm(self.Gd), cm(self.Dd), cm(self.Q), cm(self.R), cm(self.Rr), cm(self.Qf),
                    cm(self.umin), cm(self.umax), C.cast(cons, C.c_void_p), len(self.constraints),
                    int(self.reference_tracking), int(self.preprocess_mpqp), None, None,
                    int(self.reference_preview), int(self.disturbance_preview), cm(self.K),
                    npar, cm(wide(self.Ex, self.nx)), cm(self.ex), cm(wide(self.Eu, self.nu)), cm(self.eu),
                    int(self.parameter_preview), cm(self.f_offset), cm(self.h_offset), cm(self.xo), cm(self.uo),
                    int(self.reference_condensation))
        if self.move_blocks:
            mbf = np.ascontiguousarray([v for b in self.move_blocks for v in b], np.int32)
            mbn = np.ascontiguousarray([len(b) for b in self.move_blocks], np.int32)
            keep += [mbf, mbn]
            md.move_blocks, md.n_move_blocks = mbf.ctypes.data, mbn.ctypes.data
        q = C.c_void_p()
        rc = L.lmpc_gpu_condense(C.byref(md), C.byref(q))
        if rc != 0:
            raise LmpcError(rc, L.lmpc_gpu_condense_error().decode())
        try:
            d = np.zeros(9, np.int32)
            L.lmpc_gpu_mpqp_dims(q, d.ctypes.data)
            n, m, ms, nth, nu, nx, nr, nd, nup = [int(v) for v in d]
            out = dict(H=np.zeros((n, n), order="F"), f=np.zeros(n), f_theta=np.zeros((n, nth), order="F"),
                       A=np.zeros((m - ms, n), order="F"), bu=np.zeros(m), bl=np.zeros(m),
                       W=np.zeros((m, nth), order="F"), senses=np.zeros(m, np.int32))
            L.lmpc_gpu_mpqp_get(q, *[out[k].ctypes.data for k in ("H", "f", "f_theta", "A", "bu", "bl", "W", "senses")])
            out["dims"] = (nx, nr, nd, nup)
            L.lmpc_gpu_mpqp_traj2setpoint.argtypes = [C.c_void_p] * 4
            tr, tc = C.c_int32(0), C.c_int32(0)
            L.lmpc_gpu_mpqp_traj2setpoint(q, None, C.byref(tr), C.byref(tc))
            self.traj2setpoint = None
            if tr.value > 0:
                self.traj2setpoint = np.zeros((tr.value, tc.value), order="F")
                L.lmpc_gpu_mpqp_traj2setpoint(q, self.traj2setpoint.ctypes.data, None, None)
        finally:
            L.lmpc_gpu_mpqp_free(q)
        self.mpqp = out
        return out

    def setup(self, device=0, settings=None, devices=None):
        """setup!(mpc) + setup_gpu!(mpc): condense, then Cholesky / normalisation / upload
        (`devices=[...]`: one replica per listed GPU, host batches split by contiguous ranges)."""
        q = self.condense()
        args = (q["H"], q["f"], q["f_theta"], q["A"], q["bu"], q["bl"], q["W"], q["senses"])
        kw = dict(nu=self.nu, nx=self.nx, soft_weight=self.soft_weight, settings=settings, dims=q["dims"], K=self.K)
        self.gpu = (MultiGpuMPC(*args, devices=devices, **kw) if devices is not None
                    else GpuMPC(*args, device=device, **kw))
        return self.gpu

    def _g(self):
        return self.gpu if self.gpu is not None else self.setup()

    def compute_control(self, x, r=None, d=None, uprev=None, p=None, check=True):
        if self.reference_preview and self.reference_tracking:
            r = self._flatten_preview(r, self.ny)
            if self.reference_condensation:      # condense_reference (utils.jl:140-147)
                self._g()
                r = self.traj2setpoint @ r
        if self.disturbance_preview and self.nd > 0:
            d = self._flatten_preview(d, self.nd)
        nb = self.affine_parameter_dim()
        if nb > 0 and p is not None:           # format_affine_parameters, utils.jl:225-259
            p = np.asarray(p, np.float64)
            if self.parameter_preview and not (p.ndim == 1 and p.size == nb * self.Np):
                p = self._flatten_preview(p, nb)
            elif not self.parameter_preview and p.ndim == 2 and np.ndim(x) == 1:
                p = p[:, 0]
        return self._g().compute_control(x, r=r, d=d, uprev=uprev, p=p, check=check)

    def _flatten_preview(self, v, rows):
        """format_reference / format_disturbance with a preview (utils.jl:78-134, 155-205): a vector is
        held over the horizon, a rows x cols matrix is cut to Np columns or extended by holding its
        last column, then flattened column by column; rows x cols x N gives one preview per instance."""
        Np = self.Np
        if v is None:
            return np.zeros(rows * Np)
        v = np.asarray(v, np.float64)
        if v.ndim == 1:
            if v.size != rows:
                raise ValueError(f"Reference vector length ({v.size}) must match ({rows})")
            return np.tile(v, Np)
        if v.shape[0] != rows:
            raise ValueError(f"preview matrix must have {rows} rows")
        if v.shape[1] < Np:
            v = np.concatenate([v, np.repeat(v[:, -1:], Np - v.shape[1], axis=1)], axis=1)
        v = v[:, :Np]
        return v.reshape(rows * Np, -1, order="F") if v.ndim == 3 else v.reshape(-1, order="F")

    def solve(self, Theta, **kw):
        return self._g().solve(Theta, **kw)

    def clipping_bounds(self):
        """[umin umax] (nu x 2) when every input has finite simple bounds and there is no prestabilising
        feedback (then the first nu decision variables ARE the controls and clipping is exact); else None."""
        if self.K is not None and np.any(self.K != 0):
            return None
        if self.umin is None or self.umax is None:
            return None
        lo, hi = np.asarray(self.umin, float).ravel(), np.asarray(self.umax, float).ravel()
        if len(lo) != self.nu or len(hi) != self.nu or np.any(np.abs(lo) > 1e20) or np.any(np.abs(hi) > 1e20):
            return None
        return np.stack([lo, hi], axis=1)

    def explicit(self, lo, hi, n_samples=2_000_000, seed=0, device=0, clipping=True, gpu=None, **kw):
        """ExplicitMPC(mpc; range, build_tree=true) (src/explicit.jl:25-52): explicit solution over the
        parameter box [lo, hi], critical regions discovered by batched implicit solves on the GPU
        (linearmpc.jl_b200/explicit.py); soft constraints included (make_singlesided's slack columns).
        clipping=True (the reference's default, src/explicit.jl:65): the tree only has to be right after the
        control is clipped to the input bounds, which absorbs the saturated regions; an array gives the bounds
        explicitly, None / False builds an exact tree.  Returns (GpuExplicitMPC, info)."""
        from . import explicit as ex
        q = self.condense()
        clip = self.clipping_bounds() if clipping is True else (None if clipping in (None, False) else
                                                                np.asarray(clipping, float).reshape(self.nu, 2))
        sol = ex.build_explicit(q, gpu if gpu is not None else self._g(), lo, hi, self.nu, n_samples=n_samples,
                                seed=seed, clipping=clip, soft_weight=self.soft_weight, **kw)
        if self.K is not None and np.any(self.K != 0):
            # the laws give v; u = v - K x (build_tree! correction, src/explicit.jl:65-71)
            sol["feedbacks"][:, :, :self.nx] -= self.K
        empc = GpuExplicitMPC(sol["halfplanes"], sol["feedbacks"], sol["hp_list"], sol["jump_list"],
                              clipping=clip, dims=q["dims"], device=device)
        return empc, sol

    def simulate(self, x0, r=None, d=None, N=100, warm_start=True, want_traj=False, r_traj=None, d_traj=None,
                 plant=None, observer=None, noise=None, dynamics=None, measurement=None):
        """Simulation(mpc; x0, N, r, d) (simulation.jl:37-126).  With `dynamics` (a user-defined true plant, see
        GpuMPC.simulate_stepped) or an observer set by set_state_observer / set_offset_free_observer the loop is
        stepped from the host; otherwise it runs fused on chip."""
        obs = observer if observer is not None else self.observer
        affine_plant = plant is not None and plant.get("f_offset") is not None
        if dynamics is not None or affine_plant or (obs is not None and "F" in obs):
            if self.reference_preview or self.disturbance_preview:
                raise ValueError("the stepped closed loop does not slide preview windows")
            return self._g().simulate_stepped(self.F, self.G, x0, r=r, d=d, N=N, Gd=self.Gd, warm_start=warm_start,
                                              want_traj=want_traj, r_traj=r_traj, d_traj=d_traj, dynamics=dynamics,
                                              plant=plant, measurement=measurement, observer=obs, noise=noise)
        rp = self.Np if (self.reference_preview and self.reference_tracking and r_traj is not None) else 0
        dp = self.Np if (self.disturbance_preview and self.nd > 0 and d_traj is not None) else 0
        if self.reference_preview and self.reference_tracking and not rp:
            r = self._flatten_preview(r, self.ny)
        if self.disturbance_preview and self.nd > 0 and not dp:
            d = self._flatten_preview(d, self.nd)
        return self._g().simulate(self.F, self.G, x0, r=r, d=d, N=N, Gd=self.Gd, warm_start=warm_start,
                                  want_traj=want_traj, r_traj=r_traj, d_traj=d_traj, plant=plant,
                                  observer=observer, noise=noise, r_preview=rp, d_preview=dp)
