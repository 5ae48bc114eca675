#!/usr/bin/env python
"""bench.py -- batched condensed-MPC QP solves/s on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic inputs.  The headline workload at every
N is BASELINE.json configs[1]: the inverted-pendulum controller (README.md:36-53) on 10^7 uniformly
sampled theta = (x, r, u_prev) PER GPU (weak scaling: contiguous theta ranges, one process per GPU, no
data-path collective; the final result gather is measured separately, fused into the kernel's stores).

At N = 1 the same run also measures the other four BASELINE configs and prints them in the line's
`secondary` array, each with its own parity gate, roofline, cpu_baseline and e2e:

  invpend_1e4            configs[0]  10^4 sampled states, the reference's own CPU-runnable case
  aircraft_cl_1e6x40     configs[2]  AFTI-16 closed loop, 10^6 initial states x 40 steps (warm start)
  chain_1e6              configs[3]  chain of masses n = 180: 10^6 instances as 10 steps of 10^5
  explicit_invpend_1e8   configs[4]  explicit-MPC point location + affine law over 10^8 theta

  python bench.py --gpus 1 --steps 100 --warmup 10
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the CPU oracle port on the host cores, SAME inputs
  python bench.py --single-process --gpus N # one process, N GPUs through lmpc_gpu_create_multi (host buffers)

Inputs are generated on the HOST with one RNG (numpy PCG64, seed [1234, rank, chunk]) and uploaded, so the
reference arm solves exactly the leading part of what the GPU arm solves.  Prints ONE JSON line on rank 0.
`value` is device-resident throughput (CUDA events on the launching stream, max over ranks); `e2e` goes
through the public C-ABI call with pinned HOST buffers, H2D and D2H inside the timed region.
"""
import argparse
import concurrent.futures as cf
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 1234  # /root/reference/test/runtests.jl:8
CHUNK = 1_000_000
METRIC = "batched MPC QP solves/sec"
SETTINGS = {"primal_tol": 1e-6, "dual_tol": 1e-12, "zero_tol": 1e-11, "progress_tol": 1e-6, "cycle_tol": 10,
            "iter_limit": 10000}
KERNEL_NAMES = {1: "thread-per-QP", 2: "warp-per-QP", 3: "CTA-per-QP"}

# name: kind, controller, theta-range scale, units per GPU per step, default (steps, warmup), cpu sample, extras
WORKLOADS = {
    "invpend_1e7": dict(kind="solve", ctrl="invpend", scale=1.0, n=10_000_000, steps=100, warmup=10, cpu_n=10_000_000),
    "invpend_1e4": dict(kind="solve", ctrl="invpend", scale=1.0, n=10_000, steps=50, warmup=5, cpu_n=10_000,
                        flush_l2=True),
    "aircraft_cl_1e6x40": dict(kind="sim", ctrl="aircraft", scale=0.25, n=1_000_000, T=40, steps=5, warmup=3,
                               cpu_n=200_000),
    "chain_1e6": dict(kind="solve", ctrl="chain", scale=1.0, n=100_000, steps=10, warmup=3, cpu_n=20_000),
    "explicit_invpend_1e8": dict(kind="explicit", ctrl="invpend", scale=1.0, n=100_000_000, steps=10, warmup=3,
                                 cpu_n=20_000_000),
}
SECONDARY = ["invpend_1e4", "aircraft_cl_1e6x40", "chain_1e6", "explicit_invpend_1e8"]
REF_SAMPLE = {"invpend_1e7": 2_000_000, "invpend_1e4": 10_000, "aircraft_cl_1e6x40": 100_000, "chain_1e6": 8_000,
              "explicit_invpend_1e8": 20_000_000}


# ------------------------------------------------------------------------------------------------
# inputs: one host RNG for both arms
# ------------------------------------------------------------------------------------------------
def fill_uniform(out, lo, hi, rank, first_chunk=0):
    """out (n x len(lo), C-contiguous == len(lo) x n column-major) <- U(lo, hi); chunk c of CHUNK rows comes
    from PCG64([SEED, rank, c]), so any leading part of the batch can be regenerated on its own."""
    n = out.shape[0]
    span = (hi - lo)[None, :]
    def one(c):
        a, b = c * CHUNK, min(n, (c + 1) * CHUNK)
        rng = np.random.Generator(np.random.PCG64([SEED, rank, first_chunk + c]))
        v = out[a:b]
        rng.random(v.shape, out=v)
        v *= span
        v += lo[None, :]
    nch = (n + CHUNK - 1) // CHUNK
    with cf.ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        list(ex.map(one, range(nch)))
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# checker side (oracle/): parity gates, CPU baselines, the reference arm.  Never on the product path.
# ------------------------------------------------------------------------------------------------
def checker_controller(ctrl):
    from oracle import mpqp
    mpc = {"invpend": mpqp.invpend, "aircraft": mpqp.aircraft, "chain": mpqp.chain_of_masses}[ctrl]()
    return mpc, mpqp.mpc2mpqp(mpc)


def oracle_problem(mpc, qp, cst=None):
    """cst: the PRODUCT's own setup constants (so that a difference seen afterwards is the online path's);
    None: the oracle's independent setup (the reference arm)."""
    from oracle import mpqp, ldp
    L = mpqp.qp2ldp(qp, mpc.nu, soft_weight=mpc.soft_weight)
    kw = {}
    if cst is not None:
        L.Mbar, L.Dth, L.du, L.dl = cst["Mbar"], cst["Dth"], cst["du"], cst["dl"]
        Rf = np.zeros_like(L.Rinv); Rf[:mpc.nu] = cst["Rinv"]; L.Rinv = Rf
        L.Uth_offset, L.u_offset = cst["Uth"], cst["uoff"]
        L.senses = np.asarray(cst["senses"], np.int32)
        kw["cap"] = int(cst["cap"])
    P = ldp.OracleProblem(L, mpc.nu, **kw)
    if cst is not None:
        P.Gram[:] = cst["Gram"]
    return P


def explicit_tree():
    """the committed search tree of the invpend controller over the example's range (built by the product's
    explicit builder; tests/golden/make_explicit.py)"""
    z = np.load(os.path.join(ROOT, "tests", "golden", "invpend_explicit.npz"))
    return z


def count_degenerate(P, Theta, ref, tol=1e-6):
    """Operational non-degeneracy: instances whose optimum has an inactive row within 10*tol of a
    bound or an active multiplier below 10*tol -- where two correct solvers may legitimately report
    different active sets.  Re-solved in numpy from the oracle's active set (single-word masks only)."""
    m = P.m
    if m > 64:
        return None
    n_deg = 0
    au = (ref["act_u"][:, 0][:, None] >> np.arange(m, dtype=np.uint64)) & np.uint64(1)
    al = (ref["act_l"][:, 0][:, None] >> np.arange(m, dtype=np.uint64)) & np.uint64(1)
    sh = P.Dth @ Theta
    for q in range(Theta.shape[1]):
        act = np.where((au[q] | al[q]) != 0)[0]
        du, dl = P.du + sh[:, q], P.dl + sh[:, q]
        if len(act):
            Mw = P.Mbar[act]
            dw = np.where(au[q][act] != 0, du[act], dl[act])
            lam = np.linalg.solve(Mw @ Mw.T, -dw)
            v = -Mw.T @ lam
            margin = np.min(np.abs(lam))
        else:
            v = np.zeros(P.n); margin = np.inf
        inact = np.setdiff1d(np.arange(m), act)
        if len(inact):
            mv = P.Mbar[inact] @ v
            margin = min(margin, np.min(du[inact] - mv), np.min(mv - dl[inact]))
        n_deg += int(margin < 10 * tol)
    return n_deg


def sim_args(mpc):
    nx, nr, nd, nup = mpc.parameter_dims()
    return nx, nr, nd, nup


def cpu_run(w, P, mpc, inputs, nthreads, want_flops=False, dyn=None):
    """one pass of the oracle port over `inputs` (host arrays, instances in rows); returns (units, result).
    dyn: the plant matrices (F, G, Gd) the PRODUCT was given (parity gate: same plant on both sides; the
    reference arm uses the checker's own)."""
    from oracle import ldp  # noqa: F401
    if w["kind"] == "solve":
        out = P.solve_batch(inputs.T, nthreads=nthreads, want_flops=want_flops)
        return inputs.shape[0], out
    if w["kind"] == "sim":
        nx, nr, nd, nup = sim_args(mpc)
        n = inputs.shape[0]
        R = np.repeat(np.array([[0.1], [0.0]]), n, axis=1)          # the example's scenario, mpc_examples.jl:206
        D = np.zeros((nd, n))
        F, G, Gd = dyn if dyn is not None else (mpc.F, mpc.G, mpc.Gd if nd else None)
        out = P.simulate_batch(F, G, Gd if nd else None, inputs.T, R, D if nd else None, None, w["T"],
                               nr, nd, nup, warm=True, nthreads=nthreads, want_flops=want_flops)
        return n * w["T"], out
    U, reg = P.evaluate_batch(inputs.T, nthreads=nthreads)
    return inputs.shape[0], dict(U=U, region=reg)


def run_reference_one(name, steps, warmup):
    """the reference's CPU path for one workload.  The real one (Julia compute_control -> DAQPBase -> libdaqp,
    ParametricDAQP.evaluate) cannot run here (no Julia, libdaqp not under /root/reference), so this times the
    oracle port of it, independent setup, all host cores; each step is a bounded sample: the LEADING rows of
    rank 0's batch, the very numbers the GPU arm solves."""
    from oracle import ldp
    w = WORKLOADS[name]
    cores = ldp.num_threads()
    n = min(w["n"], REF_SAMPLE[name])
    if w["kind"] == "explicit":
        z = explicit_tree()
        P, mpc, lo, hi = ldp.OracleBst(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clip=z["clipping"]), None, z["lo"], z["hi"]
    else:
        mpc, qp = checker_controller(w["ctrl"])
        P = oracle_problem(mpc, qp)
        lo, hi = theta_range(w["ctrl"], w["scale"])
        if w["kind"] == "sim":
            lo, hi = lo[:mpc.nx], hi[:mpc.nx]
    X = fill_uniform(np.empty((n, len(lo))), lo, hi, 0)
    times, units = [], 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        units, _ = cpu_run(w, P, mpc, X, cores)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    val = units * len(times) / total
    return {"impl": "reference", "metric": METRIC, "value": val, "unit": "solves/s", "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "controller": w["ctrl"], "theta_scale": w["scale"],
                       "instances_per_step": n, "horizon_steps": w.get("T"),
                       "inputs": "leading rows of rank 0's batch: numpy PCG64 seed [1234, 0, chunk], as the GPU arm",
                       "note": "bounded sample of the workload"},
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": f"{n} instances per step, oracle/ldp_oracle.c cold start (closed loop: warm), "
                                       f"{cores} pthreads; the real reference (Julia + libdaqp) cannot run here"},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def theta_range(ctrl, scale):
    from oracle import mpqp
    return {"invpend": mpqp.invpend_range, "aircraft": mpqp.aircraft_range,
            "chain": lambda s: mpqp.chain_range(10, s)}[ctrl](scale)


def run_reference(args, rank):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    line = run_reference_one(args.workload, args.steps if args.steps else max(3, min(w["steps"], 10)),
                             args.warmup if args.warmup is not None else 1)
    line["n_gpus"] = args.gpus
    if args.secondary and args.workload == "invpend_1e7":
        line["secondary"] = [run_reference_one(nm, 3, 1) for nm in SECONDARY]
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run this rank (and first-touch its pinned staging buffers) on the CPUs next to its
    GPU, as NVML reports them.  Returns the number of CPUs bound to, or None when unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
# product side: one object per workload
# ------------------------------------------------------------------------------------------------
class Bench:
    """Builds the controller with the PRODUCT's own builder (lmpc_b200.mpc_examples -> C++ condensing behind
    the C ABI -> GPU upload), holds the inputs in pinned host memory and in HBM, and exposes the device-resident
    step, the end-to-end step through the host-pointer ABI call, the parity gate and the CPU leg."""

    def __init__(self, name, dev_index, rank, torch):
        import lmpc_b200
        self.lm, self.torch = lmpc_b200, torch
        self.name, self.w, self.rank = name, WORKLOADS[name], rank
        self.dev = torch.device("cuda", dev_index)
        w = self.w
        self.n = w["n"]
        if w["kind"] == "explicit":
            z = explicit_tree()
            self.g = lmpc_b200.GpuExplicitMPC(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clipping=z["clipping"],
                                              dims=(4, 2, 0, 1), device=dev_index)
            self.tree = z
            lo, hi = z["lo"], z["hi"]
            self.nin, self.nu = self.g.nth, self.g.nu
        else:
            self.model, rng, self.scen = lmpc_b200.mpc_examples(w["ctrl"])
            self.g = self.model.setup(device=dev_index)
            lo, hi = np.array(rng[0], float), np.array(rng[1], float)
            nx = self.model.nx
            # ParameterRange scaled by s, except the previous-control block (its range is the input bounds)
            nup = self.g.dims[3]
            sc = np.full(len(lo), w["scale"])
            if nup:
                sc[-nup:] = 1.0
            lo, hi = lo * sc, hi * sc
            if w["kind"] == "sim":
                lo, hi = lo[:nx], hi[:nx]
            self.nin, self.nu = len(lo), self.g.nu
        self.lo, self.hi = lo, hi
        # inputs: pinned host (one RNG for both arms) -> HBM
        self.X_h = torch.empty((self.n, self.nin), dtype=torch.float64, pin_memory=True)
        fill_uniform(self.X_h.numpy(), lo, hi, rank)
        self.X = self.X_h.to(self.dev, non_blocking=True)
        self.U = torch.empty((self.n, self.nu), dtype=torch.float64, device=self.dev)
        self.flag = torch.empty(self.n, dtype=torch.int32, device=self.dev)
        self.U_h = torch.empty((self.n, self.nu), dtype=torch.float64, pin_memory=True)
        self.flag_h = torch.empty(self.n, dtype=torch.int32, pin_memory=True)
        if w["kind"] == "sim":
            nx, nr, nd, nup = self.g.dims
            T = w["T"]
            self.Fm, self.Gm = np.asfortranarray(self.model.F), np.asfortranarray(self.model.G)
            self.Gdm = np.asfortranarray(self.model.Gd) if nd else None
            r = np.asarray(self.scen["r"], float)
            self.R_h = torch.as_tensor(np.repeat(r[None, :], self.n, axis=0).copy()).pin_memory()
            self.D_h = torch.zeros((self.n, nd), dtype=torch.float64).pin_memory() if nd else None
            self.R, self.D = self.R_h.to(self.dev), (self.D_h.to(self.dev) if nd else None)
            self.Xend = torch.empty((self.n, nx), dtype=torch.float64, device=self.dev)
            self.iters = torch.empty(self.n, dtype=torch.int32, device=self.dev)
            self.Xend_h = torch.empty((self.n, nx), dtype=torch.float64, pin_memory=True)
            self.units = self.n * T
            self.h2d = 8 * (nx + nr + nd) * self.n
            self.d2h = (8 * nx + 8 * self.nu + 4) * self.n
        else:
            self.units = self.n
            self.h2d = 8 * self.nin * self.n
            self.d2h = (8 * self.nu + (4 if w["kind"] == "solve" else 0)) * self.n
        torch.cuda.synchronize(self.dev)

    # -- device-resident step -----------------------------------------------------------------
    def step(self, stream):
        L, g, w = self.lm.lib(), self.g, self.w
        if w["kind"] == "solve":
            g.solve_async(self.X, self.U, self.flag, N=self.n, stream=stream)
        elif w["kind"] == "explicit":
            g.evaluate_async(self.X, self.U, N=self.n, stream=stream)
        else:
            nx, nr, nd, nup = g.dims
            rc = L.lmpc_gpu_simulate_batch(g._h, self.Fm.ctypes.data, self.Gm.ctypes.data,
                                           self.Gdm.ctypes.data if nd else None, nx, nr, nd, nup,
                                           self.X.data_ptr(), self.R.data_ptr(), self.D.data_ptr() if nd else None, None,
                                           self.n, w["T"], 1, self.Xend.data_ptr(), self.U.data_ptr(), None,
                                           self.flag.data_ptr(), self.iters.data_ptr())
            assert rc == 0, L.lmpc_gpu_last_error()

    def launch_stream(self, user_stream):
        """the stream the step's kernel runs on: the caller's for the async entry points, the handle's own for
        the synchronous closed-loop call (device pointers: no copies, one launch)"""
        if self.w["kind"] == "sim":
            return None
        return user_stream

    # -- end to end through the host-pointer ABI call -----------------------------------------
    def e2e_step(self):
        L, g, w = self.lm.lib(), self.g, self.w
        if w["kind"] == "solve":
            rc = L.lmpc_gpu_solve_batch(g._h, self.X_h.data_ptr(), self.n, self.U_h.data_ptr(), self.flag_h.data_ptr(),
                                        None, None, None)
        elif w["kind"] == "explicit":
            rc = L.lmpc_gpu_evaluate_batch(g._t, self.X_h.data_ptr(), self.n, self.U_h.data_ptr(), None)
        else:
            nx, nr, nd, nup = g.dims
            rc = L.lmpc_gpu_simulate_batch(g._h, self.Fm.ctypes.data, self.Gm.ctypes.data,
                                           self.Gdm.ctypes.data if nd else None, nx, nr, nd, nup,
                                           self.X_h.data_ptr(), self.R_h.data_ptr(),
                                           self.D_h.data_ptr() if nd else None, None, self.n, w["T"], 1,
                                           self.Xend_h.data_ptr(), self.U_h.data_ptr(), None,
                                           self.flag_h.data_ptr(), None)
        assert rc == 0, L.lmpc_gpu_last_error()

    # -- parity gate + CPU leg (rank 0) ---------------------------------------------------------
    def checker(self):
        from oracle import ldp
        w = self.w
        if w["kind"] == "explicit":
            z = self.tree
            return ldp.OracleBst(z["halfplanes"], z["feedbacks"], z["hp_list"], z["jump_list"], clip=z["clipping"]), None
        mpc, qp = checker_controller(w["ctrl"])
        return oracle_problem(mpc, qp, self.g.ldp_constants()), mpc

    def parity(self, P, mpc, n_gate):
        """the product's results of the timed step (already in self.U / self.flag) against the oracle on the
        leading n_gate instances"""
        w, torch = self.w, self.torch
        sub = self.X_h.numpy()[:n_gate]
        dyn = (self.model.F, self.model.G, self.model.Gd) if w["kind"] == "sim" else None
        _, ref = cpu_run(w, P, mpc, sub, 0, want_flops=(w["kind"] != "explicit"), dyn=dyn)
        U = self.U[:n_gate].cpu().numpy().T
        out = {"n": int(n_gate)}
        if w["kind"] == "explicit":
            out["bit_identical_u"] = bool(np.array_equal(U, ref["U"]))
            out["max_abs_err_u"] = float(np.max(np.abs(U - ref["U"])))
            return out, None
        fl = self.flag[:n_gate].cpu().numpy()
        ok = ref["flag"] >= 1
        out["identical_flags"] = int(np.count_nonzero(fl == ref["flag"]))
        rel = np.abs(U - (ref["U"] if w["kind"] == "solve" else ref["Ulast"])) / \
            np.maximum(1.0, np.abs(ref["U"] if w["kind"] == "solve" else ref["Ulast"]))
        out["max_rel_err_u"] = float(rel[:, ok].max()) if ok.any() else 0.0
        out["bit_identical_u"] = bool(np.array_equal(U[:, ok], (ref["U"] if w["kind"] == "solve" else ref["Ulast"])[:, ok]))
        if w["kind"] == "sim":
            out["bit_identical_xend"] = bool(np.array_equal(self.Xend[:n_gate].cpu().numpy().T, ref["Xend"]))
            out["identical_iteration_totals"] = bool(np.array_equal(self.iters[:n_gate].cpu().numpy(), ref["iters"]))
        else:
            # the same launches as the timed step (box problems: fast kernel + REDO launch at any batch size)
            os.environ["LMPC_BOX_FAST_MIN_BATCH"] = "0" if self.n >= 200000 else str(1 << 62)
            got = self.g.solve(np.asfortranarray(sub.T))
            del os.environ["LMPC_BOX_FAST_MIN_BATCH"]
            same = (got["exitflag"] == ref["flag"]) & np.all(got["active_upper"] == ref["act_u"], axis=1) & \
                np.all(got["active_lower"] == ref["act_l"], axis=1) & (got["iters"] == ref["iters"])
            out["identical_flag_iters_and_active_set"] = int(same.sum())
            if n_gate <= 10000:
                out["degenerate"] = count_degenerate(P, sub.T, ref)
        return out, float(ref["flops"].mean())


def device_timing(b, steps, warmup, world, dist, torch, flush=False):
    """W warm-up steps, then K steps timed with CUDA events on the launching stream, barrier + synchronize on
    both sides, max over ranks.  flush: rewrite a 512 MB buffer between timed steps (inputs smaller than L2)
    and sum the per-step times instead."""
    dev = b.dev
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    use = b.launch_stream(stream.cuda_stream)
    ev_stream = stream if use is not None else torch.cuda.ExternalStream(b.lm.lib().lmpc_gpu_stream(b.g._h), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    scrub = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device=dev) if flush else None
    for _ in range(warmup):
        b.step(use)
    barrier()
    l0 = b.lm.launch_count()
    per = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(ev_stream)
    for a, e in per:
        if flush:
            with torch.cuda.stream(ev_stream):
                scrub.fill_(1.0)
        a.record(ev_stream); b.step(use); e.record(ev_stream)
    ev1.record(ev_stream)
    barrier()
    launches = b.lm.launch_count() - l0
    step_ms = [a.elapsed_time(e) for a, e in per]
    total = float(sum(step_ms)) if flush else ev0.elapsed_time(ev1)
    if world > 1:
        total = b.lm.max_over_ranks(total, device=dev)
    return total, float(np.mean(step_ms)), int(launches)


def e2e_timing(b, steps, world, dist, torch):
    for _ in range(2):
        b.e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        b.e2e_step()
    torch.cuda.synchronize()
    s = time.perf_counter() - t0
    if world > 1:
        s = b.lm.max_over_ranks(s, device=b.dev)
    return s


def measure(name, dev_index, rank, world, dist, torch, steps, warmup, cpu_baseline=True, numa=None):
    """everything for one workload; returns the JSON dict on rank 0 (None elsewhere)"""
    import lmpc_b200
    w = WORKLOADS[name]
    b = Bench(name, dev_index, rank, torch)
    total_ms, kernel_ms, launches = device_timing(b, steps, warmup, world, dist, torch, flush=w.get("flush_l2", False))
    ok_all = bool((b.flag >= 1).all().item()) if w["kind"] != "explicit" else None
    parity = cpu = flops = None
    if rank == 0:
        P, mpc = b.checker()
        n_gate = min(b.n, 10000 if w["kind"] != "solve" or w["ctrl"] != "chain" else 2000)
        parity, flops = b.parity(P, mpc, n_gate)
        if cpu_baseline and world == 1:
            from oracle import ldp
            cores = ldp.num_threads()
            ns = min(b.n, w["cpu_n"])
            sub = b.X_h.numpy()[:ns]
            dyn = (b.model.F, b.model.G, b.model.Gd) if w["kind"] == "sim" else None
            cpu_run(w, P, mpc, sub[:max(1, ns // 50)], cores, dyn=dyn)          # warm the threads
            t0 = time.perf_counter()
            units, ref = cpu_run(w, P, mpc, sub, cores, want_flops=(w["kind"] != "explicit"), dyn=dyn)
            dt = time.perf_counter() - t0
            if w["kind"] != "explicit":
                flops = float(ref["flops"].mean())
            cpu = {"value": units / dt, "unit": "solves/s", "cores": cores, "kind": "port",
                   "sample": f"the leading {ns} instances of rank 0's batch ({units} solves), oracle/ldp_oracle.c on "
                             f"{cores} pthreads (the real reference, Julia + libdaqp, cannot run here)"}
    e2e_steps = max(3, min(steps, 10))
    e2e_s = e2e_timing(b, e2e_steps, world, dist, torch)
    e2e_ok = bool(torch.equal(b.U_h, b.U.cpu()))
    if rank != 0:
        return None, b
    value = float(b.units) * world * steps / (total_ms * 1e-3)
    hbm_peak, peak_src = peaks()
    g = b.g
    cfg = {"workload": name, "controller": w["ctrl"], "theta_scale": w["scale"], "seed": SEED,
           "instances_per_gpu_per_step": b.n, "settings": SETTINGS,
           "inputs": "host numpy PCG64 seed [1234, rank, chunk], uploaded; the reference arm solves the leading rows",
           "l2": ("512 MB scrub buffer rewritten between timed steps; per-step event times summed"
                  if w.get("flush_l2") else "inputs larger than L2 (%.0f MB per step vs 126 MB L2)" % (b.h2d / 1e6))}
    if w["kind"] == "explicit":
        nodes = int(len(b.tree["hp_list"]))
        bytes_per = 8 * (b.nin + b.nu)
        ach = bytes_per * b.n / (kernel_ms * 1e-3) / 1e9
        cfg.update({"nth": b.nin, "tree_nodes": nodes})
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic_of(name), "kernel": {2: "bst_eval_ring_kernel<%d,%d> (fp32-screened lane-refill ring)", 1: "bst_eval_kernel<%d> (fp32-screened, nu %d)",
                           0: "bst_eval_kernel<%d> (fp64 rows, nu %d)"}.get(b.g.kernel, "bst_eval_kernel<%d,%d>") % (b.nin, b.nu),
                "kernel_ms": kernel_ms,
                "bytes_per_theta": bytes_per, "peak_source": peak_src}
    else:
        fp64_peak = lmpc_b200.fp64_peak_tflops(dev_index)
        per_unit = flops if w["kind"] == "solve" else flops / w["T"]
        bytes_per = 8 * g.nth + 8 * g.nu + 4
        ach_tf = per_unit * b.units / (kernel_ms * 1e-3) / 1e12
        cfg.update({"n": g.n, "m": g.m, "nth": g.nth, "kernel": KERNEL_NAMES[g.kernel],
                    "all_exitflags_at_least_1": ok_all})
        if w["kind"] == "sim":
            cfg.update({"horizon_steps": w["T"], "warm_start": True, "reference": [0.1, 0.0],
                        "mean_iterations_per_solve": float(b.iters.float().mean().item()) / w["T"]})
            bytes_per = (8 * (b.nin + g.dims[1] + g.dims[2]) + 8 * b.nin + 8 * g.nu + 8) / w["T"]
        ach_gbs = bytes_per * b.units / (kernel_ms * 1e-3) / 1e9
        roof = {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                "traffic": traffic_of(name), "kernel": kernel_name(g, w), "kernel_ms": kernel_ms,
                "flops_per_solve": per_unit,
                "peak_source": "FP64 FMA microbenchmark in this run (lmpc_gpu_fp64_peak)",
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                        "bytes_per_solve": bytes_per, "peak_source": peak_src}}
    line = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": float(b.units) * world * e2e_steps / e2e_s, "unit": "solves/s",
                    "h2d_bytes_per_step": b.h2d, "d2h_bytes_per_step": b.d2h, "steps": e2e_steps,
                    "timer": "wall clock around the synchronous host-pointer C-ABI call",
                    "matches_device_path": e2e_ok, "cpus_bound_per_rank": numa},
            "gpu_launches": launches, "parity": parity}
    return line, b


def kernel_name(g, w):
    """from what the handle selected (lmpc_gpu_get_info), not a literal"""
    box = g.kernel == 1 and w["kind"] == "solve" and not np.any(g.senses) and g.m == g.n and g.n <= 8
    # box problems: the add-only fast kernel, then the generic box kernel over the instances it handed over
    base = {1: "ldp_box_fast_kernel<%d,%d> + ldp_box_kernel<%d,%d> REDO pass" % (g.n, g.nu, g.n, g.nu) if box
            else "ldp_thread_kernel<%d>" % g.n,
            2: "ldp_warp_kernel", 3: "ldp_cta_kernel / ldp_big_kernel"}[g.kernel]
    if w["kind"] == "sim":
        base = base.replace("_kernel", "_sim_kernel")
    return base + " (" + KERNEL_NAMES[g.kernel] + ", capacity %d)" % g.cap


def traffic_of(name):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        return json.load(open(tp)).get(name)
    return None


# ------------------------------------------------------------------------------------------------
# the final result gather (N > 1): fused into the kernel's stores through a CUDA-IPC window on rank 0
# ------------------------------------------------------------------------------------------------
class _Cai:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def gather_leg(b, steps, warmup, rank, world, dist, torch):
    lm = b.lm
    ntot = b.n * world
    handle = [None]
    win = None
    if rank == 0:
        win = lm.ResultWindow(b.dev.index, nbytes=12 * ntot)
        handle[0] = win.handle_bytes()
    dist.broadcast_object_list(handle, src=0)
    if rank != 0:
        win = lm.ResultWindow(b.dev.index, handle=handle[0])
    off = rank * b.n
    stream = torch.cuda.Stream(device=b.dev)
    L = lm.lib()

    def step():
        rc = L.lmpc_gpu_solve_batch_async(b.g._h, b.X.data_ptr(), b.n, win.ptr + 8 * off, win.ptr + 8 * ntot + 4 * off,
                                          None, None, None, stream.cuda_stream)
        assert rc == 0
    for _ in range(warmup):
        step()
    dist.barrier(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    dist.barrier(); torch.cuda.synchronize()
    ms = lm.max_over_ranks(ev0.elapsed_time(ev1), device=b.dev)
    # check: every rank's slice in rank 0's window equals what that rank computed locally
    sums = [torch.zeros(2, dtype=torch.float64, device=b.dev) for _ in range(world)]
    mine = torch.stack([b.U.sum(), b.flag.double().sum()])
    dist.all_gather(sums, mine)
    ok = None
    if rank == 0:
        Uw = torch.as_tensor(_Cai(win.ptr, ntot, "<f8"), device=b.dev)
        Fw = torch.as_tensor(_Cai(win.ptr + 8 * ntot, ntot, "<i4"), device=b.dev)
        ok = bool(torch.equal(Uw[:b.n], b.U.view(-1)))
        for r in range(world):
            ok = ok and bool(Uw[r * b.n:(r + 1) * b.n].sum() == sums[r][0]) and \
                bool(Fw[r * b.n:(r + 1) * b.n].double().sum() == sums[r][1])
    # pipelined variant: the kernel stores locally (double-buffered) and a copy engine pushes step i's results into
    # rank 0's window on a second stream while step i + 1 computes.  Rank 0's NVLink ingress is the bound at large N:
    # (N - 1) x 12 B x n per step must enter one GPU (840 MB at 8 GPUs, >= 0.93 ms at 900 GB/s vs 0.62 ms of compute)
    comm = torch.cuda.Stream(device=b.dev)
    bufs = [(torch.empty_like(b.U), torch.empty_like(b.flag)) for _ in range(2)]
    wu = torch.as_tensor(_Cai(win.ptr + 8 * off, b.n * b.nu, "<f8"), device=b.dev)
    wf = torch.as_tensor(_Cai(win.ptr + 8 * ntot + 4 * off, b.n, "<i4"), device=b.dev)
    evk = [torch.cuda.Event() for _ in range(2)]
    evc = [torch.cuda.Event() for _ in range(2)]

    def pstep(i):
        j = i & 1
        if i >= 2:
            stream.wait_event(evc[j])          # buffer j has been pushed out
        rc = L.lmpc_gpu_solve_batch_async(b.g._h, b.X.data_ptr(), b.n, bufs[j][0].data_ptr(), bufs[j][1].data_ptr(),
                                          None, None, None, stream.cuda_stream)
        assert rc == 0
        evk[j].record(stream)
        comm.wait_event(evk[j])
        with torch.cuda.stream(comm):
            wu.copy_(bufs[j][0].view(-1), non_blocking=True)
            wf.copy_(bufs[j][1], non_blocking=True)
        evc[j].record(comm)
    for i in range(4):
        pstep(i)
    dist.barrier(); torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    p0.record(stream)
    for i in range(steps):
        pstep(i)
    stream.wait_event(evc[(steps - 1) & 1])
    p1.record(stream)
    dist.barrier(); torch.cuda.synchronize()
    pipe_ms = lm.max_over_ranks(p0.elapsed_time(p1), device=b.dev) / steps
    if rank == 0:
        for r in range(world):
            ok = ok and bool(Uw[r * b.n:(r + 1) * b.n].sum() == sums[r][0]) and \
                bool(Fw[r * b.n:(r + 1) * b.n].double().sum() == sums[r][1])
    # NCCL for comparison: preallocated all_gather_into_tensor of the unpadded slices, warmed, timed
    flat = torch.empty(ntot * b.nu, dtype=torch.float64, device=b.dev)
    for _ in range(3):
        dist.all_gather_into_tensor(flat, b.U.view(-1))
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_gather_into_tensor(flat, b.U.view(-1))
    e1.record(); torch.cuda.synchronize()
    nccl_ms = lm.max_over_ranks(e0.elapsed_time(e1) / 10, device=b.dev)
    dist.barrier()
    del wu, wf
    win.close()
    return ms / steps, ok, nccl_ms, pipe_ms


def run_single_process(args):
    """one process, N GPUs: lmpc_gpu_create_multi + lmpc_gpu_solve_batch_multi on pinned HOST buffers (contiguous
    theta ranges, one worker thread per device, results land in disjoint slices of one host array: the gather of
    SURVEY 8e done by the D2H copies themselves)."""
    import torch
    import lmpc_b200
    w = WORKLOADS[args.workload]
    assert w["kind"] == "solve"
    G = args.gpus
    model, rng, _ = lmpc_b200.mpc_examples(w["ctrl"])
    mg = model.setup(devices=list(range(G)))
    lo, hi = np.array(rng[0], float), np.array(rng[1], float)
    sc = np.full(len(lo), w["scale"]); nup = mg.dims[3]
    if nup:
        sc[-nup:] = 1.0
    n = w["n"] * G
    X = torch.empty((n, len(lo)), dtype=torch.float64).pin_memory()
    for r in range(G):
        fill_uniform(X.numpy()[r * w["n"]:(r + 1) * w["n"]], lo * sc, hi * sc, r)
    U = torch.empty((n, mg.nu), dtype=torch.float64).pin_memory()
    fl = torch.empty(n, dtype=torch.int32).pin_memory()
    L = lmpc_b200.lib()
    steps = args.steps or 10

    def step():
        rc = L.lmpc_gpu_solve_batch_multi(mg._mh, X.data_ptr(), n, U.data_ptr(), fl.data_ptr(), None, None, None)
        assert rc == 0, L.lmpc_gpu_last_error()
    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    val = n * steps / dt
    print(json.dumps({"arm": "single-process", "metric": METRIC, "value": val, "unit": "solves/s", "n_gpus": G,
                      "steps": steps, "warmup": 3, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": args.workload, "instances_per_gpu_per_step": w["n"],
                                 "api": "lmpc_gpu_create_multi + lmpc_gpu_solve_batch_multi, pinned host buffers"},
                      "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 8 * len(lo) * n,
                              "d2h_bytes_per_step": (8 * mg.nu + 4) * n},
                      "all_exitflags_at_least_1": bool((fl >= 1).all().item())}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="invpend_1e7", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false")
    ap.add_argument("--single-process", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.single_process:
        run_single_process(args)
        return
    w = WORKLOADS[args.workload]
    steps = args.steps if args.steps else w["steps"]
    warmup = max(args.warmup if args.warmup is not None else w["warmup"], 3)

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None

    import torch
    import torch.distributed as dist
    import lmpc_b200  # noqa: F401

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()      # runs through warm-up, the timed region and the e2e leg of the headline workload
    line, b = measure(args.workload, local_rank, rank, world, dist, torch, steps, warmup,
                      cpu_baseline=not args.no_cpu_baseline, numa=numa)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1 and w["kind"] == "solve":
        g_ms, g_ok, nccl_ms, p_ms = gather_leg(b, steps, warmup, rank, world, dist, torch)
        if rank == 0:
            best = min(g_ms, p_ms)
            ingress = 12.0 * b.n * (world - 1)            # bytes that must enter rank 0 per step
            line["gather"] = {
                "how": "results land in rank 0's CUDA-IPC window (lmpc_gpu_window_*), 12 B per solve; two ways, the faster "
                       "one counts: 'fused' = every rank's kernel stores u and the exit flag straight into the window "
                       "over NVLink (no second kernel); 'pipelined' = local stores, a copy engine pushes step i into "
                       "the window on a second stream while step i + 1 computes",
                "fused_ms_per_step": g_ms, "pipelined_ms_per_step": p_ms,
                "ms_per_step_with_gather": best, "gather_ms": max(0.0, best - line["ms_per_step"]),
                "value_with_gather": float(b.n) * world / (best * 1e-3),
                "efficiency_with_gather_vs_no_gather": line["ms_per_step"] / best,
                "rank0_ingress_bytes_per_step": ingress,
                "rank0_ingress_floor_ms_at_900GBps": ingress / 900e9 * 1e3,
                "window_matches_every_rank": g_ok,
                "nccl_all_gather_into_tensor_ms": nccl_ms,
                "nccl_note": "all_gather of U alone (8 B per solve), not overlapped with compute"}
            line["gather_ms"] = line["gather"]["gather_ms"]
    if rank == 0:
        line["clocks"] = clocks
    del b
    torch.cuda.empty_cache()
    if world == 1 and args.secondary and args.workload == "invpend_1e7":
        sec = []
        for nm in SECONDARY:
            ws = WORKLOADS[nm]
            s2 = ClockSampler(local_rank); s2.start()
            try:
                l2, b2 = measure(nm, local_rank, rank, world, dist, torch, ws["steps"], ws["warmup"],
                                 cpu_baseline=not args.no_cpu_baseline)
                l2["clocks"] = s2.stop()
                del b2
            except Exception as e:      # a failing secondary workload must not take the headline line down
                s2.stop()
                l2 = {"config": {"workload": nm}, "error": repr(e)}
            torch.cuda.empty_cache()
            sec.append(l2)
        line["secondary"] = sec
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
