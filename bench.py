#!/usr/bin/env python
"""bench.py -- batched condensed-MPC QP solves/s on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (theta -> affine update -> dual active-set solve -> u)
over one batch of synthetic theta.  Workload at every N: BASELINE.json configs[1], the
inverted-pendulum controller (README.md:36-53) on 10^7 uniformly sampled theta = (x, r, u_prev)
PER GPU (weak scaling: contiguous theta ranges, one process per GPU, no data-path collective;
only a final result gather).

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...   # the CPU oracle port on the host cores

Prints ONE JSON line on rank 0.  `value` is device-resident throughput (CUDA events on the
launching stream, max over ranks); `e2e` goes through the public C-ABI call with pinned HOST
buffers, H2D and D2H inside the timed region (wall clock around the synchronous call).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (controller, theta-range scale, instances per GPU)
    "invpend_1e7": ("invpend", 1.0, 10_000_000),
}
SEED = 1234  # /root/reference/test/runtests.jl:8


def product_model(name):
    """The controller built with the PRODUCT's own builder (host-side C++ condensing behind the C ABI
    + GPU upload): README.md:36-53 inverted pendulum, Ts = 0.01 zoh, Np = 50, Nc = 5."""
    import lmpc_b200
    from scipy.linalg import expm
    if name != "invpend":
        raise KeyError(name)
    A = np.array([[0, 1, 0, 0], [0, -10, 9.81, 0], [0, 0, 0, 1], [0, -20, 39.24, 0.0]])
    B = 100 * np.array([[0], [1.0], [0], [2.0]])
    E = expm(np.block([[A * 0.01, B * 0.01], [np.zeros((1, 5))]]))      # zoh, utils.jl:291-295
    mpc = lmpc_b200.MPC(E[:4, :4], E[:4, 4:], C=[[1.0, 0, 0, 0], [0, 0, 1.0, 0]], Np=50, Nc=5)
    mpc.set_objective(Q=[1.2 ** 2, 1], R=[0.0], Rr=[1.0])
    mpc.set_bounds(umin=[-2.0], umax=[2.0])
    return mpc


def product_controller(name, device):
    return product_model(name).setup(device=device)


def build_problem(name):
    """CHECKER side only: the same controller through the test-side restatement of mpc2mpqp
    (oracle/), used for the parity gate and the CPU baseline."""
    from oracle import mpqp
    if name == "invpend":
        mpc = mpqp.invpend()
    else:
        raise KeyError(name)
    return mpc, mpqp.mpc2mpqp(mpc)


def theta_range(name, scale):
    from oracle import mpqp
    return {"invpend": mpqp.invpend_range}[name](scale)


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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_problem(mpc, qp, cst=None):
    from oracle import mpqp, ldp
    L = mpqp.qp2ldp(qp, mpc.nu, soft_weight=mpc.soft_weight)
    if cst is not None:
        L.Mbar, L.Dth, L.du, L.dl = cst["Mbar"], cst["Dth"], cst["du"], cst["dl"]
        Rf = np.zeros_like(L.Rinv); Rf[:mpc.nu] = cst["Rinv"]; L.Rinv = Rf
        L.Uth_offset, L.u_offset = cst["Uth"], cst["uoff"]
    P = ldp.OracleProblem(L, mpc.nu)
    if cst is not None:
        P.Gram[:] = cst["Gram"]
    return P


def count_degenerate(P, Theta, ref, tol=1e-6):
    """Operational non-degeneracy: instances whose optimum has an inactive row within
    10*tol of a bound or an active multiplier below 10*tol -- where two correct solvers may
    legitimately report different active sets.  Re-solved in numpy from the oracle's active set."""
    m = P.m
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


def cpu_leg(P, Theta_host, passes=2):
    """The oracle port on all host cores over a bounded sample (about 10-30 core-seconds):
    returns (solves/s, cores, n, flops/solve, last result)."""
    from oracle import ldp
    cores = ldp.num_threads()
    P.solve_batch(Theta_host[:, :20000], nthreads=cores)          # warm the threads / caches
    t0 = time.perf_counter()
    for _ in range(passes - 1):
        P.solve_batch(Theta_host, nthreads=cores)
    out = P.solve_batch(Theta_host, nthreads=cores, want_flops=True)
    dt = time.perf_counter() - t0
    return passes * Theta_host.shape[1] / dt, cores, Theta_host.shape[1], float(out["flops"].mean()), out


def run_reference(args, rank):
    """--impl reference: the reference's CPU path.  The real one (Julia compute_control -> DAQPBase
    -> libdaqp) cannot run here (no Julia, libdaqp not under /root/reference), so this times the
    oracle port of it on all host cores; each step is a bounded 2e6-theta sample of the workload."""
    if rank != 0:
        return
    ctrl, scale, n_per = WORKLOADS[args.workload]
    from oracle import mpqp
    mpc, qp = build_problem(ctrl)
    P = oracle_problem(mpc, qp)
    lo, hi = theta_range(ctrl, scale)
    n = min(n_per, 2_000_000)
    Th = mpqp.sample_theta(lo, hi, n, seed=SEED)
    from oracle import ldp
    cores = ldp.num_threads()
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        P.solve_batch(Th, nthreads=cores)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    val = n * len(times) / total
    line = {"impl": "reference", "metric": "batched MPC QP solves/sec", "value": val, "unit": "solves/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "controller": ctrl, "theta_scale": scale,
                       "instances_per_step": n, "note": "bounded sample of the 1e7-theta workload"},
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": f"{n} theta per step, oracle/ldp_oracle.c cold start, {cores} pthreads"},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run this rank (and first-touch its pinned staging buffers) on the CPUs
    next to its GPU, as NVML reports them, so that the end-to-end leg's H2D / D2H traffic does not
    cross the socket interconnect.  Returns the number of CPUs bound to, or None when unavailable."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="invpend_1e7", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None

    import torch
    import torch.distributed as dist
    import lmpc_b200
    from oracle import mpqp  # fixture generator + checker only

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    ctrl, scale, n_per = WORKLOADS[args.workload]
    g = product_controller(ctrl, local_rank)          # product path: no oracle involved
    mpc, qp = build_problem(ctrl)                      # checker path
    nth, nu = g.nth, g.nu
    lo, hi = theta_range(ctrl, scale)

    # this rank's contiguous theta range, generated on the device (Philox, seed 1234 + rank)
    gen = torch.Generator(device=dev)
    gen.manual_seed(SEED + rank)
    Th = torch.rand((n_per, nth), dtype=torch.float64, device=dev, generator=gen)
    Th = Th * torch.tensor(hi - lo, device=dev) + torch.tensor(lo, device=dev)   # N x nth == nth x N col-major
    U = torch.empty((n_per, nu), dtype=torch.float64, device=dev)
    flag = torch.empty(n_per, dtype=torch.int32, device=dev)

    # ---- parity gate on the cfg-1 set (10^4 theta) and CPU baseline, rank 0 only -----------
    parity, cpu_baseline, flops_per_solve = None, None, None
    if rank == 0:
        sub = Th[:10000].cpu().numpy().T
        cst = g.ldp_constants()
        P = oracle_problem(mpc, qp, cst)
        got = g.solve(np.asfortranarray(sub))
        ref = P.solve_batch(sub, want_flops=True)
        same = (got["exitflag"] == ref["flag"]) & np.all(got["active_upper"] == ref["act_u"], axis=1) & \
               np.all(got["active_lower"] == ref["act_l"], axis=1)
        rel = float(np.max(np.abs(got["U"] - ref["U"]) / np.maximum(1.0, np.abs(ref["U"]))))
        parity = {"n": 10000, "identical_flag_and_active_set": int(same.sum()), "max_rel_err_u": rel,
                  "bit_identical_u": bool(np.array_equal(got["U"], ref["U"])),
                  "degenerate": count_degenerate(P, sub, ref)}
        flops_per_solve = float(ref["flops"].mean())
        if not args.no_cpu_baseline and world == 1:
            host = Th.cpu().numpy().T                      # the whole 1e7-theta batch of rank 0
            val, cores, ns, fl, _ = cpu_leg(P, host, passes=2)
            flops_per_solve = fl
            cpu_baseline = {"value": val, "unit": "solves/s", "cores": cores, "kind": "port",
                            "sample": f"2 passes over rank 0's {ns} theta, oracle/ldp_oracle.c cold start, "
                                      f"{cores} pthreads (the real reference, Julia+libdaqp, cannot run here)"}

    # ---- device-resident timing ---------------------------------------------------------------
    # an explicit (non-null) stream: the C ABI reads a NULL stream as "the handle's own stream",
    # which torch events on the legacy default stream would not see
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    def step():
        g.solve_async(Th, U, flag, N=n_per, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()      # runs through warm-up, the timed region and the e2e leg
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = lmpc_b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_step = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev0.record(stream)
    for a, b in per_step:
        a.record(stream); step(); b.record(stream)
    ev1.record(stream)
    barrier()
    launches = lmpc_b200.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in per_step]))
    if world > 1:
        ms_total = lmpc_b200.max_over_ranks(ms_total, device=dev)
    ok_all = bool((flag >= 1).all().item())

    # ---- end-to-end through the C ABI with pinned host buffers ----------------------------------
    Th_h = torch.empty((n_per, nth), dtype=torch.float64).pin_memory()
    Th_h.copy_(Th)
    U_h = torch.empty((n_per, nu), dtype=torch.float64).pin_memory()
    fl_h = torch.empty(n_per, dtype=torch.int32).pin_memory()
    def e2e_step():
        rc = lmpc_b200.lib().lmpc_gpu_solve_batch(g._h, Th_h.data_ptr(), n_per, U_h.data_ptr(), fl_h.data_ptr(),
                                                  None, None, None)
        assert rc == 0, lmpc_b200.lib().lmpc_gpu_last_error()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        e2e_s = lmpc_b200.max_over_ranks(e2e_s, device=dev)
    e2e_ok = bool(torch.equal(U_h, U.cpu()))
    clocks = sampler.stop() if rank == 0 else None

    # ---- final result gather (the only inter-GPU traffic of the path), untimed but checked ------
    gather_ms = None
    if world > 1:
        torch.cuda.synchronize(); dist.barrier()
        tg = time.perf_counter()
        allU = lmpc_b200.gather_columns(U.T, n_per * world)      # nu x (N * world) on rank 0
        torch.cuda.synchronize()
        gather_ms = 1e3 * (time.perf_counter() - tg)
        if rank == 0:
            assert torch.equal(allU[:, :n_per], U.T)

    if rank == 0:
        total_solves = float(n_per) * world * args.steps
        value = total_solves / (ms_total * 1e-3)
        hbm_peak, peak_src = peaks()
        fp64_peak = lmpc_b200.fp64_peak_tflops(local_rank)
        bytes_per_solve = 8 * nth + 8 * nu + 4
        ach_tf = flops_per_solve * n_per / (kernel_ms * 1e-3) / 1e12
        ach_gbs = bytes_per_solve * n_per / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(args.workload)
        line = {
            "metric": "batched MPC QP solves/sec", "value": value, "unit": "solves/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "controller": ctrl, "n": g.n, "m": g.m, "nth": nth,
                       "theta_scale": scale, "instances_per_gpu_per_step": n_per, "seed": SEED,
                       "kernel": {1: "thread-per-QP", 2: "warp-per-QP", 3: "CTA-per-QP"}[g.kernel],
                       "l2": "inputs larger than L2 (560 MB theta per step vs 126 MB L2)",
                       "settings": {"primal_tol": 1e-6, "dual_tol": 1e-12, "zero_tol": 1e-11,
                                    "progress_tol": 1e-6, "cycle_tol": 10, "iter_limit": 10000},
                       "all_exitflags_optimal": ok_all},
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach_tf / fp64_peak, "traffic": traffic,
                         "kernel": "ldp_box_kernel<5,1> (thread-per-QP, box-only fast path)", "kernel_ms": kernel_ms,
                         "flops_per_solve": flops_per_solve,
                         "peak_source": "FP64 FMA microbenchmark in this run (lmpc_gpu_fp64_peak)",
                         "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                                 "bytes_per_solve": bytes_per_solve, "peak_source": peak_src}},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": float(n_per) * world * e2e_steps / e2e_s, "unit": "solves/s",
                    "h2d_bytes_per_step": 8 * nth * n_per, "d2h_bytes_per_step": (8 * nu + 4) * n_per,
                    "steps": e2e_steps, "timer": "wall clock around lmpc_gpu_solve_batch (host pointers)",
                    "matches_device_path": e2e_ok,
                    "cpus_bound_per_rank": numa},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": parity,
        }
        if gather_ms is not None:
            line["gather_ms"] = gather_ms
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
