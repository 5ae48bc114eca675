/* lmpc_gpu.h -- C ABI of liblmpcgpu.so: batched evaluation of a condensed linear-MPC law
 * on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE path of LinearMPC.jl: theta -> (affine QP update)
 * -> DAQP dual active-set solve -> u, and the ExplicitMPC point-location alternative.
 * Each entry point names the reference interface it replaces.  All matrices cross the
 * ABI COLUMN-MAJOR (Julia's native layout) unless stated; all pointers are plain
 * double/int32/uint64 arrays owned by the caller.  Batched arrays hold one instance per
 * COLUMN (Theta is nth x N, U is nu x N), so instance q is contiguous.
 *
 * Pointers passed to the *_batch calls may be host or device pointers (detected with
 * cudaPointerGetAttributes).  Host buffers are staged through the handle's pinned/device
 * scratch inside the call; device buffers are used in place on the handle's stream.
 * The handle's stream is non-blocking: work that PRODUCES a device buffer on another stream must have
 * completed (or be ordered by an event) before a synchronous entry point is handed that buffer; the
 * call itself returns only after its own kernels and copies have finished.
 *
 * Return value of every function: 0 on success, negative LMPC_E_* on API failure
 * (never an exception).  Per-instance solver status uses DAQP's exit-flag integers
 * (reference: src/setup.jl:15-20, src/utils.jl:46).
 */
#ifndef LMPC_GPU_H
#define LMPC_GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* per-instance exit flags (DAQP convention) */
#define LMPC_EXIT_OPTIMAL          1
#define LMPC_EXIT_SOFT_OPTIMAL     2
#define LMPC_EXIT_INFEASIBLE     (-1)
#define LMPC_EXIT_CYCLE          (-2)
#define LMPC_EXIT_UNBOUNDED      (-3)
#define LMPC_EXIT_ITERLIMIT      (-4)
#define LMPC_EXIT_NONCONVEX      (-5)
#define LMPC_EXIT_OVERDETERMINED (-6)
#define LMPC_EXIT_CAPACITY       (-7)  /* working set outgrew the kernel's on-chip capacity */

/* API-level errors */
#define LMPC_E_INVALID    (-100)  /* bad argument / dimension */
#define LMPC_E_UNSUPPORTED (-101) /* binaries, non-symmetric H (AVI), hierarchical break points */
#define LMPC_E_SETUP      (-102)  /* Cholesky failed: H not positive definite (setup flag -5) */
#define LMPC_E_CUDA       (-103)  /* CUDA runtime error; see lmpc_gpu_last_error() */
#define LMPC_E_NOGPU      (-104)  /* no usable sm_100 device: there is NO CPU fallback */
/* the two other flags DAQP.setup can return (src/setup.jl:14-23), found once at create instead of per instance */
#define LMPC_E_SETUP_INFEASIBLE     (-105)  /* setup flag -1: a row's lower bound lies above its upper bound */
#define LMPC_E_SETUP_OVERDETERMINED (-106)  /* setup flag -6: equality rows dependent or more than n */

/* constraint sense bits (DAQPBase constants, set at src/mpc2mpqp.jl:868-885) */
#define LMPC_ACTIVE 1
#define LMPC_LOWER 2
#define LMPC_IMMUTABLE 4
#define LMPC_SOFT 8
#define LMPC_BINARY 16

/* kernel selection (LMPC_KERNEL_AUTO picks by n: <=8 thread-per-QP, <=32 warp-per-QP, else CTA-per-QP) */
#define LMPC_KERNEL_AUTO 0
#define LMPC_KERNEL_THREAD 1
#define LMPC_KERNEL_WARP 2
#define LMPC_KERNEL_CTA 3

/* Threading.  A handle's problem data never changes after create, so any number of host threads
 * may use one handle: the *_async entry points (device pointers, caller's stream) run concurrently;
 * calls that take HOST buffers share the handle's staging buffers and are serialised by a mutex
 * inside the handle (use one handle per thread, or lmpc_gpu_create_multi, to overlap them).
 * lmpc_gpu_last_error() is thread-local.  (The reference's workspace and mpc.uprev are mutated
 * in place and not re-entrant at all: SURVEY 8b "Threading".) */
typedef struct lmpc_handle lmpc_handle;   /* one implicit controller on one GPU */
typedef struct lmpc_bst lmpc_bst;         /* one explicit controller (search tree) on one GPU */

/* The dense mpQP exactly as LinearMPC.MPQP holds it (src/types.jl:75-98) plus the
 * runtime fields of MPC the path reads (src/types.jl:108-160: K, settings.soft_weight). */
typedef struct {
  int32_t n;        /* decision variables                      size(H,1)            */
  int32_t m;        /* two-sided rows                          length(bu)           */
  int32_t ms;       /* leading simple bounds                   m - size(A,1)        */
  int32_t nth;      /* parameter length                        size(f_theta,2)      */
  int32_t nu;       /* controls returned (first nu variables)  mpc.model.nu         */
  int32_t nx;       /* leading theta entries K multiplies      mpc.model.nx         */
  const double *H;        /* n x n                                                  */
  const double *f;        /* n                                                      */
  const double *f_theta;  /* n x nth                                                */
  const double *A;        /* (m-ms) x n                                             */
  const double *bu;       /* m   (+-1e30 = infinite, src/mpc2mpqp.jl:887-888)       */
  const double *bl;       /* m                                                      */
  const double *W;        /* m x nth                                                */
  const int32_t *senses;  /* m                                                      */
  const double *K;        /* nu x nx prestabilising feedback, NULL = 0              */
  double soft_weight;     /* mpc.settings.soft_weight; rho_soft = 1/soft_weight (src/setup.jl:26) */
  int32_t has_binaries;   /* MPQP.has_binaries   -> refused                         */
  int32_t is_symmetric;   /* MPQP.is_symmetric   -> 0 refused                       */
  int32_t n_break_points; /* length(break_points) -> >0 refused                     */
} lmpc_problem;

/* DAQP settings used by the loop (docs/src/manual/solver.md:49-56); zero-initialise
 * and call lmpc_gpu_default_settings() to get the defaults. */
typedef struct {
  double primal_tol;    /* 1e-6  on normalised slacks */
  double dual_tol;      /* 1e-12 */
  double zero_tol;      /* 1e-11 pivot below which an added row is treated as dependent */
  double progress_tol;  /* 1e-6  */
  double fval_bound;    /* 1e30  */
  int32_t cycle_tol;    /* 10    */
  int32_t iter_limit;   /* 10000 */
  int32_t kernel;       /* LMPC_KERNEL_* */
  int32_t certified_iterations; /* > 0: a proven bound on DAQP's iteration count over the parameter range,
                                 * CertificationResult.max_iterations of certify(mpc; range) (src/certify.jl:18-29).
                                 * Becomes a hard iteration cap (iter_limit = min(iter_limit, bound + 1)) and shrinks
                                 * the on-chip working-set capacity to n_eq + bound rows (every iteration adds at most
                                 * one), i.e. smaller factor tiles and more QPs resident per SM.  0: no certificate */
} lmpc_settings;

void lmpc_gpu_default_settings(lmpc_settings *s);
const char *lmpc_gpu_last_error(void);
int lmpc_gpu_device_count(void);

/* Replaces setup!(mpc) -> DAQP.setup (src/setup.jl:7-29) and qp2ldp
 * (src/codegen.jl:239-280): Cholesky, R^-1, M = A R^-1, row normalisation, the
 * theta-absorbed bounds Dth/du/dl and the output map; uploads them once to `device`. */
int lmpc_gpu_create(const lmpc_problem *p, const lmpc_settings *s, int device, lmpc_handle **out);
int lmpc_gpu_destroy(lmpc_handle *h);

/* Read back the LDP constants the handle computed (row-major, as generated C stores
 * them, src/codegen.jl:275): Mbar m x n, Gram m x m (= Mbar Mbar'), Dth m x nth, du m, dl m,
 * Rinv nu x n, Uth nu x nth, uoff nu.  Any pointer may be NULL.  Lets a caller (and the
 * parity tests) check the setup step against qp2ldp. */
int lmpc_gpu_get_ldp(const lmpc_handle *h, double *Mbar, double *Gram, double *Dth, double *du,
                     double *dl, double *Rinv, double *Uth, double *uoff);
/* The same constants straight from a problem description, host only (no GPU touched, no
 * handle): the once-per-controller setup step in isolation.  senses m, cap 1 (nullable). */
int lmpc_gpu_host_ldp(const lmpc_problem *p, double *Mbar, double *Gram, double *Dth, double *du,
                      double *dl, double *Rinv, double *Uth, double *uoff, int32_t *senses,
                      int32_t *cap);
/* which kernel create() selected (LMPC_KERNEL_*), the working-set capacity it provisions,
 * and the effective sense bits (m; zero rows are marked IMMUTABLE).  Pointers may be NULL. */
int lmpc_gpu_get_info(const lmpc_handle *h, int32_t *kernel, int32_t *cap, int32_t *senses);

/* CTA-per-QP controllers whose factor nearly fills an SM's shared memory (chain of masses: 129 KB) run
 * in two passes: a first pass with two CTAs per SM and a working-set capacity `fast_cap` < capacity, then
 * a full-capacity pass over the instances that outgrew it (results are those of the full capacity).
 * fast_cap = 0: single pass.  redone: instances the second pass has handled on this handle's device since
 * the process started.  Either pointer may be NULL. */
int lmpc_gpu_get_tiering(const lmpc_handle *h, int32_t *fast_cap, int64_t *redone);

/* Replaces solve(mpc, theta) + the compute_control epilogue (src/utils.jl:43-51,268-283)
 * for N parameters at once.  Theta nth x N; U nu x N receives z[1:nu] - K theta[1:nx];
 * exitflag N; iters N (nullable); active_upper/active_lower ceil(m/64) words per
 * instance, bit j of word j/64 = row j active at its upper/lower bound (nullable).
 * Cold start per instance (as the generated C without DAQP_WARMSTART,
 * codegen/mpc_update_qp.c:44-47). */
int lmpc_gpu_solve_batch(lmpc_handle *h, const double *Theta, int64_t N, double *U,
                         int32_t *exitflag, int32_t *iters,
                         uint64_t *active_upper, uint64_t *active_lower);

/* Same, device pointers only, asynchronous on `stream` (a cudaStream_t, NULL = the
 * handle's own stream); the caller synchronises. */
int lmpc_gpu_solve_batch_async(lmpc_handle *h, const double *Theta, int64_t N, double *U,
                               int32_t *exitflag, int32_t *iters,
                               uint64_t *active_upper, uint64_t *active_lower, void *stream);

/* Warm-started solve: the batch form of DAQP's warm start (the Julia path keeps the workspace
 * between solve() calls, src/utils.jl:278-282; generated C: DAQP_WARMSTART,
 * codegen/mpc_update_qp.c:44-47).  active_upper / active_lower are IN/OUT (required): on entry the
 * working set to start from (normally what the previous call returned for the same closed
 * loop), on return the final one.  Equality rows enter first as in a cold start, then the named
 * rows in ascending order with zero multipliers; a named row that would make the factorisation
 * singular is left out.  Results equal the cold solve's up to rounding (a different
 * factorisation order); the iteration count is what changes.  Host or device pointers. */
int lmpc_gpu_solve_batch_warm(lmpc_handle *h, const double *Theta, int64_t N, double *U,
                              int32_t *exitflag, int32_t *iters,
                              uint64_t *active_upper, uint64_t *active_lower);

/* Batched analogue of the generated mpc_compute_control(control, state, reference,
 * disturbance) (codegen/mpc_update_qp.h:3-7, mpc_update_parameter.c:1-29): theta is
 * assembled on the fly as [state; reference; disturbance; control] with
 * nx + nr + nd + nuprev = nth; `control` (nu x N) is in/out: it holds u_prev on entry
 * (first nuprev rows are read) and the new control on return.  reference/disturbance
 * may be NULL when nr/nd is 0.  Returns per-instance exit flags in `exitflag`. */
int lmpc_gpu_compute_control_batch(lmpc_handle *h, double *control, const double *state,
                                   const double *reference, const double *disturbance,
                                   int32_t nr, int32_t nd, int32_t nuprev, int64_t N,
                                   int32_t *exitflag);

/* Monte-Carlo form of the same call: any of reference / disturbance / uprev may be ONE vector shared by
 * all N instances (bit set in `bcast`: the block is then nr x 1 / nd x 1 / nuprev x 1 and is uploaded
 * once instead of N times -- compute_control(mpc, x; r, d, uprev) called in a loop over x with the other
 * arguments fixed, src/utils.jl:43-51).  `control` (nu x N) is output only here; uprev is its own
 * argument.  exitflag may be NULL when n_failed is given: the flags then stay on the device and only
 * the number of instances with exitflag < 1 (what the reference's @assert would trip on) comes back.
 * Host or device pointers.  End to end this moves 8 nx + 8 nu bytes per instance instead of
 * 8 nth + 8 nu + 4. */
#define LMPC_BCAST_REFERENCE 1
#define LMPC_BCAST_DISTURBANCE 2
#define LMPC_BCAST_UPREV 4
int lmpc_gpu_compute_control_batch_bcast(lmpc_handle *h, double *control, const double *state,
                                         const double *reference, const double *disturbance,
                                         const double *uprev, int32_t nr, int32_t nd, int32_t nuprev,
                                         int64_t N, int32_t bcast, int32_t *exitflag, int64_t *n_failed);

/* Replaces N independent copies of the Simulation loop (src/simulation.jl:93-113) with the
 * default linear dynamics x+ = F x + G u + Gd d (src/model.jl:70): T steps of
 * compute_control + dynamics, entirely on the GPU.  F nx x nx, G nx x nu, Gd nx x nd
 * column-major.  X0 nx x N; R nr x N and D nd x N are held constant over the horizon;
 * U0 nuprev x N (NULL = 0).  warm_start != 0 carries the working set between steps (the
 * Julia path's behaviour); 0 cold-starts every step (generated C default).
 * Outputs: Xend nx x N, Ulast nu x N, Utraj nu x T x N (nullable), exitflag N = worst
 * flag over the T steps, iters N = total iterations (nullable). */
int lmpc_gpu_simulate_batch(lmpc_handle *h, const double *F, const double *G, const double *Gd,
                            int32_t nx, int32_t nr, int32_t nd, int32_t nuprev,
                            const double *X0, const double *R, const double *D, const double *U0,
                            int64_t N, int32_t T, int32_t warm_start,
                            double *Xend, double *Ulast, double *Utraj,
                            int32_t *exitflag, int32_t *iters);

/* Same, with a reference and / or measured-disturbance TRAJECTORY shared by all N closed loops:
 * Rtraj nr x T, Dtraj nd x T, column-major HOST arrays (column t = rs[:, t+1], ds[:, t+1] of
 * src/simulation.jl:70-80, i.e. already extended by "hold the last value"); either may be NULL, and
 * where given it replaces the per-trajectory constant R / D.  The disturbance of step t enters both
 * theta and the dynamics, as in src/simulation.jl:93-110. */
int lmpc_gpu_simulate_batch_traj(lmpc_handle *h, const double *F, const double *G, const double *Gd,
                                 int32_t nx, int32_t nr, int32_t nd, int32_t nuprev,
                                 const double *X0, const double *R, const double *D, const double *U0,
                                 const double *Rtraj, const double *Dtraj,
                                 int64_t N, int32_t T, int32_t warm_start,
                                 double *Xend, double *Ulast, double *Utraj,
                                 int32_t *exitflag, int32_t *iters);

/* Extended closed loop: the scenario trajectories of simulate_batch_traj plus
 *  - a TRUE PLANT x+ = Ft x + Gt u + Gdt d that may differ from the controller's prediction model
 *    (scenario.dynamics, src/simulation.jl:39,110); NULL = the model's matrix;
 *  - the steady-state Kalman observer of src/observer.jl:104-118 run inside the loop exactly as
 *    src/simulation.jl:92-108 does: estimate = X0 at the start (set_state!), every step
 *    y = C x + Dd d (+ v_t), correct! with gain K, compute_control on the ESTIMATE, predict! with the
 *    model; C ny x nx, Dd ny x nd (NULL = 0), K nx x ny column-major; C == NULL: no observer, the
 *    controller sees the true state;
 *  - Vtraj ny x T: additive measurement noise shared by all closed loops (what a scenario's
 *    get_measurement would add), NULL = none.
 * All matrices are HOST arrays.  Xend returns the TRUE final states, Xhat_end (nullable) the final
 * estimates.  ext == NULL or all-NULL fields: identical to lmpc_gpu_simulate_batch. */
typedef struct {
  const double *Rtraj, *Dtraj;
  const double *Ft, *Gt, *Gdt;
  const double *C, *Dd, *K;
  const double *Vtraj;
  int32_t ny;
  double *Xhat_end;
  /* Preview windows (reference_preview / disturbance_preview, src/simulation.jl:102-104,128-134):
   * with rstride > 0, Rtraj is a flat HOST array and step t reads its nr = ny_ref * Np entries at
   * offset t * rstride (rstride = ny_ref slides the window by one column per step; the caller
   * appends the held last column Np times); likewise dstride for Dtraj.  0 = nr / nd (plain
   * one-column-per-step trajectories, length nr * T). */
  int32_t rstride, dstride;
} lmpc_sim_ext;
int lmpc_gpu_simulate_batch_ex(lmpc_handle *h, const double *F, const double *G, const double *Gd,
                               int32_t nx, int32_t nr, int32_t nd, int32_t nuprev,
                               const double *X0, const double *R, const double *D, const double *U0,
                               const lmpc_sim_ext *ext,
                               int64_t N, int32_t T, int32_t warm_start,
                               double *Xend, double *Ulast, double *Utraj,
                               int32_t *exitflag, int32_t *iters);

/* Stepped closed loop: the same Simulation loop, one short kernel sequence per time step with the
 * loop state in HBM, for what the fused on-chip loop above cannot express:
 *  - a NONLINEAR (or otherwise user-defined) TRUE PLANT, `scenario.dynamics` of src/simulation.jl:39,110 and
 *    Model(f, h, ...)'s true_dynamics (src/model.jl:113-124): `dynamics` is called once per step and must enqueue,
 *    on `stream`, work that writes Xnext (nx x N) from X (nx x N), U (nu x N) and D (ndm x N, NULL when ndm = 0),
 *    all DEVICE arrays, instance per column; returns 0 on success.  dynamics == NULL: the affine plant
 *    x+ = Ft x + Gt u + Gdt d + ft (NULL matrices = the controller's model F, G, Gd; ft NULL = 0);
 *  - an observer with ITS OWN STATE of dimension nxo >= nx: the KalmanFilter of src/observer.jl:1-124 (nxo = nx)
 *    or the OffsetFreeObserver's augmented estimator (src/observer.jl:12-47, 203-222), whose last nxo - nx
 *    entries are estimated disturbances that are appended to the ndm measured ones in theta
 *    (get_control_disturbance, src/observer.jl:203-222).  Every step: y = Cm x + Ddm d + hm (+ v_t),
 *    correct! xhat += Ko (y - Co xhat - Ddo d - ho), compute_control on xhat[1:nx], predict!
 *    xhat = Fo xhat + Go u + Gdo d + fo.  Ko == NULL: no observer (nxo is taken as nx).
 * theta = [xhat[1:nx]; r; d_measured; xhat[nx+1:nxo]; uprev], so nx + nr + ndm + (nxo - nx) + nuprev = nth.
 * Matrices are column-major HOST arrays: Ft nx x nx, Gt nx x nu, Gdt nx x ndm, Cm ny x nx, Ddm ny x ndm,
 * Fo nxo x nxo, Go nxo x nu, Gdo nxo x ndm, Co ny x nxo, Ddo ny x ndm, Ko nxo x ny; Rtraj nr x T, Dtraj ndm x T,
 * Vtraj ny x T shared by all closed loops (NULL: R / D per closed loop, no noise).  X0, R (nr x N), D (ndm x N),
 * U0 and the outputs are HOST arrays.  warm_start carries the working set between steps through
 * the warm-start entry of the solve kernels.  Xhat_end (nullable) nxo x N. */
typedef int (*lmpc_plant_fn)(void *user, void *stream, int64_t N, int32_t k, const double *X, const double *U,
                             const double *D, double *Xnext);
typedef struct {
  lmpc_plant_fn dynamics;
  void *user;
  const double *Ft, *Gt, *Gdt, *ft;
  int32_t ny;
  const double *Cm, *Ddm, *hm, *Vtraj;
  int32_t nxo, ndm;
  const double *Fo, *Go, *Gdo, *fo, *Co, *Ddo, *ho, *Ko;
  const double *Rtraj, *Dtraj;
  double *Xhat_end;
} lmpc_sim_stepped;
int lmpc_gpu_simulate_batch_stepped(lmpc_handle *h, const double *F, const double *G, const double *Gd,
                                    int32_t nx, int32_t nr, int32_t nuprev,
                                    const double *X0, const double *R, const double *D, const double *U0,
                                    const lmpc_sim_stepped *cfg, int64_t N, int32_t T, int32_t warm_start,
                                    double *Xend, double *Ulast, double *Utraj, int32_t *exitflag, int32_t *iters);

/* Closed loops with the EXPLICIT law: N copies of Simulation(empc, scenario)
 * (src/simulation.jl:37-116 accepts an ExplicitMPC) with the linear dynamics of src/model.jl:70;
 * every step walks the tree on chip.  Arguments as lmpc_gpu_simulate_batch_traj (Rtraj / Dtraj
 * nullable HOST trajectories shared by all loops); nth <= 16, nu <= 4. */
int lmpc_gpu_simulate_bst_batch(lmpc_bst *t, const double *F, const double *G, const double *Gd,
                                int32_t nx, int32_t nr, int32_t nd, int32_t nuprev,
                                const double *X0, const double *R, const double *D, const double *U0,
                                const double *Rtraj, const double *Dtraj, int64_t N, int32_t T,
                                double *Xend, double *Ulast, double *Utraj);

/* One process, several GPUs (BASELINE north_star: "the batch is sharded by contiguous theta ranges
 * across the 8 B200s of one box with no inter-GPU traffic beyond a final result gather").
 * create_multi builds one replica of the controller per listed device (devices NULL = 0..n-1);
 * solve_batch_multi takes HOST buffers, gives device g the contiguous range
 * [g N / G, (g+1) N / G), runs the per-device host pipelines concurrently (one worker thread per
 * GPU) and returns when every slice of U / exitflag has landed in the caller's arrays. */
typedef struct lmpc_multi lmpc_multi;
int lmpc_gpu_create_multi(const lmpc_problem *p, const lmpc_settings *s, const int32_t *devices,
                          int32_t n_devices, lmpc_multi **out);
int lmpc_gpu_destroy_multi(lmpc_multi *m);
int lmpc_gpu_solve_batch_multi(lmpc_multi *m, const double *Theta, int64_t N, double *U,
                               int32_t *exitflag, int32_t *iters,
                               uint64_t *active_upper, uint64_t *active_lower);

/* Closed loops and explicit controllers on every GPU of the box from one process: the same contiguous
 * split (device g owns [g N / G, (g+1) N / G)), HOST buffers, one worker thread per device; arguments
 * otherwise as lmpc_gpu_simulate_batch_ex / lmpc_gpu_evaluate_batch. */
int lmpc_gpu_simulate_batch_multi(lmpc_multi *m, const double *F, const double *G, const double *Gd,
                                  int32_t nx, int32_t nr, int32_t nd, int32_t nuprev,
                                  const double *X0, const double *R, const double *D, const double *U0,
                                  const lmpc_sim_ext *ext, int64_t N, int32_t T, int32_t warm_start,
                                  double *Xend, double *Ulast, double *Utraj,
                                  int32_t *exitflag, int32_t *iters);

/* The final result gather of BASELINE's north_star, fused into the solve: a "window" is a device buffer
 * on ONE GPU (rank 0's) that every other process of the box maps through CUDA IPC; each rank then passes
 * `window + its offset` as the U / exitflag output of lmpc_gpu_solve_batch_async, so the kernel's own
 * stores travel over NVLink into the gathered array (8 nu + 4 bytes per solve, no second kernel, no
 * NCCL call).  create: cudaMalloc on `device` + a 64-byte handle to send to the peers (any transport);
 * open: map a peer's window into this process on `device`; close: unmap (owner = 0) or free (owner = 1). */
int lmpc_gpu_window_create(int device, size_t bytes, void **devptr, unsigned char handle[64]);
int lmpc_gpu_window_open(int device, const unsigned char handle[64], void **devptr);
int lmpc_gpu_window_close(void *devptr, int owner);
/* copy `bytes` of a window (at devptr, any offset into it) to host memory; synchronous */
int lmpc_gpu_window_read(void *dst_host, const void *devptr, size_t bytes);

/* Replaces ParametricDAQP.build_tree's result as used by compute_control(empc, ...)
 * (src/utils.jl:53-60, src/explicit.jl:65-71).  halfplanes (nth+1) x n_hp column-major
 * (column h = [a_h; b_h]); feedbacks n_fb blocks of (nth+1) x nu column-major (u = F'[theta;1]);
 * hp_list / jump_list n_nodes (0-based here): at node i, jump_list[i] == 0 marks a leaf whose
 * region is hp_list[i]; otherwise next = i + jump_list[i] and the walk goes to `next` if
 * a'theta <= b for halfplane hp_list[i], else to next+1.  clipping nu x 2 column-major
 * [min max] or NULL. */
int lmpc_gpu_create_bst(const double *halfplanes, int32_t n_hp, const double *feedbacks,
                        int32_t n_fb, const int32_t *hp_list, const int32_t *jump_list,
                        int32_t n_nodes, const double *clipping, int32_t nth, int32_t nu,
                        int device, lmpc_bst **out);
/* Same, for a tree whose producer uses the OTHER branch orientation.  le_child names the child taken
 * when a'theta <= b: 0 = `next` (this library's convention, what lmpc_gpu_create_bst assumes and what the
 * repo's own builder emits), 1 = `next + 1`.  ParametricDAQP is not vendored in the reference, so its
 * convention cannot be read off here; from its generated C (`id = next_id + (val <= b)`) it is believed to
 * be 1, which is what the Julia glue passes -- and then VERIFIES against ParametricDAQP.evaluate on sampled
 * theta before trusting the upload (LinearMPCGPU.jl: setup_gpu!(empc)).  The children are exchanged once,
 * on the host, at create time: the kernels always walk one convention. */
int lmpc_gpu_create_bst_oriented(const double *halfplanes, int32_t n_hp, const double *feedbacks,
                                 int32_t n_fb, const int32_t *hp_list, const int32_t *jump_list,
                                 int32_t n_nodes, const double *clipping, int32_t nth, int32_t nu,
                                 int32_t le_child, int device, lmpc_bst **out);
int lmpc_gpu_destroy_bst(lmpc_bst *t);
/* Diagnostics: which evaluate kernel this tree gets (0 = generic fp64 rows, 1 = fp32-screened walk, 2 = fp32-screened
 * lane-refill ring); no reference counterpart (ParametricDAQP.evaluate is one scalar loop, src/utils.jl:57). */
int lmpc_gpu_bst_kernel(const lmpc_bst *t);
/* Replaces ParametricDAQP.evaluate(bst, theta) (src/utils.jl:57) / {fname}_evaluate
 * (src/codegen.jl:113-124) for N parameters.  region (nullable) receives the leaf's region id. */
int lmpc_gpu_evaluate_batch(lmpc_bst *t, const double *Theta, int64_t N, double *U,
                            int32_t *region);
int lmpc_gpu_evaluate_batch_async(lmpc_bst *t, const double *Theta, int64_t N, double *U,
                                  int32_t *region, void *stream);
/* one process, several GPUs: a replica of the tree per listed device (devices NULL = 0..n-1), a HOST
 * batch split into contiguous ranges */
typedef struct lmpc_bst_multi lmpc_bst_multi;
int lmpc_gpu_create_bst_multi(const double *halfplanes, int32_t n_hp, const double *feedbacks,
                              int32_t n_fb, const int32_t *hp_list, const int32_t *jump_list,
                              int32_t n_nodes, const double *clipping, int32_t nth, int32_t nu,
                              const int32_t *devices, int32_t n_devices, lmpc_bst_multi **out);
int lmpc_gpu_destroy_bst_multi(lmpc_bst_multi *t);
int lmpc_gpu_evaluate_batch_multi(lmpc_bst_multi *t, const double *Theta, int64_t N, double *U,
                                  int32_t *region);

/* ---------------------------------------------------------------------------------------------
 * Host-side condensing: build the dense mpQP from a discrete-time model without
 * Julia.  Replaces mpc2mpqp(mpc) (src/mpc2mpqp.jl:612-647) for the feature subset the batched
 * path supports: move blocking, previews of r and d, a prestabilising feedback, generalized (affine)
 * parameters with or without a parameter preview, model offsets and operating points, reference condensation;
 * no binaries.
 * Pure host code; needs no GPU.
 * --------------------------------------------------------------------------------------------- */
typedef struct {            /* lb <= Ax x_k + Au u_k + Ar r + Ad d + Aup u^- <= ub for k in ks (src/types.jl:4-18) */
  int32_t rows;
  const double *Au;         /* rows x nu, NULL = 0 (all matrices column-major) */
  const double *Ax;         /* rows x nx, NULL = 0 */
  const double *Ar;         /* rows x ny, NULL = 0 */
  const double *Ad;         /* rows x nd, NULL = 0 */
  const double *Aup;        /* rows x nu, NULL = 0 */
  const double *ub, *lb;    /* rows, NULL = +-1e30 */
  const int32_t *ks;        /* 1-based steps as in the reference: k = 1 is x0 (default there: 2:Np) */
  int32_t nks;
  int32_t soft;             /* src/setup.jl:94: output bounds are soft by default */
  int32_t prio;
  const double *Ap;         /* rows x np: generalized (affine) parameter term Ap p on the constrained side
                             * (add_constraint!(...; Ap), src/setup.jl:61-84; W gets -Ap, mpc2mpqp.jl:125-143), NULL = 0 */
} lmpc_constraint;

typedef struct {            /* x+ = F x + G u + Gd d, y = C x + Dd d (src/model.jl:17-41) */
  int32_t nx, nu, ny, nd, Np, Nc;     /* Nc <= 0 means Nc = Np */
  const double *F, *G;
  const double *C;          /* ny x nx, NULL = identity (ny = nx) */
  const double *Gd, *Dd;    /* nx x nd, ny x nd, NULL = 0 */
  const double *Q, *R, *Rr, *Qf;      /* ny x ny, nu x nu, nu x nu, ny x ny; NULL = I, I, 0, 0 (src/types.jl:35-38) */
  const double *umin, *umax;          /* nu each or NULL (src/setup.jl:36-46) */
  const lmpc_constraint *constraints;
  int32_t n_constraints;
  int32_t reference_tracking;         /* MPCSettings.reference_tracking (default true) */
  int32_t preprocess;                 /* MPCSettings.preprocess_mpqp (default true) */
  /* move blocking (move_block!, src/setup.jl:196-251; apply_move_block, src/mpc2mpqp.jl:830-857):
   * the formatted blocks of all nu inputs back to back (each input's blocks sum to Np) and how many
   * each input has; Nc must be max_i(sum of all but the last block of input i) + 1 as move_block!
   * sets it.  NULL = no blocking. */
  const int32_t *move_blocks;
  const int32_t *n_move_blocks;
  /* MPCSettings.reference_preview / disturbance_preview (src/types.jl:56-57; reference condensation
   * off): theta carries ny*Np reference values / nd*Np disturbance values, column k of the preview
   * first (format_reference / format_disturbance, src/utils.jl:78-205).  Constraints with Ar != 0
   * are refused under reference_preview. */
  int32_t reference_preview;
  int32_t disturbance_preview;
  /* prestabilising feedback u = v - K x (set_prestabilizing_feedback!, src/setup.jl:186-199): nu x nx,
   * NULL = 0.  The QP is then over v, the control bounds become general rows (ms = 0), and the same K
   * must be given to lmpc_gpu_create for the output map u = z[1:nu] - K x (src/utils.jl:46-50). */
  const double *K;
  /* generalized (affine) parameters p, constant over the horizon (parameter_preview off): the linear cost
   * (Ex p + ex)'x_k + (Eu p + eu)'u_k of set_objective!(...; Ex, ex, Eu, eu) (src/setup.jl:118-132,
   * mpc2mpqp.jl:478-508) and the constraints' Ap.  theta = [x; r; d; uprev; p].  Ex nx x np, ex nx,
   * Eu nu x np, eu nu (any NULL = 0; every block given must have exactly np columns, as the reference
   * demands, mpc2mpqp.jl:486-490); np = 0: none. */
  int32_t np;
  const double *Ex, *ex, *Eu, *eu;
  /* settings.parameter_preview (src/types.jl:58): theta ends with Np blocks p_0 .. p_{Np-1} of np values instead
   * of one; stage k pairs u_k and x_{k+1} with p_k (stage_parameter_matrix, mpc2mpqp.jl:148), a constraint at
   * ks = k sees block min(k, Np) (parameter_preview_direct, mpc2mpqp.jl:125-143) */
  int32_t parameter_preview;
  /* affine model x+ = F x + G u + Gd d + f_offset, y = C x + Dd d + h_offset around the operating point (xo, uo)
   * (Model(...; f_offset, h_offset, xo, uo), src/model.jl:45-72; a linearised model, src/model.jl:113-124):
   * f_offset (nx) is a constant extra state collapsed into f / bu / bl (mpc2mpqp.jl:395-399, 518-522, 684-689),
   * h_offset (ny) shifts the tracked reference (mpc2mpqp.jl:523-530), xo (nx) / uo (nu) shift the cost of a
   * regulator (reference_tracking = 0, mpc2mpqp.jl:433-443, 457-459).  Any NULL = 0.  f_offset cannot be combined
   * with generalized parameters (LMPC_E_UNSUPPORTED). */
  const double *f_offset, *h_offset, *xo, *uo;
  /* settings.reference_condensation (src/types.jl:54,65) together with reference_preview: the previewed reference
   * trajectory is collapsed to ONE setpoint, theta carries ny reference values again, and the library computes the
   * ny x (ny Np) map traj2setpoint = (W H^-1 Fr Is) \ (W H^-1 Fr) (src/mpc2mpqp.jl:550-568; W = I with 1e6 on the first
   * move) that the caller applies to a flattened trajectory before compute_control (condense_reference,
   * src/utils.jl:140-147; read it with lmpc_gpu_mpqp_traj2setpoint) */
  int32_t reference_condensation;
} lmpc_model;

typedef struct lmpc_mpqp lmpc_mpqp;   /* the dense mpQP (src/types.jl:75-98) owned by the library */

int lmpc_gpu_condense(const lmpc_model *model, lmpc_mpqp **out);
/* dims9 = {n, m, ms, nth, nu, nx, nr, nd, nuprev}: theta = [x(nx); r(nr); d(nd); uprev(nuprev); p(nth - the rest)] */
int lmpc_gpu_mpqp_dims(const lmpc_mpqp *q, int32_t *dims9);
/* copy out, column-major: H n x n, f n, f_theta n x nth, A (m-ms) x n, bu m, bl m, W m x nth, senses m */
/* traj2setpoint (column-major rows x cols = ny x ny*Np; 0 x 0 without reference_condensation); T may be NULL */
int lmpc_gpu_mpqp_traj2setpoint(const lmpc_mpqp *q, double *T, int32_t *rows, int32_t *cols);
int lmpc_gpu_mpqp_get(const lmpc_mpqp *q, double *H, double *f, double *f_theta, double *A,
                      double *bu, double *bl, double *W, int32_t *senses);
int lmpc_gpu_mpqp_free(lmpc_mpqp *q);
const char *lmpc_gpu_condense_error(void);

/* Struct layouts as THIS build of the library sees them, so that a binding written in another
 * language (the Julia structs in linearmpc.jl_b200/julia/LinearMPCGPU.jl, the ctypes mirrors in
 * host.py) can be checked field by field instead of trusted: for each struct, in the order
 * lmpc_problem, lmpc_settings, lmpc_sim_ext, lmpc_constraint, lmpc_model the record
 *   { struct id (LMPC_ABI_*), sizeof, number of fields, offsetof(field 0), offsetof(field 1), ... }
 * with the fields in declaration order.  Writes at most `cap` int32 values to `out` (nullable)
 * and returns the number of values the full description has.  Host only, no GPU needed. */
#define LMPC_ABI_PROBLEM 1
#define LMPC_ABI_SETTINGS 2
#define LMPC_ABI_SIM_EXT 3
#define LMPC_ABI_CONSTRAINT 4
#define LMPC_ABI_MODEL 5
#define LMPC_ABI_SIM_STEPPED 6
int lmpc_gpu_abi_layout(int32_t *out, int32_t cap);

/* Measurement helpers used by bench.py (not part of the reference's interface):
 * sustained FP64 FMA throughput of the device in TFLOP/s, and the stream of a handle. */
int lmpc_gpu_fp64_peak(int device, double *tflops);
void *lmpc_gpu_stream(lmpc_handle *h);
/* number of kernel launches this library has issued in this process (for bench's gpu_launches) */
int64_t lmpc_gpu_launch_count(void);
/* pinned host memory for callers that want DMA-speed staging of Theta / U (cudaHostAlloc) */
void *lmpc_gpu_host_alloc(size_t bytes);
void lmpc_gpu_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* LMPC_GPU_H */
