/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's online path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  The product (liblmpcgpu.so) never does.
 *
 * What it restates
 *   - mpc_update_qp            /root/reference/codegen/mpc_update_qp.c:1-10
 *   - daqp_ldp (dual active set, recursive LDL' add/remove)
 *       NOT in /root/reference: DAQP C library via DAQPBase ~0.4.4
 *       (/root/reference/Project.toml:27; call sites src/utils.jl:281-282,
 *       codegen/mpc_update_qp.c:38-51).  Restated from the published
 *       algorithm (Arnstrom, Bemporad, Axehill, IEEE TAC 67(8) 2022, cited at
 *       /root/reference/README.md:71-83) with the concrete rules of
 *       SURVEY.md Appendix A.
 *   - mpc_get_solution         /root/reference/codegen/mpc_update_qp.c:14-22
 *   - Simulation loop          /root/reference/src/simulation.jl:93-113
 *   - ParametricDAQP.evaluate  call site /root/reference/src/utils.jl:57
 *
 * Parity status: pinned on the reference's known answers (runtests.jl:62-66,
 * simple.md:107,203) and on brute-force KKT enumeration; final active set and
 * exit flag versus the real libdaqp are UNPINNED (libdaqp cannot run here).
 */
#ifndef LDP_ORACLE_H
#define LDP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  ORC_OPTIMAL = 1, ORC_SOFT_OPTIMAL = 2, ORC_INFEASIBLE = -1, ORC_CYCLE = -2,
  ORC_UNBOUNDED = -3, ORC_ITERLIMIT = -4, ORC_NONCONVEX = -5, ORC_OVERDETERMINED = -6,
  ORC_CAPACITY = -7   /* not a DAQP flag: see orc_problem.cap */
};
enum { ORC_ACTIVE = 1, ORC_LOWER = 2, ORC_IMMUTABLE = 4, ORC_SOFT = 8, ORC_BINARY = 16 };

typedef struct {
  int n, m, ms, nth, nu;
  const double *Mbar;   /* m x n row-major, unit rows */
  const double *Gram;   /* m x m row-major, Mbar Mbar' */
  const double *Dth;    /* m x nth row-major */
  const double *du, *dl;/* m */
  const double *Rinv;   /* nu x n row-major: first nu rows of R^-1 (un-normalised) */
  const double *Uth;    /* nu x nth row-major */
  const double *uoff;   /* nu */
  const int32_t *sense; /* m */
  double rho_soft, primal_tol, dual_tol, zero_tol, progress_tol, fval_bound;
  int cycle_tol, iter_limit;
  int no_repair;        /* != 0: switch the soft-row repair (rebuild) off -- tests only, to show what it fixes */
  int cap;              /* > 0: working-set capacity; a row that would not fit ends the solve with
                         * ORC_CAPACITY, as the kernels do with the capacity they provision on chip
                         * (lmpc_gpu_get_info).  0: unbounded (max(m, n) + 1), DAQP's own behaviour */
} orc_problem;

/* one solve; returns exit flag.  act_u/act_l: ceil(m/64) words each (nullable).
 * flops (nullable) accumulates the algorithmic flop count of this solve. */
int orc_solve(const orc_problem *p, const double *theta, double *u, int32_t *iters,
              uint64_t *act_u, uint64_t *act_l, double *flops);

/* batch, Theta nth x N column-major, U nu x N column-major; nthreads <= 0 -> all cores */
void orc_solve_batch(const orc_problem *p, const double *Theta, int64_t N, double *U,
                     int32_t *flag, int32_t *iters, uint64_t *act_u, uint64_t *act_l,
                     double *flops, int nthreads);
/* warm start from the working sets in act_u / act_l (in/out, required): see activate_rows */
void orc_solve_batch_warm(const orc_problem *p, const double *Theta, int64_t N, double *U,
                          int32_t *flag, int32_t *iters, uint64_t *act_u, uint64_t *act_l,
                          int nthreads);

/* closed loop: x+ = Fx x + Gx u + Gdx d (simulation.jl:93-113, model.jl:70).
 * X0 nx x N, R nr x N, D nd x N (held constant), U0 nuprev x N (nullable -> 0).
 * Writes Xend nx x N, Ulast nu x N, optional Utraj nu x T x N, flags = worst flag,
 * iters = total iterations over the T steps.  warm != 0 carries the working set. */
void orc_simulate_batch(const orc_problem *p, int nx, int nr, int nd, int nuprev,
                        const double *Fx, const double *Gx, const double *Gdx,
                        const double *X0, const double *R, const double *D, const double *U0,
                        int64_t N, int T, int warm, double *Xend, double *Ulast, double *Utraj,
                        int32_t *flag, int32_t *iters, int nthreads);
/* true plant x+ = Ft x + Gt u + Gdt d (row-major; NULL = the model), steady-state observer
 * (observer.jl:104-118: C ny x nx, Dd ny x nd or NULL, K nx x ny; C == NULL: the controller sees
 * the true state) and additive measurement noise Vtraj (ny x T, shared) */
typedef struct {
  const double *Ft, *Gt, *Gdt, *C, *Dd, *K, *Vtraj;
  int ny;
  double *Xhat_end;   /* nx x N final estimates, nullable */
  int rstride, dstride;   /* > 0: step t reads its nr / nd entries at t * stride (preview windows) */
  double *flops;          /* N: algorithmic flops of each closed loop (all T solves), nullable */
} orc_sim_ext;
void orc_simulate_batch_ex(const orc_problem *p, int nx, int nr, int nd, int nuprev,
                           const double *Fx, const double *Gx, const double *Gdx,
                           const double *X0, const double *R, const double *D, const double *U0,
                           const double *Rtraj, const double *Dtraj, const orc_sim_ext *ext,
                           int64_t N, int T, int warm, double *Xend, double *Ulast, double *Utraj,
                           int32_t *flag, int32_t *iters, int nthreads);
/* same with reference / disturbance trajectories shared by all closed loops (nr x T, nd x T; NULL = use R / D) */
void orc_simulate_batch_traj(const orc_problem *p, int nx, int nr, int nd, int nuprev,
                             const double *Fx, const double *Gx, const double *Gdx,
                             const double *X0, const double *R, const double *D, const double *U0,
                             const double *Rtraj, const double *Dtraj,
                             int64_t N, int T, int warm, double *Xend, double *Ulast, double *Utraj,
                             int32_t *flag, int32_t *iters, int nthreads);

/* General closed loop (simulation.jl:37-116 with scenario.dynamics, observer.jl:1-148,203-222):
 *  - true plant: `dyn` (any function, e.g. a nonlinear model) or the affine x+ = Ft x + Gt u + Gdt d + ft;
 *  - measurement y = Cm x + Ddm d + hm + v_t;
 *  - observer state of dimension nxo >= nx (KalmanFilter on the model, or the OffsetFreeObserver's augmented
 *    estimator: the last nxo - nx entries are estimated disturbances, appended to the ndm measured ones in theta);
 *    correct! with (Co, Ddo, ho, Ko), compute_control on the estimate, predict! with (Fo, Go, Gdo, fo).
 *    Ko == NULL: no observer, the controller sees the true state (nxo == nx).
 * theta = [xhat[0:nx]; r; d_measured; xhat[nx:nxo]; uprev].  All matrices row-major. */
typedef void (*orc_dyn_fn)(void *user, int k, const double *x, const double *u, const double *d, double *xnext);
typedef struct {
  orc_dyn_fn dyn; void *user;
  const double *Ft, *Gt, *Gdt, *ft;
  int ny; const double *Cm, *Ddm, *hm, *Vtraj;
  int nxo, ndm;
  const double *Fo, *Go, *Gdo, *fo, *Co, *Ddo, *ho, *Ko;
  const double *Rtraj, *Dtraj;   /* nr x T, ndm x T shared, or NULL */
  double *Xhat_end;              /* nxo x N, nullable */
} orc_sim_stepped;
void orc_simulate_batch_stepped(const orc_problem *p, int nx, int nr, int nuprev,
                                const double *X0, const double *R, const double *D, const double *U0,
                                const orc_sim_stepped *c, int64_t N, int T, int warm,
                                double *Xend, double *Ulast, double *Utraj, int32_t *flag, int32_t *iters,
                                int nthreads);

/* explicit MPC: binary search tree walk + affine law.
 * halfplanes: n_hp x (nth+1) row-major [a; b]; feedbacks: n_fb x nu x (nth+1)
 * row-major [F | g]; node i: next = i + jump[i]; leaf iff jump[i]==0, region
 * hp_list[i]; else test a_h'theta <= b_h -> next, else next+1. clip (nullable) nu x 2. */
typedef struct {
  int nth, nu, n_hp, n_fb, n_nodes;
  const double *halfplanes, *feedbacks, *clip;
  const int32_t *hp_list, *jump_list;
} orc_bst;
void orc_bst_evaluate_batch(const orc_bst *t, const double *Theta, int64_t N, double *U,
                            int32_t *region, int nthreads);
/* closed loop with the explicit law (row-major Fx nx x nx, Gx nx x nu, Gdx nx x nd; trajectories as in
 * orc_simulate_batch_traj) */
void orc_bst_simulate_batch(const orc_bst *t, int nx, int nr, int nd, int nuprev,
                            const double *Fx, const double *Gx, const double *Gdx,
                            const double *X0, const double *R, const double *D, const double *U0,
                            const double *Rtraj, const double *Dtraj, int64_t N, int T,
                            double *Xend, double *Ulast, double *Utraj, int nthreads);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
