// Shared host/device definitions for liblmpcgpu.so.
//
// Data layout in HBM (one "constant block" per controller, uploaded once by
// lmpc_gpu_create, staged into shared memory by every CTA with one TMA bulk copy):
//
//   doubles:  Mbar   m  x n    row-major, unit rows (first ms rows: normalised rows of R^-1)
//             MbarT  n  x mp   row-major transpose, mp = m rounded up to 4 (lane-per-row scans)
//             Gram   m  x m    Mbar Mbar'
//             Dth    m  x nth  row-major (generated-C layout, /root/reference/src/codegen.jl:275)
//             du, dl m
//             Rinv   nu x n    first nu rows of the un-normalised R^-1
//             Uth    nu x nth  -(H^-1 f_theta)[1:nu,:] - [K 0]
//             uoff   nu
//   int32:    sense  m         DAQP sense bits
//
// Every section starts 16-byte aligned so the whole block can move with cp.async.bulk.
#pragma once
#include <cstdint>
#include "lmpc_gpu.h"

struct LdpLayout {
  int n, m, ms, nth, nu, mp;
  int cap;        // working-set capacity (rows) the kernels provision on chip
  int nsoft;
  int has_general; // m > ms or any soft row: singular branch reachable
  int neq;         // rows with the ACTIVE sense bit (equalities): activated at cold start
  int plain;       // box-only problem: m == ms, every sense word is 0 (no soft/equality/immutable)
  // offsets in doubles into the constant block
  int oM, oMT, oG, oDth, odu, odl, oRinv, oUth, ouoff, oSense;
  int total;      // block size in doubles (multiple of 2 -> 16 bytes)
};

struct LdpSettings {
  double primal_tol, dual_tol, zero_tol, progress_tol, fval_bound, rho_soft;
  int cycle_tol, iter_limit;
};

// theta may arrive as up to four row-segments (state / reference / disturbance / control),
// each a (len x N) column-major array: the batched mpc_update_parameter
// (/root/reference/codegen/mpc_update_parameter.c:1-29) is fused into the kernel's load.
struct ThetaSegs {
  const double *ptr[4];
  int len[4];
  int stride[4];   // doubles between consecutive instances (>= len)
  int nseg;
};

struct SolveOut {
  double *U;        // nu x N
  int32_t *flag;    // N
  int32_t *iters;   // N or null
  uint64_t *act_u;  // words x N or null
  uint64_t *act_l;
  // warm start (lmpc_gpu_solve_batch_warm): working sets to start from, words x N, or null = cold.
  // May alias act_u / act_l (every instance reads its words before it writes them).
  const uint64_t *ws_u;
  const uint64_t *ws_l;
};

// closed-loop description for the simulate kernels (x+ = F x + G u + Gd d), row-major
struct SimDesc {
  const double *F, *G, *Gd;     // device, row-major nx x nx, nx x nu, nx x nd
  const double *X0, *R, *D, *U0;
  const double *Rtraj, *Dtraj;  // device; step t reads nr / nd entries at + t * rstride / dstride: shared by all trajectories, override R / D
  int rstride, dstride;         // nr / nd for plain trajectories, ny / nd_base for sliding preview windows
  double *Xend, *Ulast, *Utraj;
  int nx, nr, nd, nuprev, T, warm;
  // extended closed loop (lmpc_gpu_simulate_batch_ex): true plant, observer, measurement noise
  int ext, ny;                            // ext != 0: the fields below are in use
  const double *Ft, *Gt, *Gdt;            // device, row-major true plant (never null when ext)
  const double *Cm, *Ddm, *Kobs;          // device, row-major ny x nx, ny x nd (nullable), nx x ny; Cm null = no observer
  const double *Vtraj;                    // device, ny x T (nullable)
  double *Xhat;                           // nx x N final estimates (nullable)
};

enum {
  EXIT_OPTIMAL = 1, EXIT_SOFT_OPTIMAL = 2, EXIT_INFEASIBLE = -1, EXIT_CYCLE = -2,
  EXIT_UNBOUNDED = -3, EXIT_ITERLIMIT = -4, EXIT_NONCONVEX = -5, EXIT_OVERDETERMINED = -6,
  EXIT_CAPACITY = -7
};
enum { S_ACTIVE = 1, S_LOWER = 2, S_IMMUTABLE = 4, S_SOFT = 8, S_BINARY = 16 };

// host-side setup result
struct LdpHost {
  LdpLayout lay;
  double *blk;   // lay.total doubles (malloc)
};

// setup.cpp
int lmpc_host_setup(const lmpc_problem *p, LdpHost *out, char *err, int errlen);

// a launcher that issues more than one kernel per call reports the extra ones here (lmpc_gpu_launch_count)
void lmpc_count_kernel(int n);
// kernel launchers (each .cu exports one); all return cudaError_t as int
struct cudaDeviceProp;
int launch_ldp_thread(const LdpLayout &lay, const LdpSettings &st, const double *cblk,
                      const ThetaSegs &th, int64_t N, const SolveOut &out, const SimDesc *sim,
                      int sm_count, void *stream);
int launch_ldp_warp(const LdpLayout &lay, const LdpSettings &st, const double *cblk,
                    const ThetaSegs &th, int64_t N, const SolveOut &out, const SimDesc *sim,
                    int sm_count, void *stream);
int launch_ldp_cta(const LdpLayout &lay, const LdpSettings &st, const double *cblk,
                   const ThetaSegs &th, int64_t N, const SolveOut &out, const SimDesc *sim,
                   int sm_count, void *stream);
bool ldp_thread_supports(const LdpLayout &lay);
bool ldp_box_supports(const LdpLayout &lay);
// hblk: host copy of the constant block (compile-time-indexed constants ride in the launch parameters)
int launch_ldp_box(const LdpLayout &lay, const LdpSettings &st, const double *cblk, const double *hblk,
                   const ThetaSegs &th, int64_t N, const SolveOut &out, int sm_count, void *stream);
bool ldp_warp_supports(const LdpLayout &lay);
bool ldp_cta_supports(const LdpLayout &lay);
long long ldp_cta_redo_count(int device);
int ldp_cta_fast_cap(const LdpLayout &lay);

struct BstDev {
  const double *blk;     // packed tree block on the device (layout: see bst_eval.cu)
  int onodes, ohp, ofb, oclip, total;        // offsets / sizes in doubles
  // fp32 screening copy of the half-planes (bst_eval.cu): [a | b] rows as 16-byte quads, replicated like hp; 0 = none.
  // The block is ordered [nodes | fb | clip | hp32 | bound constants | hp]: with the screen, only the prefix before
  // hp (`staged` doubles) is staged into shared memory, the fp64 rows stay in global memory for the rare fallback
  int ohp32, quads, oscr, staged;
  int pairs, rep, packed;                    // 16-byte pairs per row, table replicas, 32-bit node records
  int nth, nu, n_hp, n_fb, n_nodes, has_clip;
};
int bst_eval_kind(const BstDev &t);
int launch_bst_eval(const BstDev &t, const double *Theta, int64_t N, double *U, int32_t *region,
                    int sm_count, void *stream);
// closed loop with the explicit law; dyn = [F | G | Gd] row-major on the device; all pointers device
int launch_bst_sim(const BstDev &t, const double *dyn, int nx, int nr, int nd, int nup, const double *X0,
                   const double *R, const double *D, const double *U0, const double *Rtraj, const double *Dtraj,
                   int64_t N, int T, double *Xend, double *Ulast, double *Utraj, int sm_count, void *stream);
// stepped closed loop (sim_step.cu): everything on the device, matrices row-major, batches instance-contiguous
struct StepDesc {
  const double *Ft, *Gt, *Gdt, *ft;                       // affine plant (unused when the caller steps the plant)
  const double *Cm, *Ddm, *hm, *Vtraj;                    // measurement
  const double *Fo, *Go, *Gdo, *fo, *Co, *Ddo, *ho, *Ko;  // observer
  const double *Rtraj, *Dtraj, *R, *D;
  int nx, nu, nr, ndm, nxo, ny, nuprev, nth, T, has_obs, user_plant;
  double *Xs, *Xn, *Xh, *Uprev, *Theta, *U, *Dm, *Utraj;
  int32_t *flag_step, *iters_step, *flag_worst, *iters_tot;
};
bool step_dims_ok(const StepDesc &d);
int launch_step_pre(const StepDesc &d, int t, int64_t N, void *stream);
int launch_step_post(const StepDesc &d, int t, int64_t N, void *stream);
int launch_fp64_peak(int iters, double *d_out, int sm_count, void *stream);
// counter += number of flags < 1 (exit flags that the reference's `@assert exitflag >= 1` would trip on)
int launch_count_failed(const int32_t *flag, int64_t n, unsigned long long *counter, void *stream);
