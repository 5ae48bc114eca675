// K1: thread-per-QP dual active-set LDP kernel for tiny problems (n <= 8).
//
// Why a thread and not a warp per QP: with n = 5 every vector in the algorithm has <= 5
// entries, so a warp would idle 27 lanes and pay shuffle latency on every 5-term dot
// product.  Here each thread owns one QP; the working-set factorisation (packed L, D, 1/D),
// the multipliers and the primal iterate live in registers (all indices are compile-time
// after unrolling to the template capacity, with predication on the run-time working-set
// size), the constant matrices sit in shared memory where the slack scan reads them as
// warp-wide broadcasts, and the only per-thread shared-memory traffic is theta and the
// theta-dependent bound shifts (column layout [entry][thread]: conflict-free).
//
// The arithmetic (operation order, fma placement, tie-breaking) is identical to
// oracle/ldp_oracle.c, so results are comparable bit for bit.
//
// Reference path replaced: solve(mpc, theta) -> DAQP.update/solve (src/utils.jl:268-283),
// generated-C twin mpc_compute_control (codegen/mpc_update_qp.c:29-54).
#include "ldp_device.cuh"
#include <math_constants.h>

namespace {

constexpr int K1_BLOCK = 128;
// resident CTAs per SM the register allocation is capped for (measured r01: uncapped = 2 CTAs/SM,
// 8 warps, latency-bound; 3 is a win for every size, 4 for n <= 5 but spills at n = 6:
// dint_hard 2.7 / 3.5 / 3.9e9, dint_eq_small 2.19 / 2.23 / 1.63e9, invpend closed loop 5.4 / 6.4 / 6.6e9;
// n = 7, 8 keep the uncapped allocation: at 168 registers their general variants spill 1-2 KB)
constexpr int k1_minb(int n) { return n <= 5 ? 4 : (n == 6 ? 3 : 1); }

template <int N, int CAP>
struct QpState {
  double L[CAP * (CAP - 1) / 2 > 0 ? CAP * (CAP - 1) / 2 : 1];
  double D[CAP], Dinv[CAP], lam[CAP], lams[CAP], dW[CAP];
  int row[CAP];
  unsigned lowmask;   // bit i: position i is active at its LOWER bound
  unsigned immmask;   // bit i: position i is immutable (equality row)
  unsigned softmask;  // bit i: position i is a soft row
  unsigned long long inws;  // bit j: row j is in the working set
  int k;
  bool sing;
};

template <int N, int CAP>
LMPC_DEVINL void qp_reset(QpState<N, CAP> &s) {
  s.k = 0; s.sing = false; s.lowmask = 0u; s.immmask = 0u; s.softmask = 0u; s.inws = 0ull;
  SFOR(i, 0, CAP)
    s.D[i] = 1.0; s.Dinv[i] = 1.0; s.lam[i] = 0.0; s.lams[i] = 0.0; s.dW[i] = 0.0; s.row[i] = 0;
  SEND
  SFOR(i, 0, CAP * (CAP - 1) / 2) s.L[i] = 0.0; SEND
}

// append row j (oracle: ldl_add)
template <int N, int CAP>
LMPC_DEVINL void qp_add(QpState<N, CAP> &s, const ConstView &cv, int m, const LdpSettings &st,
                        int j, bool lower, double dval) {
  const double *g = cv.G + j * m;
  const int sj = cv.sense[j];
  double dnew = g[j];
  if (sj & S_SOFT) dnew += st.rho_soft;
  double y[CAP], l[CAP];
  const int k = s.k;
  SFOR(i, 0, CAP)
    y[i] = 0.0;
    if (i < k) {
      double yi = g[s.row[i]];
      SFOR(c, 0, i) yi = fma(-s.L[LIDX(i, c)], y[c], yi); SEND
      y[i] = yi;
    }
  SEND
  double pr[CAP];
  SFOR(i, 0, CAP)
    l[i] = (i < k) ? y[i] * s.Dinv[i] : 0.0;
    pr[i] = __dmul_rn(y[i], l[i]);   // never contracted into the tree's adds
  SEND
  dnew = __dsub_rn(dnew, tree_sum<CAP, 0, 1>(pr));
  const bool sing = dnew < st.zero_tol;
  if (sing) dnew = 0.0;
  const double dinv = sing ? 0.0 : 1.0 / dnew;
  SFOR(kk, 0, CAP)
    const bool here = (kk == k);
    SFOR(i, 0, kk) s.L[LIDX(kk, i)] = here ? l[i] : s.L[LIDX(kk, i)]; SEND
    s.D[kk] = here ? dnew : s.D[kk];
    s.Dinv[kk] = here ? dinv : s.Dinv[kk];
    s.lam[kk] = here ? 0.0 : s.lam[kk];
    s.dW[kk] = here ? dval : s.dW[kk];
    s.row[kk] = here ? j : s.row[kk];
  SEND
  const unsigned bit = 1u << k;
  if (lower) s.lowmask |= bit; else s.lowmask &= ~bit;
  if (sj & S_IMMUTABLE) s.immmask |= bit; else s.immmask &= ~bit;
  if (sj & S_SOFT) s.softmask |= bit; else s.softmask &= ~bit;
  s.inws |= 1ull << j;
  s.sing = sing;
  s.k = k + 1;
}

LMPC_DEVINL unsigned drop_bit(unsigned mask, int r) {
  const unsigned lo = mask & ((1u << r) - 1u);
  return lo | ((mask >> (r + 1)) << r);
}

// delete position r (oracle: ldl_remove)
template <int N, int CAP>
LMPC_DEVINL void qp_remove(QpState<N, CAP> &s, const LdpSettings &st, int r) {
  int k = s.k;
  double alpha = 0.0;
  double wv[CAP];
  int rrow = 0;
  SFOR(i, 0, CAP)
    wv[i] = 0.0;
    alpha = (i == r) ? s.D[i] : alpha;
    rrow = (i == r) ? s.row[i] : rrow;
    SFOR(c, 0, i) wv[i] = (c == r) ? s.L[LIDX(i, c)] : wv[i]; SEND
  SEND
  s.inws &= ~(1ull << rrow);
  // shift rows up / columns left; position p takes old position p+1 for p >= r
  SFOR(p, 0, CAP - 1)
    const bool mv = (p >= r && p + 1 < k);
    SFOR(c, 0, p)
      const double src = (c < r) ? s.L[LIDX(p + 1, c)] : s.L[LIDX(p + 1, c + 1)];
      s.L[LIDX(p, c)] = mv ? src : s.L[LIDX(p, c)];
    SEND
    s.D[p] = mv ? s.D[p + 1] : s.D[p];
    s.Dinv[p] = mv ? s.Dinv[p + 1] : s.Dinv[p];
    s.lam[p] = mv ? s.lam[p + 1] : s.lam[p];
    s.dW[p] = mv ? s.dW[p + 1] : s.dW[p];
    s.row[p] = mv ? s.row[p + 1] : s.row[p];
    wv[p] = mv ? wv[p + 1] : wv[p];
  SEND
  s.lowmask = drop_bit(s.lowmask, r);
  s.immmask = drop_bit(s.immmask, r);
  s.softmask = drop_bit(s.softmask, r);
  k -= 1;
  bool sing = false;
  SFOR(j, 0, CAP - 1)
    if (j >= r && j < k) {
      const double pj = wv[j];
      const double t = alpha * pj;
      double dbar = fma(t, pj, s.D[j]);
      double rinv = 1.0 / dbar;
      const double beta = t * rinv;
      alpha = (s.D[j] * alpha) * rinv;
      if (j == k - 1 && !(dbar >= st.zero_tol)) { sing = true; dbar = 0.0; rinv = 0.0; }
      s.D[j] = dbar; s.Dinv[j] = rinv;
      SFOR(i, j + 1, CAP - 1)
        if (i < k) {
          wv[i] = fma(-pj, s.L[LIDX(i, j)], wv[i]);
          s.L[LIDX(i, j)] = fma(beta, wv[i], s.L[LIDX(i, j)]);
        }
      SEND
    }
  SEND
  s.k = k;
  s.sing = sing;
}

// back substitution x <- L^-T x on the first kk positions (descending axpy order, as the oracle)
template <int N, int CAP>
LMPC_DEVINL void back_sub(const QpState<N, CAP> &s, double (&x)[CAP], int kk) {
  SFOR_DOWN(j, CAP, 1)
    if (j < kk) {
      const double xj = x[j];
      SFOR(i, 0, j) x[i] = fma(-s.L[LIDX(j, i)], xj, x[i]); SEND
    }
  SEND
}

// The dual active-set loop for one QP (oracle: ldp_loop).  shv: this thread's column of
// theta-dependent shifts sh_j (stride K1_BLOCK).  Returns the exit flag.
template <int N, int CAP, bool GENERAL>
LMPC_DEVINL int qp_loop(QpState<N, CAP> &s, const ConstView &cv, int m, const LdpSettings &st,
                        const double *shv, double (&u)[N], int &iters) {
  int flag = EXIT_ITERLIMIT, iter, cycle = 0;
  double best_fval = -1.0, fval = 0.0, soft_slack = 0.0;
  for (iter = 1; iter < st.iter_limit; ++iter) {
    const int k = s.k;
    double alpha = CUDART_INF;
    int rm = -1;
    if (!GENERAL || !s.sing) {
      // 1. CSP: L D L' lam* = -d_W
      double z[CAP];
      SFOR(i, 0, CAP)
        z[i] = 0.0;
        if (i < k) {
          double zi = -s.dW[i];
          SFOR(c, 0, i) zi = fma(-s.L[LIDX(i, c)], z[c], zi); SEND
          z[i] = zi;
        }
      SEND
      SFOR(i, 0, CAP) s.lams[i] = (i < k) ? z[i] * s.Dinv[i] : 0.0; SEND
      back_sub<N, CAP>(s, s.lams, k);
      // 2. blocking set + ratio test (first minimum wins)
      SFOR(i, 0, CAP)
        if (i < k && !((s.immmask >> i) & 1u)) {
          const bool low = (s.lowmask >> i) & 1u;
          const bool blk = low ? (s.lams[i] > st.dual_tol) : (s.lams[i] < -st.dual_tol);
          if (blk) {
            const double a = -s.lam[i] / (s.lams[i] - s.lam[i]);
            if (a < alpha) { alpha = a; rm = i; }
          }
        }
      SEND
      if (rm < 0) {
        // 3. primal iterate, dual objective, most violated row
        SFOR(c, 0, N) u[c] = 0.0; SEND
        SFOR(i, 0, CAP)
          if (i < k) {
            const double *mr = cv.M + s.row[i] * N;
            const double li = s.lams[i];
            SFOR(c, 0, N) u[c] = fma(-li, mr[c], u[c]); SEND
          }
        SEND
        soft_slack = 0.0;
        {
          double pu[N];
          SFOR(c, 0, N) pu[c] = __dmul_rn(u[c], u[c]); SEND
          fval = tree_sum<N, 0, 1>(pu);
        }
        if (GENERAL) {
          double ps[CAP];
          SFOR(i, 0, CAP)
            const bool sf = (i < k) && ((s.softmask >> i) & 1u);
            ps[i] = sf ? __dmul_rn(s.lams[i], s.lams[i]) : 0.0;
            const double sl = st.rho_soft * fabs(s.lams[i]);
            if (sf && sl > soft_slack) soft_slack = sl;
          SEND
          fval = fma(st.rho_soft, tree_sum<CAP, 0, 1>(ps), fval);
        }
        if (fval > st.fval_bound) { flag = EXIT_INFEASIBLE; break; }
        double best = -st.primal_tol, addd = 0.0;
        int add = -1;
        bool addlow = false;
        for (int j = 0; j < m; ++j) {
          const double *mr = cv.M + j * N;   // warp-uniform address: shared-memory broadcast
          double mu = 0.0;
          SFOR(c, 0, N) mu = fma(mr[c], u[c], mu); SEND
          const bool skip = ((s.inws >> j) & 1ull) || (cv.sense[j] & S_IMMUTABLE);
          const double shj = shv[j * K1_BLOCK];
          const double dU = cv.du[j] + shj;
          const double su = dU - mu;
          if (!skip) {
            if (su < best) { best = su; add = j; addlow = false; addd = dU; }
            else {
              const double dLo = cv.dl[j] + shj;
              const double sl = mu - dLo;
              if (sl < best) { best = sl; add = j; addlow = true; addd = dLo; }
            }
          }
        }
        if (add < 0) { flag = soft_slack > st.primal_tol ? EXIT_SOFT_OPTIMAL : EXIT_OPTIMAL; break; }
        // cycle guard where the dual objective is fresh (see oracle/ldp_oracle.c)
        if (fval - st.progress_tol > best_fval) { best_fval = fval; cycle = 0; }
        else if (cycle++ > st.cycle_tol) { flag = EXIT_CYCLE; break; }
        if (k >= CAP) { flag = EXIT_CAPACITY; break; }
        SFOR(i, 0, CAP) s.lam[i] = (i < k) ? s.lams[i] : s.lam[i]; SEND
        qp_add<N, CAP>(s, cv, m, st, add, addlow, addd);
        continue;
      }
      // 4. blocking: step, remove
      if (!(alpha < CUDART_INF)) { flag = EXIT_CYCLE; break; }   // blocked but no finite ratio
      SFOR(i, 0, CAP)
        s.lam[i] = (i < k) ? fma(alpha, s.lams[i] - s.lam[i], s.lam[i]) : s.lam[i];
      SEND
      qp_remove<N, CAP>(s, st, rm);
    } else {
      // 5. singular: p = side * [-L^-T l ; 1] on the first k-1 positions
      const int kk = k - 1;
      const bool lastlow = (s.lowmask >> kk) & 1u;
      SFOR(i, 0, CAP) s.lams[i] = 0.0; SEND
      SFOR(i, 0, CAP)
        SFOR(c, 0, i) s.lams[c] = (i == kk) ? -s.L[LIDX(i, c)] : s.lams[c]; SEND
      SEND
      back_sub<N, CAP>(s, s.lams, kk);
      SFOR(i, 0, CAP)
        s.lams[i] = (i == kk) ? 1.0 : s.lams[i];
        s.lams[i] = (lastlow && i <= kk) ? -s.lams[i] : s.lams[i];
      SEND
      SFOR(i, 0, CAP)
        if (i < k && !((s.immmask >> i) & 1u)) {
          const bool low = (s.lowmask >> i) & 1u;
          const bool blk = low ? (s.lams[i] > st.dual_tol) : (s.lams[i] < -st.dual_tol);
          if (blk) {
            const double a = -s.lam[i] / s.lams[i];
            if (a < alpha) { alpha = a; rm = i; }
          }
        }
      SEND
      if (rm < 0) { flag = EXIT_INFEASIBLE; break; }
      SFOR(i, 0, CAP) s.lam[i] = (i < k) ? fma(alpha, s.lams[i], s.lam[i]) : s.lam[i]; SEND
      qp_remove<N, CAP>(s, st, rm);
    }
  }
  iters = iter;
  return flag;
}

// One solve given theta in this thread's shared-memory column thv (stride K1_BLOCK).
// Computes the bound shifts, runs the loop, writes nu controls to ctl[] (stride cstride).
template <int N, int CAP, bool GENERAL>
LMPC_DEVINL int qp_solve(QpState<N, CAP> &s, const ConstView &cv, const LdpLayout &lay,
                         const LdpSettings &st, const double *thv, double *shv, bool cold,
                         double (&u)[N], int &iters, bool warm = false,
                         unsigned long long wsu = 0ull, unsigned long long wsl = 0ull) {
  const int m = lay.m, nth = lay.nth;
  // mpc_update_qp: sh_j = Dth[j,:] theta
  for (int j = 0; j < m; ++j) {
    const double *dr = cv.Dth + j * nth;
    double acc = 0.0;
    for (int t = 0; t < nth; ++t) acc = fma(dr[t], thv[t * K1_BLOCK], acc);
    shv[j * K1_BLOCK] = acc;
  }
  int flag = 0;
  if (cold) {
    qp_reset<N, CAP>(s);
    SFOR(c, 0, N) u[c] = 0.0; SEND
    for (int j = 0; lay.neq > 0 && j < m; ++j) {
      const int sj = cv.sense[j];
      if (sj & S_ACTIVE) {
        if (s.k >= CAP) { flag = EXIT_OVERDETERMINED; break; }
        const bool low = sj & S_LOWER;
        const double dv = (low ? cv.dl[j] : cv.du[j]) + shv[j * K1_BLOCK];
        qp_add<N, CAP>(s, cv, m, st, j, low, dv);
        if (s.sing) { flag = EXIT_OVERDETERMINED; break; }
      }
    }
    // warm start from a caller-supplied working set (oracle: activate_rows)
    for (int j = 0; warm && flag == 0 && j < m; ++j) {
      const int sj = cv.sense[j];
      const bool up = (wsu >> j) & 1ull, lo = (wsl >> j) & 1ull;
      if ((sj & (S_ACTIVE | S_IMMUTABLE)) || !(up || lo) || s.k >= CAP) continue;
      qp_add<N, CAP>(s, cv, m, st, j, !up, (up ? cv.du[j] : cv.dl[j]) + shv[j * K1_BLOCK]);
      if (s.sing) { s.k -= 1; s.inws &= ~(1ull << j); s.sing = false; }   // undo the trial add
    }
  } else {
    // warm start: same working set, new right-hand sides
    SFOR(i, 0, CAP)
      if (i < s.k) {
        const int j = s.row[i];
        s.dW[i] = (((s.lowmask >> i) & 1u) ? cv.dl[j] : cv.du[j]) + shv[j * K1_BLOCK];
      }
    SEND
  }
  iters = 0;
  if (flag == 0) flag = qp_loop<N, CAP, GENERAL>(s, cv, m, st, shv, u, iters);
  return flag;
}

// mpc_get_solution: control_i = (R^-1 u)_i + uoff_i + Uth[i,:] theta
template <int N>
LMPC_DEVINL double control_out(const ConstView &cv, int nth, int i, const double (&u)[N],
                               const double *thv) {
  const double *rr = cv.Rinv + i * N;
  double pr[N];
  SFOR(c, 0, N) pr[c] = __dmul_rn(rr[c], u[c]); SEND
  const double x = tree_sum<N, 0, 1>(pr);
  double sh = cv.uoff[i];
  const double *ur = cv.Uth + i * nth;
  for (int t = 0; t < nth; ++t) sh = fma(ur[t], thv[t * K1_BLOCK], sh);
  return __dadd_rn(x, sh);
}

template <int N, int CAP, bool GENERAL>
__global__ void __launch_bounds__(K1_BLOCK, k1_minb(N))
ldp_thread_kernel(LdpLayout lay, LdpSettings st, const double *__restrict__ cblk, ThetaSegs th,
                  int64_t Ninst, SolveOut out) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar;
  stage_constants_tma(smem, cblk, lay.total, &mbar);
  const ConstView cv = make_view(smem, lay);
  double *thv = smem + lay.total + threadIdx.x;            // [nth][K1_BLOCK]
  double *shv = thv + (size_t)lay.nth * K1_BLOCK;           // [m][K1_BLOCK]
  const int words = (lay.m + 63) >> 6;
  QpState<N, CAP> s;
  double u[N];
  for (int64_t q = (int64_t)blockIdx.x * K1_BLOCK + threadIdx.x; q < Ninst;
       q += (int64_t)gridDim.x * K1_BLOCK) {
    // batched mpc_update_parameter: gather theta from its row segments
    int o = 0;
    SFOR(sgi, 0, 4)
      if (sgi < th.nseg) {
        const int len = th.len[sgi];
        const double *src = th.ptr[sgi] + q * th.stride[sgi];
        for (int t = 0; t < len; ++t) thv[(o + t) * K1_BLOCK] = ld_stream(src + t);
        o += len;
      }
    SEND
    int iters;
    const bool warm = out.ws_u != nullptr;
    const int flag = qp_solve<N, CAP, GENERAL>(s, cv, lay, st, thv, shv, true, u, iters, warm,
                                               warm ? out.ws_u[q * words] : 0ull, warm ? out.ws_l[q * words] : 0ull);
    for (int i = 0; i < lay.nu; ++i) out.U[q * lay.nu + i] = control_out<N>(cv, lay.nth, i, u, thv);
    out.flag[q] = flag;
    if (out.iters) out.iters[q] = iters;
    if (out.act_u) {
      unsigned long long au = 0ull, al = 0ull;
      SFOR(i, 0, CAP)
        if (i < s.k) {
          if ((s.lowmask >> i) & 1u) al |= 1ull << s.row[i]; else au |= 1ull << s.row[i];
        }
      SEND
      out.act_u[q * words] = au;
      out.act_l[q * words] = al;
    }
  }
}

// closed loop: T steps of (theta = [x; r; d; u_prev]) -> solve -> x+ = F x + G u + Gd d
template <int N, int CAP, bool GENERAL>
__global__ void __launch_bounds__(K1_BLOCK, k1_minb(N))
ldp_thread_sim_kernel(LdpLayout lay, LdpSettings st, const double *__restrict__ cblk,
                      int64_t Ninst, SimDesc sd, int32_t *flagout, int32_t *itersout) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar;
  stage_constants_tma(smem, cblk, lay.total, &mbar);
  const ConstView cv = make_view(smem, lay);
  double *thv = smem + lay.total + threadIdx.x;            // [nth][K1_BLOCK]
  double *shv = thv + (size_t)lay.nth * K1_BLOCK;           // [m][K1_BLOCK]
  double *xnv = shv + (size_t)lay.m * K1_BLOCK;             // [nx][K1_BLOCK]
  double *ucv = xnv + (size_t)sd.nx * K1_BLOCK;             // [nu][K1_BLOCK]
  double *xsv = ucv + (size_t)lay.nu * K1_BLOCK;            // ext: [nx][K1_BLOCK] true plant state
  double *ev = xsv + (size_t)sd.nx * K1_BLOCK;              // ext: [ny][K1_BLOCK] innovation
  const int nx = sd.nx, nr = sd.nr, nd = sd.nd, nup = sd.nuprev, nu = lay.nu;
  const bool ext = sd.ext != 0;
  const int ny = ext ? sd.ny : 0;
  QpState<N, CAP> s;
  double u[N];
  for (int64_t q = (int64_t)blockIdx.x * K1_BLOCK + threadIdx.x; q < Ninst;
       q += (int64_t)gridDim.x * K1_BLOCK) {
    for (int i = 0; i < nx; ++i) { thv[i * K1_BLOCK] = ld_stream(sd.X0 + q * nx + i); if (ext) xsv[i * K1_BLOCK] = thv[i * K1_BLOCK]; }
    for (int i = 0; i < nr; ++i) thv[(nx + i) * K1_BLOCK] = sd.R ? ld_stream(sd.R + q * nr + i) : 0.0;
    for (int i = 0; i < nd; ++i) thv[(nx + nr + i) * K1_BLOCK] = sd.D ? ld_stream(sd.D + q * nd + i) : 0.0;
    for (int i = 0; i < nu; ++i) ucv[i * K1_BLOCK] = (sd.U0 && i < nup) ? ld_stream(sd.U0 + q * nup + i) : 0.0;
    int worst = EXIT_OPTIMAL, tot = 0;
    for (int t = 0; t < sd.T; ++t) {
      for (int i = 0; i < nup; ++i) thv[(nx + nr + nd + i) * K1_BLOCK] = ucv[i * K1_BLOCK];
      // reference / disturbance of step t (rs[:, k], ds[:, k] of src/simulation.jl:93-104)
      if (sd.Rtraj) for (int i = 0; i < nr; ++i) thv[(nx + i) * K1_BLOCK] = sd.Rtraj[(size_t)t * sd.rstride + i];
      if (sd.Dtraj) for (int i = 0; i < nd; ++i) thv[(nx + nr + i) * K1_BLOCK] = sd.Dtraj[(size_t)t * sd.dstride + i];
      if (ext) {
        const double *dk = thv + (size_t)(nx + nr) * K1_BLOCK;
        if (sd.Cm) {
          // y = C x + Dd d + v, innovation against the estimate, correction (observer.jl:114-118)
          for (int i = 0; i < ny; ++i) {
            double y = sd.Vtraj ? sd.Vtraj[(size_t)t * ny + i] : 0.0, yh = 0.0;
            for (int c = 0; c < nx; ++c) y = fma(sd.Cm[i * nx + c], xsv[c * K1_BLOCK], y);
            if (sd.Ddm) for (int c = 0; c < nd; ++c) y = fma(sd.Ddm[i * nd + c], dk[c * K1_BLOCK], y);
            for (int c = 0; c < nx; ++c) yh = fma(sd.Cm[i * nx + c], thv[c * K1_BLOCK], yh);
            if (sd.Ddm) for (int c = 0; c < nd; ++c) yh = fma(sd.Ddm[i * nd + c], dk[c * K1_BLOCK], yh);
            ev[i * K1_BLOCK] = __dsub_rn(y, yh);
          }
          for (int j = 0; j < nx; ++j) {
            double acc = thv[j * K1_BLOCK];
            for (int i = 0; i < ny; ++i) acc = fma(sd.Kobs[j * ny + i], ev[i * K1_BLOCK], acc);
            xnv[j * K1_BLOCK] = acc;
          }
          for (int j = 0; j < nx; ++j) thv[j * K1_BLOCK] = xnv[j * K1_BLOCK];
        } else {
          for (int j = 0; j < nx; ++j) thv[j * K1_BLOCK] = xsv[j * K1_BLOCK];   // the controller measures the true state
        }
      }
      int iters;
      const int f = qp_solve<N, CAP, GENERAL>(s, cv, lay, st, thv, shv, !(sd.warm && t > 0), u, iters);
      tot += iters;
      if (f < 0) { if (worst > 0 || f < worst) worst = f; }
      else if (worst > 0 && f > worst) worst = f;
      for (int i = 0; i < nu; ++i) {
        const double c = control_out<N>(cv, lay.nth, i, u, thv);
        ucv[i * K1_BLOCK] = c;
        if (sd.Utraj) sd.Utraj[(q * sd.T + t) * nu + i] = c;
      }
      for (int i = 0; i < nx; ++i) {
        double acc = 0.0;
        for (int c = 0; c < nx; ++c) acc = fma(sd.F[i * nx + c], thv[c * K1_BLOCK], acc);
        for (int c = 0; c < nu; ++c) acc = fma(sd.G[i * nu + c], ucv[c * K1_BLOCK], acc);
        for (int c = 0; c < nd; ++c) acc = fma(sd.Gd[i * nd + c], thv[(nx + nr + c) * K1_BLOCK], acc);
        xnv[i * K1_BLOCK] = acc;
      }
      for (int i = 0; i < nx; ++i) thv[i * K1_BLOCK] = xnv[i * K1_BLOCK];
      if (ext) {   // true plant step
        for (int i = 0; i < nx; ++i) {
          double acc = 0.0;
          for (int c = 0; c < nx; ++c) acc = fma(sd.Ft[i * nx + c], xsv[c * K1_BLOCK], acc);
          for (int c = 0; c < nu; ++c) acc = fma(sd.Gt[i * nu + c], ucv[c * K1_BLOCK], acc);
          for (int c = 0; c < nd; ++c) acc = fma(sd.Gdt[i * nd + c], thv[(nx + nr + c) * K1_BLOCK], acc);
          xnv[i * K1_BLOCK] = acc;
        }
        for (int i = 0; i < nx; ++i) xsv[i * K1_BLOCK] = xnv[i * K1_BLOCK];
      }
    }
    if (ext && sd.Xhat) for (int i = 0; i < nx; ++i) sd.Xhat[q * nx + i] = thv[i * K1_BLOCK];
    for (int i = 0; i < nx; ++i) sd.Xend[q * nx + i] = ext ? xsv[i * K1_BLOCK] : thv[i * K1_BLOCK];
    for (int i = 0; i < nu; ++i) sd.Ulast[q * nu + i] = ucv[i * K1_BLOCK];
    flagout[q] = worst;
    if (itersout) itersout[q] = tot;
  }
}

template <int N, int CAP, bool GENERAL>
int launch_inst(const LdpLayout &lay, const LdpSettings &st, const double *cblk, const ThetaSegs &th,
                int64_t Ninst, const SolveOut &out, const SimDesc *sim, int sm_count,
                cudaStream_t stream) {
  size_t smem = ((size_t)lay.total + (size_t)(lay.nth + lay.m) * K1_BLOCK) * sizeof(double);
  if (sim) smem += (size_t)(sim->nx + lay.nu + (sim->ext ? sim->nx + sim->ny : 0)) * K1_BLOCK * sizeof(double);
  cudaError_t e;
  int64_t blocks = (Ninst + K1_BLOCK - 1) / K1_BLOCK;
  int per_sm = 0;
  if (sim) {
    auto kern = ldp_thread_sim_kernel<N, CAP, GENERAL>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K1_BLOCK, smem);
    if (per_sm < 1) per_sm = 1;
    const int64_t cap = (int64_t)sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, K1_BLOCK, smem, stream>>>(lay, st, cblk, Ninst, *sim, out.flag, out.iters);
  } else {
    auto kern = ldp_thread_kernel<N, CAP, GENERAL>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K1_BLOCK, smem);
    if (per_sm < 1) per_sm = 1;
    // persistent grid: a whole number of resident waves over the 148 SMs
    const int64_t cap = (int64_t)sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, K1_BLOCK, smem, stream>>>(lay, st, cblk, th, Ninst, out);
  }
  return (int)cudaGetLastError();
}

}  // namespace

bool ldp_thread_supports(const LdpLayout &lay) {
  if (lay.n < 1 || lay.n > 8 || lay.m > 64) return false;
  if (!lay.has_general) return lay.cap <= lay.n;
  return lay.cap <= lay.n + 1;   // hard general rows only (soft rows need more capacity)
}

#define K1_CASE(NN)                                                                         \
  case NN:                                                                                  \
    return lay.has_general                                                                  \
               ? launch_inst<NN, NN + 1, true>(lay, st, cblk, th, N, out, sim, sm_count, s) \
               : launch_inst<NN, NN, false>(lay, st, cblk, th, N, out, sim, sm_count, s);

int launch_ldp_thread(const LdpLayout &lay, const LdpSettings &st, const double *cblk,
                      const ThetaSegs &th, int64_t N, const SolveOut &out, const SimDesc *sim,
                      int sm_count, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (N <= 0) return 0;
  switch (lay.n) {
    K1_CASE(1) K1_CASE(2) K1_CASE(3) K1_CASE(4) K1_CASE(5) K1_CASE(6) K1_CASE(7) K1_CASE(8)
    default: return (int)cudaErrorInvalidValue;
  }
}
