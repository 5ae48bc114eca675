// K2: warp-per-QP dual active-set LDP kernel for 8 < n <= 32 (working-set capacity <= 32,
// m <= 64).  The configuration north_star names for the AFTI-16 controller (n = 20, m = 22).
//
// Mapping.  One warp owns one QP at a time and loops over its share of the batch.
//   * working-set position i  -> lane i : D_i, 1/D_i, lambda_i, lambda*_i, d_W,i, row id
//   * decision variable c     -> lane c : u_c
//   * constraint row j        -> lane j mod 32 (two rows per lane when m > 32): shift sh_j
// The unit-lower factor L lives in the warp's shared-memory tile as a square array indexed by
// POSITION (odd leading dimension: column reads and row reads are both conflict free), so a
// sweep step addresses it with the loop counter alone.  Removing a row shifts the lanes above
// it down by one (__shfl_down); the trailing block of L is rewritten by the rank-1 update
// anyway, which stores it one row / column up, and only the r columns left of the removed
// position are moved explicitly.  (r01 used a position -> physical-slot indirection instead:
// no data moved on removal, but every sweep step paid a shuffle and an address multiply for it.)
// Triangular solves are column sweeps: one shuffle broadcast + one fma per step.  Slack argmin, ratio-test argmin and the three cross-lane
// sums (pivot, dual objective, output map) are warp reductions; the sums use the canonical
// butterfly order shared with the oracle, so results stay bit-identical to
// oracle/ldp_oracle.c and to the thread-per-QP kernel.
//
// The constant block (Mbar, Mbar', Gram, Dth, du, dl, R^-1 rows, Uth, uoff, senses) is staged
// into shared memory once per CTA by a TMA bulk copy and shared by its warps.
//
// Reference path replaced: solve(mpc, theta) -> DAQP.update/solve (src/utils.jl:268-283),
// Simulation loop (src/simulation.jl:93-113).
#include "ldp_device.cuh"
#include <math_constants.h>

namespace {

constexpr int K2_WARPS = 8;
constexpr int K2_BLOCK = K2_WARPS * 32;
#ifndef K2_MINB
#define K2_MINB 3
#endif
// Unroll factor of the sweep / gather loops, per kernel (measured on AFTI-16, r01): single solves
// run ~40 iterations per QP and like 4 (1.36e7 vs 1.31e7 / 1.19e7 solves/s at 2 / 1); the closed
// loop runs ~2 iterations per warm-started solve, is instruction-fetch bound, and likes 2
// (3.35e8 vs 2.95e8 solves/s at 4).
constexpr int K2_UF_SOLVE = 4;
constexpr int K2_UF_SIM = 2;
constexpr unsigned FULL = 0xffffffffu;
// internal: this instance needs the factorisation repair (wq_rebuild); the fast kernel leaves it to the second,
// repair-capable launch (never seen by the caller)
constexpr int EXIT_REDO = -100;

LMPC_DEVINL double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }
LMPC_DEVINL double shfl_down_d(double v, int d) { return __shfl_down_sync(FULL, v, d); }

// canonical lane sum for len <= 32: one element per lane, xor-butterfly 16,8,4,2,1
LMPC_DEVINL double bfly_sum(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(FULL, v, off));
  return v;
}
LMPC_DEVINL double bfly_max(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const double o = __shfl_xor_sync(FULL, v, off);
    v = (o > v) ? o : v;
  }
  return v;
}
// Lanes holding the warp-wide minimum of v (NaN-free), as a ballot mask.  The minimum of a set
// does not depend on how it is found, so this uses two 32-bit REDUX operations on an
// order-preserving integer image of the double instead of a 5-step shuffle butterfly
// (-0.0 is folded onto +0.0 first so that equal values share one key).
LMPC_DEVINL unsigned min_lanes(double v) {
  const long long b = __double_as_longlong(__dadd_rn(v, 0.0));
  const unsigned long long key = (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned hmin = __reduce_min_sync(FULL, hi);
  const unsigned lmin = __reduce_min_sync(FULL, hi == hmin ? lo : 0xffffffffu);
  return __ballot_sync(FULL, hi == hmin && lo == lmin);
}

struct WarpQp {
  // per lane (position = lane)
  double D, Dinv, lam, lams, dW;
  int row;
  bool low, imm, soft;
  // warp-uniform
  int k;
  bool sing;
  int rc;                     // exit flag raised by wq_rebuild (0: none)
  unsigned long long inws;    // rows in the working set
};

struct WarpCtx {
  ConstView cv;
  const LdpLayout *lay;
  const LdpSettings *st;
  double *Lw;      // cap x LS tile, indexed [position_row * LS + position_col]
  double *uw;      // n: primal iterate, broadcast source for the slack scan
  double *thw;     // nth: theta
  int LS;
  int lane;
};

LMPC_DEVINL void wq_reset(WarpQp &s) {
  s.D = 1.0; s.Dinv = 1.0; s.lam = 0.0; s.lams = 0.0; s.dW = 0.0; s.row = 0;
  s.low = false; s.imm = false; s.soft = false;
  s.k = 0; s.sing = false; s.rc = 0; s.inws = 0ull;
}

// Invariant of the tile: row i < k holds L[i][j] for j < i and exact ZEROS for i <= j < LS (the
// add writes the whole row, the remove only ever rewrites strictly-lower entries).  A sweep step
// is then an unpredicated fma on every lane: lanes at or above the diagonal multiply by zero and
// keep their value bit for bit; lanes >= kk compute garbage nobody reads (callers mask).
//
// column sweep: z <- L^-1 z on positions [0, kk)
template <int UF>
LMPC_DEVINL double sweep_fwd(const WarpCtx &c, const WarpQp &s, double z, int kk) {
  const double *lrow = c.Lw + (c.lane < kk ? c.lane : 0) * c.LS;   // idle lanes re-read row 0 (in bounds)
#pragma unroll UF
  for (int j = 0; j + 1 < kk; ++j) z = fma(-lrow[j], shfl_d(z, j), z);
  return z;
}
// column sweep: x <- L^-T x on positions [0, kk), descending (the oracle's axpy order)
template <int UF>
LMPC_DEVINL double sweep_bwd(const WarpCtx &c, const WarpQp &s, double x, int kk) {
  const double *lcol = c.Lw + (c.lane < c.LS ? c.lane : c.LS - 1);
#pragma unroll UF
  for (int j = kk - 1; j > 0; --j) x = fma(-lcol[j * c.LS], shfl_d(x, j), x);
  return x;
}

// append row j (oracle: ldl_add)
template <int UF>
LMPC_DEVINL void wq_add(const WarpCtx &c, WarpQp &s, int j, bool lower, double dval) {
  const int m = c.lay->m, k = s.k, lane = c.lane;
  const double *g = c.cv.G + j * m;
  const int sj = c.cv.sense[j];
  double dnew = g[j];
  if (sj & S_SOFT) dnew += c.st->rho_soft;
  double y = (lane < k) ? g[s.row] : 0.0;
  y = sweep_fwd<UF>(c, s, y, k);
  y = (lane < k) ? y : 0.0;                  // lanes >= k carried garbage through the sweep
  const double l = y * s.Dinv;
  dnew = __dsub_rn(dnew, bfly_sum(__dmul_rn(y, l)));
  const bool sing = dnew < c.st->zero_tol;
  if (sing) dnew = 0.0;
  const double dinv = sing ? 0.0 : 1.0 / dnew;
  __syncwarp();                              // idle lanes may still be reading row k in the sweep
  if (lane < c.LS) c.Lw[k * c.LS + lane] = l;   // the whole row: l on [0, k), zeros from the diagonal on
  if (lane == k) {
    s.D = dnew; s.Dinv = dinv; s.lam = 0.0; s.dW = dval; s.row = j;
    s.low = lower; s.imm = (sj & S_IMMUTABLE) != 0; s.soft = (sj & S_SOFT) != 0;
  }
  s.inws |= 1ull << j;
  s.sing = sing;
  s.k = k + 1;
  __syncwarp();
}

// delete position r (oracle: ldl_remove): positions above r move down a lane, L moves up a row
// and (right of r) a column while the rank-1 update rewrites it
LMPC_DEVINL void wq_remove(const WarpCtx &c, WarpQp &s, int r) {
  const int lane = c.lane, LS = c.LS;
  int k = s.k;
  double alpha = shfl_d(s.D, r);
  const int rrow = __shfl_sync(FULL, s.row, r);
  // column r below the diagonal, still at old positions
  double wv = (lane > r && lane < k) ? c.Lw[lane * LS + r] : 0.0;
  s.inws &= ~(1ull << rrow);
  // positions p >= r take position p+1
  {
    const double D = shfl_down_d(s.D, 1), Dinv = shfl_down_d(s.Dinv, 1), lam = shfl_down_d(s.lam, 1);
    const double dW = shfl_down_d(s.dW, 1), w2 = shfl_down_d(wv, 1);
    const int row = __shfl_down_sync(FULL, s.row, 1);
    const unsigned fl = __shfl_down_sync(FULL, (unsigned)s.low | ((unsigned)s.imm << 1) | ((unsigned)s.soft << 2), 1);
    if (lane >= r && lane + 1 < k) {
      s.D = D; s.Dinv = Dinv; s.lam = lam; s.dW = dW; s.row = row; wv = w2;
      s.low = fl & 1u; s.imm = (fl >> 1) & 1u; s.soft = (fl >> 2) & 1u;
    }
  }
  k -= 1;
  const bool below = lane >= r && lane < k;    // new positions whose row came from old row lane + 1
  // columns left of r: same values, one row up (all lanes read, then all lanes write)
#pragma unroll 1
  for (int cc = 0; cc < r; ++cc) {
    const double v = below ? c.Lw[(lane + 1) * LS + cc] : 0.0;
    __syncwarp();
    if (below) c.Lw[lane * LS + cc] = v;
  }
  bool sing = false;
#pragma unroll 1
  for (int j = r; j < k; ++j) {
    const double pj = shfl_d(wv, j);
    const double Dj = shfl_d(s.D, j);
    const double t = alpha * pj;
    double dbar = fma(t, pj, Dj);
    double rinv = 1.0 / dbar;
    const double beta = t * rinv;
    alpha = (Dj * alpha) * rinv;
    if (j == k - 1 && !(dbar >= c.st->zero_tol)) { sing = true; dbar = 0.0; rinv = 0.0; }
    if (lane == j) { s.D = dbar; s.Dinv = rinv; }
    // new L[i][j] from old L[i+1][j+1]; old column j was consumed one step ago
    const bool upd = lane > j && lane < k;
    const double lij = upd ? c.Lw[(lane + 1) * LS + j + 1] : 0.0;
    __syncwarp();
    if (upd) {
      wv = fma(-pj, lij, wv);
      c.Lw[lane * LS + j] = fma(beta, wv, lij);
    }
  }
  s.k = k;
  s.sing = sing;
  __syncwarp();
}

// lowest lane holding the minimum of a (NaN-free), or -1 when every lane is +inf
LMPC_DEVINL int argmin_first(double a, double &amin) {
  const int lane = __ffs(min_lanes(a)) - 1;
  amin = shfl_d(a, lane);                    // the first minimiser's own value, as the oracle keeps it
  return (amin < CUDART_INF) ? lane : -1;
}

// Problems with SOFT rows only (oracle: rebuild): refactorise from scratch with the hard rows first, in
// their current order, and the soft rows last, multipliers kept.  Taken when a pivot fell below
// zero_tol while a soft row (pivot contribution rho = 1e-6) is in the working set, and once per
// solve when the cycle guard fires.  A hard row that is dependent on the hard block itself goes to
// the very end with a zero pivot (the genuinely singular case); a second one: EXIT_CYCLE in s.rc.
// Rare and long: kept out of line, state passed by value so the caller's copy stays in registers.
template <int UF>
__device__ __noinline__ WarpQp wq_rebuild(const WarpCtx c, WarpQp s) {
  const int lane = c.lane, k = s.k;
  const int row0 = s.row;
  const bool low0 = s.low;
  const double lam0 = s.lam, dW0 = s.dW;
  const unsigned hard = __ballot_sync(FULL, lane < k && !s.soft), softm = __ballot_sync(FULL, lane < k && s.soft);
  const int nh = __popc(hard);
  s.k = 0; s.sing = false; s.rc = 0;
  int defer = -1;
  for (int t = 0; t < k; ++t) {
    const int src = (t < nh) ? (int)__fns(hard, 0, t + 1) : (int)__fns(softm, 0, t - nh + 1);
    const int j = __shfl_sync(FULL, row0, src);
    const bool lo = __shfl_sync(FULL, (int)low0, src);
    const double dv = shfl_d(dW0, src), lm = shfl_d(lam0, src);
    wq_add<UF>(c, s, j, lo, dv);
    if (lane == s.k - 1) s.lam = lm;
    if (s.sing) {
      if (t >= nh || defer >= 0) { s.rc = EXIT_CYCLE; return s; }
      defer = src; s.k -= 1; s.sing = false;
    }
  }
  if (defer >= 0) {
    const int j = __shfl_sync(FULL, row0, defer);
    const bool lo = __shfl_sync(FULL, (int)low0, defer);
    const double dv = shfl_d(dW0, defer), lm = shfl_d(lam0, defer);
    wq_add<UF>(c, s, j, lo, dv);
    if (lane == s.k - 1) { s.lam = lm; s.D = 0.0; s.Dinv = 0.0; }
    s.sing = true;
  }
  return s;
}
// a soft row is in the working set: a pivot below zero_tol is then suspect (see wq_rebuild)
#ifndef K2_REPAIR
#define K2_REPAIR 1
#endif
LMPC_DEVINL bool wq_has_soft(const WarpCtx &c, const WarpQp &s) {
  return K2_REPAIR && c.lay->nsoft > 0 && __any_sync(FULL, c.lane < s.k && s.soft);
}

// The dual active-set loop for one QP (oracle: ldp_loop).  sh0/sh1: shifts of rows lane, lane+32.
// REPAIR = false: the hot variant.  Where the repair would run (problems with soft rows only: a pivot below
// zero_tol next to a soft row, or the cycle guard's second chance) it stops with EXIT_REDO instead, and the
// instance is solved again from scratch by the REPAIR = true kernel launched right after (r02: keeping the
// out-of-line rebuild call in the one kernel cost every solve 6-9 % in spills around the call site).
template <int UF, bool REPAIR>
LMPC_DEVINL int wq_loop(const WarpCtx &c, WarpQp &s, double sh0, double sh1, double &u, int &iters) {
  const LdpLayout &lay = *c.lay;
  const LdpSettings &st = *c.st;
  const int lane = c.lane, n = lay.n, m = lay.m, mp = lay.mp;
  int flag = EXIT_ITERLIMIT, iter, cycle = 0;
  bool repaired = false;
  double best_fval = -1.0, fval = 0.0, soft_slack = 0.0;
  for (iter = 1; iter < st.iter_limit; ++iter) {
    const int k = s.k;
    double a = CUDART_INF, dir = 0.0;
    int noblock = EXIT_CYCLE;
    if (!s.sing) {
      // 1. CSP: L D L' lam* = -d_W
      double z = (lane < k) ? -s.dW : 0.0;
      z = sweep_fwd<UF>(c, s, z, k);
      s.lams = (lane < k) ? z * s.Dinv : 0.0;
      s.lams = sweep_bwd<UF>(c, s, s.lams, k);
      // 2. blocking set + ratio test
      const bool blk = lane < k && !s.imm && (s.low ? (s.lams > st.dual_tol) : (s.lams < -st.dual_tol));
      if (!__any_sync(FULL, blk)) {
        // 3. primal iterate u = -M_W' lam*
        double uc = 0.0;
        const double *mcol = c.cv.M + (lane < n ? lane : 0);   // lanes >= n shadow lane 0 (masked below)
        const int rown = s.row * n;
#pragma unroll UF
        for (int i = 0; i < k; ++i) uc = fma(-shfl_d(s.lams, i), mcol[__shfl_sync(FULL, rown, i)], uc);
        u = uc;
        if (lane < n) c.uw[lane] = uc;
        __syncwarp();
        fval = k > 0 ? bfly_sum(lane < n ? __dmul_rn(uc, uc) : 0.0) : 0.0;   // k == 0: a sum of +0 products
        const bool sf = lay.nsoft > 0 && lane < k && s.soft;
        if (lay.nsoft > 0) fval = fma(st.rho_soft, bfly_sum(sf ? __dmul_rn(s.lams, s.lams) : 0.0), fval);
        if (fval > st.fval_bound) { flag = EXIT_INFEASIBLE; break; }
        // most violated row: candidate per row, then first-index argmin over the warp
        double best = CUDART_INF, bestd = 0.0;
        int bestj = -1;
        bool bestlow = false;
#pragma unroll 1
        for (int half = 0; half < (m > 32 ? 2 : 1); ++half) {
          const int j = lane + 32 * half;
          if (j < m) {
            double mu = 0.0;
            // empty working set (three steps out of four of a settled closed loop): u is exactly +0, and so is
            // every fma(M, +0, +0) of this chain
            if (k > 0) {
#pragma unroll UF
              for (int cc = 0; cc < n; ++cc) mu = fma(c.cv.MT[cc * mp + j], c.uw[cc], mu);
            }
            const bool skip = ((s.inws >> j) & 1ull) || (c.cv.sense[j] & S_IMMUTABLE);
            const double shj = half ? sh1 : sh0;
            const double dU = c.cv.du[j] + shj, dLo = c.cv.dl[j] + shj;
            const double su = dU - mu, sl = mu - dLo;
            if (!skip) {
              // the sequential rule (upper tested first against the running best) picks, for
              // consistent bounds, the negative one of su / sl; strict '<' keeps the lower row
              if (su < -st.primal_tol && su < best) { best = su; bestd = dU; bestj = j; bestlow = false; }
              else if (!(su < -st.primal_tol) && sl < -st.primal_tol && sl < best) { best = sl; bestd = dLo; bestj = j; bestlow = true; }
            }
          }
        }
        // ties: lowest row index wins; rows lane (half 0) come before lane+32 (half 1)
        const bool hit = ((min_lanes(best) >> lane) & 1u) && bestj >= 0;
        const unsigned hit0 = __ballot_sync(FULL, hit && bestj < 32);
        const unsigned hit1 = __ballot_sync(FULL, hit && bestj >= 32);
        if (!(hit0 | hit1)) {
          // largest soft slack rho |lam*|, only needed to tell OPTIMAL from SOFT_OPTIMAL
          if (lay.nsoft > 0) soft_slack = bfly_max(sf ? st.rho_soft * fabs(s.lams) : 0.0);
          flag = soft_slack > st.primal_tol ? EXIT_SOFT_OPTIMAL : EXIT_OPTIMAL;
          break;
        }
        // cycle guard where the dual objective is fresh (see oracle/ldp_oracle.c)
        bool rebuilt = false;
        if (fval - st.progress_tol > best_fval) { best_fval = fval; cycle = 0; }
        else if (cycle++ > st.cycle_tol) {
          // second chance for problems with soft rows: refactorise once, restart the count
          if (repaired || !wq_has_soft(c, s)) { flag = EXIT_CYCLE; break; }
          if (!REPAIR) { flag = EXIT_REDO; break; }
          repaired = true; rebuilt = true; best_fval = -1.0; cycle = 0;
        }
        if (k >= lay.cap || k >= 32) { flag = EXIT_CAPACITY; break; }
        const int src = hit0 ? (__ffs(hit0) - 1) : (__ffs(hit1) - 1);
        const int add = __shfl_sync(FULL, bestj, src);
        const bool addlow = __shfl_sync(FULL, (int)bestlow, src);
        const double addd = shfl_d(bestd, src);
        if (lane < k) s.lam = s.lams;
        if constexpr (REPAIR) { if (rebuilt) { s = wq_rebuild<UF>(c, s); if (s.rc) { flag = s.rc; break; } } }
        wq_add<UF>(c, s, add, addlow, addd);
        if (s.sing && wq_has_soft(c, s)) {
          if constexpr (REPAIR) { s = wq_rebuild<UF>(c, s); if (s.rc) { flag = s.rc; break; } }
          else { flag = EXIT_REDO; break; }
        }
        continue;
      }
      // 4. blocking: ratio test against lam* - lam
      if (blk) {
        a = -s.lam / (s.lams - s.lam);
        if (!(a == a)) a = CUDART_INF;
      }
      dir = s.lams - s.lam;
    } else {
      // 5. singular: p = side * [-L^-T l ; 1] on the first k-1 positions
      const int kk = k - 1;
      const bool lastlow = __shfl_sync(FULL, (int)s.low, kk);
      double p = (lane < kk) ? -c.Lw[kk * c.LS + lane] : 0.0;
      p = sweep_bwd<UF>(c, s, p, kk);
      if (lane == kk) p = 1.0;
      if (lastlow && lane <= kk) p = -p;
      s.lams = (lane < k) ? p : 0.0;
      const bool blk = lane < k && !s.imm && (s.low ? (p > st.dual_tol) : (p < -st.dual_tol));
      if (blk) {
        a = -s.lam / p;
        if (!(a == a)) a = CUDART_INF;
      }
      dir = p;
      noblock = EXIT_INFEASIBLE;
    }
    // step to the first blocking multiplier and drop its row (one site for both branches)
    double alpha;
    const int rm = argmin_first(a, alpha);
    if (rm < 0) { flag = noblock; break; }   // regular: blocked but no finite ratio (NaN guard)
    if (lane < k) s.lam = fma(alpha, dir, s.lam);
    wq_remove(c, s, rm);
    if (s.sing && wq_has_soft(c, s)) {
      if constexpr (REPAIR) { s = wq_rebuild<UF>(c, s); if (s.rc) { flag = s.rc; break; } }
      else { flag = EXIT_REDO; break; }
    }
  }
  iters = iter;
  return flag;
}

// one solve for theta in c.thw; returns the flag; u (lane c holds u_c) and iters by reference
template <int UF, bool EQ, bool REPAIR>
LMPC_DEVINL int wq_solve(const WarpCtx &c, WarpQp &s, bool cold, double &u, int &iters, double &sho,
                         bool warm = false, unsigned long long wsu = 0ull, unsigned long long wsl = 0ull) {
  const LdpLayout &lay = *c.lay;
  const int lane = c.lane, m = lay.m, nth = lay.nth;
  // mpc_update_qp: sh_j = Dth[j,:] theta for rows lane and lane + 32.  When lanes are left over
  // (m + nu <= 32) lanes m .. m+nu-1 ride along and accumulate the output shifts
  // uoff_i + Uth[i,:] theta of mpc_get_solution in the same loop (sho on lane m + i).
  double sh0 = 0.0, sh1 = 0.0;
  const bool fold = m + lay.nu <= 32;
  if (lane < m || (fold && lane < m + lay.nu)) {
    const bool mine = lane < m;
    const double *dr = mine ? c.cv.Dth + lane * nth : c.cv.Uth + (lane - m) * nth;
    double acc = mine ? 0.0 : c.cv.uoff[lane - m];
#pragma unroll 2
    for (int t = 0; t < nth; ++t) acc = fma(dr[t], c.thw[t], acc);
    sh0 = mine ? acc : 0.0;
    sho = acc;
  }
  if (lane + 32 < m) {
    const double *dr = c.cv.Dth + (lane + 32) * nth;
#pragma unroll 2
    for (int t = 0; t < nth; ++t) sh1 = fma(dr[t], c.thw[t], sh1);
  }
  int flag = 0;
  if (cold) {
    wq_reset(s);
    u = 0.0;
    // equality rows enter the working set before the first iteration; controllers without any
    // (EQ == false) are compiled without this second copy of the add
    for (int j = 0; EQ && lay.neq > 0 && j < m; ++j) {
      const int sj = c.cv.sense[j];
      if (sj & S_ACTIVE) {
        if (s.k >= lay.cap || s.k >= 32) { flag = EXIT_OVERDETERMINED; break; }
        const bool low = sj & S_LOWER;
        const double shj = shfl_d(j < 32 ? sh0 : sh1, j & 31);
        wq_add<UF>(c, s, j, low, (low ? c.cv.dl[j] : c.cv.du[j]) + shj);
        if (s.sing) { flag = EXIT_OVERDETERMINED; break; }
      }
    }
    // warm start from a caller-supplied working set (oracle: activate_rows): named rows in
    // ascending order, lam = 0, upper wins; a row that would be singular is left out
    for (int j = 0; EQ && warm && flag == 0 && j < m; ++j) {
      const int sj = c.cv.sense[j];
      const bool up = (wsu >> j) & 1ull, lo = (wsl >> j) & 1ull;
      if ((sj & (S_ACTIVE | S_IMMUTABLE)) || !(up || lo)) continue;
      if (s.k >= lay.cap || s.k >= 32) continue;
      const double shj = shfl_d(j < 32 ? sh0 : sh1, j & 31);
      wq_add<UF>(c, s, j, !up, (up ? c.cv.du[j] : c.cv.dl[j]) + shj);
      if (s.sing) { s.k -= 1; s.inws &= ~(1ull << j); s.sing = false; }   // undo the trial add
    }
  } else {
    // warm start: same working set, new right-hand sides
    const int j = s.row;
    const double shj_lo = shfl_d(sh0, j & 31), shj_hi = shfl_d(sh1, j & 31);
    if (lane < s.k) s.dW = (s.low ? c.cv.dl[j] : c.cv.du[j]) + (j < 32 ? shj_lo : shj_hi);
  }
  iters = 0;
  if (flag == 0) flag = wq_loop<UF, REPAIR>(c, s, sh0, sh1, u, iters);
  return flag;
}

// mpc_get_solution: control_i = (R^-1 u)_i + uoff_i + Uth[i,:] theta, canonical lane sum over c
LMPC_DEVINL double control_out_w(const WarpCtx &c, int i, double u, double sho) {
  const LdpLayout &lay = *c.lay;
  const double x = bfly_sum(c.lane < lay.n ? __dmul_rn(c.cv.Rinv[i * lay.n + c.lane], u) : 0.0);
  if (lay.m + lay.nu <= 32) return __dadd_rn(x, shfl_d(sho, lay.m + i));   // accumulated by wq_solve
  double sh = c.cv.uoff[i];
  const double *ur = c.cv.Uth + i * lay.nth;
#pragma unroll 2
  for (int t = 0; t < lay.nth; ++t) sh = fma(ur[t], c.thw[t], sh);
  return __dadd_rn(x, sh);
}

LMPC_DEVINL WarpCtx make_ctx(double *smem, const LdpLayout &lay, const LdpSettings &st, int extra) {
  WarpCtx c;
  c.cv = make_view(smem, lay);
  c.lay = &lay; c.st = &st;
  c.LS = lay.cap | 1;
  const int warp = threadIdx.x >> 5;
  const int per_warp = ((lay.cap * c.LS + lay.n + lay.nth + extra) + 1) & ~1;
  double *w = smem + lay.total + (size_t)warp * per_warp;
  c.Lw = w; c.uw = w + lay.cap * c.LS; c.thw = c.uw + lay.n;
  c.lane = threadIdx.x & 31;
  return c;
}

template <bool EQ, bool REPAIR>
__global__ void __launch_bounds__(K2_BLOCK, K2_MINB)
ldp_warp_kernel(const __grid_constant__ LdpLayout lay, const __grid_constant__ LdpSettings st,
                const double *__restrict__ cblk, ThetaSegs th,
                int64_t Ninst, SolveOut out, int redo) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar;
  stage_constants_tma(smem, cblk, lay.total, &mbar);
  const WarpCtx c = make_ctx(smem, lay, st, 0);
  const int lane = c.lane, words = (lay.m + 63) >> 6;
  const int64_t wid = (int64_t)blockIdx.x * K2_WARPS + (threadIdx.x >> 5);
  const int64_t wstride = (int64_t)gridDim.x * K2_WARPS;
  WarpQp s;
  double u = 0.0;
  for (int64_t q = wid; q < Ninst; q += wstride) {
    if (redo && out.flag[q] != EXIT_REDO) continue;      // second launch: only what the fast kernel handed over
    // batched mpc_update_parameter: gather theta from its row segments (coalesced per segment)
    int o = 0;
#pragma unroll
    for (int sg = 0; sg < 4; ++sg) {
      if (sg < th.nseg) {
        const int len = th.len[sg];
#pragma unroll 1
        for (int t = lane; t < len; t += 32) c.thw[o + t] = ld_stream(th.ptr[sg] + q * th.stride[sg] + t);
        o += len;
      }
    }
    __syncwarp();
    int iters;
    const bool warm = EQ && out.ws_u != nullptr;
    double sho = 0.0;
    const int flag = wq_solve<K2_UF_SOLVE, EQ, REPAIR>(c, s, true, u, iters, sho, warm, warm ? out.ws_u[q * words] : 0ull,
                                               warm ? out.ws_l[q * words] : 0ull);
    for (int i = 0; i < lay.nu; ++i) {
      const double ctl = control_out_w(c, i, u, sho);
      if (lane == 0) out.U[q * lay.nu + i] = ctl;
    }
    // active set as bit masks (m <= 64: one word each)
    unsigned long long au = 0ull, al = 0ull;
    if (lane < s.k) { if (s.low) al = 1ull << s.row; else au = 1ull << s.row; }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      au |= __shfl_xor_sync(FULL, au, off);
      al |= __shfl_xor_sync(FULL, al, off);
    }
    if (lane == 0) {
      out.flag[q] = flag;
      if (out.iters) out.iters[q] = iters;
      // (an instance handed over to the repair launch keeps its input working-set words: they may be the
      // warm start it has to begin from again)
      if (out.act_u && (REPAIR || flag != EXIT_REDO)) { out.act_u[q * words] = au; out.act_l[q * words] = al; }
    }
    __syncwarp();
  }
}

// closed loop: T steps of (theta = [x; r; d; u_prev]) -> solve -> x+ = F x + G u + Gd d
// x+ = A x + B u + Bd d for the rows lane, lane + 32, ... (fma chains in index order)
LMPC_DEVINL void step_rows(const double *A, const double *B, const double *Bd, const double *x, const double *uu,
                           const double *d, double *out, int nx, int nu, int nd, int lane) {
  _Pragma("unroll 1")
  for (int i = lane; i < nx; i += 32) {
    double acc = 0.0;
    _Pragma("unroll 2")
    for (int cc = 0; cc < nx; ++cc) acc = fma(A[i * nx + cc], x[cc], acc);
    _Pragma("unroll 2")
    for (int cc = 0; cc < nu; ++cc) acc = fma(B[i * nu + cc], uu[cc], acc);
    _Pragma("unroll 2")
    for (int cc = 0; cc < nd; ++cc) acc = fma(Bd[i * nd + cc], d[cc], acc);
    out[i] = acc;
  }
}

// one row of the same update (row i on the calling lane), returned instead of stored
LMPC_DEVINL double step_row(const double *A, const double *B, const double *Bd, const double *x, const double *uu,
                            const double *d, int nx, int nu, int nd, int i) {
  double acc = 0.0;
  _Pragma("unroll 4")
  for (int cc = 0; cc < nx; ++cc) acc = fma(A[i * nx + cc], x[cc], acc);
  _Pragma("unroll 2")
  for (int cc = 0; cc < nu; ++cc) acc = fma(B[i * nu + cc], uu[cc], acc);
  _Pragma("unroll 2")
  for (int cc = 0; cc < nd; ++cc) acc = fma(Bd[i * nd + cc], d[cc], acc);
  return acc;
}

// EXT: separate true plant, steady-state observer and measurement noise (lmpc_gpu_simulate_batch_ex);
// compiled as its own kernel so the plain closed loop keeps its instruction footprint
template <bool EQ, bool EXT, bool REPAIR>
__global__ void __launch_bounds__(K2_BLOCK, K2_MINB)
ldp_warp_sim_kernel(const __grid_constant__ LdpLayout lay, const __grid_constant__ LdpSettings st,
                    const double *__restrict__ cblk,
                    int64_t Ninst, SimDesc sd, int32_t *flagout, int32_t *itersout, int redo) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar;
  stage_constants_tma(smem, cblk, lay.total, &mbar);
  const int nx = sd.nx, nr = sd.nr, nd = sd.nd, nup = sd.nuprev, nu = lay.nu;
  const int ny = EXT ? sd.ny : 0;
  const WarpCtx c = make_ctx(smem, lay, st, nx + nu + (EXT ? nx + ny : 0));
  double *xnw = c.thw + lay.nth;      // [nx] next state
  double *ucw = xnw + nx;             // [nu] control
  double *xsw = ucw + nu;             // EXT: [nx] true plant state (thw[0, nx) is then what the controller sees)
  double *ew = xsw + nx;              // EXT: [ny] innovation
  const int lane = c.lane;
  const int64_t wid = (int64_t)blockIdx.x * K2_WARPS + (threadIdx.x >> 5);
  const int64_t wstride = (int64_t)gridDim.x * K2_WARPS;
  WarpQp s;
  double u = 0.0;
  for (int64_t q = wid; q < Ninst; q += wstride) {
    if (redo && flagout[q] != EXIT_REDO) continue;       // second launch: whole closed loops handed over
    _Pragma("unroll 1")
    for (int i = lane; i < nx; i += 32) { c.thw[i] = ld_stream(sd.X0 + q * nx + i); if (EXT) xsw[i] = c.thw[i]; }
    _Pragma("unroll 1")
    for (int i = lane; i < nr; i += 32) c.thw[nx + i] = sd.R ? ld_stream(sd.R + q * nr + i) : 0.0;
    _Pragma("unroll 1")
    for (int i = lane; i < nd; i += 32) c.thw[nx + nr + i] = sd.D ? ld_stream(sd.D + q * nd + i) : 0.0;
    _Pragma("unroll 1")
    for (int i = lane; i < nu; i += 32) ucw[i] = (sd.U0 && i < nup) ? ld_stream(sd.U0 + q * nup + i) : 0.0;
    __syncwarp();
    int worst = EXIT_OPTIMAL, tot = 0;
    // lean step bookkeeping (plain closed loop, nx <= 32): the controls of a step are collected on lanes 0..nu-1 and
    // the next state on lanes 0..nx-1, in registers, and go to theta with one predicated store each -- no per-element
    // copy loops, no staging vector for x+, two warp barriers per step fewer (r02 ncu: the step's bookkeeping was 14 %
    // of the executed instructions of the AFTI-16 closed loop)
    const bool lean = !EXT && nx <= 32;
    if (lean) {
      if (lane < nup) c.thw[nx + nr + nd + lane] = ucw[lane];
      __syncwarp();
    }
    for (int t = 0; t < sd.T; ++t) {
      if (!lean) {
        _Pragma("unroll 1")
        for (int i = lane; i < nup; i += 32) c.thw[nx + nr + nd + i] = ucw[i];
      }
      // reference / disturbance of step t (rs[:, k], ds[:, k] of src/simulation.jl:93-104)
      if (sd.Rtraj) { _Pragma("unroll 1") for (int i = lane; i < nr; i += 32) c.thw[nx + i] = sd.Rtraj[(size_t)t * sd.rstride + i]; }
      if (sd.Dtraj) { _Pragma("unroll 1") for (int i = lane; i < nd; i += 32) c.thw[nx + nr + i] = sd.Dtraj[(size_t)t * sd.dstride + i]; }
      __syncwarp();
      if constexpr (EXT) {
        const double *dk = c.thw + nx + nr;
        if (sd.Cm) {
          // y = C x + Dd d + v, innovation against the estimate, correction (observer.jl:114-118)
          _Pragma("unroll 1")
          for (int i = lane; i < ny; i += 32) {
            double y = sd.Vtraj ? sd.Vtraj[(size_t)t * ny + i] : 0.0, yh = 0.0;
            for (int cc = 0; cc < nx; ++cc) y = fma(sd.Cm[i * nx + cc], xsw[cc], y);
            if (sd.Ddm) for (int cc = 0; cc < nd; ++cc) y = fma(sd.Ddm[i * nd + cc], dk[cc], y);
            for (int cc = 0; cc < nx; ++cc) yh = fma(sd.Cm[i * nx + cc], c.thw[cc], yh);
            if (sd.Ddm) for (int cc = 0; cc < nd; ++cc) yh = fma(sd.Ddm[i * nd + cc], dk[cc], yh);
            ew[i] = __dsub_rn(y, yh);
          }
          __syncwarp();
          _Pragma("unroll 1")
          for (int j = lane; j < nx; j += 32) {
            double acc = c.thw[j];
            for (int i = 0; i < ny; ++i) acc = fma(sd.Kobs[j * ny + i], ew[i], acc);
            xnw[j] = acc;
          }
          __syncwarp();
          _Pragma("unroll 1")
          for (int j = lane; j < nx; j += 32) c.thw[j] = xnw[j];
        } else {
          _Pragma("unroll 1")
          for (int j = lane; j < nx; j += 32) c.thw[j] = xsw[j];      // the controller measures the true state
        }
        __syncwarp();
      }
      int iters;
      double sho = 0.0;
      const int f = wq_solve<K2_UF_SIM, EQ, REPAIR>(c, s, !(sd.warm && t > 0), u, iters, sho);
      tot += iters;
      if (!REPAIR && f == EXIT_REDO) { worst = EXIT_REDO; break; }   // the repair-capable launch redoes this loop
      if (f < 0) { if (worst > 0 || f < worst) worst = f; }
      else if (worst > 0 && f > worst) worst = f;
      if (lean) {
        double myctl = 0.0;
        for (int i = 0; i < nu; ++i) {
          const double ctl = control_out_w(c, i, u, sho);        // the same value on every lane
          if (lane == i) myctl = ctl;
        }
        if (lane < nu) {
          ucw[lane] = myctl;
          if (sd.Utraj) sd.Utraj[(q * sd.T + t) * nu + lane] = myctl;
          if (lane < nup) c.thw[nx + nr + nd + lane] = myctl;     // u_prev of the next step (not read below)
        }
        __syncwarp();
        double xplus = 0.0;
        if (lane < nx) xplus = step_row(sd.F, sd.G, sd.Gd, c.thw, ucw, c.thw + nx + nr, nx, nu, nd, lane);   // model (predict!)
        __syncwarp();
        if (lane < nx) c.thw[lane] = xplus;
        __syncwarp();
      } else {
        for (int i = 0; i < nu; ++i) {
          const double ctl = control_out_w(c, i, u, sho);
          if (lane == 0) {
            ucw[i] = ctl;
            if (sd.Utraj) sd.Utraj[(q * sd.T + t) * nu + i] = ctl;
          }
        }
        __syncwarp();
        step_rows(sd.F, sd.G, sd.Gd, c.thw, ucw, c.thw + nx + nr, xnw, nx, nu, nd, lane);   // model (predict!)
        __syncwarp();
        _Pragma("unroll 1")
        for (int i = lane; i < nx; i += 32) c.thw[i] = xnw[i];
        __syncwarp();
      }
      if constexpr (EXT) {
        step_rows(sd.Ft, sd.Gt, sd.Gdt, xsw, ucw, c.thw + nx + nr, xnw, nx, nu, nd, lane);   // true plant
        __syncwarp();
        _Pragma("unroll 1")
        for (int i = lane; i < nx; i += 32) xsw[i] = xnw[i];
        __syncwarp();
      }
    }
    if (EXT && sd.Xhat) {
      _Pragma("unroll 1")
      for (int i = lane; i < nx; i += 32) sd.Xhat[q * nx + i] = c.thw[i];
    }
    _Pragma("unroll 1")
    for (int i = lane; i < nx; i += 32) sd.Xend[q * nx + i] = EXT ? xsw[i] : c.thw[i];
    _Pragma("unroll 1")
    for (int i = lane; i < nu; i += 32) sd.Ulast[q * nu + i] = ucw[i];
    if (lane == 0) {
      flagout[q] = worst;
      if (itersout) itersout[q] = tot;
    }
    __syncwarp();
  }
}

size_t k2_smem(const LdpLayout &lay, int extra) {
  const int LS = lay.cap | 1;
  const size_t per_warp = (((size_t)lay.cap * LS + lay.n + lay.nth + extra) + 1) & ~(size_t)1;
  return ((size_t)lay.total + per_warp * K2_WARPS) * sizeof(double);
}

}  // namespace

bool ldp_warp_supports(const LdpLayout &lay) {
  return lay.n >= 1 && lay.n <= 32 && lay.cap <= 32 && lay.m <= 64 && k2_smem(lay, 64) <= 200 * 1024;
}

template <bool EQ, bool REPAIR>
static int launch_warp_t(const LdpLayout &lay, const LdpSettings &st, const double *cblk,
                         const ThetaSegs &th, int64_t N, const SolveOut &out, const SimDesc *sim,
                         int sm_count, cudaStream_t s, int redo) {
  const size_t smem = k2_smem(lay, sim ? sim->nx + lay.nu + (sim->ext ? sim->nx + sim->ny : 0) : 0);
  int per_sm = 0;
  int64_t blocks = (N + K2_WARPS - 1) / K2_WARPS;
  cudaError_t e;
  if (sim) {
    auto kern = sim->ext ? ldp_warp_sim_kernel<EQ, true, REPAIR> : ldp_warp_sim_kernel<EQ, false, REPAIR>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K2_BLOCK, smem);
    if (per_sm < 1) per_sm = 1;
    if (blocks > (int64_t)sm_count * per_sm) blocks = (int64_t)sm_count * per_sm;
    kern<<<(unsigned)blocks, K2_BLOCK, smem, s>>>(lay, st, cblk, N, *sim, out.flag, out.iters, redo);
  } else {
    auto kern = ldp_warp_kernel<EQ, REPAIR>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K2_BLOCK, smem);
    if (per_sm < 1) per_sm = 1;
    if (blocks > (int64_t)sm_count * per_sm) blocks = (int64_t)sm_count * per_sm;
    kern<<<(unsigned)blocks, K2_BLOCK, smem, s>>>(lay, st, cblk, th, N, out, redo);
  }
  return (int)cudaGetLastError();
}

// Problems without soft rows never repair: one launch of the fast kernel.  With soft rows: the fast kernel,
// then the repair-capable one over the instances (closed loops) it handed over -- none on the BASELINE
// AFTI-16 controller, a few percent with output bounds on the whole horizon.
template <bool EQ>
static int launch_warp_2(const LdpLayout &lay, const LdpSettings &st, const double *cblk,
                         const ThetaSegs &th, int64_t N, const SolveOut &out, const SimDesc *sim,
                         int sm_count, cudaStream_t s) {
  int e = launch_warp_t<EQ, false>(lay, st, cblk, th, N, out, sim, sm_count, s, 0);
  if (e != 0 || lay.nsoft == 0 || !K2_REPAIR) return e;
  lmpc_count_kernel(1);
  return launch_warp_t<EQ, true>(lay, st, cblk, th, N, out, sim, sm_count, s, 1);
}

int launch_ldp_warp(const LdpLayout &lay, const LdpSettings &st, const double *cblk,
                    const ThetaSegs &th, int64_t N, const SolveOut &out, const SimDesc *sim,
                    int sm_count, void *stream) {
  if (N <= 0) return 0;
  return (lay.neq > 0 || out.ws_u) ? launch_warp_2<true>(lay, st, cblk, th, N, out, sim, sm_count, (cudaStream_t)stream)
                     : launch_warp_2<false>(lay, st, cblk, th, N, out, sim, sm_count, (cudaStream_t)stream);
}
