// K1b: thread-per-QP fast path for BOX-ONLY problems (m == n simple bounds, no soft /
// equality / general rows, n <= 8, nu <= 2) -- the shape of the headline workload
// (inverted pendulum: n = m = 5, nth = 7, nu = 1).
//
// Same algorithm and the same arithmetic as ldp_thread.cu (bit-identical results), restructured
// to cut the instruction count, which is what bounds this kernel (ncu r01 of the general kernel:
// issue slots 54 % busy with 12 warps/SM, FP64 pipe 18 %, 7350 thread-instructions per solve for
// 578 algorithmic flops; this kernel: 3200, FP64 pipe 39 %).  Resident CTAs per SM were swept
// (3 / 4 / 5 / 6 -> 7.0 / 7.0 / 5.1 / 4.1e9 solves/s): registers beat occupancy, hence MINB = 3.
//   * "unused slots are zero rows": positions >= k always hold L row 0, D = 1/D = 1,
//     d_W = lam = 0, so the CSP solve, the step and the primal update run unpredicated over
//     the full capacity and produce exact zeros there (no per-slot selects / predicates);
//   * the slack scan is unrolled over the m == N rows, uses the triangular structure of the
//     normalised R^-1 rows (row j starts at column j) and keeps sh_j in registers;
//   * theta is consumed as it is loaded: one pass accumulates every sh_j and the output
//     shift, so theta never sits in shared memory;
//   * no immutable / soft / singular handling at all (cannot occur for this class).
#include "ldp_device.cuh"
#include <math_constants.h>
#include <cstdlib>

namespace {

#ifndef KB_BLOCK_SIZE
#define KB_BLOCK_SIZE 128
#endif
constexpr int KB_BLOCK = KB_BLOCK_SIZE;
#ifndef KB_TH_UNROLL
#define KB_TH_UNROLL 4
#endif
constexpr int KB_THU = KB_TH_UNROLL;   // unroll factor of the loop over theta

// Constants whose index is known at compile time travel as a kernel parameter: the compiler
// reads them straight from the constant bank as an FMA operand (no load instruction).  Rows of
// M-bar picked by a run-time row id stay in shared memory, padded to an even stride so that a
// row is fetched with 128-bit loads.
template <int N, int NU>
struct BoxConst {
  double M[N * N];
  double du[N], dl[N];
  double Rinv[NU * N];
};
template <int N>
struct BoxPad { static constexpr int NP = (N + 1) & ~1; };

// u <- u - li * M[row, :]  (row stride NP doubles, 16-byte aligned)
template <int N>
LMPC_DEVINL void box_axpy_row(const double *Mp, int row, double li, double (&u)[N]) {
  constexpr int NP = BoxPad<N>::NP;
  const double2 *r2 = reinterpret_cast<const double2 *>(Mp + row * NP);
  double mr[NP];
  SFOR(h, 0, NP / 2) const double2 v = r2[h]; mr[2 * h] = v.x; mr[2 * h + 1] = v.y; SEND
  SFOR(c, 0, N) u[c] = fma(-li, mr[c], u[c]); SEND
}

template <int N>
struct BoxState {
  double L[N * (N - 1) / 2 > 0 ? N * (N - 1) / 2 : 1];
  double D[N], Dinv[N], lam[N], lams[N], dW[N];
  int row[N];
  unsigned lowmask;  // bit i: position i active at its LOWER bound
  unsigned inws;     // bit j: row j in the working set
  int k;
};

template <int N>
LMPC_DEVINL void box_reset(BoxState<N> &s) {
  s.k = 0; s.lowmask = 0u; s.inws = 0u;
  SFOR(i, 0, N) s.D[i] = 1.0; s.Dinv[i] = 1.0; s.lam[i] = 0.0; s.lams[i] = 0.0; s.dW[i] = 0.0; s.row[i] = 0; SEND
  SFOR(i, 0, N * (N - 1) / 2) s.L[i] = 0.0; SEND
}

// append row j (oracle: ldl_add).  Rows are independent for a box problem: never singular.
template <int N>
LMPC_DEVINL void box_add(BoxState<N> &s, const double *G, int j, bool lower, double dval) {
  const double *g = G + j * N;
  const int k = s.k;
  double y[N], l[N], pr[N];
  SFOR(i, 0, N)
    double yi = (i < k) ? g[s.row[i]] : 0.0;
    SFOR(c, 0, i) yi = fma(-s.L[LIDX(i, c)], y[c], yi); SEND
    y[i] = yi;
  SEND
  SFOR(i, 0, N) l[i] = y[i] * s.Dinv[i]; pr[i] = __dmul_rn(y[i], l[i]); SEND
  const double dnew = __dsub_rn(g[j], tree_sum<N, 0, 1>(pr));
  const double dinv = 1.0 / dnew;
  SFOR(kk, 0, N)
    const bool here = (kk == k);
    SFOR(i, 0, kk) s.L[LIDX(kk, i)] = here ? l[i] : s.L[LIDX(kk, i)]; SEND
    s.D[kk] = here ? dnew : s.D[kk];
    s.Dinv[kk] = here ? dinv : s.Dinv[kk];
    s.dW[kk] = here ? dval : s.dW[kk];
    s.row[kk] = here ? j : s.row[kk];
  SEND
  const unsigned bit = 1u << k;
  s.lowmask = lower ? (s.lowmask | bit) : (s.lowmask & ~bit);
  s.inws |= 1u << j;
  s.k = k + 1;
}

// delete position r (oracle: ldl_remove); vacated slots become zero rows again
template <int N>
LMPC_DEVINL void box_remove(BoxState<N> &s, int r) {
  const int k = s.k - 1;   // size after removal
  double alpha = 0.0, wv[N];
  int rrow = 0;
  SFOR(i, 0, N)
    wv[i] = 0.0;
    alpha = (i == r) ? s.D[i] : alpha;
    rrow = (i == r) ? s.row[i] : rrow;
    SFOR(c, 0, i) wv[i] = (c == r) ? s.L[LIDX(i, c)] : wv[i]; SEND
  SEND
  s.inws &= ~(1u << rrow);
  // position p takes old position p + 1 for every p >= r; the last slot becomes a zero row
  SFOR(p, 0, N)
    const bool mv = (p >= r);
    if constexpr (p + 1 < N) {
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
    } else {
      SFOR(c, 0, p) s.L[LIDX(p, c)] = 0.0; SEND
      s.D[p] = 1.0; s.Dinv[p] = 1.0; s.lam[p] = 0.0; s.dW[p] = 0.0; wv[p] = 0.0;
    }
  SEND
  const unsigned lo = s.lowmask & ((1u << r) - 1u);
  s.lowmask = lo | ((s.lowmask >> (r + 1)) << r);
  SFOR(j, 0, N - 1)
    if (j >= r && j < k) {
      const double pj = wv[j];
      const double t = alpha * pj;
      const double dbar = fma(t, pj, s.D[j]);
      const double rinv = 1.0 / dbar;
      const double beta = t * rinv;
      alpha = (s.D[j] * alpha) * rinv;
      s.D[j] = dbar; s.Dinv[j] = rinv;
      SFOR(i, j + 1, N - 1)
        // rows >= k are zero rows with wv == 0: the update is an exact no-op there
        wv[i] = fma(-pj, s.L[LIDX(i, j)], wv[i]);
        s.L[LIDX(i, j)] = fma(beta, wv[i], s.L[LIDX(i, j)]);
      SEND
    }
  SEND
  s.k = k;
}

struct IterCtx {
  double best_fval, fval;
  int cycle, flag, iter;
  bool done;
};

// Iteration K + 1 for a thread whose working set is exactly positions 0..K-1 (it has only added
// rows so far).  Same operations in the same order as the generic iteration, with exact loop
// bounds and the new row written to the static slot K.  Does nothing for threads that are done,
// behind (iter != K + 1) or whose multipliers block.
template <int N, int K, class BC>
LMPC_DEVINL void box_iter_static(BoxState<N> &s, const ConstView &cv, const BC &bc, const double *Mp,
                                 const LdpSettings &st,
                                 const double (&dU)[N], const double (&dLo)[N], double (&u)[N], IterCtx &cx) {
  static_assert(K >= 1 && K < N, "static iterations append at slot K");
  if (cx.done || cx.iter != K + 1 || s.k != K) return;
  double z[K], lams[K];
  SFOR(i, 0, K)
    double zi = -s.dW[i];
    SFOR(c, 0, i) zi = fma(-s.L[LIDX(i, c)], z[c], zi); SEND
    z[i] = zi;
  SEND
  SFOR(i, 0, K) lams[i] = z[i] * s.Dinv[i]; SEND
  SFOR_DOWN(j, K, 1)
    const double xj = lams[j];
    SFOR(i, 0, j) lams[i] = fma(-s.L[LIDX(j, i)], xj, lams[i]); SEND
  SEND
  bool blocked = false;
  SFOR(i, 0, K)
    const bool low = (s.lowmask >> i) & 1u;
    blocked |= low ? (lams[i] > st.dual_tol) : (lams[i] < -st.dual_tol);
  SEND
  if (blocked) return;                     // the generic loop redoes this iteration
  SFOR(i, 0, K) s.lams[i] = lams[i]; SEND
  SFOR(c, 0, N) u[c] = 0.0; SEND
  SFOR(i, 0, K) box_axpy_row<N>(Mp, s.row[i], lams[i], u); SEND
  double pu[N];
  SFOR(c, 0, N) pu[c] = __dmul_rn(u[c], u[c]); SEND
  cx.fval = tree_sum<N, 0, 1>(pu);
  if (cx.fval > st.fval_bound) { cx.flag = EXIT_INFEASIBLE; cx.done = true; return; }
  double best = -st.primal_tol;
  int code = -1;
  SFOR(j, 0, N)
    double mu = 0.0;
    SFOR(c, j, N) mu = fma(bc.M[j * N + c], u[c], mu); SEND
    const bool free_row = !((s.inws >> j) & 1u);
    const double su = dU[j] - mu, sl = mu - dLo[j];
    const bool tu = free_row && (su < best);
    const bool tl = free_row && !tu && (sl < best);
    best = tu ? su : (tl ? sl : best);
    code = tu ? 2 * j : (tl ? 2 * j + 1 : code);
  SEND
  const int add = code >> 1;
  const bool addlow = code & 1;
  if (add < 0) { cx.flag = EXIT_OPTIMAL; cx.done = true; return; }
  if (cx.fval - st.progress_tol > cx.best_fval) { cx.best_fval = cx.fval; cx.cycle = 0; }
  else if (cx.cycle++ > st.cycle_tol) { cx.flag = EXIT_CYCLE; cx.done = true; return; }
  double addd = 0.0;
  SFOR(j, 0, N) addd = (add == j) ? (addlow ? dLo[j] : dU[j]) : addd; SEND
  const double *g = cv.G + add * N;
  double y[K], l[K], pr[K];
  SFOR(i, 0, K)
    s.lam[i] = lams[i];
    double yi = g[s.row[i]];
    SFOR(c, 0, i) yi = fma(-s.L[LIDX(i, c)], y[c], yi); SEND
    y[i] = yi;
  SEND
  SFOR(i, 0, K) l[i] = y[i] * s.Dinv[i]; pr[i] = __dmul_rn(y[i], l[i]); SEND
  const double dnew = __dsub_rn(g[add], tree_sum<K, 0, 1>(pr));
  SFOR(i, 0, K) s.L[LIDX(K, i)] = l[i]; SEND
  s.D[K] = dnew; s.Dinv[K] = 1.0 / dnew; s.dW[K] = addd; s.row[K] = add;
  if (addlow) s.lowmask |= 1u << K;
  s.inws |= 1u << add;
  s.k = K + 1;
  cx.iter = K + 2;
}


// ---------------------------------------------------------------------------------------------
// K1c: the add-only fast path.  98 % of the headline batch (s = 1; 100 % at s <= 0.25) never remove a
// row: iteration K + 1 then works on exactly the slots 0..K-1, so the whole solve is straight-line
// code with compile-time loop bounds and slots -- no generic loop, no "which slot" selects, no lam /
// lam* / D kept between iterations, and the forward sweep of the CSP solve collapses to the one new
// entry of z (z_0..z_{K-2} are the same fma chains as in the previous iteration).  An instance whose
// multipliers block is handed over to the second launch through its CTA's queue (a shared-memory counter and
// a stream-ordered global scratch; nothing is written to the caller's arrays): the generic kernel in REDO
// mode solves the queued instances from scratch.  Same operations in the same order as the generic
// iteration -> bit-identical.
// ---------------------------------------------------------------------------------------------
constexpr int EXIT_REDO_BOX = -100;   // internal, never seen by the caller

template <int N>
struct FastState {
  double L[N * (N - 1) / 2 > 0 ? N * (N - 1) / 2 : 1];
  double Dinv[N], dW[N], z[N];
  int row[N];
  unsigned lowmask, inws;
};

// most violated free row: code = 2 * row + (1 if lower), -1 if none (first minimum wins)
template <int N, class BC>
LMPC_DEVINL int box_scan(const BC &bc, unsigned inws, double primal_tol, const double (&dU)[N],
                         const double (&dLo)[N], const double (&u)[N]) {
  double best = -primal_tol;
  int code = -1;
  SFOR(j, 0, N)
    double mu = 0.0;
    SFOR(c, j, N) mu = fma(bc.M[j * N + c], u[c], mu); SEND
    const bool free_row = !((inws >> j) & 1u);
    const double su = dU[j] - mu, sl = mu - dLo[j];
    const bool tu = free_row && (su < best);
    const bool tl = free_row && !tu && (sl < best);
    best = tu ? su : (tl ? sl : best);
    code = tu ? 2 * j : (tl ? 2 * j + 1 : code);
  SEND
  return code;
}

// the same scan with the bounds read back from the thread's shared-memory column (bnd[(2 j + lower) * KB_BLOCK]),
// where the add needs them anyway.  (Measured and dropped: a tournament form -- independent per-row candidates
// min(su, sl) + a comparison tree instead of the serial compare-select chain -- was 2 % slower.)
template <int N, class BC>
LMPC_DEVINL int box_scan_smem(const BC &bc, unsigned inws, double primal_tol, const double *bnd, const double (&u)[N]) {
  double best = -primal_tol;
  int code = -1;
  SFOR(j, 0, N)
    double mu = 0.0;
    SFOR(c, j, N) mu = fma(bc.M[j * N + c], u[c], mu); SEND
    const bool free_row = !((inws >> j) & 1u);
    const double su = bnd[(2 * j) * KB_BLOCK] - mu, sl = mu - bnd[(2 * j + 1) * KB_BLOCK];
    const bool tu = free_row && (su < best);
    const bool tl = free_row && !tu && (sl < best);
    best = tu ? su : (tl ? sl : best);
    code = tu ? 2 * j : (tl ? 2 * j + 1 : code);
  SEND
  return code;
}

// iteration K + 1 on the working set {slots 0..K-1}; true = this instance is finished (flag, iter set)
template <int N, int K, class BC>
LMPC_DEVINL bool fast_iter(FastState<N> &s, const ConstView &cv, const BC &bc, const double *Mp, const LdpSettings &st,
                           const double *bnd, double (&u)[N],
                           double &best_fval, int &cycle, int &flag, int &iter) {
  iter = K + 1;
  {
    double zi = -s.dW[K - 1];
    SFOR(c, 0, K - 1) zi = fma(-s.L[LIDX(K - 1, c)], s.z[c], zi); SEND
    s.z[K - 1] = zi;
  }
  double lams[K];
  SFOR(i, 0, K) lams[i] = s.z[i] * s.Dinv[i]; SEND
  SFOR_DOWN(j, K, 1)
    const double xj = lams[j];
    SFOR(i, 0, j) lams[i] = fma(-s.L[LIDX(j, i)], xj, lams[i]); SEND
  SEND
  bool blocked = false;
  SFOR(i, 0, K)
    const bool low = (s.lowmask >> i) & 1u;
    blocked |= low ? (lams[i] > st.dual_tol) : (lams[i] < -st.dual_tol);
  SEND
  if (blocked) { flag = EXIT_REDO_BOX; return true; }
  SFOR(c, 0, N) u[c] = 0.0; SEND
  SFOR(i, 0, K) box_axpy_row<N>(Mp, s.row[i], lams[i], u); SEND
  double pu[N];
  SFOR(c, 0, N) pu[c] = __dmul_rn(u[c], u[c]); SEND
  const double fval = tree_sum<N, 0, 1>(pu);
  if (fval > st.fval_bound) { flag = EXIT_INFEASIBLE; return true; }
  if constexpr (K == N) {          // every row is in the working set: nothing left to violate
    flag = EXIT_OPTIMAL;
    return true;
  } else {
    const int code = box_scan_smem<N>(bc, s.inws, st.primal_tol, bnd, u);
    const int add = code >> 1;
    const bool addlow = code & 1;
    if (add < 0) { flag = EXIT_OPTIMAL; return true; }
    if (fval - st.progress_tol > best_fval) { best_fval = fval; cycle = 0; }
    else if (cycle++ > st.cycle_tol) { flag = EXIT_CYCLE; return true; }
    const double addd = bnd[code * KB_BLOCK];   // the thread's own copy of dU / dLo: one load instead of 2 N selects
    const double *g = cv.G + add * N;
    double y[K], l[K], pr[K];
    SFOR(i, 0, K)
      double yi = g[s.row[i]];
      SFOR(c, 0, i) yi = fma(-s.L[LIDX(i, c)], y[c], yi); SEND
      y[i] = yi;
    SEND
    SFOR(i, 0, K) l[i] = y[i] * s.Dinv[i]; pr[i] = __dmul_rn(y[i], l[i]); SEND
    const double dnew = __dsub_rn(g[add], tree_sum<K, 0, 1>(pr));
    SFOR(i, 0, K) s.L[LIDX(K, i)] = l[i]; SEND
    s.Dinv[K] = 1.0 / dnew; s.dW[K] = addd; s.row[K] = add;
    if (addlow) s.lowmask |= 1u << K;
    s.inws |= 1u << add;
    return false;
  }
}

template <int N, int K, class BC>
LMPC_DEVINL void fast_chain(FastState<N> &s, const ConstView &cv, const BC &bc, const double *Mp, const LdpSettings &st,
                            const double *bnd, double (&u)[N],
                            double &best_fval, int &cycle, int &flag, int &iter) {
  if (fast_iter<N, K>(s, cv, bc, Mp, st, bnd, u, best_fval, cycle, flag, iter)) return;
  if constexpr (K < N) fast_chain<N, K + 1>(s, cv, bc, Mp, st, bnd, u, best_fval, cycle, flag, iter);
}

// theta pass shared by the box kernels: sh = Dth theta, sho = uoff + Uth theta, from the warp's TMA tile
// (refilled for the warp's next batch) or straight from the row segments
template <int N, int NU, bool CP>
LMPC_DEVINL void box_theta_pass(const ConstView &cv, const double *Cp, double *tile, uint64_t *tbar_w, uint32_t &tphase,
                                const ThetaSegs &th, int nth, int64_t q, int lane, int64_t qstride, int64_t Ninst,
                                bool use_tiles, double (&sh)[N], double (&sho)[NU]) {
  constexpr int NQ = (N + NU + 1) & ~1;
  SFOR(j, 0, N) sh[j] = 0.0; SEND
  SFOR(i, 0, NU) sho[i] = cv.uoff[i]; SEND
  const int64_t qt = q - lane;                 // first instance of this warp's tile (warp-uniform)
  if (use_tiles && qt + 32 <= Ninst) {
    tile_wait(tbar_w, tphase);
    tphase ^= 1u;
    const double *src = tile + lane * nth;     // stride nth doubles: conflict-free for odd and even nth alike
    if constexpr (!CP) {
#pragma unroll KB_THU
      for (int tt = 0; tt < nth; ++tt) {
        const double tv = src[tt];
        SFOR(j, 0, N) sh[j] = fma(cv.Dth[j * nth + tt], tv, sh[j]); SEND
        SFOR(i, 0, NU) sho[i] = fma(cv.Uth[i * nth + tt], tv, sho[i]); SEND
      }
    } else
#pragma unroll KB_THU
    for (int tt = 0; tt < nth; ++tt) {
      const double tv = src[tt];
      const double2 *c2 = reinterpret_cast<const double2 *>(Cp + tt * NQ);
      double cc[NQ];
      SFOR(h, 0, NQ / 2) const double2 v = c2[h]; cc[2 * h] = v.x; cc[2 * h + 1] = v.y; SEND
      SFOR(j, 0, N) sh[j] = fma(cc[j], tv, sh[j]); SEND
      SFOR(i, 0, NU) sho[i] = fma(cc[N + i], tv, sho[i]); SEND
    }
    __syncwarp();                              // everybody has read the slot: refill it
    if (lane == 0 && qt + qstride + 32 <= Ninst)
      tile_issue(tbar_w, tile, th.ptr[0] + (qt + qstride) * nth, 32u * nth * 8u);
  } else {
    int o = 0;
#pragma unroll 1
    for (int sg = 0; sg < th.nseg; ++sg) {      // one copy of the loop body for all segments
      const int len = th.len[sg];
      const double *src = th.ptr[sg] + q * th.stride[sg];
#pragma unroll KB_THU
      for (int t = 0; t < len; ++t) {
        const double tv = ld_stream(src + t);
        const int tt = o + t;
        SFOR(j, 0, N) sh[j] = fma(cv.Dth[j * nth + tt], tv, sh[j]); SEND
        SFOR(i, 0, NU) sho[i] = fma(cv.Uth[i * nth + tt], tv, sho[i]); SEND
      }
      o += len;
    }
  }
}

// shared prologue of the box kernels: constants by TMA, padded M-bar, transposed [Dth; Uth], theta tiles
template <int N, int NU, bool CP>
struct BoxSmem {
  ConstView cv;
  double *Mp, *Cp, *tile, *end;
};
template <int N, int NU, bool CP>
LMPC_DEVINL BoxSmem<N, NU, CP> box_prologue(double *smem, uint64_t *mbar, uint64_t *tbar, const LdpLayout &lay,
                                            const double *cblk, const ThetaSegs &th, int64_t Ninst, bool use_tiles) {
  BoxSmem<N, NU, CP> b;
  stage_constants_tma(smem, cblk, lay.total, mbar);
  b.cv = make_view(smem, lay);
  constexpr int NP = BoxPad<N>::NP;
  constexpr int NQ = (N + NU + 1) & ~1;
  b.Mp = smem + ((lay.total + 1) & ~1);
  for (int i = threadIdx.x; i < N * NP; i += KB_BLOCK) b.Mp[i] = (i % NP < N) ? b.cv.M[(i / NP) * N + i % NP] : 0.0;
  b.Cp = b.Mp + N * NP;
  const int ncp = CP ? lay.nth * NQ : 0;
  for (int i = threadIdx.x; i < ncp; i += KB_BLOCK) {
    const int tt = i / NQ, c = i % NQ;
    b.Cp[i] = c < N ? b.cv.Dth[c * lay.nth + tt] : (c < N + NU ? b.cv.Uth[(c - N) * lay.nth + tt] : 0.0);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  b.tile = b.Cp + ncp + (size_t)warp * 32 * lay.nth;
  b.end = b.Cp + ncp;
  if (use_tiles) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbar[warp])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t q0 = (int64_t)blockIdx.x * KB_BLOCK + warp * 32;
    if (lane == 0 && q0 + 32 <= Ninst) tile_issue(&tbar[warp], b.tile, th.ptr[0] + q0 * lay.nth, 32u * lay.nth * 8u);
  }
  __syncthreads();
  return b;
}

// instances a CTA of a persistent grid of `nblocks` CTAs can own: passes of KB_BLOCK instances, rounded up
__host__ __device__ inline int64_t box_percta(int64_t Ninst, int nblocks) {
  const int64_t passes = (Ninst + KB_BLOCK - 1) / KB_BLOCK;
  return (passes + nblocks - 1) / nblocks * KB_BLOCK;
}

template <int N, int NU, int MINB, bool CP>
__global__ void __launch_bounds__(KB_BLOCK, MINB)
ldp_box_fast_kernel(LdpLayout lay, LdpSettings st, const __grid_constant__ BoxConst<N, NU> bc,
                    const double *__restrict__ cblk, const __grid_constant__ ThetaSegs th, int64_t Ninst,
                    SolveOut out, int flags, int32_t *hq) {
  const bool use_tiles = flags & 1;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ __align__(8) uint64_t tbar[KB_BLOCK / 32];
  __shared__ int hcnt;                    // instances this CTA has handed over so far
  if (threadIdx.x == 0) hcnt = 0;         // (the prologue ends with a __syncthreads)
  const BoxSmem<N, NU, CP> b = box_prologue<N, NU, CP>(smem, &mbar, tbar, lay, cblk, th, Ninst, use_tiles);
  // hand-over queues: hq = [count per fast CTA (gridDim.x) | gridDim.x queues of (flags >> 1) * KB_BLOCK instance ids]
  // (nothing about them is kept in registers across a solve: the queue address is rebuilt on the rare path)
  const ConstView &cv = b.cv;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t qstride = (int64_t)gridDim.x * KB_BLOCK;
  uint32_t tphase = 0;
  const int nth = lay.nth;
  double *bnd = b.end + (use_tiles ? (size_t)KB_BLOCK * nth : 0) + threadIdx.x;
  int pend = -1;                          // instance handed over in the previous pass, not yet queued
  auto push = [&](unsigned act) {         // one shared-memory atomic per warp and pass that has hand-overs (2 % of the threads)
    const unsigned bal = __ballot_sync(act, pend >= 0);
    if (bal) {
      const int leader = __ffs(bal) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&hcnt, __popc(bal));
      base = __shfl_sync(act, base, leader);
      if (pend >= 0)
        hq[gridDim.x + (size_t)blockIdx.x * ((size_t)(flags >> 1) * KB_BLOCK) + base + __popc(bal & ((1u << lane) - 1u))] = pend;
      pend = -1;
    }
  };
  for (int64_t q = (int64_t)blockIdx.x * KB_BLOCK + threadIdx.x; q < Ninst; q += qstride) {
    push(__activemask());
    double sh[N], sho[NU];
    box_theta_pass<N, NU, CP>(cv, b.Cp, b.tile, &tbar[warp], tphase, th, nth, q, lane, qstride, Ninst, use_tiles, sh, sho);
    double dU[N], dLo[N];
    SFOR(j, 0, N) dU[j] = bc.du[j] + sh[j]; dLo[j] = bc.dl[j] + sh[j]; SEND
    // bnd[(2 j + lower) * KB_BLOCK]: the bound a scan result `code` refers to (column of this thread)
    SFOR(j, 0, N) bnd[(2 * j) * KB_BLOCK] = dU[j]; bnd[(2 * j + 1) * KB_BLOCK] = dLo[j]; SEND
    FastState<N> s;
    s.lowmask = 0u; s.inws = 0u;
    double u[N];
    SFOR(c, 0, N) u[c] = 0.0; SEND
    int flag = EXIT_OPTIMAL, iter = 1, cycle = 0;
    double best_fval = -1.0;
    // iteration 1: empty working set, u = 0, every slack is a bound, the dual objective is 0
    {
      double best = -st.primal_tol;
      int code = -1;
      SFOR(j, 0, N)
        const double su = dU[j] - 0.0, sl = 0.0 - dLo[j];
        const bool tu = su < best;
        const bool tl = !tu && (sl < best);
        best = tu ? su : (tl ? sl : best);
        code = tu ? 2 * j : (tl ? 2 * j + 1 : code);
      SEND
      const int add = code >> 1;
      const bool addlow = code & 1;
      if (add >= 0) {
        if (0.0 - st.progress_tol > best_fval) { best_fval = 0.0; cycle = 0; }
        else cycle++;
        const double addd = bnd[code * KB_BLOCK];
        const double g = cv.G[add * N + add];
        s.Dinv[0] = 1.0 / g; s.dW[0] = addd; s.row[0] = add;
        s.lowmask = addlow ? 1u : 0u;
        s.inws = 1u << add;
        fast_chain<N, 1>(s, cv, bc, b.Mp, st, bnd, u, best_fval, cycle, flag, iter);
      }
    }
    // hand-over: nothing is written to the caller's arrays, the REDO launch solves the instance from scratch.  The id is
    // parked in a register and pushed at the top of the next pass, where the warp is converged again (pushing here, inside
    // the divergent exit, made the early lanes run the next pass's prologue on their own: +8 % instructions).
    if (flag == EXIT_REDO_BOX) { pend = (int)q; continue; }
    SFOR(i, 0, NU)
      double pr[N];
      SFOR(c, 0, N) pr[c] = __dmul_rn(bc.Rinv[i * N + c], u[c]); SEND
      out.U[q * NU + i] = __dadd_rn(tree_sum<N, 0, 1>(pr), sho[i]);
    SEND
    out.flag[q] = flag;
    if (out.iters) out.iters[q] = iter;
    if (out.act_u) {
      // slots 0..k-1 hold the working set; k = popcount(inws)
      unsigned long long au = 0ull, al = 0ull;
      const int k = __popc(s.inws);
      SFOR(i, 0, N)
        if (i < k) {
          if ((s.lowmask >> i) & 1u) al |= 1ull << s.row[i]; else au |= 1ull << s.row[i];
        }
      SEND
      out.act_u[q] = au;
      out.act_l[q] = al;
    }
  }
  push(__activemask());
  __syncthreads();
  if (threadIdx.x == 0) hq[blockIdx.x] = hcnt;
}

// REDO = true: the second launch behind the add-only fast kernel.  The fast kernel's CTAs leave the instances they
// handed over in per-CTA queues (hq: counts, then one queue of local indices per fast CTA; `flags` = number of fast
// CTAs); every CTA here takes the queues c = blockIdx.x, blockIdx.x + gridDim.x, ... and solves their entries from
// scratch 128 at a time, so the lanes stay full although only ~2 % of the batch comes back.  (First r02 version: the
// fast kernel marked the flag array and this launch re-read all N flags and compacted them per CTA -- 40 MB of DRAM
// reads and 62 us per 1e7-instance step; with the queues the launch only touches what it solves.)
template <int N, int NU, int MINB, bool CP = false, bool REDO = false>
__global__ void __launch_bounds__(KB_BLOCK, MINB)
ldp_box_kernel(LdpLayout lay, LdpSettings st, const __grid_constant__ BoxConst<N, NU> bc,
               const double *__restrict__ cblk, const __grid_constant__ ThetaSegs th, int64_t Ninst,
               SolveOut out, int flags, int32_t *hq) {
  const bool use_tiles = !REDO && (flags & 1);
  constexpr bool use_cp = CP;   // compile-time: with both loops in one kernel the gain on short theta is lost
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ __align__(8) uint64_t tbar[KB_BLOCK / 32];
  const BoxSmem<N, NU, CP> bsm = box_prologue<N, NU, CP>(smem, &mbar, tbar, lay, cblk, th, Ninst, use_tiles);
  const ConstView &cv = bsm.cv;
  const double *Mp = bsm.Mp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t qstride = (int64_t)gridDim.x * KB_BLOCK;
  uint32_t tphase = 0;
  const int nth = lay.nth;
  BoxState<N> s;
  auto solve_one = [&](const int64_t q) {
    // ---- mpc_update_parameter + mpc_update_qp + output shift in one pass over theta ----------
    double sh[N], sho[NU];
    box_theta_pass<N, NU, use_cp>(cv, bsm.Cp, bsm.tile, &tbar[warp], tphase, th, nth, q, lane, qstride, Ninst,
                                  use_tiles, sh, sho);
    // theta-dependent bounds, once per solve (mpc_update_qp): d_u = du + sh, d_l = dl + sh
    double dU[N], dLo[N];
    SFOR(j, 0, N) dU[j] = bc.du[j] + sh[j]; dLo[j] = bc.dl[j] + sh[j]; SEND
    // ---- dual active-set loop, cold start --------------------------------------------------
    box_reset<N>(s);
    double u[N];
    SFOR(c, 0, N) u[c] = 0.0; SEND
    int flag = EXIT_ITERLIMIT, iter = 1, cycle = 0;
    double best_fval = -1.0, fval = 0.0;
    bool done = false;
    // Iteration 1 is the same for every thread: empty working set, u = 0, so every slack is
    // just a bound (d_u - 0, 0 - d_l), the dual objective is 0, and the first row is appended at
    // the static position 0.  Same values as the generic iteration, none of its padding.
    if (st.iter_limit > 2) {
      double best = -st.primal_tol;
      int code = -1;
      SFOR(j, 0, N)
        const double su = dU[j] - 0.0, sl = 0.0 - dLo[j];
        const bool tu = su < best;
        const bool tl = !tu && (sl < best);
        best = tu ? su : (tl ? sl : best);
        code = tu ? 2 * j : (tl ? 2 * j + 1 : code);
      SEND
      const int add = code >> 1;
      const bool addlow = code & 1;
      if (add < 0) { flag = EXIT_OPTIMAL; done = true; }
      else {
        if (fval - st.progress_tol > best_fval) { best_fval = fval; cycle = 0; }
        else cycle++;
        double addd = 0.0;
        SFOR(j, 0, N) addd = (add == j) ? (addlow ? dLo[j] : dU[j]) : addd; SEND
        const double g = cv.G[add * N + add];
        s.D[0] = g; s.Dinv[0] = 1.0 / g; s.dW[0] = addd; s.row[0] = add;
        s.lowmask = addlow ? 1u : 0u;
        s.inws = 1u << add;
        s.k = 1;
        iter = 2;
      }
    }
    // Iterations 2, 3, 4: a thread that has only ever added rows holds exactly K = iter - 1 of
    // them, so its iteration can use exact loop bounds and static slots (box_iter_static).  A
    // thread whose multipliers block skips these and is picked up, from the same state, by the
    // generic loop below.  (Measured: static K up to 3 -> 8.8e9 solves/s at s = 1; extending to
    // K = 4, 5 is slower, 7.4e9, because the rare blocked threads then trail the others by more
    // generic iterations; blocking happens late, at |W| >= 3.)
    if (!done && st.iter_limit > N + 2) {
      IterCtx cx{best_fval, fval, cycle, flag, iter, done};
      if constexpr (N >= 2) box_iter_static<N, 1>(s, cv, bc, Mp, st, dU, dLo, u, cx);
      if constexpr (N >= 3) box_iter_static<N, 2>(s, cv, bc, Mp, st, dU, dLo, u, cx);
      if constexpr (N >= 4) box_iter_static<N, 3>(s, cv, bc, Mp, st, dU, dLo, u, cx);
      best_fval = cx.best_fval; fval = cx.fval; cycle = cx.cycle; flag = cx.flag; iter = cx.iter; done = cx.done;
    }
    for (; !done && iter < st.iter_limit; ++iter) {
      // 1. CSP over the full capacity (slots >= k are zero rows -> exact zeros)
      double z[N];
      SFOR(i, 0, N)
        double zi = -s.dW[i];
        SFOR(c, 0, i) zi = fma(-s.L[LIDX(i, c)], z[c], zi); SEND
        z[i] = zi;
      SEND
      SFOR(i, 0, N) s.lams[i] = z[i] * s.Dinv[i]; SEND
      SFOR_DOWN(j, N, 1)
        const double xj = s.lams[j];
        SFOR(i, 0, j) s.lams[i] = fma(-s.L[LIDX(j, i)], xj, s.lams[i]); SEND
      SEND
      // 2. blocking set (an unused slot has lam* == 0 and can never block)
      unsigned blkmask = 0u;
      SFOR(i, 0, N)
        const bool low = (s.lowmask >> i) & 1u;
        const bool blk = low ? (s.lams[i] > st.dual_tol) : (s.lams[i] < -st.dual_tol);
        blkmask |= blk ? (1u << i) : 0u;
      SEND
      if (blkmask == 0u) {
        // 3. primal iterate, dual objective, most violated row
        SFOR(c, 0, N) u[c] = 0.0; SEND
        SFOR(i, 0, N) box_axpy_row<N>(Mp, s.row[i], s.lams[i], u); SEND
        double pu[N];
        SFOR(c, 0, N) pu[c] = __dmul_rn(u[c], u[c]); SEND
        fval = tree_sum<N, 0, 1>(pu);
        if (fval > st.fval_bound) { flag = EXIT_INFEASIBLE; break; }
        // every row already in the working set: the scan cannot find a violated free row
        if (s.inws == (1u << N) - 1u) { flag = EXIT_OPTIMAL; break; }
        double best = -st.primal_tol;
        int code = -1;                       // 2 * row + (1 if lower)
        SFOR(j, 0, N)
          // row j of the normalised R^-1 is zero before column j
          double mu = 0.0;
          SFOR(c, j, N) mu = fma(bc.M[j * N + c], u[c], mu); SEND
          const bool free_row = !((s.inws >> j) & 1u);
          const double su = dU[j] - mu, sl = mu - dLo[j];
          const bool tu = free_row && (su < best);
          const bool tl = free_row && !tu && (sl < best);
          best = tu ? su : (tl ? sl : best);
          code = tu ? 2 * j : (tl ? 2 * j + 1 : code);
        SEND
        const int add = code >> 1;
        const bool addlow = code & 1;
        double addd = 0.0;
        SFOR(j, 0, N) addd = (add == j) ? (addlow ? dLo[j] : dU[j]) : addd; SEND
        if (add < 0) { flag = EXIT_OPTIMAL; break; }
        if (fval - st.progress_tol > best_fval) { best_fval = fval; cycle = 0; }
        else if (cycle++ > st.cycle_tol) { flag = EXIT_CYCLE; break; }
        if (s.k >= N) { flag = EXIT_CAPACITY; break; }
        SFOR(i, 0, N) s.lam[i] = s.lams[i]; SEND
        box_add<N>(s, cv.G, add, addlow, addd);
      } else {
        // 4. ratio test (first minimum wins), step, remove
        double alpha = CUDART_INF;
        int rm = -1;
        SFOR(i, 0, N)
          if ((blkmask >> i) & 1u) {
            const double a = -s.lam[i] / (s.lams[i] - s.lam[i]);
            if (a < alpha) { alpha = a; rm = i; }
          }
        SEND
        if (!(alpha < CUDART_INF)) { flag = EXIT_CYCLE; break; }
        SFOR(i, 0, N) s.lam[i] = fma(alpha, s.lams[i] - s.lam[i], s.lam[i]); SEND
        box_remove<N>(s, rm);
      }
    }
    // ---- mpc_get_solution -------------------------------------------------------------------
    SFOR(i, 0, NU)
      double pr[N];
      SFOR(c, 0, N) pr[c] = __dmul_rn(bc.Rinv[i * N + c], u[c]); SEND
      out.U[q * NU + i] = __dadd_rn(tree_sum<N, 0, 1>(pr), sho[i]);
    SEND
    out.flag[q] = flag;
    if (out.iters) out.iters[q] = iter;
    if (out.act_u) {
      unsigned long long au = 0ull, al = 0ull;
      SFOR(i, 0, N)
        if (i < s.k) {
          if ((s.lowmask >> i) & 1u) al |= 1ull << s.row[i]; else au |= 1ull << s.row[i];
        }
      SEND
      out.act_u[q] = au;
      out.act_l[q] = al;
    }
  };
  if constexpr (!REDO) {
    for (int64_t q = (int64_t)blockIdx.x * KB_BLOCK + threadIdx.x; q < Ninst; q += qstride) solve_one(q);
  } else {
    const int nfast = flags;                                  // CTAs of the fast launch
    const int64_t percta = box_percta(Ninst, nfast);
    for (int cq = blockIdx.x; cq < nfast; cq += gridDim.x) {
      const int cnt = hq[cq];
      const int32_t *qv = hq + nfast + (size_t)cq * percta;
      for (int i = threadIdx.x; i < cnt; i += KB_BLOCK) solve_one((int64_t)qv[i]);
    }
  }
}

#ifndef KB_MINB
#define KB_MINB 3
#endif
#ifndef KB_FAST
#define KB_FAST 1      // 0: the generic kernel alone, as before r02
#endif
#ifndef KB_FAST_MIN_BATCH
#define KB_FAST_MIN_BATCH 200000
#endif
#ifndef KB_FAST_MINB
#define KB_FAST_MINB 3
#endif
// resident CTAs per SM: 3 up to n = 5 (145-159 registers, no spills); from n = 6 on the state no longer fits
// 168 registers (80 - 1700 B of spills per thread), so two CTAs with up to 255 registers each
template <int N, int NU, int MINB = (N <= 5 ? KB_MINB : 2)>
int launch_box(const LdpLayout &lay, const LdpSettings &st, const double *cblk, const double *hblk,
               const ThetaSegs &th,
               int64_t Ninst, const SolveOut &out, int sm_count, cudaStream_t stream) {
  // the add-only fast kernel + the REDO launch of the generic one, or the generic kernel alone when the
  // settings could stop a solve inside the fast kernel's N + 1 iterations
  // (small batches: the second launch costs ~8 us, more than the fast kernel saves below ~2e5 instances)
  // LMPC_BOX_FAST_MIN_BATCH overrides the threshold (tests and sanitizer runs push small batches through both launches)
  const char *env_min = getenv("LMPC_BOX_FAST_MIN_BATCH");
  const int64_t min_batch = env_min ? atoll(env_min) : (int64_t)KB_FAST_MIN_BATCH;
  // (instance ids travel through the hand-over queues as 32-bit integers)
  const bool fast = KB_FAST && st.iter_limit > N + 2 && Ninst >= min_batch && Ninst < (1ll << 31) - 1;
  constexpr int FMINB = N <= 5 ? KB_FAST_MINB : 2;
  using kern_t = void (*)(LdpLayout, LdpSettings, const BoxConst<N, NU>, const double *, const ThetaSegs, int64_t,
                          SolveOut, int, int32_t *);
  kern_t kern = fast ? (kern_t)ldp_box_fast_kernel<N, NU, FMINB, false> : (kern_t)ldp_box_kernel<N, NU, MINB, false>;
  kern_t kern_cp = fast ? (kern_t)ldp_box_fast_kernel<N, NU, FMINB, true> : (kern_t)ldp_box_kernel<N, NU, MINB, true>;
  const size_t smem_base = ((size_t)((lay.total + 1) & ~1) + N * BoxPad<N>::NP) * sizeof(double);
  size_t smem = smem_base;
  // one 32-theta slot per warp when Theta is a single contiguous, 16-byte aligned block
  const int use_tiles = th.nseg == 1 && th.len[0] == lay.nth && th.stride[0] == lay.nth &&
                        (reinterpret_cast<uintptr_t>(th.ptr[0]) & 15u) == 0 && Ninst >= 32;
  if (use_tiles) smem += (size_t)KB_BLOCK * lay.nth * sizeof(double);
  if (fast) smem += (size_t)KB_BLOCK * 2 * N * sizeof(double);   // per-thread columns of dU / dLo
  // transposed constants for the theta mat-vec: taken only when they do not cost a resident CTA
  const size_t smem_cp = smem + (size_t)lay.nth * ((N + NU + 1) & ~1) * sizeof(double);
  int use_cp = 0;
  if (use_tiles && smem_cp <= 227u * 1024u &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
      cudaFuncSetAttribute(kern_cp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cp) == cudaSuccess) {
    int a = 0, b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, kern, KB_BLOCK, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern_cp, KB_BLOCK, smem_cp);
    if (b >= a && b >= 1) { use_cp = 1; smem = smem_cp; kern = kern_cp; }
  }
  (void)use_cp;
  BoxConst<N, NU> bc;
  for (int i = 0; i < N * N; ++i) bc.M[i] = hblk[lay.oM + i];
  for (int i = 0; i < N; ++i) { bc.du[i] = hblk[lay.odu + i]; bc.dl[i] = hblk[lay.odl + i]; }
  for (int i = 0; i < NU * N; ++i) bc.Rinv[i] = hblk[lay.oRinv + i];
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, KB_BLOCK, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = (Ninst + KB_BLOCK - 1) / KB_BLOCK;
  const int64_t cap = (int64_t)sm_count * per_sm;   // persistent: whole resident waves over the SMs
  if (blocks > cap) blocks = cap;
  // hand-over queues of the fast kernel's CTAs (stream-ordered scratch, kept in the pool between calls)
  int32_t *hq = nullptr;
  if (fast) {
    int dev = 0;
    cudaGetDevice(&dev);
    static bool pool_kept[64] = {false};
    if (dev < 64 && !pool_kept[dev]) {   // keep freed scratch in the pool instead of returning it to the OS every call
      cudaMemPool_t mp;
      unsigned long long keep = ~0ull;
      if (cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess) cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
      pool_kept[dev] = true;
    }
    const size_t words = (size_t)blocks * (1 + (size_t)box_percta(Ninst, (int)blocks));
    if (cudaMallocAsync((void **)&hq, words * sizeof(int32_t), stream) != cudaSuccess) {
      // no scratch: the generic kernel alone
      cudaGetLastError();
      hq = nullptr;
      kern_t gk = (kern_t)ldp_box_kernel<N, NU, MINB, false>;
      cudaError_t e2 = cudaFuncSetAttribute(gk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_base);
      if (e2 != cudaSuccess) return (int)e2;
      int ps = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, gk, KB_BLOCK, smem_base);
      if (ps < 1) ps = 1;
      int64_t gb = (Ninst + KB_BLOCK - 1) / KB_BLOCK;
      if (gb > (int64_t)sm_count * ps) gb = (int64_t)sm_count * ps;
      gk<<<(unsigned)gb, KB_BLOCK, smem_base, stream>>>(lay, st, bc, cblk, th, Ninst, out, 0, nullptr);
      return (int)cudaGetLastError();
    }
  }
  const int kflags = use_tiles | (fast ? (int)(box_percta(Ninst, (int)blocks) / KB_BLOCK) << 1 : 0);
  kern<<<(unsigned)blocks, KB_BLOCK, smem, stream>>>(lay, st, bc, cblk, th, Ninst, out, kflags, hq);
  e = cudaGetLastError();
  if (e != cudaSuccess || !fast) { if (hq) cudaFreeAsync(hq, stream); return (int)e; }
  // second launch: whatever the fast kernel's CTAs left in their queues
  kern_t redo = (kern_t)ldp_box_kernel<N, NU, MINB, false, true>;
  const size_t smem_redo = smem_base;
  e = cudaFuncSetAttribute(redo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_redo);
  if (e != cudaSuccess) return (int)e;
  per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, redo, KB_BLOCK, smem_redo);
  if (per_sm < 1) per_sm = 1;
  int64_t rblocks = (Ninst + KB_BLOCK - 1) / KB_BLOCK;
  if (rblocks > (int64_t)sm_count * per_sm) rblocks = (int64_t)sm_count * per_sm;
  if (rblocks > blocks) rblocks = blocks;
  redo<<<(unsigned)rblocks, KB_BLOCK, smem_redo, stream>>>(lay, st, bc, cblk, th, Ninst, out, (int)blocks, hq);
  lmpc_count_kernel(1);
  cudaFreeAsync(hq, stream);
  return (int)cudaGetLastError();
}

}  // namespace

bool ldp_box_supports(const LdpLayout &lay) {
  return lay.plain && lay.m == lay.n && lay.n >= 1 && lay.n <= 8 && lay.nu >= 1 && lay.nu <= 2 &&
         lay.nu <= lay.n && lay.cap <= lay.n;
}

#define KB_CASE(NN)                                                                              \
  case NN:                                                                                       \
    if (lay.nu == 1) return launch_box<NN, 1>(lay, st, cblk, hblk, th, N, out, sm_count, s);           \
    if constexpr (NN >= 2) return launch_box<NN, 2>(lay, st, cblk, hblk, th, N, out, sm_count, s);     \
    return (int)cudaErrorInvalidValue;

int launch_ldp_box(const LdpLayout &lay, const LdpSettings &st, const double *cblk, const double *hblk,
                   const ThetaSegs &th, int64_t N, const SolveOut &out, int sm_count, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (N <= 0) return 0;
  switch (lay.n) {
#ifdef KB_ONLY5          // A/B builds: the headline instantiation only (compiles in seconds)
    KB_CASE(5)
#else
    KB_CASE(1) KB_CASE(2) KB_CASE(3) KB_CASE(4) KB_CASE(5) KB_CASE(6) KB_CASE(7) KB_CASE(8)
#endif
    default: return (int)cudaErrorInvalidValue;
  }
}
