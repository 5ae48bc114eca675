// K3: the large-problem dual active-set LDP kernel (32 < n <= 256, capacity <= 256, m <= 512),
// e.g. the chain-of-masses controller (n = m = 180) and AFTI-16 with output bounds on the
// whole horizon (m = 38, capacity 38).
//
// Mapping.  One warp drives one QP; a lane owns T working-set positions (i = lane + 32 t),
// T decision variables and 2T constraint rows, all in statically indexed register tiles.
// The unit-lower factor L is PACKED (row i holds i entries at i(i-1)/2) in the warp's
// shared-memory region -- 129 KB at capacity 180, which is why a CTA holds as many QPs as
// shared memory allows (1 at n = 180, 8 at capacity 38) -- and Mbar / Mbar' / Gram stay in
// global memory (L2-resident, read coalesced).  Triangular solves are column sweeps over the
// register tiles (one shuffle broadcast per step), blocked by 32 so that every register index
// is a compile-time constant.  Cross-element sums follow the canonical lane order of the
// oracle (element i on lane i mod 32, then xor-butterfly), so results are bit-identical to
// oracle/ldp_oracle.c and to K1/K2.
//
// Reference path replaced: solve(mpc, theta) -> DAQP.update/solve (src/utils.jl:268-283).
#include "ldp_device.cuh"
#include <math_constants.h>

namespace {

constexpr unsigned FULL = 0xffffffffu;

LMPC_DEVINL double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }
LMPC_DEVINL double bfly_sum(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(FULL, v, off));
  return v;
}
// warp-wide minimum of v (NaN-free): two 32-bit REDUX operations on an order-preserving integer
// image of the double (-0.0 folded onto +0.0) instead of a 5-step shuffle butterfly; the minimum
// of a set does not depend on how it is found
LMPC_DEVINL double bfly_min(double v) {
  const long long b = __double_as_longlong(__dadd_rn(v, 0.0));
  const unsigned long long key = (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned hmin = __reduce_min_sync(FULL, hi);
  const unsigned lmin = __reduce_min_sync(FULL, hi == hmin ? lo : 0xffffffffu);
  const unsigned long long kmin = ((unsigned long long)hmin << 32) | lmin;
  const long long bm = (long long)kmin;
  return __longlong_as_double(bm < 0 ? (bm ^ (long long)0x8000000000000000ull) : ~bm);
}
LMPC_DEVINL double bfly_max(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const double o = __shfl_xor_sync(FULL, v, off);
    v = (o > v) ? o : v;
  }
  return v;
}
LMPC_DEVINL int bfly_min_i(int v) { return __reduce_min_sync(FULL, v); }

template <int T>
struct BigQp {
  double D[T], Dinv[T], lam[T], lams[T], dW[T];
  int row[T];
  unsigned low, imm, soft;   // bit t: flag of position lane + 32 t
  unsigned inws;             // bit t2: row lane + 32 t2 is in the working set (t2 < 2T)
  int k;
  bool sing;
};

// SQ_: the warp's L tile is a square cap x LS array indexed by position (rows hold exact zeros on
// and above the diagonal, as in ldp_warp.cu) instead of the packed lower triangle; chosen by the
// launcher when the square tile still leaves a full set of warps per CTA.
template <bool SQ_>
struct BigCtxT {
  static constexpr bool SQ = SQ_;
  int LS;                            // SQ: row stride of the tile (odd)
  const double *M, *MT, *G;          // global (L2) copies
  ConstView cv;                      // shared-memory copies of the small sections
  const LdpLayout *lay;
  const LdpSettings *st;
  double *Lw;      // packed L
  double *uw;      // n
  double *thw;     // nth
  double *lamw;    // cap: lam* broadcast copy
  int *roww;       // cap: row ids broadcast copy
  unsigned char *inwsb;   // CTA mode: m flags "row is in the working set" (else null)
  int lane;
};
using BigCtx = BigCtxT<false>;

#ifndef K3_BC_UNROLL
#define K3_BC_UNROLL 8
#endif
constexpr int K3_BCU = K3_BC_UNROLL;
#ifndef K3_INTERLEAVE
#define K3_INTERLEAVE 0
#endif   // unroll of the CTA-mode u / slack-scan loops (loads in flight per thread)
// rows per TMA batch of the CTA kernel's Mbar / Mbar' stream (two batches in flight: 2 * K3_RB row slots)
// Measured (r02, chain n = 180): a TMA bulk copy of one 1.4 KB row lands ~390 cycles after issue while the fma
// chains consume a row every 8 cycles, so a ring that fits next to the factor (8 rows) starves: 3.3e4 solves/s
// against 7.9e4 with direct loads.  Kept for problems whose constants leave room for a deep ring; off by default.
#ifndef K3_TRIM_U
#define K3_TRIM_U 0
#endif
#ifndef K3_PAIR
#define K3_PAIR 1            // a thread's two columns / rows run as interleaved chains
#endif
#ifndef K3_PAIR_B
#define K3_PAIR_B 8          // block of the paired chains (2 * K3_PAIR_B loads in flight per thread)
#endif
#ifndef K3_USE_RING
#define K3_USE_RING 0
#endif
#ifndef K3_RING_ROWS
#define K3_RING_ROWS 4
#endif
constexpr int K3_RB = K3_RING_ROWS;
LMPC_DEVINL void bulk_expect(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
LMPC_DEVINL void bulk_copy(uint64_t *bar, double *dst, const double *src, uint32_t bytes) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#define PIDX(i, c) ((size_t)(i) * ((i)-1) / 2 + (c))
#define TFOR(t) _Pragma("unroll") for (int t = 0; t < T; ++t)

template <int T>
LMPC_DEVINL void bq_reset(BigQp<T> &s) {
  TFOR(t) { s.D[t] = 1.0; s.Dinv[t] = 1.0; s.lam[t] = 0.0; s.lams[t] = 0.0; s.dW[t] = 0.0; s.row[t] = 0; }
  s.low = s.imm = s.soft = 0u; s.inws = 0u; s.k = 0; s.sing = false;
}

// value held at position j (tile j>>5, lane j&31), broadcast to the warp; j is warp-uniform
template <int T>
LMPC_DEVINL double bcast_pos(const double (&v)[T], int j) {
  double r = 0.0;
  TFOR(t) { const double x = shfl_d(v[t], j & 31); r = ((j >> 5) == t) ? x : r; }
  return r;
}
template <int T>
LMPC_DEVINL int bcast_pos_i(const int (&v)[T], int j) {
  int r = 0;
  TFOR(t) { const int x = __shfl_sync(FULL, v[t], j & 31); r = ((j >> 5) == t) ? x : r; }
  return r;
}

// z <- L^-1 z on positions [0, kk): column sweep, blocked by tile so register indices are static
template <int T, class C>
LMPC_DEVINL void sweep_fwd(const C &c, double (&z)[T], int kk) {
  if constexpr (C::SQ) {
    // unpredicated: rows at or above the diagonal multiply by stored zeros; idle positions re-read row 0
    const double *rowp[T];
    TFOR(t) { const int i = c.lane + 32 * t; rowp[t] = c.Lw + (size_t)(i < kk ? i : 0) * c.LS; }
#pragma unroll
    for (int tt = 0; tt < T; ++tt) {
      const int jend = min(32, kk - 1 - 32 * tt);
#pragma unroll 4
      for (int jj = 0; jj < jend; ++jj) {
        const double zj = shfl_d(z[tt], jj);
#pragma unroll
        for (int t = tt; t < T; ++t) z[t] = fma(-rowp[t][32 * tt + jj], zj, z[t]);
      }
    }
    return;
  }
#pragma unroll
  for (int tt = 0; tt < T; ++tt) {
    const int jend = min(32, kk - 1 - 32 * tt);
    for (int jj = 0; jj < jend; ++jj) {
      const int j = 32 * tt + jj;
      const double zj = shfl_d(z[tt], jj);
#pragma unroll
      for (int t = tt; t < T; ++t) {
        const int i = c.lane + 32 * t;
        if (i > j && i < kk) z[t] = fma(-c.Lw[PIDX(i, j)], zj, z[t]);
      }
    }
  }
}
// x <- L^-T x on positions [0, kk), descending
template <int T, class C>
LMPC_DEVINL void sweep_bwd(const C &c, double (&x)[T], int kk) {
  if constexpr (C::SQ) {
    const double *colp[T];
    TFOR(t) { const int i = c.lane + 32 * t; colp[t] = c.Lw + (i < c.LS ? i : c.LS - 1); }
#pragma unroll
    for (int tt = T - 1; tt >= 0; --tt) {
      const int jhi = min(31, kk - 1 - 32 * tt);
#pragma unroll 4
      for (int jj = jhi; jj >= 0; --jj) {
        const int j = 32 * tt + jj;
        if (j == 0) break;
        const double xj = shfl_d(x[tt], jj);
#pragma unroll
        for (int t = 0; t <= tt; ++t) x[t] = fma(-colp[t][(size_t)j * c.LS], xj, x[t]);
      }
    }
    return;
  }
#pragma unroll
  for (int tt = T - 1; tt >= 0; --tt) {
    const int jhi = min(31, kk - 1 - 32 * tt);
    for (int jj = jhi; jj >= 0; --jj) {
      const int j = 32 * tt + jj;
      if (j == 0) break;
      const double xj = shfl_d(x[tt], jj);
#pragma unroll
      for (int t = 0; t <= tt; ++t) {
        const int i = c.lane + 32 * t;
        if (i < j) x[t] = fma(-c.Lw[PIDX(j, i)], xj, x[t]);
      }
    }
  }
}

// canonical lane sum of a[i] * b[i] over positions i < len (element i on lane i mod 32)
template <int T, class C>
LMPC_DEVINL double lanesum_pos(const C &c, const double (&a)[T], const double (&b)[T], int len) {
  double s = (c.lane < len) ? __dmul_rn(a[0], b[0]) : 0.0;
#pragma unroll
  for (int t = 1; t < T; ++t)
    if (c.lane + 32 * t < len) s = fma(a[t], b[t], s);
  return bfly_sum(s);
}

// append row j (oracle: ldl_add)
// swept: y = L^-1 g_W already computed by the whole CTA (coop_sweep_fwd), read it from there
template <int T, class C>
LMPC_DEVINL void bq_add(const C &c, BigQp<T> &s, int j, bool lower, double dval, const double *swept = nullptr) {
  const int m = c.lay->m, k = s.k, lane = c.lane;
  const double *g = c.G + (size_t)j * m;
  const int sj = c.cv.sense[j];
  double dnew = g[j];
  if (sj & S_SOFT) dnew += c.st->rho_soft;
  double y[T], l[T];
  if (swept) {
    TFOR(t) y[t] = (lane + 32 * t < k) ? swept[lane + 32 * t] : 0.0;
  } else {
    TFOR(t) y[t] = (lane + 32 * t < k) ? g[s.row[t]] : 0.0;
    sweep_fwd<T>(c, y, k);
  }
  TFOR(t) l[t] = (lane + 32 * t < k) ? y[t] * s.Dinv[t] : 0.0;
  dnew = __dsub_rn(dnew, lanesum_pos<T>(c, y, l, k));
  const bool sing = dnew < c.st->zero_tol;
  if (sing) dnew = 0.0;
  const double dinv = sing ? 0.0 : 1.0 / dnew;
  TFOR(t) {
    const int i = lane + 32 * t;
    if constexpr (C::SQ) { if (i < c.LS) c.Lw[(size_t)k * c.LS + i] = (i < k) ? l[t] : 0.0; }   // whole row, zeros from the diagonal on
    else if (i < k) c.Lw[PIDX(k, i)] = l[t];
    if (i == k) {
      s.D[t] = dnew; s.Dinv[t] = dinv; s.lam[t] = 0.0; s.dW[t] = dval; s.row[t] = j;
      const unsigned bit = 1u << t;
      s.low = lower ? (s.low | bit) : (s.low & ~bit);
      s.imm = (sj & S_IMMUTABLE) ? (s.imm | bit) : (s.imm & ~bit);
      s.soft = (sj & S_SOFT) ? (s.soft | bit) : (s.soft & ~bit);
      c.roww[k] = j;
    }
  }
  if ((j & 31) == lane) s.inws |= 1u << (j >> 5);
  if (c.inwsb && lane == 0) c.inwsb[j] = 1;
  s.sing = sing;
  s.k = k + 1;
  __syncwarp();
}

// take back a trial add of row j (warm start from masks: the row would have been singular)
template <int T, class C>
LMPC_DEVINL void bq_undo_add(const C &c, BigQp<T> &s, int j) {
  s.k -= 1;
  if ((j & 31) == c.lane) s.inws &= ~(1u << (j >> 5));
  if (c.inwsb && c.lane == 0) c.inwsb[j] = 0;
  s.sing = false;
  __syncwarp();
}

// position i takes the value of position i + 1 (for r <= i < k - 1)
template <int T>
LMPC_DEVINL void shift_down(double (&v)[T], int lane, int r, int k) {
  double nxt[T];
  TFOR(t) {
    const double same = __shfl_down_sync(FULL, v[t], 1);
    const double wrap = (t + 1 < T) ? shfl_d(v[(t + 1 < T) ? t + 1 : t], 0) : 0.0;
    nxt[t] = (lane < 31) ? same : wrap;
  }
  TFOR(t) { const int i = lane + 32 * t; if (i >= r && i + 1 < k) v[t] = nxt[t]; }
}
template <int T>
LMPC_DEVINL void shift_down_i(int (&v)[T], int lane, int r, int k) {
  int nxt[T];
  TFOR(t) {
    const int same = __shfl_down_sync(FULL, v[t], 1);
    const int wrap = (t + 1 < T) ? __shfl_sync(FULL, v[(t + 1 < T) ? t + 1 : t], 0) : 0;
    nxt[t] = (lane < 31) ? same : wrap;
  }
  TFOR(t) { const int i = lane + 32 * t; if (i >= r && i + 1 < k) v[t] = nxt[t]; }
}
template <int T>
LMPC_DEVINL unsigned shift_down_bits(unsigned bits, int lane, int r, int k) {
  int v[T];
  TFOR(t) v[t] = (bits >> t) & 1u;
  shift_down_i<T>(v, lane, r, k);
  unsigned o = 0u;
  TFOR(t) o |= (unsigned)v[t] << t;
  return o;
}

// delete position r (oracle: ldl_remove): packed rows are physically shifted, as in the oracle
template <int T, class C>
LMPC_DEVINL void bq_remove(const C &c, BigQp<T> &s, int r) {
  const int lane = c.lane;
  int k = s.k;
  double alpha = bcast_pos<T>(s.D, r);
  const int rrow = bcast_pos_i<T>(s.row, r);
  double wv[T];
  if constexpr (C::SQ) {
    // square tile (see ldp_warp.cu: wq_remove): positions above r move down one, L moves up a
    // row and, right of r, a column while the rank-1 update rewrites it; zeros on and above the
    // diagonal are never touched
    const int LS = c.LS;
    TFOR(t) { const int i = lane + 32 * t; wv[t] = (i > r && i < k) ? c.Lw[(size_t)i * LS + r] : 0.0; }
    if ((rrow & 31) == lane) s.inws &= ~(1u << (rrow >> 5));
    if (c.inwsb && lane == 0) c.inwsb[rrow] = 0;
    shift_down<T>(s.D, lane, r, k); shift_down<T>(s.Dinv, lane, r, k); shift_down<T>(s.lam, lane, r, k);
    shift_down<T>(s.dW, lane, r, k); shift_down<T>(wv, lane, r, k); shift_down_i<T>(s.row, lane, r, k);
    s.low = shift_down_bits<T>(s.low, lane, r, k);
    s.imm = shift_down_bits<T>(s.imm, lane, r, k);
    s.soft = shift_down_bits<T>(s.soft, lane, r, k);
    k -= 1;
    TFOR(t) { const int i = lane + 32 * t; if (i >= r && i < k) c.roww[i] = s.row[t]; }
#pragma unroll 1
    for (int cc = 0; cc < r; ++cc) {           // columns left of r: same values, one row up
      double v[T];
      TFOR(t) { const int i = lane + 32 * t; v[t] = (i >= r && i < k) ? c.Lw[(size_t)(i + 1) * LS + cc] : 0.0; }
      __syncwarp();
      TFOR(t) { const int i = lane + 32 * t; if (i >= r && i < k) c.Lw[(size_t)i * LS + cc] = v[t]; }
    }
    bool sing = false;
#pragma unroll
    for (int tt = 0; tt < T; ++tt) {
#pragma unroll 1
      for (int jj = 0; jj < 32; ++jj) {
        const int j = 32 * tt + jj;
        if (j < r) continue;
        if (j >= k) break;
        const double pj = shfl_d(wv[tt], jj);
        const double Dj = shfl_d(s.D[tt], jj);
        const double tq = alpha * pj;
        double dbar = fma(tq, pj, Dj);
        double rinv = 1.0 / dbar;
        const double beta = tq * rinv;
        alpha = (Dj * alpha) * rinv;
        if (j == k - 1 && !(dbar >= c.st->zero_tol)) { sing = true; dbar = 0.0; rinv = 0.0; }
        if (lane == jj) { s.D[tt] = dbar; s.Dinv[tt] = rinv; }
        double lij[T];
#pragma unroll
        for (int t = tt; t < T; ++t) {           // new L[i][j] from old L[i+1][j+1]
          const int i = lane + 32 * t;
          lij[t] = (i > j && i < k) ? c.Lw[(size_t)(i + 1) * LS + j + 1] : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int t = tt; t < T; ++t) {
          const int i = lane + 32 * t;
          if (i > j && i < k) {
            wv[t] = fma(-pj, lij[t], wv[t]);
            c.Lw[(size_t)i * LS + j] = fma(beta, wv[t], lij[t]);
          }
        }
      }
    }
    s.k = k;
    s.sing = sing;
    __syncwarp();
    return;
  }
  TFOR(t) { const int i = lane + 32 * t; wv[t] = (i > r && i < k) ? c.Lw[PIDX(i, r)] : 0.0; }
  if ((rrow & 31) == lane) s.inws &= ~(1u << (rrow >> 5));
  if (c.inwsb && lane == 0) c.inwsb[rrow] = 0;
  // rows r+1 .. k-1 move up one row and lose column r
  for (int p = r; p + 1 < k; ++p) {
    double tmp[T];
    TFOR(t) { const int cc = lane + 32 * t; tmp[t] = (cc < p) ? c.Lw[PIDX(p + 1, cc < r ? cc : cc + 1)] : 0.0; }
    __syncwarp();
    TFOR(t) { const int cc = lane + 32 * t; if (cc < p) c.Lw[PIDX(p, cc)] = tmp[t]; }
    __syncwarp();
  }
  shift_down<T>(s.D, lane, r, k); shift_down<T>(s.Dinv, lane, r, k); shift_down<T>(s.lam, lane, r, k);
  shift_down<T>(s.dW, lane, r, k); shift_down<T>(wv, lane, r, k); shift_down_i<T>(s.row, lane, r, k);
  s.low = shift_down_bits<T>(s.low, lane, r, k);
  s.imm = shift_down_bits<T>(s.imm, lane, r, k);
  s.soft = shift_down_bits<T>(s.soft, lane, r, k);
  k -= 1;
  TFOR(t) { const int i = lane + 32 * t; if (i >= r && i < k) c.roww[i] = s.row[t]; }
  bool sing = false;
#pragma unroll
  for (int tt = 0; tt < T; ++tt) {
    for (int jj = 0; jj < 32; ++jj) {
      const int j = 32 * tt + jj;
      if (j < r) continue;
      if (j >= k) break;
      const double pj = shfl_d(wv[tt], jj);
      const double Dj = shfl_d(s.D[tt], jj);
      const double tq = alpha * pj;
      double dbar = fma(tq, pj, Dj);
      double rinv = 1.0 / dbar;
      const double beta = tq * rinv;
      alpha = (Dj * alpha) * rinv;
      if (j == k - 1 && !(dbar >= c.st->zero_tol)) { sing = true; dbar = 0.0; rinv = 0.0; }
      if (lane == jj) { s.D[tt] = dbar; s.Dinv[tt] = rinv; }
#pragma unroll
      for (int t = tt; t < T; ++t) {
        const int i = lane + 32 * t;
        if (i > j && i < k) {
          double *lp = c.Lw + PIDX(i, j);
          const double lij = *lp;
          wv[t] = fma(-pj, lij, wv[t]);
          *lp = fma(beta, wv[t], lij);
        }
      }
    }
  }
  s.k = k;
  s.sing = sing;
  __syncwarp();
}

// Problems with SOFT rows only (oracle: rebuild; see ldp_warp.cu: wq_rebuild): refactorise from scratch
// with the hard rows first, in their current order, and the soft rows last, multipliers kept.  A hard
// row that is dependent on the hard block goes to the very end with a zero pivot; a second one:
// returns EXIT_CYCLE.  Warp-local (in CTA mode warp 0 runs it alone, with its own forward sweeps).
template <int T, class C>
LMPC_DEVINL int bq_rebuild(const C &c, BigQp<T> &s) {
  const int lane = c.lane, k = s.k;
  int row0[T];
  double lam0[T], dW0[T];
  unsigned hm[T], sm[T];
  TFOR(t) { row0[t] = s.row[t]; lam0[t] = s.lam[t]; dW0[t] = s.dW[t]; }
  const unsigned low0 = s.low;
  int nh = 0;
  TFOR(t) {
    const bool live = lane + 32 * t < k, sf = (s.soft >> t) & 1u;
    hm[t] = __ballot_sync(FULL, live && !sf);
    sm[t] = __ballot_sync(FULL, live && sf);
    nh += __popc(hm[t]);
  }
  s.k = 0; s.sing = false;
  int defer = -1;
  for (int p = 0; p <= k; ++p) {
    int src;
    if (p < k) {
      int rem = (p < nh) ? p : p - nh;
      src = -1;
      TFOR(t) {
        const unsigned mk = (p < nh) ? hm[t] : sm[t];
        const int cnt = __popc(mk);
        if (src < 0) { if (rem < cnt) src = 32 * t + (int)__fns(mk, 0, rem + 1); else rem -= cnt; }
      }
    } else {
      if (defer < 0) break;
      src = defer;
    }
    const int j = bcast_pos_i<T>(row0, src);
    int lowv[T];
    TFOR(t) lowv[t] = (low0 >> t) & 1u;
    const bool lo = bcast_pos_i<T>(lowv, src) != 0;
    const double dv = bcast_pos<T>(dW0, src), lm = bcast_pos<T>(lam0, src);
    bq_add<T>(c, s, j, lo, dv);
    TFOR(t) if (lane + 32 * t == s.k - 1) s.lam[t] = lm;
    if (p == k) {
      TFOR(t) if (lane + 32 * t == s.k - 1) { s.D[t] = 0.0; s.Dinv[t] = 0.0; }
      s.sing = true;
    } else if (s.sing) {
      if (p >= nh || defer >= 0) return EXIT_CYCLE;
      defer = src; s.k -= 1; s.sing = false;
    }
  }
  return 0;
}
template <int T, class C>
LMPC_DEVINL bool bq_has_soft(const C &c, const BigQp<T> &s) {
  if (c.lay->nsoft == 0) return false;
  bool any = false;
  TFOR(t) any |= (c.lane + 32 * t < s.k) && ((s.soft >> t) & 1u);
  return __any_sync(FULL, any);
}

// lowest position holding the minimum of a[] (NaN-free), or -1 when all +inf
template <int T, class C>
LMPC_DEVINL int argmin_first(const C &c, const double (&a)[T], double &amin) {
  double mine = CUDART_INF;
  TFOR(t) mine = (a[t] < mine) ? a[t] : mine;
  amin = bfly_min(mine);
  int pos = 1 << 30;
#pragma unroll
  for (int t = T - 1; t >= 0; --t)
    if (a[t] == amin && a[t] < CUDART_INF) pos = c.lane + 32 * t;
  pos = bfly_min_i(pos);
  return pos == (1 << 30) ? -1 : pos;
}

template <int T, class C>
LMPC_DEVINL int bq_loop(const C &c, BigQp<T> &s, const double (&sh)[2 * T], double (&u)[T], int &iters) {
  const LdpLayout &lay = *c.lay;
  const LdpSettings &st = *c.st;
  const int lane = c.lane, n = lay.n, m = lay.m, mp = lay.mp;
  int flag = EXIT_ITERLIMIT, iter, cycle = 0, rc;
  bool repaired = false;
  double best_fval = -1.0, fval = 0.0, soft_slack = 0.0;
  for (iter = 1; iter < st.iter_limit; ++iter) {
    const int k = s.k;
    double a[T];
    TFOR(t) a[t] = CUDART_INF;
    if (!s.sing) {
      // 1. CSP
      double z[T];
      TFOR(t) z[t] = (lane + 32 * t < k) ? -s.dW[t] : 0.0;
      sweep_fwd<T>(c, z, k);
      TFOR(t) s.lams[t] = (lane + 32 * t < k) ? z[t] * s.Dinv[t] : 0.0;
      sweep_bwd<T>(c, s.lams, k);
      // 2. blocking
      bool anyblk = false;
      bool blk[T];
      TFOR(t) {
        const bool low = (s.low >> t) & 1u, imm = (s.imm >> t) & 1u;
        blk[t] = (lane + 32 * t < k) && !imm && (low ? (s.lams[t] > st.dual_tol) : (s.lams[t] < -st.dual_tol));
        anyblk |= blk[t];
      }
      if (!__any_sync(FULL, anyblk)) {
        // 3. u = -M_W' lam*
        TFOR(t) { const int i = lane + 32 * t; if (i < k) c.lamw[i] = s.lams[t]; }
        __syncwarp();
        TFOR(t) u[t] = 0.0;
        for (int i = 0; i < k; ++i) {
          const double li = c.lamw[i];
          const double *mr = c.M + (size_t)c.roww[i] * n;
          TFOR(t) { const int cc = lane + 32 * t; if (cc < n) u[t] = fma(-li, mr[cc], u[t]); }
        }
        TFOR(t) { const int cc = lane + 32 * t; if (cc < n) c.uw[cc] = u[t]; }
        __syncwarp();
        fval = lanesum_pos<T>(c, u, u, n);
        soft_slack = 0.0;
        if (lay.nsoft > 0) {
          double sl[T];
          double smax = 0.0;
          TFOR(t) {
            const bool sf = (lane + 32 * t < k) && ((s.soft >> t) & 1u);
            sl[t] = sf ? s.lams[t] : 0.0;
            const double v = st.rho_soft * fabs(s.lams[t]);
            if (sf && v > smax) smax = v;
          }
          fval = fma(st.rho_soft, lanesum_pos<T>(c, sl, s.lams, k), fval);
          soft_slack = bfly_max(smax);
        }
        if (fval > st.fval_bound) { flag = EXIT_INFEASIBLE; break; }
        // slack scan: rows lane + 32 t2; the loop over columns is outermost for ILP across rows
        double mu[2 * T];
#pragma unroll
        for (int t2 = 0; t2 < 2 * T; ++t2) mu[t2] = 0.0;
        for (int cc = 0; cc < n; ++cc) {
          const double uc = c.uw[cc];
          const double *mt = c.MT + (size_t)cc * mp + lane;
#pragma unroll
          for (int t2 = 0; t2 < 2 * T; ++t2)
            if (lane + 32 * t2 < m) mu[t2] = fma(mt[32 * t2], uc, mu[t2]);
        }
        double best = CUDART_INF, bestd = 0.0;
        int bestj = 1 << 30;
        bool bestlow = false;
#pragma unroll
        for (int t2 = 0; t2 < 2 * T; ++t2) {
          const int j = lane + 32 * t2;
          if (j < m && !((s.inws >> t2) & 1u) && !(c.cv.sense[j] & S_IMMUTABLE)) {
            const double dU = c.cv.du[j] + sh[t2], dLo = c.cv.dl[j] + sh[t2];
            const double su = dU - mu[t2], sl = mu[t2] - dLo;
            if (su < -st.primal_tol && su < best) { best = su; bestd = dU; bestj = j; bestlow = false; }
            else if (!(su < -st.primal_tol) && sl < -st.primal_tol && sl < best) { best = sl; bestd = dLo; bestj = j; bestlow = true; }
          }
        }
        const double vmin = bfly_min(best);
        const int add = bfly_min_i((best == vmin && bestj < (1 << 30)) ? bestj : (1 << 30));
        if (add == (1 << 30)) { flag = soft_slack > st.primal_tol ? EXIT_SOFT_OPTIMAL : EXIT_OPTIMAL; break; }
        // cycle guard where the dual objective is fresh (see oracle/ldp_oracle.c)
        bool rebuilt = false;
        if (fval - st.progress_tol > best_fval) { best_fval = fval; cycle = 0; }
        else if (cycle++ > st.cycle_tol) {
          // second chance for problems with soft rows: refactorise once, restart the count
          if (repaired || !bq_has_soft<T>(c, s)) { flag = EXIT_CYCLE; break; }
          repaired = true; rebuilt = true; best_fval = -1.0; cycle = 0;
        }
        if (k >= lay.cap || k >= 32 * T) { flag = EXIT_CAPACITY; break; }
        const int src = add & 31;
        const bool addlow = __shfl_sync(FULL, (int)(bestj == add && bestlow), src);
        const double addd = shfl_d(bestd, src);
        TFOR(t) if (lane + 32 * t < k) s.lam[t] = s.lams[t];
        if (rebuilt && (rc = bq_rebuild<T>(c, s)) != 0) { flag = rc; break; }
        bq_add<T>(c, s, add, addlow, addd);
        if (s.sing && bq_has_soft<T>(c, s) && (rc = bq_rebuild<T>(c, s)) != 0) { flag = rc; break; }
        continue;
      }
      // 4. blocking
      TFOR(t) if (blk[t]) { a[t] = -s.lam[t] / (s.lams[t] - s.lam[t]); if (!(a[t] == a[t])) a[t] = CUDART_INF; }
      double alpha;
      const int rm = argmin_first<T>(c, a, alpha);
      if (rm < 0) { flag = EXIT_CYCLE; break; }
      TFOR(t) if (lane + 32 * t < k) s.lam[t] = fma(alpha, s.lams[t] - s.lam[t], s.lam[t]);
      bq_remove<T>(c, s, rm);
      if (s.sing && bq_has_soft<T>(c, s) && (rc = bq_rebuild<T>(c, s)) != 0) { flag = rc; break; }
    } else {
      // 5. singular direction
      const int kk = k - 1;
      bool lastlow = false;
      {
        int lowv[T];
        TFOR(t) lowv[t] = (s.low >> t) & 1u;
        lastlow = bcast_pos_i<T>(lowv, kk) != 0;
      }
      double p[T];
      TFOR(t) { const int i = lane + 32 * t; p[t] = (i < kk) ? -(C::SQ ? c.Lw[(size_t)kk * c.LS + i] : c.Lw[PIDX(kk, i)]) : 0.0; }
      sweep_bwd<T>(c, p, kk);
      TFOR(t) {
        const int i = lane + 32 * t;
        if (i == kk) p[t] = 1.0;
        if (lastlow && i <= kk) p[t] = -p[t];
        s.lams[t] = (i < k) ? p[t] : 0.0;
        const bool low = (s.low >> t) & 1u, imm = (s.imm >> t) & 1u;
        const bool b = (i < k) && !imm && (low ? (p[t] > st.dual_tol) : (p[t] < -st.dual_tol));
        if (b) { a[t] = -s.lam[t] / p[t]; if (!(a[t] == a[t])) a[t] = CUDART_INF; }
      }
      double alpha;
      const int rm = argmin_first<T>(c, a, alpha);
      if (rm < 0) { flag = EXIT_INFEASIBLE; break; }
      TFOR(t) if (lane + 32 * t < k) s.lam[t] = fma(alpha, p[t], s.lam[t]);
      bq_remove<T>(c, s, rm);
      if (s.sing && bq_has_soft<T>(c, s) && (rc = bq_rebuild<T>(c, s)) != 0) { flag = rc; break; }
    }
  }
  iters = iter;
  return flag;
}

template <int T, class C>
LMPC_DEVINL int bq_solve(const C &c, BigQp<T> &s, double (&u)[T], int &iters, bool cold = true,
                         const uint64_t *wsu = nullptr, const uint64_t *wsl = nullptr) {
  const LdpLayout &lay = *c.lay;
  const int lane = c.lane, m = lay.m, nth = lay.nth;
  double sh[2 * T];
#pragma unroll
  for (int t2 = 0; t2 < 2 * T; ++t2) {
    sh[t2] = 0.0;
    const int j = lane + 32 * t2;
    if (j < m) {
      const double *dr = c.cv.Dth + (size_t)j * nth;
      double acc = 0.0;
      for (int t = 0; t < nth; ++t) acc = fma(dr[t], c.thw[t], acc);
      sh[t2] = acc;
    }
  }
  int flag = 0;
  if (cold) {
    bq_reset<T>(s);
    TFOR(t) u[t] = 0.0;
    for (int j = 0; lay.neq > 0 && j < m; ++j) {
      const int sj = c.cv.sense[j];
      if (sj & S_ACTIVE) {
        if (s.k >= lay.cap || s.k >= 32 * T) { flag = EXIT_OVERDETERMINED; break; }
        const bool low = sj & S_LOWER;
        double shj = 0.0;
#pragma unroll
        for (int t2 = 0; t2 < 2 * T; ++t2) { const double x = shfl_d(sh[t2], j & 31); shj = ((j >> 5) == t2) ? x : shj; }
        bq_add<T>(c, s, j, low, (low ? c.cv.dl[j] : c.cv.du[j]) + shj);
        if (s.sing) { flag = EXIT_OVERDETERMINED; break; }
      }
    }
    // warm start from a caller-supplied working set (oracle: activate_rows)
    for (int j = 0; wsu && flag == 0 && j < m; ++j) {
      const int sj = c.cv.sense[j];
      const bool up = (wsu[j >> 6] >> (j & 63)) & 1ull, lo = (wsl[j >> 6] >> (j & 63)) & 1ull;
      if ((sj & (S_ACTIVE | S_IMMUTABLE)) || !(up || lo)) continue;
      if (s.k >= lay.cap || s.k >= 32 * T) continue;
      double shj = 0.0;
#pragma unroll
      for (int t2 = 0; t2 < 2 * T; ++t2) { const double x = shfl_d(sh[t2], j & 31); shj = ((j >> 5) == t2) ? x : shj; }
      bq_add<T>(c, s, j, !up, (up ? c.cv.du[j] : c.cv.dl[j]) + shj);
      if (s.sing) bq_undo_add<T>(c, s, j);
    }
  } else {
    // warm start: same working set, new right-hand sides (row j's shift lives on lane j & 31, tile j >> 5)
    TFOR(t) {
      const int j = s.row[t];
      double shj = 0.0;
#pragma unroll
      for (int t2 = 0; t2 < 2 * T; ++t2) { const double x = shfl_d(sh[t2], j & 31); shj = ((j >> 5) == t2) ? x : shj; }
      if (lane + 32 * t < s.k) s.dW[t] = (((s.low >> t) & 1u) ? c.cv.dl[j] : c.cv.du[j]) + shj;
    }
  }
  iters = 0;
  if (flag == 0) flag = bq_loop<T>(c, s, sh, u, iters);
  return flag;
}

struct K3Geom { int warps; size_t per_warp; size_t small_total; size_t smem; };

// shared memory holds only the small constant sections (Dth, du, dl, Rinv, Uth, uoff, sense):
// they are contiguous at the tail of the block starting at oDth
LMPC_DEVINL int small_off(const LdpLayout &lay) { return lay.oDth; }

template <int T, bool SQ>
__global__ void __launch_bounds__(256)
ldp_big_kernel(const __grid_constant__ LdpLayout lay, const __grid_constant__ LdpSettings st,
               const double *__restrict__ cblk, ThetaSegs th, int64_t Ninst, SolveOut out,
               int warps, int per_warp_doubles) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar;
  const int soff = small_off(lay);
  stage_constants_tma(smem, cblk + soff, lay.total - soff, &mbar);
  BigCtxT<SQ> c;
  c.M = cblk + lay.oM; c.MT = cblk + lay.oMT; c.G = cblk + lay.oG;
  const double *sb = smem - soff;    // so that section offsets stay valid
  c.cv.M = c.M; c.cv.MT = c.MT; c.cv.G = c.G;
  c.cv.Dth = sb + lay.oDth; c.cv.du = sb + lay.odu; c.cv.dl = sb + lay.odl; c.cv.Rinv = sb + lay.oRinv;
  c.cv.Uth = sb + lay.oUth; c.cv.uoff = sb + lay.ouoff;
  c.cv.sense = reinterpret_cast<const int32_t *>(sb + lay.oSense);
  c.lay = &lay; c.st = &st;
  const int warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  double *w = smem + (lay.total - soff) + (size_t)warp * per_warp_doubles;
  const int cap = lay.cap;
  c.LS = cap | 1;
  c.Lw = w; w += SQ ? (((size_t)cap * c.LS + 1) & ~(size_t)1) : (((size_t)cap * (cap - 1) / 2 + 1) & ~(size_t)1);
  c.uw = w; w += (lay.n + 1) & ~1;
  c.thw = w; w += (lay.nth + 1) & ~1;
  c.lamw = w; w += (cap + 1) & ~1;
  c.roww = reinterpret_cast<int *>(w);
  c.inwsb = nullptr;
  const int lane = c.lane, words = (lay.m + 63) >> 6;
  const int64_t wid = (int64_t)blockIdx.x * warps + warp;
  const int64_t wstride = (int64_t)gridDim.x * warps;
  BigQp<T> s;
  double u[T];
  for (int64_t q = wid; q < Ninst; q += wstride) {
    int o = 0;
#pragma unroll
    for (int sg = 0; sg < 4; ++sg) {
      if (sg < th.nseg) {
        const int len = th.len[sg];
        for (int t = lane; t < len; t += 32) c.thw[o + t] = ld_stream(th.ptr[sg] + q * th.stride[sg] + t);
        o += len;
      }
    }
    __syncwarp();
    int iters;
    const int flag = bq_solve<T>(c, s, u, iters, true, out.ws_u ? out.ws_u + q * words : nullptr,
                                 out.ws_u ? out.ws_l + q * words : nullptr);
    // output map: canonical lane sum over the n variables, one control at a time
    for (int i = 0; i < lay.nu; ++i) {
      double rr[T];
      TFOR(t) { const int cc = lane + 32 * t; rr[t] = (cc < lay.n) ? c.cv.Rinv[(size_t)i * lay.n + cc] : 0.0; }
      const double x = lanesum_pos<T>(c, rr, u, lay.n);
      double shv = c.cv.uoff[i];
      const double *ur = c.cv.Uth + (size_t)i * lay.nth;
      for (int t = 0; t < lay.nth; ++t) shv = fma(ur[t], c.thw[t], shv);
      if (lane == 0) out.U[q * lay.nu + i] = __dadd_rn(x, shv);
    }
    if (out.act_u) {
      for (int wd = lane; wd < words; wd += 32) { out.act_u[q * words + wd] = 0ull; out.act_l[q * words + wd] = 0ull; }
      __syncwarp();
      TFOR(t) {
        const int i = lane + 32 * t;
        if (i < s.k) {
          const int j = s.row[t];
          unsigned long long *dst = reinterpret_cast<unsigned long long *>(((s.low >> t) & 1u) ? out.act_l : out.act_u) + q * words + (j >> 6);
          atomicOr(dst, 1ull << (j & 63));
        }
      }
    }
    if (lane == 0) {
      out.flag[q] = flag;
      if (out.iters) out.iters[q] = iters;
    }
    __syncwarp();
  }
}

// closed loop in warp mode: T steps of (theta = [x; r; d; u_prev]) -> solve -> x+ = F x + G u + Gd d
template <int T, bool SQ>
__global__ void __launch_bounds__(256)
ldp_big_sim_kernel(const __grid_constant__ LdpLayout lay, const __grid_constant__ LdpSettings st,
                   const double *__restrict__ cblk, int64_t Ninst, SimDesc sd, int32_t *flagout,
                   int32_t *itersout, int warps, int per_warp_doubles) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar;
  const int soff = small_off(lay);
  stage_constants_tma(smem, cblk + soff, lay.total - soff, &mbar);
  BigCtxT<SQ> c;
  c.M = cblk + lay.oM; c.MT = cblk + lay.oMT; c.G = cblk + lay.oG;
  const double *sb = smem - soff;
  c.cv.M = c.M; c.cv.MT = c.MT; c.cv.G = c.G;
  c.cv.Dth = sb + lay.oDth; c.cv.du = sb + lay.odu; c.cv.dl = sb + lay.odl; c.cv.Rinv = sb + lay.oRinv;
  c.cv.Uth = sb + lay.oUth; c.cv.uoff = sb + lay.ouoff;
  c.cv.sense = reinterpret_cast<const int32_t *>(sb + lay.oSense);
  c.lay = &lay; c.st = &st;
  const int warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  double *w = smem + (lay.total - soff) + (size_t)warp * per_warp_doubles;
  const int cap = lay.cap;
  c.LS = cap | 1;
  c.Lw = w; w += SQ ? (((size_t)cap * c.LS + 1) & ~(size_t)1) : (((size_t)cap * (cap - 1) / 2 + 1) & ~(size_t)1);
  c.uw = w; w += (lay.n + 1) & ~1;
  c.thw = w; w += (lay.nth + 1) & ~1;
  c.lamw = w; w += (cap + 1) & ~1;
  c.roww = reinterpret_cast<int *>(w); w += (cap + 1) / 2 + 1;
  c.inwsb = nullptr;
  const int nx = sd.nx, nr = sd.nr, nd = sd.nd, nup = sd.nuprev, nu = lay.nu;
  double *xnw = w; w += (nx + 1) & ~1;
  double *ucw = w; w += (nu + 1) & ~1;
  const bool ext = sd.ext != 0;
  const int ny = ext ? sd.ny : 0;
  double *xsw = w; w += (nx + 1) & ~1;     // ext: true plant state
  double *ew = w;                          // ext: innovation [ny]
  const int lane = c.lane;
  const int64_t wid = (int64_t)blockIdx.x * warps + warp;
  const int64_t wstride = (int64_t)gridDim.x * warps;
  BigQp<T> s;
  double u[T];
  for (int64_t q = wid; q < Ninst; q += wstride) {
    for (int i = lane; i < nx; i += 32) { c.thw[i] = ld_stream(sd.X0 + q * nx + i); if (ext) xsw[i] = c.thw[i]; }
    for (int i = lane; i < nr; i += 32) c.thw[nx + i] = sd.R ? ld_stream(sd.R + q * nr + i) : 0.0;
    for (int i = lane; i < nd; i += 32) c.thw[nx + nr + i] = sd.D ? ld_stream(sd.D + q * nd + i) : 0.0;
    for (int i = lane; i < nu; i += 32) ucw[i] = (sd.U0 && i < nup) ? ld_stream(sd.U0 + q * nup + i) : 0.0;
    __syncwarp();
    int worst = EXIT_OPTIMAL, tot = 0;
    for (int t = 0; t < sd.T; ++t) {
      for (int i = lane; i < nup; i += 32) c.thw[nx + nr + nd + i] = ucw[i];
      // reference / disturbance of step t (rs[:, k], ds[:, k] of src/simulation.jl:93-104)
      if (sd.Rtraj) for (int i = lane; i < nr; i += 32) c.thw[nx + i] = sd.Rtraj[(size_t)t * sd.rstride + i];
      if (sd.Dtraj) for (int i = lane; i < nd; i += 32) c.thw[nx + nr + i] = sd.Dtraj[(size_t)t * sd.dstride + i];
      __syncwarp();
      if (ext) {
        const double *dk = c.thw + nx + nr;
        if (sd.Cm) {
          // y = C x + Dd d + v, innovation against the estimate, correction (observer.jl:114-118)
          for (int i = lane; i < ny; i += 32) {
            double y = sd.Vtraj ? sd.Vtraj[(size_t)t * ny + i] : 0.0, yh = 0.0;
            for (int cc = 0; cc < nx; ++cc) y = fma(sd.Cm[i * nx + cc], xsw[cc], y);
            if (sd.Ddm) for (int cc = 0; cc < nd; ++cc) y = fma(sd.Ddm[i * nd + cc], dk[cc], y);
            for (int cc = 0; cc < nx; ++cc) yh = fma(sd.Cm[i * nx + cc], c.thw[cc], yh);
            if (sd.Ddm) for (int cc = 0; cc < nd; ++cc) yh = fma(sd.Ddm[i * nd + cc], dk[cc], yh);
            ew[i] = __dsub_rn(y, yh);
          }
          __syncwarp();
          for (int j = lane; j < nx; j += 32) {
            double acc = c.thw[j];
            for (int i = 0; i < ny; ++i) acc = fma(sd.Kobs[j * ny + i], ew[i], acc);
            xnw[j] = acc;
          }
          __syncwarp();
          for (int j = lane; j < nx; j += 32) c.thw[j] = xnw[j];
        } else {
          for (int j = lane; j < nx; j += 32) c.thw[j] = xsw[j];      // the controller measures the true state
        }
        __syncwarp();
      }
      int iters;
      const int f = bq_solve<T>(c, s, u, iters, !(sd.warm && t > 0));
      tot += iters;
      if (f < 0) { if (worst > 0 || f < worst) worst = f; }
      else if (worst > 0 && f > worst) worst = f;
      for (int i = 0; i < nu; ++i) {
        double rr[T];
        TFOR(tt) { const int cc = lane + 32 * tt; rr[tt] = (cc < lay.n) ? c.cv.Rinv[(size_t)i * lay.n + cc] : 0.0; }
        const double x = lanesum_pos<T>(c, rr, u, lay.n);
        double shv = c.cv.uoff[i];
        const double *ur = c.cv.Uth + (size_t)i * lay.nth;
        for (int tt = 0; tt < lay.nth; ++tt) shv = fma(ur[tt], c.thw[tt], shv);
        const double ctl = __dadd_rn(x, shv);
        if (lane == 0) {
          ucw[i] = ctl;
          if (sd.Utraj) sd.Utraj[(q * sd.T + t) * nu + i] = ctl;
        }
      }
      __syncwarp();
      for (int i = lane; i < nx; i += 32) {
        double acc = 0.0;
        for (int cc = 0; cc < nx; ++cc) acc = fma(sd.F[i * nx + cc], c.thw[cc], acc);
        for (int cc = 0; cc < nu; ++cc) acc = fma(sd.G[i * nu + cc], ucw[cc], acc);
        for (int cc = 0; cc < nd; ++cc) acc = fma(sd.Gd[i * nd + cc], c.thw[nx + nr + cc], acc);
        xnw[i] = acc;
      }
      __syncwarp();
      for (int i = lane; i < nx; i += 32) c.thw[i] = xnw[i];
      __syncwarp();
      if (ext) {   // true plant step
        for (int i = lane; i < nx; i += 32) {
          double acc = 0.0;
          for (int cc = 0; cc < nx; ++cc) acc = fma(sd.Ft[i * nx + cc], xsw[cc], acc);
          for (int cc = 0; cc < nu; ++cc) acc = fma(sd.Gt[i * nu + cc], ucw[cc], acc);
          for (int cc = 0; cc < nd; ++cc) acc = fma(sd.Gdt[i * nd + cc], c.thw[nx + nr + cc], acc);
          xnw[i] = acc;
        }
        __syncwarp();
        for (int i = lane; i < nx; i += 32) xsw[i] = xnw[i];
        __syncwarp();
      }
    }
    if (ext && sd.Xhat) for (int i = lane; i < nx; i += 32) sd.Xhat[q * nx + i] = c.thw[i];
    for (int i = lane; i < nx; i += 32) sd.Xend[q * nx + i] = ext ? xsw[i] : c.thw[i];
    for (int i = lane; i < nu; i += 32) sd.Ulast[q * nu + i] = ucw[i];
    if (lane == 0) {
      flagout[q] = worst;
      if (itersout) itersout[q] = tot;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-per-QP mode (problems whose packed L fills most of an SM's shared memory, e.g. n = 180).
// Warp 0 is the factorisation warp: it owns the working-set state in its register tiles and runs
// the triangular sweeps, LDL' add/remove and the ratio test exactly as in warp mode.  The two
// O(k n) / O(n m) phases of an iteration -- the primal iterate u = -M_W' lam* and the slack scan --
// are spread over all 256 threads (thread c owns u_c, thread j owns row j and its shift sh_j),
// each thread keeping the oracle's sequential accumulation order, with the Mbar / Mbar' reads
// coalesced and unrolled so eight loads are in flight per thread.
// ---------------------------------------------------------------------------------------------
struct CoopShared {
  double *shw;        // m: shifts sh_j (for equality activation / warm rows)
  double *zs;         // cap: z = L^-1 (-d_W), kept between iterations (see "z cache" below)
  double *red_v;      // 8: per-warp scan minima
  double *red_d;      // 8: bound value of the candidate
  int *red_j;         // 8: row of the candidate (1<<30: none)
  int *red_low;       // 8
  int *ctl;           // [0] branch (0 add / 1 remove / 2 singular), [1] exit flag or 0, [2] k, [3] pending add, [4] z mode
};

// Forward substitution z <- L^-1 z with the whole CTA (CTA mode, packed L): columns are taken in
// blocks of 32; warp 0 finishes the 32 x 32 diagonal block with shuffle broadcasts (the serial
// part), then every thread owns one trailing row and runs its 32 fmas of that block on its own.
// Every z_i still accumulates its terms in ascending column order, so the result is the same bits
// as the single-warp sweep (and the oracle); what changes is that the O(k^2) trailing work no
// longer sits on one warp (r01: the single-warp sweeps were ~85 % of the critical path at |W| ~ 130).
// zs: kk values in shared memory, in/out.  Must be called by all NT threads.
template <int NT>
LMPC_DEVINL void coop_sweep_fwd(const BigCtx &c, double *zs, int kk) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (kk <= 1) { __syncthreads(); return; }          // always ends with a barrier (callers rely on it)
  for (int j0 = 0; j0 + 1 < kk; j0 += 32) {
    const int nb = min(32, kk - 1 - j0);            // columns j0 .. j0 + nb - 1 have rows below them
    if (warp == 0) {
      const int i = j0 + lane;
      const bool live = i < kk;
      double zi = live ? zs[i] : 0.0;
      const double *lrow = c.Lw + PIDX(i, j0);        // only dereferenced for i > j0
      for (int jj = 0; jj < nb; ++jj) {
        const double zj = shfl_d(zi, jj);
        if (lane > jj && live) zi = fma(-lrow[jj], zj, zi);
      }
      if (live) zs[i] = zi;
    }
    __syncthreads();
    for (int i = j0 + 32 + tid; i < kk; i += NT) {
      double zi = zs[i];
      const double *lrow = c.Lw + PIDX(i, j0);
#pragma unroll 8
      for (int jj = 0; jj < nb; ++jj) zi = fma(-lrow[jj], zs[j0 + jj], zi);
      zs[i] = zi;
    }
    __syncthreads();
  }
}

// Backward substitution x <- L^-T x with the whole CTA, the mirror image of coop_sweep_fwd: rows are
// taken in blocks of 32 from the bottom; warp 0 finishes the diagonal block (x_i -= L[j][i] x_j for
// i, j inside the block, j descending), then every thread owns trailing positions i < j0 and applies
// the block's 32 rows to them, j descending.  Each x_i accumulates j = kk-1 .. i+1 in descending
// order exactly as the oracle's axpy loop does, so the bits are the single-warp sweep's.
// Row j of the packed L is contiguous in i: both parts read shared memory conflict free.
template <int NT>
LMPC_DEVINL void coop_sweep_bwd(const BigCtx &c, double *xs, int kk) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (kk <= 1) { __syncthreads(); return; }
  for (int j0 = ((kk - 1) >> 5) << 5; j0 >= 0; j0 -= 32) {
    const int j1 = min(kk, j0 + 32);
    if (warp == 0) {
      const int i = j0 + lane;
      const bool live = i < j1;
      double xi = live ? xs[i] : 0.0;
      for (int j = j1 - 1; j > j0; --j) {
        const double xj = shfl_d(xi, j - j0);
        if (i < j && live) xi = fma(-c.Lw[PIDX(j, i)], xj, xi);
      }
      if (live) xs[i] = xi;
    }
    __syncthreads();
    for (int i = tid; i < j0; i += NT) {
      double xi = xs[i];
#pragma unroll 8
      for (int j = j1 - 1; j >= j0; --j) xi = fma(-c.Lw[PIDX(j, i)], xs[j], xi);
      xs[i] = xi;
    }
    __syncthreads();
  }
}

// The two L2-bound fma chains of the CTA kernel, blocked by B with the TAIL blocked as well.  `#pragma unroll B` alone
// leaves up to B - 1 remainder iterations that run one load at a time: each exposes a full L2 latency (~600 cycles under
// load), 3.5 of them on average per chain and iteration.  Here the remainder's loads are issued together (indices
// clamped into range, the surplus fmas predicated off): same operations on the same operands in the same order.
#ifndef K3_PREFETCH
#define K3_PREFETCH 0       // 1: L1 prefetch of the next block's operands behind the current block's loads (measured, r02:
                            // chain s = 1 8.5e4 -> 5.9e4 solves/s, s = 0.25 8.8e5 -> 4.1e5: the ~30 KB of L1 left next to 2 x 110 KB
                            // of shared memory thrash; kept off)
#endif
LMPC_DEVINL void pf_l1(const double *p) {
#if K3_PREFETCH
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
template <int B>
LMPC_DEVINL double chain_scan(const double *__restrict__ mt, const double *uw, int cc, int n, int mp, double mu) {
  for (; cc + B <= n; cc += B) {
    double v[B];
#pragma unroll
    for (int t = 0; t < B; ++t) v[t] = mt[(size_t)t * mp];
    if (K3_PREFETCH && cc + 2 * B <= n) {
#pragma unroll
      for (int t = 0; t < B; ++t) pf_l1(mt + (size_t)(B + t) * mp);
    }
#pragma unroll
    for (int t = 0; t < B; ++t) mu = fma(v[t], uw[cc + t], mu);
    mt += (size_t)B * mp;
  }
  if (cc < n) {
    const int r = n - cc;
    double v[B];
#pragma unroll
    for (int t = 0; t < B; ++t) v[t] = mt[(size_t)(t < r ? t : r - 1) * mp];
#pragma unroll
    for (int t = 0; t < B; ++t) if (t < r) mu = fma(v[t], uw[cc + t], mu);
  }
  return mu;
}
template <int B>
LMPC_DEVINL double chain_u(const double *__restrict__ mcol, const int *roww, const double *lamw, int k, int n, double uc) {
  int i = 0;
  for (; i + B <= k; i += B) {
    double v[B];
#pragma unroll
    for (int t = 0; t < B; ++t) v[t] = mcol[(size_t)roww[i + t] * n];
    if (K3_PREFETCH && i + 2 * B <= k) {
#pragma unroll
      for (int t = 0; t < B; ++t) pf_l1(mcol + (size_t)roww[i + B + t] * n);
    }
#pragma unroll
    for (int t = 0; t < B; ++t) uc = fma(-lamw[i + t], v[t], uc);
  }
  if (i < k) {
    const int r = k - i;
    double v[B];
#pragma unroll
    for (int t = 0; t < B; ++t) v[t] = mcol[(size_t)roww[i + (t < r ? t : r - 1)] * n];
#pragma unroll
    for (int t = 0; t < B; ++t) if (t < r) uc = fma(-lamw[i + t], v[t], uc);
  }
  return uc;
}

// Two chains of one thread run together (a thread of a 128-thread CTA owns columns / rows tid and tid + 128 when n, m > 128):
// the same two sequences of fmas, interleaved, so the thread's critical path is one chain instead of two.
template <int B>
LMPC_DEVINL void chain_scan2(const double *__restrict__ ma, const double *__restrict__ mb, const double *uw, int cc, int n,
                             int mp, double &mua, double &mub) {
  for (; cc + B <= n; cc += B) {
    double va[B], vb[B];
#pragma unroll
    for (int t = 0; t < B; ++t) { va[t] = ma[(size_t)t * mp]; vb[t] = mb[(size_t)t * mp]; }
#pragma unroll
    for (int t = 0; t < B; ++t) { const double ucc = uw[cc + t]; mua = fma(va[t], ucc, mua); mub = fma(vb[t], ucc, mub); }
    ma += (size_t)B * mp; mb += (size_t)B * mp;
  }
  if (cc < n) {
    const int r = n - cc;
    double va[B], vb[B];
#pragma unroll
    for (int t = 0; t < B; ++t) { const size_t o = (size_t)(t < r ? t : r - 1) * mp; va[t] = ma[o]; vb[t] = mb[o]; }
#pragma unroll
    for (int t = 0; t < B; ++t) if (t < r) { const double ucc = uw[cc + t]; mua = fma(va[t], ucc, mua); mub = fma(vb[t], ucc, mub); }
  }
}
template <int B>
LMPC_DEVINL void chain_u2(const double *__restrict__ ca, const double *__restrict__ cb, const int *roww, const double *lamw,
                          int k, int n, double &ua, double &ub) {
  int i = 0;
  for (; i + B <= k; i += B) {
    double va[B], vb[B];
#pragma unroll
    for (int t = 0; t < B; ++t) { const size_t o = (size_t)roww[i + t] * n; va[t] = ca[o]; vb[t] = cb[o]; }
#pragma unroll
    for (int t = 0; t < B; ++t) { const double li = -lamw[i + t]; ua = fma(li, va[t], ua); ub = fma(li, vb[t], ub); }
  }
  if (i < k) {
    const int r = k - i;
    double va[B], vb[B];
#pragma unroll
    for (int t = 0; t < B; ++t) { const size_t o = (size_t)roww[i + (t < r ? t : r - 1)] * n; va[t] = ca[o]; vb[t] = cb[o]; }
#pragma unroll
    for (int t = 0; t < B; ++t) if (t < r) { const double li = -lamw[i + t]; ua = fma(li, va[t], ua); ub = fma(li, vb[t], ub); }
  }
}

// CTA-per-QP kernel.  NT threads per CTA (256: one QP per SM at full capacity; 128: two CTAs per SM
// with a reduced capacity `cap`, overflowing instances are redone by a full-capacity pass, `redo`).
//
// z cache.  The CSP's forward substitution z = L^-1 (-d_W) is kept in shared memory between
// iterations (DAQP's reuse_ind): appending a row leaves the first k rows of
// L and d_W untouched, so after an add only the new last entry is missing -- one sequential fma chain
// of length k on one thread, the same operations in the same order as the sweep would do for that row
// (bit-identical), instead of an O(k^2) sweep with k serial broadcast steps.  A removal at position r
// invalidates the entries from r on; the chain of masses removes in < 1 % of its iterations.
template <int T, int NT>
__global__ void __launch_bounds__(NT, (NT <= 128) ? 2 : 1)
ldp_cta_kernel(const __grid_constant__ LdpLayout lay, const __grid_constant__ LdpSettings st,
               const double *__restrict__ cblk, ThetaSegs th, int64_t Ninst, SolveOut out, int cap, int redo,
               unsigned long long *redo_count) {
  constexpr int RH = (512 + NT - 1) / NT;      // rows per thread (m <= 512)
  constexpr int CH = (256 + NT - 1) / NT;      // columns per thread (n <= 256)
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ __align__(8) uint64_t rbar[2];     // "batch landed" barriers of the two halves of the row ring
  // shared-memory constants: du, dl (contiguous from odu) and Uth, uoff, sense (contiguous from oUth);
  // Dth and R^-1 are read once per QP and stay in global memory / L2
  const int nA = lay.oRinv - lay.odu, nB = lay.total - lay.oUth;
  stage_constants_tma(smem, cblk + lay.odu, nA, &mbar[0]);
  stage_constants_tma(smem + nA, cblk + lay.oUth, nB, &mbar[1]);
  BigCtx c;
  c.LS = 0;
  c.M = cblk + lay.oM; c.MT = cblk + lay.oMT; c.G = cblk + lay.oG;
  c.cv.M = c.M; c.cv.MT = c.MT; c.cv.G = c.G;
  c.cv.Dth = cblk + lay.oDth; c.cv.Rinv = cblk + lay.oRinv;
  c.cv.du = smem; c.cv.dl = smem + (lay.odl - lay.odu);
  c.cv.Uth = smem + nA; c.cv.uoff = smem + nA + (lay.ouoff - lay.oUth);
  c.cv.sense = reinterpret_cast<const int32_t *>(smem + nA + (lay.oSense - lay.oUth));
  c.lay = &lay; c.st = &st;
  const int tid = threadIdx.x, warp = tid >> 5;
  c.lane = tid & 31;
  const int n = lay.n, m = lay.m, mp = lay.mp, nth = lay.nth;
  double *w = smem + nA + nB;
  c.Lw = w; w += ((size_t)cap * (cap - 1) / 2 + 1) & ~(size_t)1;
  c.uw = w; w += (n + 1) & ~1;
  c.thw = w; w += (nth + 1) & ~1;
  c.lamw = w; w += (cap + 1) & ~1;
  CoopShared cs;
  cs.shw = w; w += (m + 1) & ~1;
  cs.zs = w; w += (cap + 1) & ~1;
  cs.red_v = w; w += 8;
  cs.red_d = w; w += 8;
  cs.red_j = reinterpret_cast<int *>(w); w += 4;
  cs.red_low = reinterpret_cast<int *>(w); w += 4;
  cs.ctl = reinterpret_cast<int *>(w); w += 4;
  // Row ring: the two O(k n) / O(n m) phases stream rows of Mbar / Mbar' from L2 through 2 * K3_RB row slots
  // with TMA bulk copies, one batch ahead of the fma chains that consume them (r02 ncu: those chains waited
  // on L2 for ~55 cycles per step with 8 loads in flight; from shared memory a step costs the DFMA latency, 8).
  // Needs 16-byte aligned rows: n even (Mbar rows) -- mp is a multiple of 4 by construction.
  double *ring = w; w += K3_USE_RING ? 2 * K3_RB * mp : 0;
  const bool use_ring = K3_USE_RING && (n & 1) == 0 && (mp & 1) == 0 && n <= mp;
  c.roww = reinterpret_cast<int *>(w); w += (cap + 1) / 2 + 1;
  c.inwsb = reinterpret_cast<unsigned char *>(w);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&rbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&rbar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t rph0 = 0, rph1 = 0;                  // phase parity of rbar[0] / rbar[1] (uniform over the CTA)
  const bool boxonly = (m == lay.ms);           // every row is a (triangular) row of R^-1: copies can be trimmed
  const int lane = c.lane, words = (m + 63) >> 6;
  BigQp<T> s;          // meaningful in warp 0 only
  double u[T];
  for (int64_t q = blockIdx.x; q < Ninst; q += gridDim.x) {
    if (redo) {         // second pass: only the instances the reduced-capacity pass could not hold
      if (out.flag[q] != EXIT_CAPACITY) continue;
      if (tid == 0 && redo_count) atomicAdd(redo_count, 1ull);
    }
    // theta, shifts (thread j <-> rows j, j + NT, ...), reset
    int o = 0;
#pragma unroll
    for (int sg = 0; sg < 4; ++sg) {
      if (sg < th.nseg) {
        const int len = th.len[sg];
        for (int t = tid; t < len; t += NT) c.thw[o + t] = ld_stream(th.ptr[sg] + q * th.stride[sg] + t);
        o += len;
      }
    }
    for (int j = tid; j < m; j += NT) c.inwsb[j] = 0;
    __syncthreads();
    double shr[RH];
#pragma unroll
    for (int h = 0; h < RH; ++h) {
      shr[h] = 0.0;
      const int j = tid + NT * h;
      if (j < m) {
        const double *dr = c.cv.Dth + (size_t)j * nth;
        double acc = 0.0;
        for (int t = 0; t < nth; ++t) acc = fma(dr[t], c.thw[t], acc);
        shr[h] = acc;
        cs.shw[j] = acc;
      }
    }
    if (tid == 0) { cs.ctl[1] = 0; cs.ctl[3] = 0; }
    __syncthreads();
    int flag = EXIT_ITERLIMIT, iter = 0, cycle = 0, zvalid = 0;
    bool repaired = false;
    double best_fval = -1.0, fval = 0.0, soft_slack = 0.0;
    if (warp == 0) {
      bq_reset<T>(s);
      TFOR(t) u[t] = 0.0;
      for (int j = 0; lay.neq > 0 && j < m; ++j) {
        const int sj = c.cv.sense[j];
        if (sj & S_ACTIVE) {
          if (s.k >= cap || s.k >= 32 * T) { flag = EXIT_OVERDETERMINED; break; }
          const bool low = sj & S_LOWER;
          bq_add<T>(c, s, j, low, (low ? c.cv.dl[j] : c.cv.du[j]) + cs.shw[j]);
          if (s.sing) { flag = EXIT_OVERDETERMINED; break; }
        }
      }
      // warm start from a caller-supplied working set (oracle: activate_rows)
      for (int j = 0; out.ws_u && flag != EXIT_OVERDETERMINED && j < m; ++j) {
        const int sj = c.cv.sense[j];
        const uint64_t *wsu = out.ws_u + q * words, *wsl = out.ws_l + q * words;
        const bool up = (wsu[j >> 6] >> (j & 63)) & 1ull, lo = (wsl[j >> 6] >> (j & 63)) & 1ull;
        if ((sj & (S_ACTIVE | S_IMMUTABLE)) || !(up || lo)) continue;
        if (s.k >= cap || s.k >= 32 * T) continue;
        bq_add<T>(c, s, j, !up, (up ? c.cv.du[j] : c.cv.dl[j]) + cs.shw[j]);
        if (s.sing) bq_undo_add<T>(c, s, j);
      }
      if (flag == EXIT_OVERDETERMINED && lane == 0) cs.ctl[1] = flag;
    }
    __syncthreads();
    const bool dead = cs.ctl[1] != 0;
    int pend_add = 0;
    bool pend_low = false;
    double pend_d = 0.0;
    for (iter = 1; !dead && iter < st.iter_limit; ++iter) {
      // ---- phase A: CSP.  z from the cache (chain for one new row, cooperative sweep otherwise), then
      //      lam* = L^-T (z / D) by the whole CTA, blocking test on warp 0 ------------------------------
      double a[T];
      bool blk[T];
      if (warp == 0) {
        TFOR(t) { a[t] = CUDART_INF; blk[t] = false; }
        int zmode = 0;                               // 0: cache complete, 1: last entry missing, 2: recompute
        if (!s.sing && zvalid < s.k) {
          zmode = (zvalid == s.k - 1) ? 1 : 2;
          TFOR(t) { const int i = lane + 32 * t; if (i < s.k && (zmode == 2 || i == s.k - 1)) cs.zs[i] = -s.dW[t]; }
        }
        if (lane == 0) { cs.ctl[0] = s.sing ? 2 : 0; cs.ctl[2] = s.k; cs.ctl[4] = zmode; }
      }
      __syncthreads();
      {
        const int not_sing = cs.ctl[0] != 2, kk = cs.ctl[2], zmode = cs.ctl[4];   // read before warp 0 rewrites them
        if (not_sing && zmode == 2) coop_sweep_fwd<NT>(c, cs.zs, kk);             // ends with a __syncthreads
        else {
          if (not_sing && zmode == 1 && tid == 0) {
            const int i = kk - 1;
            double acc = cs.zs[i];
            const double *lrow = c.Lw + PIDX(i, 0);
#pragma unroll 8
            for (int cc = 0; cc < i; ++cc) acc = fma(-lrow[cc], cs.zs[cc], acc);
            cs.zs[i] = acc;
          }
          __syncthreads();
        }
        if (warp == 0 && not_sing) {
          TFOR(t) { const int i = lane + 32 * t; if (i < kk) c.lamw[i] = cs.zs[i] * s.Dinv[t]; }
          zvalid = kk;
        }
        __syncthreads();
        if (not_sing) coop_sweep_bwd<NT>(c, c.lamw, kk);                          // ends with a __syncthreads
      }
      if (warp == 0) {
        const int k = s.k;
        int branch;
        if (!s.sing) {
          TFOR(t) { const int i = lane + 32 * t; s.lams[t] = (i < k) ? c.lamw[i] : 0.0; }
          bool anyblk = false;
          TFOR(t) {
            const bool low = (s.low >> t) & 1u, imm = (s.imm >> t) & 1u;
            blk[t] = (lane + 32 * t < k) && !imm && (low ? (s.lams[t] > st.dual_tol) : (s.lams[t] < -st.dual_tol));
            anyblk |= blk[t];
          }
          branch = __any_sync(FULL, anyblk) ? 1 : 0;
        } else {
          branch = 2;
        }
        if (lane == 0) { cs.ctl[0] = branch; cs.ctl[2] = k; }
      }
      __syncthreads();
      const int branch = cs.ctl[0], k = cs.ctl[2];
      if (branch == 0) {
        // ---- phase B (all threads): u_c = -sum_i lam*_i Mbar[row_i][c], sequential in i ----------
        // (a thread's CH columns run as independent chains inside one loop over i: their loads overlap)
        {
          double uc[CH];
          const double *mcol[CH];
#pragma unroll
          for (int h = 0; h < CH; ++h) { uc[h] = 0.0; mcol[h] = c.M + min(tid + NT * h, n - 1); }
          const int nmine = (n - tid + NT - 1) / NT;          // columns this thread owns (<= CH)
          if (use_ring) {
            // batch b = working-set positions [b RB, (b+1) RB): row w_i of Mbar, trimmed to its non-zero tail
            // [w_i & ~1, n) for a simple row (zeros before the diagonal); column c of the row sits at slot[c]
            const int nb = (k + K3_RB - 1) / K3_RB;
            auto issue = [&](int b) {
              const int hf = b & 1, i0 = b * K3_RB, i1 = min(k, i0 + K3_RB);
              uint32_t tot = 0;
              for (int i = i0; i < i1; ++i) { const int wr = c.roww[i]; tot += (uint32_t)(n - (wr < lay.ms ? (wr & ~1) : 0)) * 8u; }
              bulk_expect(&rbar[hf], tot);
              for (int i = i0; i < i1; ++i) {
                const int wr = c.roww[i], cs0 = wr < lay.ms ? (wr & ~1) : 0;
                bulk_copy(&rbar[hf], ring + (size_t)(hf * K3_RB + (i - i0)) * mp + cs0, c.M + (size_t)wr * n + cs0,
                          (uint32_t)(n - cs0) * 8u);
              }
            };
            if (tid == 0 && nb > 0) issue(0);
            for (int b = 0; b < nb; ++b) {
              if (tid == 0 && b + 1 < nb) issue(b + 1);
              if (b & 1) { tile_wait(&rbar[1], rph1); rph1 ^= 1u; } else { tile_wait(&rbar[0], rph0); rph0 ^= 1u; }
              const double *slot = ring + (size_t)((b & 1) * K3_RB) * mp;
              const int i0 = b * K3_RB;
#pragma unroll
              for (int r = 0; r < K3_RB; ++r) {
                const int i = i0 + r;
                if (i < k) {
                  const double li = -c.lamw[i];
                  const int wr = c.roww[i], cs0 = wr < lay.ms ? (wr & ~1) : 0;
#pragma unroll
                  for (int h = 0; h < CH; ++h) {
                    const int cc = tid + NT * h;
                    if (cc < n && cc >= cs0) uc[h] = fma(li, slot[r * mp + cc], uc[h]);
                  }
                }
              }
              __syncthreads();               // the half is free again (it is refilled two batches later)
            }
          } else if (K3_INTERLEAVE && nmine >= 2) {
#pragma unroll K3_BCU
            for (int i = 0; i < k; ++i) {
              const double li = -c.lamw[i];
              const int ro = c.roww[i] * n;
#pragma unroll
              for (int h = 0; h < CH; ++h) uc[h] = fma(li, mcol[h][ro], uc[h]);
            }
          } else if (!K3_TRIM_U && K3_PAIR && CH == 2 && nmine == 2) {
            chain_u2<K3_PAIR_B>(mcol[0], mcol[CH - 1], c.roww, c.lamw, k, n, uc[0], uc[CH - 1]);
          } else {
#pragma unroll
            for (int h = 0; h < CH; ++h) {
              if (h < nmine) {
                const int cc = tid + NT * h;
                // a simple row w is zero before column w: fma(-lam, 0, uc) == uc, so those steps are skipped
                // (K3_TRIM_U; the loads they would have issued are half of this phase's L2 traffic)
#if K3_TRIM_U
#pragma unroll K3_BCU
                for (int i = 0; i < k; ++i) {
                  const int wr = c.roww[i];
                  if (wr <= cc || wr >= lay.ms) uc[h] = fma(-c.lamw[i], mcol[h][wr * n], uc[h]);
                }
#else
                (void)cc;
                uc[h] = chain_u<K3_BCU>(mcol[h], c.roww, c.lamw, k, n, uc[h]);
#endif
              }
            }
          }
#pragma unroll
          for (int h = 0; h < CH; ++h) { const int cc = tid + NT * h; if (cc < n) c.uw[cc] = uc[h]; }
        }
        __syncthreads();
        // ---- phase C (all threads): slack of rows tid, tid + NT, ..., sequential in c ------------
        double best = CUDART_INF, bestd = 0.0;
        int bestj = 1 << 30, bestlow = 0;
        // a simple row j is a normalised row of the upper-triangular R^-1: exact zeros before column j,
        // so its sum may start at any column <= j (fma(0, u, mu) == mu).  A thread's rows run as
        // independent chains inside one loop over the columns, from the first row's start.
        double mus[RH];
        {
          const int nmine = (m - tid + NT - 1) / NT;          // rows this thread owns (<= RH)
#pragma unroll
          for (int h = 0; h < RH; ++h) mus[h] = 0.0;
          if (use_ring) {
            // batch b = columns [b RB, (b+1) RB) of Mbar == rows of Mbar'; on a box-only problem row cc of Mbar'
            // has its non-zeros in [0, cc] (row j of Mbar starts at column j), so only that head is copied
            const int nb = (n + K3_RB - 1) / K3_RB;
            auto issue = [&](int b) {
              const int hf = b & 1, c0 = b * K3_RB, c1 = min(n, c0 + K3_RB);
              uint32_t tot = 0;
              for (int cc = c0; cc < c1; ++cc) tot += (uint32_t)(boxonly ? min(mp, (cc + 2) & ~1) : mp) * 8u;
              bulk_expect(&rbar[hf], tot);
              for (int cc = c0; cc < c1; ++cc)
                bulk_copy(&rbar[hf], ring + (size_t)(hf * K3_RB + (cc - c0)) * mp, c.MT + (size_t)cc * mp,
                          (uint32_t)(boxonly ? min(mp, (cc + 2) & ~1) : mp) * 8u);
            };
            int jc0[RH];                       // first column with a non-zero in row j (simple rows), else 0
#pragma unroll
            for (int h = 0; h < RH; ++h) { const int j = tid + NT * h; jc0[h] = (j < lay.ms) ? j : 0; }
            if (tid == 0) issue(0);
            for (int b = 0; b < nb; ++b) {
              if (tid == 0 && b + 1 < nb) issue(b + 1);
              if (b & 1) { tile_wait(&rbar[1], rph1); rph1 ^= 1u; } else { tile_wait(&rbar[0], rph0); rph0 ^= 1u; }
              const double *slot = ring + (size_t)((b & 1) * K3_RB) * mp;
              const int c0 = b * K3_RB;
#pragma unroll
              for (int r = 0; r < K3_RB; ++r) {
                const int cc = c0 + r;
                if (cc < n) {
                  const double ucc = c.uw[cc];
#pragma unroll
                  for (int h = 0; h < RH; ++h)
                    if (h < nmine && cc >= jc0[h]) mus[h] = fma(slot[r * mp + tid + NT * h], ucc, mus[h]);
                }
              }
              __syncthreads();
            }
          } else if (nmine >= 1) {
            // start at the first row's own column only when every row of this thread is a simple row
            const int c0 = (tid + NT * (nmine - 1) < lay.ms) ? tid : 0;
            const double *mt = c.MT + tid + c0 * mp;
            if (nmine == 1) {
              mus[0] = chain_scan<K3_BCU>(mt, c.uw, c0, n, mp, mus[0]);
            } else if (K3_PAIR && nmine == 2 && RH >= 2 && tid + NT < lay.ms && tid + NT <= n && NT % K3_BCU == 0 && NT % K3_PAIR_B == 0) {
              // rows j = tid and j + NT, both simple: row j alone on columns [j, j + NT), then both (row j + NT is zero before
              // its own column, so starting it at any column <= j + NT adds exact zeros)
              const double *ma = c.MT + tid + (size_t)tid * mp;
              mus[0] = chain_scan<K3_BCU>(ma, c.uw, tid, tid + NT, mp, mus[0]);
              chain_scan2<K3_PAIR_B>(ma + (size_t)NT * mp, c.MT + (tid + NT) + (size_t)(tid + NT) * mp, c.uw, tid + NT, n, mp,
                                     mus[0], mus[1]);
            } else if (!K3_INTERLEAVE) {
#pragma unroll
              for (int h = 0; h < RH; ++h) {
                if (h < nmine) {
                  const int j = tid + NT * h;
                  const int cj = (j < lay.ms) ? j : 0;
                  const double *mj = c.MT + j + cj * mp;
                  mus[h] = chain_scan<K3_BCU>(mj, c.uw, cj, n, mp, mus[h]);
                }
              }
            } else {
#pragma unroll K3_BCU
              for (int cc = c0; cc < n; ++cc) {
                const double ucc = c.uw[cc];
#pragma unroll
                for (int h = 0; h < RH; ++h) { if (h < nmine) mus[h] = fma(mt[min(NT * h, mp - 1 - tid)], ucc, mus[h]); }
                mt += mp;
              }
            }
          }
        }
#pragma unroll
        for (int h = 0; h < RH; ++h) {
          const int j = tid + NT * h;
          if (j < m) {
            const double mu = mus[h];
            if (!c.inwsb[j] && !(c.cv.sense[j] & S_IMMUTABLE)) {
              const double dU = c.cv.du[j] + shr[h], dLo = c.cv.dl[j] + shr[h];
              const double su = dU - mu, sl = mu - dLo;
              if (su < -st.primal_tol && su < best) { best = su; bestd = dU; bestj = j; bestlow = 0; }
              else if (!(su < -st.primal_tol) && sl < -st.primal_tol && sl < best) { best = sl; bestd = dLo; bestj = j; bestlow = 1; }
            }
          }
        }
        {
          const double vmin = bfly_min(best);
          const int jw = bfly_min_i((best == vmin && bestj < (1 << 30)) ? bestj : (1 << 30));
          // the owner of row jw (if any) publishes the warp's candidate
          if (jw < (1 << 30) && bestj == jw && best == vmin) {
            cs.red_v[warp] = vmin; cs.red_d[warp] = bestd; cs.red_j[warp] = jw; cs.red_low[warp] = bestlow;
          } else if (jw == (1 << 30) && lane == 0) {
            cs.red_v[warp] = CUDART_INF; cs.red_j[warp] = 1 << 30;
          }
        }
        __syncthreads();
        // ---- phase D (warp 0): objective, decision, add ----------------------------------------
        if (warp == 0) {
          TFOR(t) { const int cc = lane + 32 * t; u[t] = (cc < n) ? c.uw[cc] : 0.0; }
          fval = lanesum_pos<T>(c, u, u, n);
          soft_slack = 0.0;
          if (lay.nsoft > 0) {
            double sl[T];
            double smax = 0.0;
            TFOR(t) {
              const bool sf = (lane + 32 * t < k) && ((s.soft >> t) & 1u);
              sl[t] = sf ? s.lams[t] : 0.0;
              const double v = st.rho_soft * fabs(s.lams[t]);
              if (sf && v > smax) smax = v;
            }
            fval = fma(st.rho_soft, lanesum_pos<T>(c, sl, s.lams, k), fval);
            soft_slack = bfly_max(smax);
          }
          int done = 0;
          if (fval > st.fval_bound) done = EXIT_INFEASIBLE;
          else {
            const double v = (lane < NW) ? cs.red_v[lane] : CUDART_INF;
            const int jl = (lane < NW) ? cs.red_j[lane] : (1 << 30);
            const double vmin = bfly_min(v);
            const int add = bfly_min_i((v == vmin && jl < (1 << 30)) ? jl : (1 << 30));
            if (add == (1 << 30)) done = soft_slack > st.primal_tol ? EXIT_SOFT_OPTIMAL : EXIT_OPTIMAL;
            else {
              bool rebuilt = false;
              if (fval - st.progress_tol > best_fval) { best_fval = fval; cycle = 0; }
              else if (cycle++ > st.cycle_tol) {
                // second chance for problems with soft rows: refactorise once, restart the count
                if (repaired || !bq_has_soft<T>(c, s)) done = EXIT_CYCLE;
                else { repaired = true; rebuilt = true; best_fval = -1.0; cycle = 0; }
              }
              if (!done && (k >= cap || k >= 32 * T)) done = EXIT_CAPACITY;
              if (!done) {
                const unsigned own = __ballot_sync(FULL, lane < NW && jl == add && v == vmin);
                const int src = __ffs(own) - 1;
                const bool addlow = __shfl_sync(FULL, (lane < NW) ? cs.red_low[lane] : 0, src) != 0;
                const double addd = shfl_d((lane < NW) ? cs.red_d[lane] : 0.0, src);
                TFOR(t) if (lane + 32 * t < k) s.lam[t] = s.lams[t];
                if (rebuilt) { done = bq_rebuild<T>(c, s); zvalid = 0; }
                if (!done) {
                  // right-hand side of the add's forward sweep: the Gram column of the new row
                  const double *g = c.G + (size_t)add * m;
                  TFOR(t) { const int i = lane + 32 * t; if (i < k) c.lamw[i] = g[s.row[t]]; }
                  pend_add = add; pend_low = addlow; pend_d = addd;
                  if (lane == 0) cs.ctl[3] = 1;
                }
              }
            }
          }
          if (lane == 0) cs.ctl[1] = done;
        }
        __syncthreads();
        const int do_add = cs.ctl[3];
        if (do_add != 0) {             // the add: y = L^-1 g_W by the whole CTA, the rest on warp 0
          coop_sweep_fwd<NT>(c, c.lamw, k);
          if (warp == 0) {
            bq_add<T>(c, s, pend_add, pend_low, pend_d, c.lamw);
            if (s.sing && bq_has_soft<T>(c, s)) {
              const int rc = bq_rebuild<T>(c, s);
              zvalid = 0;
              if (rc != 0 && lane == 0) cs.ctl[1] = rc;
            }
            if (lane == 0) cs.ctl[3] = 0;
          }
        }
      } else if (warp == 0) {
        // ---- removal branches (warp 0 only) -----------------------------------------------------
        int done = 0;
        if (branch == 1) {
          TFOR(t) if (blk[t]) { a[t] = -s.lam[t] / (s.lams[t] - s.lam[t]); if (!(a[t] == a[t])) a[t] = CUDART_INF; }
          double alpha;
          const int rm = argmin_first<T>(c, a, alpha);
          if (rm < 0) done = EXIT_CYCLE;
          else {
            TFOR(t) if (lane + 32 * t < k) s.lam[t] = fma(alpha, s.lams[t] - s.lam[t], s.lam[t]);
            bq_remove<T>(c, s, rm);
            zvalid = min(zvalid, rm);
            if (s.sing && bq_has_soft<T>(c, s)) { done = bq_rebuild<T>(c, s); zvalid = 0; }
          }
        } else {
          const int kk = k - 1;
          bool lastlow = false;
          {
            int lowv[T];
            TFOR(t) lowv[t] = (s.low >> t) & 1u;
            lastlow = bcast_pos_i<T>(lowv, kk) != 0;
          }
          double p[T];
          TFOR(t) { const int i = lane + 32 * t; p[t] = (i < kk) ? -c.Lw[PIDX(kk, i)] : 0.0; }
          sweep_bwd<T>(c, p, kk);
          TFOR(t) {
            const int i = lane + 32 * t;
            if (i == kk) p[t] = 1.0;
            if (lastlow && i <= kk) p[t] = -p[t];
            s.lams[t] = (i < k) ? p[t] : 0.0;
            const bool low = (s.low >> t) & 1u, imm = (s.imm >> t) & 1u;
            const bool b = (i < k) && !imm && (low ? (p[t] > st.dual_tol) : (p[t] < -st.dual_tol));
            if (b) { a[t] = -s.lam[t] / p[t]; if (!(a[t] == a[t])) a[t] = CUDART_INF; }
          }
          double alpha;
          const int rm = argmin_first<T>(c, a, alpha);
          if (rm < 0) done = EXIT_INFEASIBLE;
          else {
            TFOR(t) if (lane + 32 * t < k) s.lam[t] = fma(alpha, p[t], s.lam[t]);
            bq_remove<T>(c, s, rm);
            zvalid = min(zvalid, rm);
            if (s.sing && bq_has_soft<T>(c, s)) { done = bq_rebuild<T>(c, s); zvalid = 0; }
          }
        }
        if (lane == 0) cs.ctl[1] = done;
      }
      __syncthreads();
      if (cs.ctl[1] != 0) break;
    }
    flag = cs.ctl[1] != 0 ? cs.ctl[1] : EXIT_ITERLIMIT;
    // ---- output (warp 0) ------------------------------------------------------------------------
    if (warp == 0) {
      for (int i = 0; i < lay.nu; ++i) {
        double rr[T];
        TFOR(t) { const int cc = lane + 32 * t; rr[t] = (cc < n) ? c.cv.Rinv[(size_t)i * n + cc] : 0.0; }
        const double x = lanesum_pos<T>(c, rr, u, n);
        double shv = c.cv.uoff[i];
        const double *ur = c.cv.Uth + (size_t)i * nth;
        for (int t = 0; t < nth; ++t) shv = fma(ur[t], c.thw[t], shv);
        if (lane == 0) out.U[q * lay.nu + i] = __dadd_rn(x, shv);
      }
      if (out.act_u) {
        for (int wd = lane; wd < words; wd += 32) { out.act_u[q * words + wd] = 0ull; out.act_l[q * words + wd] = 0ull; }
        __syncwarp();
        TFOR(t) {
          const int i = lane + 32 * t;
          if (i < s.k) {
            const int j = s.row[t];
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(((s.low >> t) & 1u) ? out.act_l : out.act_u) + q * words + (j >> 6);
            atomicOr(dst, 1ull << (j & 63));
          }
        }
      }
      if (lane == 0) {
        out.flag[q] = flag;
        if (out.iters) out.iters[q] = iter;
      }
    }
    __syncthreads();
  }
}

size_t cta_smem(const LdpLayout &lay, int cap) {
  size_t d = (size_t)(lay.oRinv - lay.odu) + (size_t)(lay.total - lay.oUth);
  d += ((size_t)cap * (cap - 1) / 2 + 1) & ~(size_t)1;
  d += (lay.n + 1) & ~1; d += (lay.nth + 1) & ~1; d += (cap + 1) & ~1; d += (lay.m + 1) & ~1;
  d += (cap + 1) & ~1;
  d += 8 + 8 + 4 + 4 + 4 + (cap + 1) / 2 + 1;
  d += K3_USE_RING ? (size_t)2 * K3_RB * lay.mp : 0;
  d += (lay.m + 7) / 8 + 1;
  return d * sizeof(double);
}
size_t cta_smem(const LdpLayout &lay) { return cta_smem(lay, lay.cap); }

// Two CTAs per SM: the largest working-set capacity whose CTA still fits twice into an SM's shared
// memory.  Used as a first pass when it is at least 80 % of the full capacity (instances that outgrow it
// end with EXIT_CAPACITY and are redone at full capacity by a second launch); 0 = not worth it.
int cta_fast_cap(const LdpLayout &lay) {
  if (lay.n > 256 || lay.m > 512) return 0;
  const size_t budget = (227u * 1024u - 2048u) / 2;
  int cap = lay.cap;
  while (cap > 0 && cta_smem(lay, cap) > budget) --cap;
  if (cap >= lay.cap) return lay.cap;            // the full problem fits twice: no second pass needed
  return (cap * 100 >= lay.cap * 80) ? cap : 0;
}

unsigned long long *g_redo_counter(int device) {
  static unsigned long long *ctr[64] = {nullptr};
  if (device < 0 || device >= 64) return nullptr;
  if (!ctr[device]) {
    if (cudaMalloc((void **)&ctr[device], 8) != cudaSuccess) { cudaGetLastError(); ctr[device] = nullptr; return nullptr; }
    cudaMemset(ctr[device], 0, 8);
  }
  return ctr[device];
}

template <int T>
int launch_cta_full(const LdpLayout &lay, const LdpSettings &st, const double *cblk, const ThetaSegs &th,
                    int64_t N, const SolveOut &out, int sm_count, cudaStream_t s, int redo) {
  auto kern = ldp_cta_kernel<T, 256>;
  const size_t smem = cta_smem(lay, lay.cap);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = N;
  if (blocks > (int64_t)sm_count * per_sm) blocks = (int64_t)sm_count * per_sm;
  kern<<<(unsigned)blocks, 256, smem, s>>>(lay, st, cblk, th, N, out, lay.cap, redo, redo ? g_redo_counter(dev) : nullptr);
  return (int)cudaGetLastError();
}

template <int T>
int launch_cta(const LdpLayout &lay, const LdpSettings &st, const double *cblk, const ThetaSegs &th,
               int64_t N, const SolveOut &out, int sm_count, cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  const int fast = cta_fast_cap(lay);
  if (fast > 0) {
    auto k2 = ldp_cta_kernel<T, 128>;
    const size_t smem2 = cta_smem(lay, fast);
    cudaError_t e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2, 128, smem2);
    if (per_sm >= 2) {
      int64_t blocks = N;
      if (blocks > (int64_t)sm_count * per_sm) blocks = (int64_t)sm_count * per_sm;
      k2<<<(unsigned)blocks, 128, smem2, s>>>(lay, st, cblk, th, N, out, fast, 0, nullptr);
      e = cudaGetLastError();
      if (e != cudaSuccess || fast >= lay.cap) return (int)e;
      // second pass at full capacity over the instances that ended with EXIT_CAPACITY
    } else {
      return launch_cta_full<T>(lay, st, cblk, th, N, out, sm_count, s, 0);
    }
    lmpc_count_kernel(1);
    return launch_cta_full<T>(lay, st, cblk, th, N, out, sm_count, s, 1);
  }
  return launch_cta_full<T>(lay, st, cblk, th, N, out, sm_count, s, 0);
}

K3Geom k3_geom(const LdpLayout &lay, int extra = 0, bool sq = false) {
  K3Geom g;
  const size_t cap = lay.cap;
  const size_t ltile = sq ? ((cap * (cap | 1) + 1) & ~(size_t)1) : ((cap * (cap - 1) / 2 + 1) & ~(size_t)1);
  size_t pw = ltile + ((lay.n + 1) & ~1) + ((lay.nth + 1) & ~1) +
              ((cap + 1) & ~1) + ((cap + 1) / 2 + 1) + extra;
  pw = (pw + 1) & ~(size_t)1;
  g.per_warp = pw;
  g.small_total = (size_t)(lay.total - lay.oDth);
  const size_t budget = 200 * 1024 / sizeof(double);
  int w = (int)((budget - g.small_total) / pw);
  if (g.small_total + pw > budget) w = 0;
  if (w > 8) w = 8;
  g.warps = w;
  g.smem = (g.small_total + pw * (w > 0 ? w : 1)) * sizeof(double);
  return g;
}

// the square (unpredicated-sweep) tile is used when it still leaves at least 6 warps per CTA
inline bool k3_square(const LdpLayout &lay, int extra) {
  return lay.cap <= 64 && k3_geom(lay, extra, true).warps >= 6;
}

template <int T, bool SQ = false>
int launch_big(const LdpLayout &lay, const LdpSettings &st, const double *cblk, const ThetaSegs &th,
               int64_t N, const SolveOut &out, int sm_count, cudaStream_t s) {
  if constexpr (T == 2 && !SQ) {
    if (k3_square(lay, 0)) return launch_big<T, true>(lay, st, cblk, th, N, out, sm_count, s);
  }
  const K3Geom g = k3_geom(lay, 0, SQ);
  auto kern = ldp_big_kernel<T, SQ>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, g.warps * 32, g.smem);
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = (N + g.warps - 1) / g.warps;
  if (blocks > (int64_t)sm_count * per_sm) blocks = (int64_t)sm_count * per_sm;
  kern<<<(unsigned)blocks, g.warps * 32, g.smem, s>>>(lay, st, cblk, th, N, out, g.warps, (int)g.per_warp);
  return (int)cudaGetLastError();
}

template <int T, bool SQ = false>
int launch_big_sim(const LdpLayout &lay, const LdpSettings &st, const double *cblk, int64_t N,
                   const SolveOut &out, const SimDesc &sim, int sm_count, cudaStream_t s) {
  const int extra = ((sim.nx + 1) & ~1) + ((lay.nu + 1) & ~1) + (sim.ext ? ((sim.nx + 1) & ~1) + ((sim.ny + 1) & ~1) : 0);
  if constexpr (T == 2 && !SQ) {
    if (k3_square(lay, extra)) return launch_big_sim<T, true>(lay, st, cblk, N, out, sim, sm_count, s);
  }
  const K3Geom g = k3_geom(lay, extra, SQ);
  if (g.warps < 1) return (int)cudaErrorNotSupported;
  auto kern = ldp_big_sim_kernel<T, SQ>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, g.warps * 32, g.smem);
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = (N + g.warps - 1) / g.warps;
  if (blocks > (int64_t)sm_count * per_sm) blocks = (int64_t)sm_count * per_sm;
  kern<<<(unsigned)blocks, g.warps * 32, g.smem, s>>>(lay, st, cblk, N, sim, out.flag, out.iters, g.warps,
                                                      (int)g.per_warp);
  return (int)cudaGetLastError();
}

}  // namespace

// instances redone at full capacity by the two-pass CTA path on `device` since the process started
long long ldp_cta_redo_count(int device) {
  unsigned long long *p = g_redo_counter(device);
  unsigned long long v = 0;
  if (!p || cudaMemcpy(&v, p, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return -1; }
  return (long long)v;
}
// capacity of the first (two CTAs per SM) pass for this layout; 0: single pass
int ldp_cta_fast_cap(const LdpLayout &lay) {
  const bool coop = k3_geom(lay).warps <= 2 && cta_smem(lay) <= 220 * 1024;
  return coop ? cta_fast_cap(lay) : 0;
}

bool ldp_cta_supports(const LdpLayout &lay) {
  if (lay.n < 1 || lay.n > 256 || lay.cap > 256 || lay.m > 512) return false;
  return k3_geom(lay).warps >= 1;
}

int launch_ldp_cta(const LdpLayout &lay, const LdpSettings &st, const double *cblk,
                   const ThetaSegs &th, int64_t N, const SolveOut &out, const SimDesc *sim,
                   int sm_count, void *stream) {
  if (N <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int need = lay.cap > lay.n ? lay.cap : lay.n;
  if ((lay.m + 1) / 2 > need) need = (lay.m + 1) / 2;
  const int T = (need + 31) / 32;
  if (sim) {   // closed loop: warp mode (one trajectory per warp)
    switch (T) {
      case 1: case 2: return launch_big_sim<2>(lay, st, cblk, N, out, *sim, sm_count, s);
      case 3: case 4: return launch_big_sim<4>(lay, st, cblk, N, out, *sim, sm_count, s);
      case 5: case 6: return launch_big_sim<6>(lay, st, cblk, N, out, *sim, sm_count, s);
      case 7: case 8: return launch_big_sim<8>(lay, st, cblk, N, out, *sim, sm_count, s);
      default: return (int)cudaErrorInvalidValue;
    }
  }
  // few QPs fit in an SM's shared memory -> give each QP a whole CTA (cooperative u / scan)
  const bool coop = k3_geom(lay).warps <= 2 && cta_smem(lay) <= 220 * 1024;
  if (coop) {
    switch (T) {
      case 1: case 2: return launch_cta<2>(lay, st, cblk, th, N, out, sm_count, s);
      case 3: case 4: return launch_cta<4>(lay, st, cblk, th, N, out, sm_count, s);
      case 5: case 6: return launch_cta<6>(lay, st, cblk, th, N, out, sm_count, s);
      case 7: case 8: return launch_cta<8>(lay, st, cblk, th, N, out, sm_count, s);
      default: return (int)cudaErrorInvalidValue;
    }
  }
  switch (T) {
    case 1: case 2: return launch_big<2>(lay, st, cblk, th, N, out, sm_count, s);
    case 3: case 4: return launch_big<4>(lay, st, cblk, th, N, out, sm_count, s);
    case 5: case 6: return launch_big<6>(lay, st, cblk, th, N, out, sm_count, s);
    case 7: case 8: return launch_big<8>(lay, st, cblk, th, N, out, sm_count, s);
    default: return (int)cudaErrorInvalidValue;
  }
}
