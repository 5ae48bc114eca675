/* TEST INFRASTRUCTURE ONLY -- see ldp_oracle.h for scope, provenance and parity status.
 *
 * Plain C, double precision, explicit fma() in a fixed summation order so that a
 * sequential GPU implementation of the same rules can be compared bit for bit.
 *
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -pthread -fPIC -shared (see Makefile).
 */
#include "ldp_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct {
  int k;          /* working-set size */
  int sing;       /* last added row is dependent (D[k-1] == 0) */
  int32_t *ws;    /* row index per position */
  int8_t *side;   /* +1 upper, -1 lower */
  double *lam, *lams, *L, *D, *Dinv, *z, *y, *u, *dup, *dlo;
  uint8_t *inws;  /* m flags */
  int32_t *tail_row; int8_t *tail_side; double *tail_lam;   /* rebuild scratch */
} work_t;

static work_t *work_new(const orc_problem *p) {
  int n = p->n, m = p->m, cap = (m > n ? m : n) + 1;
  work_t *w = (work_t *)calloc(1, sizeof(work_t));
  w->ws = (int32_t *)calloc(cap, sizeof(int32_t));
  w->side = (int8_t *)calloc(cap, 1);
  w->lam = (double *)calloc(cap, sizeof(double));
  w->lams = (double *)calloc(cap, sizeof(double));
  w->L = (double *)calloc((size_t)cap * (cap + 1) / 2, sizeof(double));
  w->D = (double *)calloc(cap, sizeof(double));
  w->Dinv = (double *)calloc(cap, sizeof(double));
  w->z = (double *)calloc(cap, sizeof(double));
  w->y = (double *)calloc(cap, sizeof(double));
  w->u = (double *)calloc(n, sizeof(double));
  w->dup = (double *)calloc(m, sizeof(double));
  w->dlo = (double *)calloc(m, sizeof(double));
  w->inws = (uint8_t *)calloc(m, 1);
  w->tail_row = (int32_t *)calloc(cap, sizeof(int32_t));
  w->tail_side = (int8_t *)calloc(cap, 1);
  w->tail_lam = (double *)calloc(cap, sizeof(double));
  return w;
}
static void work_free(work_t *w) {
  free(w->ws); free(w->side); free(w->lam); free(w->lams); free(w->L); free(w->D); free(w->Dinv);
  free(w->z); free(w->y); free(w->u); free(w->dup); free(w->dlo); free(w->inws);
  free(w->tail_row); free(w->tail_side); free(w->tail_lam); free(w);
}
static void work_reset(const orc_problem *p, work_t *w) {
  w->k = 0; w->sing = 0;
  memset(w->inws, 0, (size_t)p->m);
  memset(w->u, 0, sizeof(double) * (size_t)p->n);
}

/* Canonical cross-element sum, shared by the oracle and every kernel so that results can be
 * compared bit for bit whatever the degree of parallelism: element i is accumulated (fma
 * chain, increasing i) on "lane" i mod 32, then the 32 lane sums are combined with an
 * xor-butterfly (16, 8, 4, 2, 1).  stride lets a/b be gathered. */
static double lane_tree(const double *acc, int base, int step, int nl) {
  if (step >= 32) return acc[base];
  if (base + step >= nl) return lane_tree(acc, base, 2 * step, nl);  /* right half: padding zeros */
  return lane_tree(acc, base, 2 * step, nl) + lane_tree(acc, base + step, 2 * step, nl);
}
static double lanesum(const double *a, const double *b, int len) {
  double acc[32];
  int nl = len < 32 ? len : 32;
  if (nl <= 0) return 0.0;
  for (int l = 0; l < nl; ++l) {
    double s = a[l] * b[l];
    for (int i = l + 32; i < len; i += 32) s = fma(a[i], b[i], s);
    acc[l] = s;
  }
  /* value on lane 0 after the butterfly: S(base, step) = S(base, 2 step) + S(base + step, 2 step);
   * sub-trees holding only padding zeros are skipped (x + 0 == x) */
  return lane_tree(acc, 0, 1, nl);
}

/* packed unit-lower L: row i holds its i sub-diagonal entries at offset i(i-1)/2 */
#define LROW(w, i) ((w)->L + ((size_t)(i) * ((i) - 1)) / 2)

/* mpc_update_qp (codegen/mpc_update_qp.c:1-10): d = d0 + Dth theta */
static void update_bounds(const orc_problem *p, const double *th, work_t *w, double *fl) {
  for (int j = 0; j < p->m; ++j) {
    double acc = 0.0;
    const double *row = p->Dth + (size_t)j * p->nth;
    for (int t = 0; t < p->nth; ++t) acc = fma(row[t], th[t], acc);
    w->dup[j] = p->du[j] + acc;
    w->dlo[j] = p->dl[j] + acc;
  }
  if (fl) *fl += 2.0 * p->m * p->nth;
}

/* append row `j` (side s) to the factorisation; SURVEY Appendix A "LDL-add" */
static void ldl_add(const orc_problem *p, work_t *w, int j, int s, double *fl) {
  int k = w->k;
  double *lrow = LROW(w, k);
  const double *g = p->Gram + (size_t)j * p->m;
  double dnew = g[j];
  if (p->sense[j] & ORC_SOFT) dnew += p->rho_soft;
  for (int i = 0; i < k; ++i) {
    double yi = g[w->ws[i]];
    const double *li = LROW(w, i);
    for (int c = 0; c < i; ++c) yi = fma(-li[c], w->y[c], yi);
    w->y[i] = yi;
  }
  for (int i = 0; i < k; ++i) lrow[i] = w->y[i] * w->Dinv[i];
  dnew = dnew - lanesum(w->y, lrow, k);
  w->ws[k] = j; w->side[k] = (int8_t)s; w->lam[k] = 0.0; w->inws[j] = 1;
  if (dnew < p->zero_tol) { w->sing = 1; dnew = 0.0; }
  w->D[k] = dnew;
  w->Dinv[k] = w->sing ? 0.0 : 1.0 / dnew;  /* pivots are inverted once, when they change */
  w->k = k + 1;
  if (fl) *fl += (double)k * k + 3.0 * k;
}

/* delete position r; rank-one update of the trailing block (Gill et al. C1) */
static void ldl_remove(const orc_problem *p, work_t *w, int r, double *fl) {
  int k = w->k;
  double alpha = w->D[r];
  double *wv = w->y; /* scratch: column r below the diagonal */
  for (int i = r + 1; i < k; ++i) wv[i] = LROW(w, i)[r];
  /* shift rows up / columns left */
  for (int i = r + 1; i < k; ++i) {
    double *src = LROW(w, i), *dst = LROW(w, i - 1);
    for (int c = 0; c < r; ++c) dst[c] = src[c];
    for (int c = r + 1; c < i; ++c) dst[c - 1] = src[c];
    w->D[i - 1] = w->D[i]; w->Dinv[i - 1] = w->Dinv[i];
    wv[i - 1] = wv[i];
    w->ws[i - 1] = w->ws[i]; w->side[i - 1] = w->side[i]; w->lam[i - 1] = w->lam[i];
  }
  k -= 1;
  for (int j = r; j < k; ++j) {
    double pj = wv[j];
    double t = alpha * pj;
    double dbar = fma(t, pj, w->D[j]);
    double rinv = 1.0 / dbar;
    double beta = t * rinv;
    alpha = (w->D[j] * alpha) * rinv;
    w->D[j] = dbar; w->Dinv[j] = rinv;
    for (int i = j + 1; i < k; ++i) {
      double *li = LROW(w, i);
      wv[i] = fma(-pj, li[j], wv[i]);
      li[j] = fma(beta, wv[i], li[j]);
    }
  }
  w->k = k;
  w->sing = 0;
  if (k > 0 && !(w->D[k - 1] >= p->zero_tol)) { w->sing = 1; w->D[k - 1] = 0.0; w->Dinv[k - 1] = 0.0; }
  if (fl) *fl += 3.0 * (k - r) * (k - r) + 2.0 * k;
}

/* Repair of the factorisation for problems with SOFT rows (DAQP's "reorder and refactorise"
 * step, restated; libdaqp itself is not in /root/reference -- see ldp_oracle.h).
 *
 * A soft row contributes only rho (1e-6) to its own pivot, so a HARD row factorised after soft rows
 * it depends on gets a pivot of order rho |c_soft|^2 (1e-10 and below on AFTI-16 with output bounds
 * on every step).  Such a pivot drowns in the rounding of 1 - sum l^2 D: the zero_tol test fires at
 * random and a problem that is feasible by construction ends as INFEASIBLE.  The repair rebuilds
 * L, D from scratch with the hard rows first (their pivots are then Schur complements among hard
 * rows only) and the soft rows last (pivots >= rho), multipliers kept.  A hard row that is dependent
 * on the hard block itself is genuinely singular: it goes to the very end with a zero pivot, so
 * that nothing is factorised after it.  Two such rows cannot both be handled: ORC_CYCLE. */
static int rebuild(const orc_problem *p, work_t *w, double *fl) {
  const int k = w->k;
  int32_t *tr = w->tail_row; int8_t *ts = w->tail_side; double *tl = w->tail_lam;
  for (int i = 0; i < k; ++i) { tr[i] = w->ws[i]; ts[i] = w->side[i]; tl[i] = w->lam[i]; }
  w->k = 0; w->sing = 0;
  int defer = -1;
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < k; ++i) {
      if (((p->sense[tr[i]] & ORC_SOFT) != 0) != pass) continue;
      ldl_add(p, w, tr[i], ts[i], fl);
      w->lam[w->k - 1] = tl[i];
      if (w->sing) {
        if (pass == 1 || defer >= 0) return ORC_CYCLE;
        defer = i; w->k -= 1; w->sing = 0;
      }
    }
  if (defer >= 0) {
    ldl_add(p, w, tr[defer], ts[defer], fl);
    w->lam[w->k - 1] = tl[defer];
    w->sing = 1; w->D[w->k - 1] = 0.0; w->Dinv[w->k - 1] = 0.0;
  }
  return 0;
}

/* the working set holds a soft row: a pivot below zero_tol is then suspect (see rebuild) */
static int has_soft(const orc_problem *p, const work_t *w) {
  if (p->no_repair) return 0;
  for (int i = 0; i < w->k; ++i) if (p->sense[w->ws[i]] & ORC_SOFT) return 1;
  return 0;
}

/* append row j; returns 0 or an exit flag */
static int add_row(const orc_problem *p, work_t *w, int j, int s, double *fl) {
  ldl_add(p, w, j, s, fl);
  if (w->sing && has_soft(p, w)) return rebuild(p, w, fl);
  return 0;
}

static int remove_pos(const orc_problem *p, work_t *w, int r, double *fl) {
  w->inws[w->ws[r]] = 0;
  ldl_remove(p, w, r, fl);
  if (w->sing && has_soft(p, w)) return rebuild(p, w, fl);
  return 0;
}

/* the dual active-set loop */
static int ldp_loop(const orc_problem *p, work_t *w, int32_t *iters, double *soft_slack_out,
                    double *fl) {
  const int n = p->n, m = p->m;
  int flag = ORC_ITERLIMIT, iter, cycle = 0, repaired = 0, rc;
  double best_fval = -1.0, fval = 0.0, soft_slack = 0.0;
  for (iter = 1; iter < p->iter_limit; ++iter) {
    int k = w->k, nblock = 0;
    if (!w->sing) {
      /* 1. CSP: L D L' lam* = -d_W */
      for (int i = 0; i < k; ++i) {
        int j = w->ws[i];
        double zi = -(w->side[i] > 0 ? w->dup[j] : w->dlo[j]);
        const double *li = LROW(w, i);
        for (int c = 0; c < i; ++c) zi = fma(-li[c], w->z[c], zi);
        w->z[i] = zi;
      }
      for (int i = 0; i < k; ++i) w->lams[i] = w->z[i] * w->Dinv[i];
      for (int jx = k - 1; jx > 0; --jx) {
        const double *lj = LROW(w, jx);
        double xj = w->lams[jx];
        for (int i = 0; i < jx; ++i) w->lams[i] = fma(-lj[i], xj, w->lams[i]);
      }
      if (fl) *fl += 2.0 * k * k;
      /* 2. blocking set */
      for (int i = 0; i < k; ++i) {
        if (p->sense[w->ws[i]] & ORC_IMMUTABLE) continue;
        if (w->side[i] > 0 ? (w->lams[i] < -p->dual_tol) : (w->lams[i] > p->dual_tol)) nblock++;
      }
      if (nblock == 0) {
        /* 3. primal iterate, objective, most violated row */
        for (int c = 0; c < n; ++c) w->u[c] = 0.0;
        for (int i = 0; i < k; ++i) {
          const double *mr = p->Mbar + (size_t)w->ws[i] * n;
          double li = w->lams[i];
          for (int c = 0; c < n; ++c) w->u[c] = fma(-li, mr[c], w->u[c]);
        }
        fval = lanesum(w->u, w->u, n); soft_slack = 0.0;
        for (int i = 0; i < k; ++i) {
          int sf = (p->sense[w->ws[i]] & ORC_SOFT) != 0;
          w->z[i] = sf ? w->lams[i] : 0.0;   /* z is free here: soft multipliers, 0 elsewhere */
          if (sf) {
            double sl = p->rho_soft * fabs(w->lams[i]);
            if (sl > soft_slack) soft_slack = sl;
          }
        }
        fval = fma(p->rho_soft, lanesum(w->z, w->lams, k), fval);
        if (fl) *fl += 2.0 * k * n + 2.0 * n;
        if (fval > p->fval_bound) { flag = ORC_INFEASIBLE; break; }
        double best = -p->primal_tol;
        int add = -1, addside = 0;
        for (int j = 0; j < m; ++j) {
          if (w->inws[j] || (p->sense[j] & ORC_IMMUTABLE)) continue;
          const double *mr = p->Mbar + (size_t)j * n;
          double mu = 0.0;
          for (int c = 0; c < n; ++c) mu = fma(mr[c], w->u[c], mu);
          double su = w->dup[j] - mu;
          if (su < best) { best = su; add = j; addside = 1; }
          else {
            double sl = mu - w->dlo[j];
            if (sl < best) { best = sl; add = j; addside = -1; }
          }
        }
        if (fl) *fl += 2.0 * n * (m - k);
        if (add < 0) { flag = soft_slack > p->primal_tol ? ORC_SOFT_OPTIMAL : ORC_OPTIMAL; break; }
        /* cycle guard, evaluated where the dual objective is fresh: a new dual-feasible iterate
         * that still violates a row must have raised fval by progress_tol, or the counter runs */
        int rebuilt = 0;
        if (fval - p->progress_tol > best_fval) { best_fval = fval; cycle = 0; }
        else if (cycle++ > p->cycle_tol) {
          /* second chance for problems with soft rows: refactorise once, restart the count */
          if (repaired || !has_soft(p, w)) { flag = ORC_CYCLE; break; }
          repaired = 1; rebuilt = 1; best_fval = -1.0; cycle = 0;
        }
        if (p->cap > 0 && k >= p->cap) { flag = ORC_CAPACITY; break; }
        for (int i = 0; i < k; ++i) w->lam[i] = w->lams[i];
        if (rebuilt && (rc = rebuild(p, w, fl)) != 0) { flag = rc; break; }
        if ((rc = add_row(p, w, add, addside, fl)) != 0) { flag = rc; break; }
        continue;
      }
      /* 4. blocking: ratio test, step, remove */
      double alpha = INFINITY; int rm = -1;
      for (int i = 0; i < k; ++i) {
        if (p->sense[w->ws[i]] & ORC_IMMUTABLE) continue;
        if (!(w->side[i] > 0 ? (w->lams[i] < -p->dual_tol) : (w->lams[i] > p->dual_tol))) continue;
        double a = -w->lam[i] / (w->lams[i] - w->lam[i]);
        if (a < alpha) { alpha = a; rm = i; }
      }
      if (!(alpha < INFINITY)) { flag = ORC_CYCLE; break; }  /* blocked but no finite ratio */
      for (int i = 0; i < k; ++i) w->lam[i] = fma(alpha, w->lams[i] - w->lam[i], w->lam[i]);
      if (fl) *fl += 3.0 * k;
      if ((rc = remove_pos(p, w, rm, fl)) != 0) { flag = rc; break; }
    } else {
      /* 5. singular: direction p = s [-L^-T l ; 1] on the first k-1 rows */
      int kk = k - 1;
      const double *ll = LROW(w, kk);
      for (int i = 0; i < kk; ++i) w->lams[i] = -ll[i];
      for (int jx = kk - 1; jx > 0; --jx) {
        const double *lj = LROW(w, jx);
        double xj = w->lams[jx];
        for (int i = 0; i < jx; ++i) w->lams[i] = fma(-lj[i], xj, w->lams[i]);
      }
      w->lams[kk] = 1.0;
      if (w->side[kk] < 0) for (int i = 0; i <= kk; ++i) w->lams[i] = -w->lams[i];
      if (fl) *fl += (double)kk * kk;
      double alpha = INFINITY; int rm = -1;
      for (int i = 0; i < k; ++i) {
        if (p->sense[w->ws[i]] & ORC_IMMUTABLE) continue;
        if (!(w->side[i] > 0 ? (w->lams[i] < -p->dual_tol) : (w->lams[i] > p->dual_tol))) continue;
        double a = -w->lam[i] / w->lams[i];
        if (a < alpha) { alpha = a; rm = i; }
      }
      if (rm < 0) { flag = ORC_INFEASIBLE; break; }
      for (int i = 0; i < k; ++i) w->lam[i] = fma(alpha, w->lams[i], w->lam[i]);
      if (fl) *fl += 3.0 * k;
      if ((rc = remove_pos(p, w, rm, fl)) != 0) { flag = rc; break; }
    }
  }
  *iters = iter;
  if (soft_slack_out) *soft_slack_out = soft_slack;
  return flag;
}

/* cold start: working set = equality rows (ACTIVE bit), SURVEY Appendix A "Start state" */
static int activate_equalities(const orc_problem *p, work_t *w) {
  for (int j = 0; j < p->m; ++j)
    if (p->sense[j] & ORC_ACTIVE) {
      if (p->cap > 0 && w->k >= p->cap) return ORC_OVERDETERMINED;
      ldl_add(p, w, j, (p->sense[j] & ORC_LOWER) ? -1 : 1, NULL);
      if (w->sing) return ORC_OVERDETERMINED;
    }
  return 0;
}

/* mpc_get_solution (codegen/mpc_update_qp.c:14-22) with x = R^-1 u (ldp2qp_solution) */
static void get_solution(const orc_problem *p, const double *th, const work_t *w, double *ctl,
                         double *fl) {
  for (int i = 0; i < p->nu; ++i) {
    const double *rr = p->Rinv + (size_t)i * p->n;
    double x = lanesum(rr, w->u, p->n);
    double sh = p->uoff[i];
    const double *ur = p->Uth + (size_t)i * p->nth;
    for (int t = 0; t < p->nth; ++t) sh = fma(ur[t], th[t], sh);
    ctl[i] = x + sh;
  }
  if (fl) *fl += (double)p->nu * (2 * p->n - p->nu + 1) + 2.0 * p->nu * p->nth;
}

/* Warm start from a caller-supplied working set (the batch form of DAQP_WARMSTART,
 * codegen/mpc_update_qp.c:44-47): first the equality rows, exactly as in a cold start, then the
 * rows named in wu / wl in ascending row order with lam = 0 (an upper bit wins over a lower
 * bit); a named row that would make the factorisation singular is left out, IMMUTABLE rows are
 * never named. */
static int activate_rows(const orc_problem *p, work_t *w, const uint64_t *wu, const uint64_t *wl) {
  int flag = activate_equalities(p, w);
  if (flag) return flag;
  for (int j = 0; j < p->m; ++j) {
    if (p->sense[j] & (ORC_ACTIVE | ORC_IMMUTABLE)) continue;
    const int up = wu && ((wu[j >> 6] >> (j & 63)) & 1ull), lo = wl && ((wl[j >> 6] >> (j & 63)) & 1ull);
    if (!up && !lo) continue;
    if (p->cap > 0 && w->k >= p->cap) continue;
    ldl_add(p, w, j, up ? 1 : -1, NULL);
    if (w->sing) { w->k -= 1; w->inws[j] = 0; w->sing = 0; }   /* undo the trial add */
  }
  return 0;
}

static int solve_with(const orc_problem *p, work_t *w, const double *th, double *ctl,
                      int32_t *iters, uint64_t *au, uint64_t *al, double *fl, int cold) {
  int flag = 0, it = 0;
  update_bounds(p, th, w, fl);
  if (cold == 2) {            /* warm start from the masks in au / al (in/out) */
    work_reset(p, w);
    flag = activate_rows(p, w, au, al);
  } else if (cold) {
    work_reset(p, w);
    flag = activate_equalities(p, w);
  }
  if (flag == 0) flag = ldp_loop(p, w, &it, NULL, fl);
  get_solution(p, th, w, ctl, fl);
  if (iters) *iters = it;
  int words = (p->m + 63) / 64;
  if (au) memset(au, 0, sizeof(uint64_t) * words);
  if (al) memset(al, 0, sizeof(uint64_t) * words);
  for (int i = 0; i < w->k; ++i) {
    int j = w->ws[i];
    if (w->side[i] > 0) { if (au) au[j >> 6] |= 1ull << (j & 63); }
    else if (al) al[j >> 6] |= 1ull << (j & 63);
  }
  return flag;
}

int orc_solve(const orc_problem *p, const double *theta, double *u, int32_t *iters,
              uint64_t *act_u, uint64_t *act_l, double *flops) {
  work_t *w = work_new(p);
  if (flops) *flops = 0.0;
  int flag = solve_with(p, w, theta, u, iters, act_u, act_l, flops, 1);
  work_free(w);
  return flag;
}

int orc_num_threads(void) {
  long c = sysconf(_SC_NPROCESSORS_ONLN);
  return c > 0 ? (int)c : 1;
}

/* static contiguous chunks over pthreads (one workspace per thread) */
typedef void (*chunk_fn)(void *ctx, int64_t lo, int64_t hi);
typedef struct { chunk_fn fn; void *ctx; int64_t lo, hi; } chunk_t;
static void *chunk_tramp(void *a) { chunk_t *c = (chunk_t *)a; c->fn(c->ctx, c->lo, c->hi); return NULL; }
static void parallel_for(int64_t N, int nthreads, chunk_fn fn, void *ctx) {
  if (nthreads <= 0) nthreads = orc_num_threads();
  if (nthreads > N) nthreads = N > 0 ? (int)N : 1;
  if (nthreads <= 1) { fn(ctx, 0, N); return; }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  chunk_t *ch = (chunk_t *)malloc(sizeof(chunk_t) * nthreads);
  for (int t = 0; t < nthreads; ++t) {
    ch[t].fn = fn; ch[t].ctx = ctx;
    ch[t].lo = N * t / nthreads; ch[t].hi = N * (t + 1) / nthreads;
    pthread_create(&th[t], NULL, chunk_tramp, &ch[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th); free(ch);
}

typedef struct {
  const orc_problem *p; const double *Theta; double *U;
  int32_t *flag, *iters; uint64_t *au, *al; double *flops; int mode;   /* 1 cold, 2 warm from masks */
} batch_ctx;
static void batch_chunk(void *vc, int64_t lo, int64_t hi) {
  batch_ctx *c = (batch_ctx *)vc;
  const orc_problem *p = c->p;
  int words = (p->m + 63) / 64;
  work_t *w = work_new(p);
  for (int64_t q = lo; q < hi; ++q) {
    double fl = 0.0;
    int32_t it = 0;
    int f = solve_with(p, w, c->Theta + q * p->nth, c->U + q * p->nu, &it,
                       c->au ? c->au + q * words : NULL, c->al ? c->al + q * words : NULL,
                       c->flops ? &fl : NULL, c->mode);
    if (c->flag) c->flag[q] = f;
    if (c->iters) c->iters[q] = it;
    if (c->flops) c->flops[q] = fl;
  }
  work_free(w);
}
void orc_solve_batch(const orc_problem *p, const double *Theta, int64_t N, double *U,
                     int32_t *flag, int32_t *iters, uint64_t *act_u, uint64_t *act_l,
                     double *flops, int nthreads) {
  batch_ctx c = {p, Theta, U, flag, iters, act_u, act_l, flops, 1};
  parallel_for(N, nthreads, batch_chunk, &c);
}
void orc_solve_batch_warm(const orc_problem *p, const double *Theta, int64_t N, double *U,
                          int32_t *flag, int32_t *iters, uint64_t *act_u, uint64_t *act_l,
                          int nthreads) {
  batch_ctx c = {p, Theta, U, flag, iters, act_u, act_l, NULL, 2};
  parallel_for(N, nthreads, batch_chunk, &c);
}

typedef struct {
  const orc_problem *p; int nx, nr, nd, nuprev; const double *Fx, *Gx, *Gdx, *X0, *R, *D, *U0;
  const double *Rtraj, *Dtraj;   /* nr x T, nd x T shared trajectories (simulation.jl:70-80), or NULL */
  int T, warm; double *Xend, *Ulast, *Utraj; int32_t *flag, *iters;
  const orc_sim_ext *ext;        /* true plant / observer / measurement noise, or NULL */
} sim_ctx;
static void sim_chunk(void *vc, int64_t lo, int64_t hi) {
  sim_ctx *s = (sim_ctx *)vc;
  const orc_problem *p = s->p;
  const int nu = p->nu, nth = p->nth, nx = s->nx, nr = s->nr, nd = s->nd, nuprev = s->nuprev;
  work_t *w = work_new(p);
  double *th = (double *)calloc(nth, sizeof(double));
  double *x = (double *)calloc(nx, sizeof(double));
  double *xn = (double *)calloc(nx, sizeof(double));
  double *u = (double *)calloc(nu, sizeof(double));
  const orc_sim_ext *e = s->ext;
  const int ny = e ? e->ny : 0;
  const int obs = e && e->C && e->K;
  /* the separate true-state bookkeeping runs only when a plant / observer field is set (flops alone does not count) */
  const int ex = e && (e->Ft || e->Gt || e->Gdt || e->C || e->K || e->Vtraj || e->Xhat_end);
  /* sliding preview windows (simulation.jl:102-104, get_preview :128-134): step t reads nr entries at t * rstride */
  const int rstride = (e && e->rstride > 0) ? e->rstride : nr, dstride = (e && e->dstride > 0) ? e->dstride : nd;
  /* true plant (defaults to the prediction model) */
  const double *Ft = (e && e->Ft) ? e->Ft : s->Fx, *Gt = (e && e->Gt) ? e->Gt : s->Gx;
  const double *Gdt = (e && e->Gdt) ? e->Gdt : s->Gdx;
  double *xs = (double *)calloc(nx, sizeof(double));           /* true state when ext != NULL */
  double *inn = (double *)calloc(ny > 0 ? ny : 1, sizeof(double));
  for (int64_t q = lo; q < hi; ++q) {
    int worst = ORC_OPTIMAL, tot = 0;
    double flq = 0.0;
    for (int i = 0; i < nx; ++i) x[i] = xs[i] = s->X0[q * nx + i];   /* set_state!(mpc, x0), simulation.jl:92 */
    for (int i = 0; i < nu; ++i) u[i] = (s->U0 && i < nuprev) ? s->U0[q * nuprev + i] : 0.0;
    for (int t = 0; t < s->T; ++t) {
      /* theta = [x; r; d; uprev] (explicit.jl:54-63); x is the estimate when an observer runs */
      int o = nx;
      for (int i = 0; i < nr; ++i) th[o++] = s->Rtraj ? s->Rtraj[(size_t)t * rstride + i] : (s->R ? s->R[q * nr + i] : 0.0);
      for (int i = 0; i < nd; ++i) th[o++] = s->Dtraj ? s->Dtraj[(size_t)t * dstride + i] : (s->D ? s->D[q * nd + i] : 0.0);
      for (int i = 0; i < nuprev; ++i) th[o++] = u[i];
      const double *dk = th + nx + nr;
      if (obs) {
        /* y = C x + Dd d + v (simulation.jl:94), innovation and correction (observer.jl:114-118) */
        for (int i = 0; i < ny; ++i) {
          double y = e->Vtraj ? e->Vtraj[(size_t)t * ny + i] : 0.0, yh = 0.0;
          for (int c = 0; c < nx; ++c) y = fma(e->C[i * nx + c], xs[c], y);
          for (int c = 0; c < nd; ++c) y = fma(e->Dd ? e->Dd[i * nd + c] : 0.0, dk[c], y);
          for (int c = 0; c < nx; ++c) yh = fma(e->C[i * nx + c], x[c], yh);
          for (int c = 0; c < nd; ++c) yh = fma(e->Dd ? e->Dd[i * nd + c] : 0.0, dk[c], yh);
          inn[i] = y - yh;
        }
        for (int j = 0; j < nx; ++j) {
          double acc = x[j];
          for (int i = 0; i < ny; ++i) acc = fma(e->K[j * ny + i], inn[i], acc);
          xn[j] = acc;
        }
        for (int j = 0; j < nx; ++j) x[j] = xn[j];
      } else if (ex) {
        for (int j = 0; j < nx; ++j) x[j] = xs[j];           /* the controller measures the true state */
      }
      for (int i = 0; i < nx; ++i) th[i] = x[i];
      int32_t it = 0;
      int f = solve_with(p, w, th, u, &it, NULL, NULL, (e && e->flops) ? &flq : NULL, !(s->warm && t > 0));
      tot += it;
      if (f < 0) { if (worst > 0 || f < worst) worst = f; }
      else if (worst > 0 && f > worst) worst = f;
      if (s->Utraj) for (int i = 0; i < nu; ++i) s->Utraj[(q * s->T + t) * nu + i] = u[i];
      /* model step x+ = F x + G u + Gd d (model.jl:70; predict!, observer.jl:104-108) */
      for (int i = 0; i < nx; ++i) {
        double acc = 0.0;
        for (int c = 0; c < nx; ++c) acc = fma(s->Fx[i * nx + c], x[c], acc);
        for (int c = 0; c < nu; ++c) acc = fma(s->Gx[i * nu + c], u[c], acc);
        for (int c = 0; c < nd; ++c) acc = fma(s->Gdx[i * nd + c], dk[c], acc);
        xn[i] = acc;
      }
      for (int i = 0; i < nx; ++i) x[i] = xn[i];
      if (ex) {   /* true plant step (scenario.dynamics, simulation.jl:110) */
        for (int i = 0; i < nx; ++i) {
          double acc = 0.0;
          for (int c = 0; c < nx; ++c) acc = fma(Ft[i * nx + c], xs[c], acc);
          for (int c = 0; c < nu; ++c) acc = fma(Gt[i * nu + c], u[c], acc);
          for (int c = 0; c < nd; ++c) acc = fma(Gdt[i * nd + c], dk[c], acc);
          xn[i] = acc;
        }
        for (int i = 0; i < nx; ++i) xs[i] = xn[i];
      }
    }
    for (int i = 0; i < nx; ++i) s->Xend[q * nx + i] = ex ? xs[i] : x[i];
    if (e && e->Xhat_end) for (int i = 0; i < nx; ++i) e->Xhat_end[q * nx + i] = x[i];
    for (int i = 0; i < nu; ++i) s->Ulast[q * nu + i] = u[i];
    if (s->flag) s->flag[q] = worst;
    if (s->iters) s->iters[q] = tot;
    if (e && e->flops) e->flops[q] = flq;
  }
  free(th); free(x); free(xn); free(u); free(xs); free(inn);
  work_free(w);
}
typedef struct {
  const orc_problem *p; int nx, nr, nuprev; const double *X0, *R, *D, *U0; const orc_sim_stepped *c;
  int T, warm; double *Xend, *Ulast, *Utraj; int32_t *flag, *iters;
} step_ctx;
/* acc = A[i, :] v (row-major, fma chain in index order) added onto acc */
static double row_fma(const double *A, int i, int n, const double *v, double acc) {
  if (A) for (int c = 0; c < n; ++c) acc = fma(A[i * n + c], v[c], acc);
  return acc;
}
static void step_chunk(void *vc, int64_t lo, int64_t hi) {
  step_ctx *s = (step_ctx *)vc;
  const orc_problem *p = s->p;
  const orc_sim_stepped *c = s->c;
  const int nu = p->nu, nth = p->nth, nx = s->nx, nr = s->nr, nuprev = s->nuprev;
  const int nxo = c->Ko ? c->nxo : nx, ndm = c->ndm, ndo = nxo - nx, ny = c->ny;
  work_t *w = work_new(p);
  double *th = (double *)calloc(nth, sizeof(double));
  double *xs = (double *)calloc(nx, sizeof(double)), *xn = (double *)calloc(nxo > nx ? nxo : nx, sizeof(double));
  double *xh = (double *)calloc(nxo, sizeof(double));
  double *u = (double *)calloc(nu, sizeof(double)), *dm = (double *)calloc(ndm > 0 ? ndm : 1, sizeof(double));
  double *inn = (double *)calloc(ny > 0 ? ny : 1, sizeof(double));
  const int words = (p->m + 63) / 64;
  uint64_t *au = (uint64_t *)calloc(words, sizeof(uint64_t)), *al = (uint64_t *)calloc(words, sizeof(uint64_t));
  for (int64_t q = lo; q < hi; ++q) {
    int worst = ORC_OPTIMAL, tot = 0;
    for (int i = 0; i < nx; ++i) xs[i] = s->X0[q * nx + i];
    for (int i = 0; i < nxo; ++i) xh[i] = i < nx ? xs[i] : 0.0;     /* set_state!(observer, x0): d-hat = 0 */
    for (int i = 0; i < nu; ++i) u[i] = (s->U0 && i < nuprev) ? s->U0[q * nuprev + i] : 0.0;
    for (int t = 0; t < s->T; ++t) {
      for (int i = 0; i < ndm; ++i) dm[i] = c->Dtraj ? c->Dtraj[(size_t)t * ndm + i] : (s->D ? s->D[q * ndm + i] : 0.0);
      if (c->Ko) {
        for (int i = 0; i < ny; ++i) {
          double y = c->Vtraj ? c->Vtraj[(size_t)t * ny + i] : 0.0;
          y = row_fma(c->Cm, i, nx, xs, y); y = row_fma(c->Ddm, i, ndm, dm, y);
          if (c->hm) y += c->hm[i];
          double yh = row_fma(c->Co, i, nxo, xh, 0.0); yh = row_fma(c->Ddo, i, ndm, dm, yh);
          if (c->ho) yh += c->ho[i];
          inn[i] = y - yh;
        }
        for (int j = 0; j < nxo; ++j) xn[j] = row_fma(c->Ko, j, ny, inn, xh[j]);
        for (int j = 0; j < nxo; ++j) xh[j] = xn[j];
      } else {
        for (int j = 0; j < nx; ++j) xh[j] = xs[j];
      }
      int o = 0;
      for (int i = 0; i < nx; ++i) th[o++] = xh[i];
      for (int i = 0; i < nr; ++i) th[o++] = c->Rtraj ? c->Rtraj[(size_t)t * nr + i] : (s->R ? s->R[q * nr + i] : 0.0);
      for (int i = 0; i < ndm; ++i) th[o++] = dm[i];
      for (int i = 0; i < ndo; ++i) th[o++] = xh[nx + i];
      for (int i = 0; i < nuprev; ++i) th[o++] = u[i];
      int32_t it = 0;
      /* warm: the working set travels between the steps as bit masks and the factorisation is rebuilt from
       * them (the batch warm start, orc_solve_batch_warm), which is what a loop stepped from the host does */
      int f = solve_with(p, w, th, u, &it, au, al, NULL, (s->warm && t > 0) ? 2 : 1);
      tot += it;
      if (f < 0) { if (worst > 0 || f < worst) worst = f; }
      else if (worst > 0 && f > worst) worst = f;
      if (s->Utraj) for (int i = 0; i < nu; ++i) s->Utraj[(q * s->T + t) * nu + i] = u[i];
      if (c->Ko) {   /* predict!(observer, u, d_measured) */
        for (int i = 0; i < nxo; ++i) {
          double acc = row_fma(c->Fo, i, nxo, xh, 0.0);
          acc = row_fma(c->Go, i, nu, u, acc); acc = row_fma(c->Gdo, i, ndm, dm, acc);
          xn[i] = c->fo ? acc + c->fo[i] : acc;
        }
        for (int i = 0; i < nxo; ++i) xh[i] = xn[i];
      }
      if (c->dyn) c->dyn(c->user, t, xs, u, dm, xn);
      else
        for (int i = 0; i < nx; ++i) {
          double acc = row_fma(c->Ft, i, nx, xs, 0.0);
          acc = row_fma(c->Gt, i, nu, u, acc); acc = row_fma(c->Gdt, i, ndm, dm, acc);
          xn[i] = c->ft ? acc + c->ft[i] : acc;
        }
      for (int i = 0; i < nx; ++i) xs[i] = xn[i];
    }
    for (int i = 0; i < nx; ++i) s->Xend[q * nx + i] = xs[i];
    if (c->Xhat_end) for (int i = 0; i < nxo; ++i) c->Xhat_end[q * nxo + i] = xh[i];
    for (int i = 0; i < nu; ++i) s->Ulast[q * nu + i] = u[i];
    if (s->flag) s->flag[q] = worst;
    if (s->iters) s->iters[q] = tot;
  }
  free(th); free(xs); free(xn); free(xh); free(u); free(dm); free(inn); free(au); free(al);
  work_free(w);
}
void orc_simulate_batch_stepped(const orc_problem *p, int nx, int nr, int nuprev,
                                const double *X0, const double *R, const double *D, const double *U0,
                                const orc_sim_stepped *c, int64_t N, int T, int warm,
                                double *Xend, double *Ulast, double *Utraj, int32_t *flag, int32_t *iters,
                                int nthreads) {
  step_ctx s = {p, nx, nr, nuprev, X0, R, D, U0, c, T, warm, Xend, Ulast, Utraj, flag, iters};
  parallel_for(N, c->dyn ? 1 : nthreads, step_chunk, &s);   /* a Python callback is not thread-safe */
}
void orc_simulate_batch(const orc_problem *p, int nx, int nr, int nd, int nuprev,
                        const double *Fx, const double *Gx, const double *Gdx,
                        const double *X0, const double *R, const double *D, const double *U0,
                        int64_t N, int T, int warm, double *Xend, double *Ulast, double *Utraj,
                        int32_t *flag, int32_t *iters, int nthreads) {
  orc_simulate_batch_traj(p, nx, nr, nd, nuprev, Fx, Gx, Gdx, X0, R, D, U0, NULL, NULL, N, T, warm, Xend, Ulast,
                          Utraj, flag, iters, nthreads);
}
void orc_simulate_batch_traj(const orc_problem *p, int nx, int nr, int nd, int nuprev,
                             const double *Fx, const double *Gx, const double *Gdx,
                             const double *X0, const double *R, const double *D, const double *U0,
                             const double *Rtraj, const double *Dtraj,
                             int64_t N, int T, int warm, double *Xend, double *Ulast, double *Utraj,
                             int32_t *flag, int32_t *iters, int nthreads) {
  orc_simulate_batch_ex(p, nx, nr, nd, nuprev, Fx, Gx, Gdx, X0, R, D, U0, Rtraj, Dtraj, NULL, N, T, warm, Xend,
                        Ulast, Utraj, flag, iters, nthreads);
}
void orc_simulate_batch_ex(const orc_problem *p, int nx, int nr, int nd, int nuprev,
                           const double *Fx, const double *Gx, const double *Gdx,
                           const double *X0, const double *R, const double *D, const double *U0,
                           const double *Rtraj, const double *Dtraj, const orc_sim_ext *ext,
                           int64_t N, int T, int warm, double *Xend, double *Ulast, double *Utraj,
                           int32_t *flag, int32_t *iters, int nthreads) {
  sim_ctx s = {p, nx, nr, nd, nuprev, Fx, Gx, Gdx, X0, R, D, U0, Rtraj, Dtraj, T, warm, Xend, Ulast, Utraj, flag,
               iters, ext};
  parallel_for(N, nthreads, sim_chunk, &s);
}

typedef struct { const orc_bst *t; const double *Theta; double *U; int32_t *region; } bst_ctx;
static void bst_chunk(void *vc, int64_t lo, int64_t hi) {
  bst_ctx *b = (bst_ctx *)vc;
  const orc_bst *t = b->t;
  const int nth = t->nth, nu = t->nu, ld = nth + 1;
  for (int64_t q = lo; q < hi; ++q) {
    const double *th = b->Theta + q * nth;
    int id = 0;
    for (;;) {
      int jmp = t->jump_list[id];
      if (jmp == 0) break;
      const double *hp = t->halfplanes + (size_t)t->hp_list[id] * ld;
      double acc = 0.0;
      for (int c = 0; c < nth; ++c) acc = fma(hp[c], th[c], acc);
      id = id + jmp + (acc <= hp[nth] ? 0 : 1);
    }
    int reg = t->hp_list[id];
    const double *fb = t->feedbacks + (size_t)reg * nu * ld;
    for (int i = 0; i < nu; ++i) {
      double acc = fb[i * ld + nth];
      for (int c = 0; c < nth; ++c) acc = fma(fb[i * ld + c], th[c], acc);
      if (t->clip) {
        if (acc < t->clip[2 * i]) acc = t->clip[2 * i];
        if (acc > t->clip[2 * i + 1]) acc = t->clip[2 * i + 1];
      }
      b->U[q * nu + i] = acc;
    }
    if (b->region) b->region[q] = reg;
  }
}
/* closed loop with the explicit law: Simulation(empc, scenario), simulation.jl:37-116 */
typedef struct {
  const orc_bst *t; int nx, nr, nd, nuprev; const double *Fx, *Gx, *Gdx, *X0, *R, *D, *U0, *Rtraj, *Dtraj;
  int T; double *Xend, *Ulast, *Utraj;
} bsim_ctx;
static void bsim_chunk(void *vc, int64_t lo, int64_t hi) {
  bsim_ctx *s = (bsim_ctx *)vc;
  const orc_bst *t = s->t;
  const int nth = t->nth, nu = t->nu, ld = nth + 1, nx = s->nx, nr = s->nr, nd = s->nd, nup = s->nuprev;
  double *th = (double *)calloc(nth, sizeof(double)), *xn = (double *)calloc(nx, sizeof(double));
  double *u = (double *)calloc(nu, sizeof(double));
  for (int64_t q = lo; q < hi; ++q) {
    for (int i = 0; i < nx; ++i) th[i] = s->X0[q * nx + i];
    for (int i = 0; i < nu; ++i) u[i] = (s->U0 && i < nup) ? s->U0[q * nup + i] : 0.0;
    for (int k = 0; k < s->T; ++k) {
      for (int i = 0; i < nr; ++i) th[nx + i] = s->Rtraj ? s->Rtraj[(size_t)k * nr + i] : (s->R ? s->R[q * nr + i] : 0.0);
      for (int i = 0; i < nd; ++i) th[nx + nr + i] = s->Dtraj ? s->Dtraj[(size_t)k * nd + i] : (s->D ? s->D[q * nd + i] : 0.0);
      for (int i = 0; i < nup; ++i) th[nx + nr + nd + i] = u[i];
      int id = 0;
      for (;;) {
        int jmp = t->jump_list[id];
        if (jmp == 0) break;
        const double *hp = t->halfplanes + (size_t)t->hp_list[id] * ld;
        double acc = 0.0;
        for (int c = 0; c < nth; ++c) acc = fma(hp[c], th[c], acc);
        id = id + jmp + (acc <= hp[nth] ? 0 : 1);
      }
      const double *fb = t->feedbacks + (size_t)t->hp_list[id] * nu * ld;
      for (int i = 0; i < nu; ++i) {
        double acc = fb[i * ld + nth];
        for (int c = 0; c < nth; ++c) acc = fma(fb[i * ld + c], th[c], acc);
        if (t->clip) {
          if (acc < t->clip[2 * i]) acc = t->clip[2 * i];
          if (acc > t->clip[2 * i + 1]) acc = t->clip[2 * i + 1];
        }
        u[i] = acc;
        if (s->Utraj) s->Utraj[((size_t)q * s->T + k) * nu + i] = acc;
      }
      for (int i = 0; i < nx; ++i) {
        double acc = 0.0;
        for (int c = 0; c < nx; ++c) acc = fma(s->Fx[i * nx + c], th[c], acc);
        for (int c = 0; c < nu; ++c) acc = fma(s->Gx[i * nu + c], u[c], acc);
        for (int c = 0; c < nd; ++c) acc = fma(s->Gdx[i * nd + c], th[nx + nr + c], acc);
        xn[i] = acc;
      }
      for (int i = 0; i < nx; ++i) th[i] = xn[i];
    }
    for (int i = 0; i < nx; ++i) s->Xend[q * nx + i] = th[i];
    for (int i = 0; i < nu; ++i) s->Ulast[q * nu + i] = u[i];
  }
  free(th); free(xn); free(u);
}
void orc_bst_simulate_batch(const orc_bst *t, int nx, int nr, int nd, int nuprev,
                            const double *Fx, const double *Gx, const double *Gdx,
                            const double *X0, const double *R, const double *D, const double *U0,
                            const double *Rtraj, const double *Dtraj, int64_t N, int T,
                            double *Xend, double *Ulast, double *Utraj, int nthreads) {
  bsim_ctx s = {t, nx, nr, nd, nuprev, Fx, Gx, Gdx, X0, R, D, U0, Rtraj, Dtraj, T, Xend, Ulast, Utraj};
  parallel_for(N, nthreads, bsim_chunk, &s);
}

void orc_bst_evaluate_batch(const orc_bst *t, const double *Theta, int64_t N, double *U,
                            int32_t *region, int nthreads) {
  bst_ctx b = {t, Theta, U, region};
  parallel_for(N, nthreads, bst_chunk, &b);
}
