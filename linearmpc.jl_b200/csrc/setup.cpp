// Host-side, once-per-controller setup: what DAQP.setup + qp2ldp do in the reference
// (/root/reference/src/setup.jl:7-29, /root/reference/src/codegen.jl:239-280), written as
// plain dense loops (n <= a few hundred, runs once).  Produces the constant block described
// in lmpc_common.h.  Compiled with -ffp-contract=off so the constants do not depend on the
// host compiler's contraction choices.
#include "lmpc_common.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
inline int align2(int x) { return (x + 1) & ~1; }
}

int lmpc_host_setup(const lmpc_problem *p, LdpHost *out, char *err, int errlen) {
  const int n = p->n, m = p->m, ms = p->ms, nth = p->nth, nu = p->nu, mg = m - ms;
  auto fail = [&](int code, const char *msg) { snprintf(err, errlen, "%s", msg); return code; };
  if (n <= 0 || m < 0 || ms < 0 || ms > m || ms > n || nth < 0 || nu <= 0 || nu > n)
    return fail(LMPC_E_INVALID, "lmpc_gpu_create: inconsistent dimensions");
  if (!p->H || !p->bu || !p->bl || !p->senses || (nth > 0 && (!p->f_theta || !p->W)) || (mg > 0 && !p->A))
    return fail(LMPC_E_INVALID, "lmpc_gpu_create: null problem array");
  if (p->has_binaries) return fail(LMPC_E_UNSUPPORTED, "binary constraints (daqp_bnb) are not supported");
  if (!p->is_symmetric) return fail(LMPC_E_UNSUPPORTED, "non-symmetric H (AVI) is not supported");
  if (p->n_break_points > 0) return fail(LMPC_E_UNSUPPORTED, "hierarchical priorities are not supported");
  for (int i = 0; i < m; ++i)
    if (p->senses[i] & S_BINARY) return fail(LMPC_E_UNSUPPORTED, "binary constraints are not supported");
  if (p->K && (p->nx <= 0 || p->nx > nth)) return fail(LMPC_E_INVALID, "K given but nx out of range");

  // H = R'R with R upper (H symmetrised as qp2ldp does, codegen.jl:242)
  std::vector<double> R((size_t)n * n, 0.0), Rinv((size_t)n * n, 0.0);
  auto Hs = [&](int i, int j) { return 0.5 * (p->H[(size_t)j * n + i] + p->H[(size_t)i * n + j]); };
  for (int j = 0; j < n; ++j) {          // column j of R (Cholesky-Crout, upper)
    for (int i = 0; i <= j; ++i) {
      double s = Hs(i, j);
      for (int c = 0; c < i; ++c) s -= R[(size_t)c * n + i] * R[(size_t)c * n + j];
      if (i == j) {
        if (!(s > 0.0)) return fail(LMPC_E_SETUP, "H is not positive definite (setup flag -5)");
        R[(size_t)i * n + i] = std::sqrt(s);
      } else {
        R[(size_t)i * n + j] = s / R[(size_t)i * n + i];
      }
    }
  }
  // Rinv upper: R * Rinv = I, column by column (back substitution)
  for (int j = 0; j < n; ++j) {
    for (int i = j; i >= 0; --i) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int c = i + 1; c <= j; ++c) s -= R[(size_t)i * n + c] * Rinv[(size_t)c * n + j];
      Rinv[(size_t)i * n + j] = s / R[(size_t)i * n + i];
    }
  }
  // Vth = R^-T f_theta (n x nth), v = R^-T f : forward substitution with R' (lower)
  std::vector<double> Vth((size_t)n * (nth > 0 ? nth : 1), 0.0), v(n, 0.0);
  for (int t = 0; t <= nth; ++t) {
    for (int i = 0; i < n; ++i) {
      double s = (t < nth) ? p->f_theta[(size_t)t * n + i] : (p->f ? p->f[i] : 0.0);
      for (int c = 0; c < i; ++c)
        s -= R[(size_t)c * n + i] * ((t < nth) ? Vth[(size_t)c * nth + t] : v[c]);
      s /= R[(size_t)i * n + i];
      if (t < nth) Vth[(size_t)i * nth + t] = s; else v[i] = s;
    }
  }

  LdpLayout lay;
  memset(&lay, 0, sizeof(lay));
  lay.n = n; lay.m = m; lay.ms = ms; lay.nth = nth; lay.nu = nu; lay.mp = (m + 3) & ~3;
  int nsoft = 0;
  for (int i = 0; i < m; ++i) if (p->senses[i] & S_SOFT) nsoft++;
  lay.nsoft = nsoft;
  lay.has_general = (mg > 0 || nsoft > 0) ? 1 : 0;
  int cap = n + (lay.has_general ? 1 : 0) + nsoft;
  if (cap > m) cap = m;
  if (cap < 1) cap = 1;
  lay.cap = cap;
  int o = 0;
  lay.oM = o;     o += align2(m * n);
  lay.oMT = o;    o += align2(n * lay.mp);
  lay.oG = o;     o += align2(m * m);
  lay.oDth = o;   o += align2(m * nth);
  lay.odu = o;    o += align2(m);
  lay.odl = o;    o += align2(m);
  lay.oRinv = o;  o += align2(nu * n);
  lay.oUth = o;   o += align2(nu * nth);
  lay.ouoff = o;  o += align2(nu);
  lay.oSense = o; o += align2((m + 1) / 2);
  lay.total = o;
  double *blk = (double *)calloc((size_t)lay.total + 2, sizeof(double));
  if (!blk) return fail(LMPC_E_INVALID, "out of host memory");
  double *M = blk + lay.oM, *MT = blk + lay.oMT, *G = blk + lay.oG, *Dth = blk + lay.oDth;
  double *du = blk + lay.odu, *dl = blk + lay.odl, *Ri = blk + lay.oRinv, *Uth = blk + lay.oUth;
  double *uoff = blk + lay.ouoff;
  int32_t *sense = (int32_t *)(blk + lay.oSense);

  // Mext = [I_ms; A] R^-1 (codegen.jl:243), then Dth, du, dl (codegen.jl:244-249)
  for (int i = 0; i < m; ++i) {
    double *row = M + (size_t)i * n;
    if (i < ms) {
      for (int c = 0; c < n; ++c) row[c] = Rinv[(size_t)i * n + c];
    } else {
      const int g = i - ms;
      for (int c = 0; c < n; ++c) {
        double s = 0.0;
        for (int r = 0; r <= c; ++r) s += p->A[(size_t)r * mg + g] * Rinv[(size_t)r * n + c];
        row[c] = s;
      }
    }
    double dd = 0.0;
    for (int c = 0; c < n; ++c) dd += row[c] * v[c];
    for (int t = 0; t < nth; ++t) {
      double s = p->W[(size_t)t * m + i];
      for (int c = 0; c < n; ++c) s += row[c] * Vth[(size_t)c * nth + t];
      Dth[(size_t)i * nth + t] = s;
    }
    du[i] = p->bu[i] + dd;
    dl[i] = p->bl[i] + dd;
    sense[i] = p->senses[i];
    // row normalisation (codegen.jl:252-264)
    double nrm = 0.0;
    for (int c = 0; c < n; ++c) nrm += row[c] * row[c];
    nrm = std::sqrt(nrm);
    if (nrm > 0.0) {
      for (int c = 0; c < n; ++c) row[c] /= nrm;
      for (int t = 0; t < nth; ++t) Dth[(size_t)i * nth + t] /= nrm;
      du[i] /= nrm;
      dl[i] /= nrm;
    } else {
      sense[i] |= S_IMMUTABLE;  // a row that does not depend on U can never be activated
      sense[i] &= ~S_ACTIVE;
    }
  }
  int neq = 0;
  for (int i = 0; i < m; ++i) if (sense[i] & S_ACTIVE) neq++;
  lay.neq = neq;
  lay.plain = (m == ms) ? 1 : 0;
  for (int i = 0; i < m; ++i) if (sense[i] != 0) lay.plain = 0;
  for (int i = 0; i < m; ++i)
    for (int c = 0; c < n; ++c) MT[(size_t)c * lay.mp + i] = M[(size_t)i * n + c];
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      double s = 0.0;
      for (int c = 0; c < n; ++c) s += M[(size_t)i * n + c] * M[(size_t)j * n + c];
      G[(size_t)i * m + j] = s;
    }
  // output map (codegen.jl:269-273, :141): control = (R^-1 u)[1:nu] + uoff + Uth theta
  for (int i = 0; i < nu; ++i) {
    for (int c = 0; c < n; ++c) Ri[(size_t)i * n + c] = Rinv[(size_t)i * n + c];
    for (int t = 0; t < nth; ++t) {
      double s = 0.0;
      for (int c = i; c < n; ++c) s += Rinv[(size_t)i * n + c] * Vth[(size_t)c * nth + t];
      s = -s;
      if (p->K && t < p->nx) s -= p->K[(size_t)t * nu + i];
      Uth[(size_t)i * nth + t] = s;
    }
    double s = 0.0;
    for (int c = i; c < n; ++c) s += Rinv[(size_t)i * n + c] * v[c];
    uoff[i] = -s;
  }
  // DAQP.setup's remaining flags (src/setup.jl:14-23), both independent of theta because bu and bl shift
  // together: -1 a row whose bounds cross (d_u - d_l < -primal_tol on the normalised row), -6 equality rows
  // that are linearly dependent or outnumber the variables (found by factorising their Gram block the way
  // the cold start activates them: a pivot below zero_tol).
  for (int i = 0; i < m; ++i) {
    if (sense[i] & S_IMMUTABLE && !(sense[i] & S_ACTIVE)) continue;
    if (du[i] < 1e29 && dl[i] > -1e29 && du[i] - dl[i] < -1e-6) {
      free(blk);
      snprintf(err, errlen, "row %d: lower bound above upper bound, the problem is infeasible (setup flag -1)", i);
      return LMPC_E_SETUP_INFEASIBLE;
    }
  }
  if (neq > 0) {
    std::vector<int> eq;
    for (int i = 0; i < m; ++i) if (sense[i] & S_ACTIVE) eq.push_back(i);
    const int ne = (int)eq.size();
    std::vector<double> Lq((size_t)ne * ne, 0.0), Dq(ne, 0.0);
    bool over = ne > n;
    for (int a = 0; a < ne && !over; ++a) {          // LDL' of the equality rows' Gram block, row by row
      double d = G[(size_t)eq[a] * m + eq[a]];
      for (int b = 0; b < a; ++b) {
        double y = G[(size_t)eq[a] * m + eq[b]];
        for (int c = 0; c < b; ++c) y -= Lq[(size_t)a * ne + c] * Dq[c] * Lq[(size_t)b * ne + c];
        Lq[(size_t)a * ne + b] = y / Dq[b];
        d -= y * Lq[(size_t)a * ne + b];
      }
      if (d < 1e-11) over = true;
      Dq[a] = d;
    }
    if (over) {
      free(blk);
      snprintf(err, errlen, "equality constraints are overdetermined (setup flag -6)");
      return LMPC_E_SETUP_OVERDETERMINED;
    }
  }
  out->lay = lay;
  out->blk = blk;
  return 0;
}
