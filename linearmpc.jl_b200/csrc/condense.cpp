// Host-side mpQP condensing behind the C ABI: from a discrete-time model, weights,
// bounds and constraints to the dense mpQP {H, f, f_theta, A, bu, bl, W, senses} that
// lmpc_gpu_create consumes, so that non-Julia callers can build a controller.
//
// What it follows in the reference (subset: no offsets, binaries or generalized parameters; previews
// of r / d without reference condensation):
//   extended system / weights    /root/reference/src/mpc2mpqp.jl:649-731
//   state predictor              /root/reference/src/mpc2mpqp.jl:20-46
//   disturbance predictor        /root/reference/src/mpc2mpqp.jl:48-68
//   preview cost / constraints   /root/reference/src/mpc2mpqp.jl:70-109, 535-610
//   dense objective              /root/reference/src/mpc2mpqp.jl:407-533
//   control bounds, general rows /root/reference/src/mpc2mpqp.jl:206-245, 249-354
//   move blocking                /root/reference/src/mpc2mpqp.jl:830-857, src/setup.jl:196-251
//   sort / preprocessing         /root/reference/src/mpc2mpqp.jl:733-866
//   senses, clamping             /root/reference/src/mpc2mpqp.jl:868-899
// Runs once per controller; plain dense loops.  Checked against the numpy restatement in
// oracle/mpqp.py by tests/test_condense_cpu.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>
#include "lmpc_common.h"

namespace {

struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;   // row-major
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double &operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};

Mat from_colmajor(const double *p, int r, int c) {
  Mat m(r, c);
  if (p) for (int j = 0; j < c; ++j) for (int i = 0; i < r; ++i) m(i, j) = p[(size_t)j * r + i];
  return m;
}
Mat eye(int n) { Mat m(n, n); for (int i = 0; i < n; ++i) m(i, i) = 1.0; return m; }
Mat mul(const Mat &A, const Mat &B) {
  Mat C(A.r, B.c);
  for (int i = 0; i < A.r; ++i)
    for (int k = 0; k < A.c; ++k) {
      const double aik = A(i, k);
      if (aik == 0.0) continue;
      for (int j = 0; j < B.c; ++j) C(i, j) += aik * B(k, j);
    }
  return C;
}
Mat tr(const Mat &A) { Mat T(A.c, A.r); for (int i = 0; i < A.r; ++i) for (int j = 0; j < A.c; ++j) T(j, i) = A(i, j); return T; }
bool any_nonzero(const Mat &A) { for (double v : A.a) if (v != 0.0) return true; return false; }
void set_block(Mat &D, int r0, int c0, const Mat &S, double scale = 1.0) {
  for (int i = 0; i < S.r; ++i) for (int j = 0; j < S.c; ++j) D(r0 + i, c0 + j) = scale * S(i, j);
}
void add_block(Mat &D, int r0, int c0, const Mat &S, double scale = 1.0) {
  for (int i = 0; i < S.r; ++i) for (int j = 0; j < S.c; ++j) D(r0 + i, c0 + j) += scale * S(i, j);
}
Mat block_diag(const Mat &A, const Mat &B) {
  Mat D(A.r + B.r, A.c + B.c);
  set_block(D, 0, 0, A); set_block(D, A.r, A.c, B);
  return D;
}
Mat rows_of(const Mat &A, const std::vector<int> &idx) {
  Mat S((int)idx.size(), A.c);
  for (size_t i = 0; i < idx.size(); ++i) for (int j = 0; j < A.c; ++j) S((int)i, j) = A(idx[i], j);
  return S;
}
Mat sub(const Mat &A, const std::vector<int> &ri, const std::vector<int> &ci) {
  Mat S((int)ri.size(), (int)ci.size());
  for (size_t i = 0; i < ri.size(); ++i) for (size_t j = 0; j < ci.size(); ++j) S((int)i, (int)j) = A(ri[i], ci[j]);
  return S;
}

// X with A X = B (LU, partial pivoting); false when A is singular to working precision
bool solve_lu(Mat A, Mat B, Mat &X) {
  const int n = A.r;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    for (int i = k + 1; i < n; ++i) if (std::fabs(A(i, k)) > std::fabs(A(piv, k))) piv = i;
    if (A(piv, k) == 0.0) return false;
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(A(k, j), A(piv, j));
      for (int j = 0; j < B.c; ++j) std::swap(B(k, j), B(piv, j));
    }
    for (int i = k + 1; i < n; ++i) {
      const double l = A(i, k) / A(k, k);
      if (l == 0.0) continue;
      for (int j = k; j < n; ++j) A(i, j) -= l * A(k, j);
      for (int j = 0; j < B.c; ++j) B(i, j) -= l * B(k, j);
    }
  }
  X = Mat(n, B.c);
  for (int i = n - 1; i >= 0; --i)
    for (int j = 0; j < B.c; ++j) {
      double v = B(i, j);
      for (int k = i + 1; k < n; ++k) v -= A(i, k) * X(k, j);
      X(i, j) = v / A(i, i);
    }
  return true;
}
// least squares X = argmin |A X - B| for a tall full-column-rank A (Householder QR), what Julia's A \ B does
bool lstsq_qr(Mat A, Mat B, Mat &X) {
  const int m = A.r, n = A.c;
  if (m < n) return false;
  for (int k = 0; k < n; ++k) {
    double nrm = 0.0;
    for (int i = k; i < m; ++i) nrm += A(i, k) * A(i, k);
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) return false;
    const double alpha = A(k, k) > 0 ? -nrm : nrm;
    std::vector<double> v(m - k);
    for (int i = k; i < m; ++i) v[i - k] = A(i, k);
    v[0] -= alpha;
    double vn = 0.0;
    for (double t : v) vn += t * t;
    if (vn == 0.0) continue;
    auto reflect = [&](Mat &M) {
      for (int j = 0; j < M.c; ++j) {
        double d = 0.0;
        for (int i = k; i < m; ++i) d += v[i - k] * M(i, j);
        d = 2.0 * d / vn;
        for (int i = k; i < m; ++i) M(i, j) -= d * v[i - k];
      }
    };
    reflect(A); reflect(B);
  }
  X = Mat(n, B.c);
  for (int i = n - 1; i >= 0; --i)
    for (int j = 0; j < B.c; ++j) {
      double v = B(i, j);
      for (int k = i + 1; k < n; ++k) v -= A(i, k) * X(k, j);
      X(i, j) = v / A(i, i);
    }
  return true;
}

}  // namespace

struct lmpc_mpqp {
  int n = 0, m = 0, ms = 0, nth = 0, nx = 0, nr = 0, nd = 0, nuprev = 0, nu = 0;
  Mat H, f_theta, A, W;
  Mat traj2setpoint;   // ny x ny*Np with reference_condensation, else empty
  std::vector<double> f, bu, bl;
  std::vector<int32_t> senses;
};

static thread_local char c_err[256] = "";
extern "C" const char *lmpc_gpu_condense_error(void) { return c_err; }
static int cfail(int code, const char *msg) { snprintf(c_err, sizeof(c_err), "%s", msg); return code; }

extern "C" int lmpc_gpu_condense(const lmpc_model *md, lmpc_mpqp **out) {
  if (!md || !out) return cfail(LMPC_E_INVALID, "lmpc_gpu_condense: null argument");
  *out = nullptr;
  const int nx = md->nx, nu = md->nu, Np = md->Np, Nc = md->Nc > 0 ? md->Nc : md->Np;
  if (nx <= 0 || nu <= 0 || Np <= 0 || Nc > Np || !md->F || !md->G)
    return cfail(LMPC_E_INVALID, "lmpc_gpu_condense: bad dimensions");
  const int ny = md->C ? md->ny : nx;
  const int ndm = md->nd;
  if (ndm > 0 && !md->Gd && !md->Dd) return cfail(LMPC_E_INVALID, "nd > 0 needs Gd or Dd");
  const Mat F0 = from_colmajor(md->F, nx, nx), G0 = from_colmajor(md->G, nx, nu);
  const Mat C0 = md->C ? from_colmajor(md->C, ny, nx) : eye(nx);
  const Mat Gd = from_colmajor(md->Gd, nx, ndm), Dd = from_colmajor(md->Dd, ny, ndm);
  Mat Q = md->Q ? from_colmajor(md->Q, ny, ny) : eye(ny);               // MPCWeights(nu,nx,nr) defaults
  Mat R = md->R ? from_colmajor(md->R, nu, nu) : eye(nu);
  Mat Rr = md->Rr ? from_colmajor(md->Rr, nu, nu) : Mat(nu, nu);
  Mat Qf = md->Qf ? from_colmajor(md->Qf, ny, ny) : Mat(ny, ny);
  // prestabilising feedback u = v - K x (src/setup.jl:186-199): the QP is over v
  const Mat K = from_colmajor(md->K, nu, nx);
  const bool hasK = any_nonzero(K);
  const Mat Rorig = R;
  const bool kr = hasK && any_nonzero(Rorig);   // u'Ru then needs the output K x (mpc2mpqp.jl:679-681, 718-724)
  // parameter layout (mpc2mpqp.jl:147-164)
  // with a preview the Np future values of r / d are parameters, not extended states
  const bool rp = md->reference_preview && md->reference_tracking;
  const bool dp = md->disturbance_preview && ndm > 0;
  const int nr = (md->reference_tracking && !rp) ? ny : 0;     // reference part of the extended state
  const int nd = dp ? 0 : ndm;                                 // disturbance part of the extended state
  // reference_condensation (types.jl:54,65): the previewed trajectory is collapsed to one setpoint, theta carries
  // ny values again and traj2setpoint maps a trajectory to them (mpc2mpqp.jl:550-568, utils.jl:140-147)
  const bool rcond = rp && md->reference_condensation;
  const int nrp = rp ? (rcond ? ny : ny * Np) : nr, ndp = dp ? ndm * Np : nd; // widths of the r / d blocks of theta
  bool aup = false;
  for (int ci = 0; ci < md->n_constraints; ++ci) {
    const lmpc_constraint &c = md->constraints[ci];
    if (c.Aup) for (int i = 0; i < c.rows * nu; ++i) if (c.Aup[i] != 0.0) aup = true;
  }
  const int nup = (any_nonzero(Rr) || aup) ? nu : 0;
  // generalized parameters: the last block of theta, np values, or Np blocks of np with parameter_preview
  const int npb = md->np > 0 ? md->np : 0, pp = (md->parameter_preview && npb > 0) ? 1 : 0;
  const int npar = pp ? npb * Np : npb;
  // a constant dynamics offset is one more extended state that stays at 1 (mpc2mpqp.jl:684-689); its column of
  // f_theta / W lives in a scratch column behind theta and is collapsed into f / bu / bl before move blocking
  bool has_off = false;
  if (md->f_offset) for (int i = 0; i < nx; ++i) if (md->f_offset[i] != 0.0) has_off = true;
  const int noff = has_off ? 1 : 0;
  if (noff && npar > 0)   // the reference collapses the LAST column, which is a p column then (mpc2mpqp.jl:395-399)
    return cfail(LMPC_E_UNSUPPORTED, "f_offset together with generalized parameters is not supported");
  const int nth_out = nx + nrp + ndp + nup + npar;
  const int nxe = nx + nr + nd + nup + noff, opar = nth_out - npar;
  int nth = nth_out + noff;
  // column of theta that holds extended-state entry j (insert_preview_parameter_blocks, mpc2mpqp.jl:70-92)
  auto th_col = [&](int j) { return j < nx + nr ? j : (j < nx + nr + nd ? j - nr + nrp : j - nr - nd + nrp + ndp); };

  // ---- extended system x~ = [x; r; d; u^-] (mpc2mpqp.jl:649-690, K = 0) -------------------------
  Mat F(nxe, nxe), G(nxe, nu);
  set_block(F, 0, 0, F0); set_block(G, 0, 0, G0);
  if (hasK) add_block(F, 0, 0, mul(G0, K), -1.0);
  int off = nx;
  Mat C = C0;   // grows column-wise
  auto hcat = [](const Mat &A, const Mat &B) { Mat M(A.r, A.c + B.c); set_block(M, 0, 0, A); set_block(M, 0, A.c, B); return M; };
  if (nr > 0) {
    for (int i = 0; i < ny; ++i) F(off + i, off + i) = 1.0;
    Mat mI = eye(ny); for (double &v : mI.a) v = -v;
    C = hcat(C, mI);
    off += ny;
  }
  if (nd > 0) {
    for (int i = 0; i < nd; ++i) F(off + i, off + i) = 1.0;
    set_block(F, 0, off, Gd);
    C = hcat(C, Dd);
    off += nd;
  }
  if (nup > 0) {
    for (int i = 0; i < nu; ++i) G(off + i, i) = 1.0;      // F block stays 0 ...
    set_block(F, off, 0, K, -1.0);                         // ... except u^-+ = v - K x
    Mat C2(C.r + nu, C.c + nu);
    set_block(C2, 0, 0, C);
    set_block(C2, C.r, 0, K);
    for (int i = 0; i < nu; ++i) C2(C.r + i, C.c + i) = 1.0;
    C = C2;
  }
  if (kr) {
    Mat C2(C.r + nu, C.c);
    set_block(C2, 0, 0, C);
    set_block(C2, C.r, 0, K);
    C = C2;
  }
  if (noff) {
    F(nxe - 1, nxe - 1) = 1.0;
    for (int i = 0; i < nx; ++i) F(i, nxe - 1) = md->f_offset[i];
    C = hcat(C, Mat(C.r, 1));
  }
  // ---- extended weights (mpc2mpqp.jl:692-731, K = 0) ----------------------------------------------
  if (!any_nonzero(Qf)) Qf = Q;
  const Mat Q0 = Q, Qf0 = Qf;   // output weights before the extension (Q_full, Qf_full)
  Mat S(nxe, nu);
  if (nup > 0) {
    Q = block_diag(Q, Rr);
    Qf = block_diag(Qf, Mat(nu, nu));
    for (int i = 0; i < nu; ++i) for (int j = 0; j < nu; ++j) { S(nx + nr + nd + i, j) = -Rr(i, j); R(i, j) += Rr(i, j); }
    if (hasK) add_block(S, 0, 0, mul(tr(K), Rr), -1.0);
  }
  if (kr) {
    Q = block_diag(Q, Rorig);
    Qf = block_diag(Qf, Mat(nu, nu));
    add_block(S, 0, 0, mul(tr(K), Rorig), -1.0);
  }
  // ---- predictor X = Phi x0 + Gamma U (mpc2mpqp.jl:20-46) -----------------------------------------
  Mat Phi((Np + 1) * nxe, nxe), Gam((Np + 1) * nxe, Nc * nu);
  set_block(Phi, 0, 0, eye(nxe));
  {
    Mat Pk = eye(nxe), Gk(nxe, Nc * nu);
    for (int k = 0; k < Np; ++k) {
      Pk = mul(F, Pk);
      Gk = mul(F, Gk);
      const int j = std::min(k, Nc - 1);
      add_block(Gk, 0, j * nu, G);
      set_block(Phi, (k + 1) * nxe, 0, Pk);
      set_block(Gam, (k + 1) * nxe, 0, Gk);
    }
  }
  // disturbance predictor X += Psi [d_0; ...; d_{Np-1}] (mpc2mpqp.jl:48-68), block row by block row
  std::vector<Mat> Psi;
  if (dp) {
    Psi.assign(Np + 1, Mat(nxe, ndm * Np));
    for (int k = 1; k <= Np; ++k) {
      Psi[k] = mul(F, Psi[k - 1]);
      add_block(Psi[k], 0, (k - 1) * ndm, Gd);   // E = [Gd; 0]
    }
  }
  // ---- objective (mpc2mpqp.jl:407-533) -------------------------------------------------------------
  std::vector<int> posQ, posF, allc(nxe);
  for (int i = 0; i < nxe; ++i) allc[i] = i;
  for (int i = 0; i < Q.r; ++i) if (Q(i, i) > 0) posQ.push_back(i);
  for (int i = 0; i < Qf.r; ++i) if (Qf(i, i) > 0) posF.push_back(i);
  const Mat Cp = rows_of(C, posQ), Qp = sub(Q, posQ, posQ), Cf = rows_of(C, posF), Qfp = sub(Qf, posF, posF);
  const Mat CQC = mul(tr(Cp), mul(Qp, Cp)), CQCf = mul(tr(Cf), mul(Qfp, Cf));
  const int n = Nc * nu;
  Mat H(n, n);
  for (int b = 0; b < Nc; ++b) add_block(H, b * nu, b * nu, R);
  add_block(H, n - nu, n - nu, R, (double)(Np - Nc));
  std::vector<double> fvec(n, 0.0);
  // a regulator (reference_tracking off) penalises the distance to the operating point (xo, uo)
  // (mpc2mpqp.jl:433-443, 457-459)
  const bool regulator = !md->reference_tracking;
  bool has_xo = false, has_uo = false;
  if (regulator && md->xo) for (int i = 0; i < nx; ++i) if (md->xo[i] != 0.0) has_xo = true;
  if (regulator && md->uo) for (int i = 0; i < nu; ++i) if (md->uo[i] != 0.0) has_uo = true;
  if (has_uo) {
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) fvec[i] -= H(i, j) * md->uo[j % nu];
    if (hasK && any_nonzero(R)) {
      // f -= (GKR + GKR') Uo, GKR = Gamma' KRtot, KRtot: -K'R on stage k's plant-state rows, column block min(k, Nc-1)
      const Mat KR = mul(tr(K), R);
      Mat GKR(n, n);
      for (int k = 0; k < Np; ++k) {
        const int cb = std::min(k, Nc - 1);
        for (int r = 0; r < n; ++r) for (int i = 0; i < nx; ++i) {
          const double g = Gam(k * nxe + i, r);
          if (g != 0.0) for (int j = 0; j < nu; ++j) GKR(r, cb * nu + j) -= g * KR(i, j);
        }
      }
      for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) fvec[i] -= (GKR(i, j) + GKR(j, i)) * md->uo[j % nu];
    }
  }
  // Gamma' CQCtot Gamma and Gamma' CQCtot Phi, block row by block row (CQCtot is block diagonal)
  Mat fth(n, nth);
  Mat Frp(n, rp ? ny * Np : 0);     // Fr of ref_preview_cost: column block k-1 pairs y_k with preview column k-1
  for (int k = 0; k <= Np; ++k) {
    const Mat &W = (k == Np) ? CQCf : CQC;
    Mat Gk(nxe, n), Pk(nxe, nxe);
    for (int i = 0; i < nxe; ++i) {
      for (int j = 0; j < n; ++j) Gk(i, j) = Gam(k * nxe + i, j);
      for (int j = 0; j < nxe; ++j) Pk(i, j) = Phi(k * nxe + i, j);
    }
    const Mat GtW = mul(tr(Gk), W);
    add_block(H, 0, 0, mul(GtW, Gk));
    if (has_xo) for (int i = 0; i < n; ++i) for (int j = 0; j < nx; ++j) fvec[i] -= GtW(i, j) * md->xo[j];
    const Mat fk = mul(GtW, Pk);
    for (int i = 0; i < n; ++i) for (int j = 0; j < nxe; ++j) fth(i, th_col(j)) += fk(i, j);
    if (k >= 1 && (rp || dp)) {
      // ref_preview_cost / disturbance_preview_cost (mpc2mpqp.jl:535-610): y_k = C x_k + Dd d_{k-1}
      // is weighted against preview column k-1 with Q_full (Qf_full at k = Np)
      Mat Cfull(ny, nxe);
      for (int i = 0; i < ny; ++i) for (int j = 0; j < nxe; ++j) Cfull(i, j) = C(i, j);
      const Mat GCQ = mul(tr(Gk), mul(tr(Cfull), (k == Np) ? Qf0 : Q0));   // n x ny
      if (rp) add_block(Frp, 0, (k - 1) * ny, GCQ, -1.0);
      if (dp) {
        Mat Yd = mul(Cfull, Psi[k]);
        add_block(Yd, 0, (k - 1) * ndm, Dd);
        add_block(fth, 0, nx + nrp, mul(GCQ, Yd));
      }
    }
  }
  if (any_nonzero(S)) {
    // Stot: S on blocks k < Nc (column block k), and on Nc <= k < Np into the last column block
    Mat GS(n, n), StP(n, nxe);
    for (int k = 0; k < Np; ++k) {
      const int cb = std::min(k, Nc - 1);
      Mat Gk(nxe, n), Pk(nxe, nxe);
      for (int i = 0; i < nxe; ++i) {
        for (int j = 0; j < n; ++j) Gk(i, j) = Gam(k * nxe + i, j);
        for (int j = 0; j < nxe; ++j) Pk(i, j) = Phi(k * nxe + i, j);
      }
      add_block(GS, 0, cb * nu, mul(tr(Gk), S));
      add_block(StP, cb * nu, 0, mul(tr(S), Pk));
    }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) H(i, j) += GS(i, j) + GS(j, i);
    for (int i = 0; i < n; ++i) for (int j = 0; j < nxe; ++j) fth(i, th_col(j)) += StP(i, j);
  }
  Mat T2S;
  if (rp && !rcond) add_block(fth, 0, nx, Frp);
  if (rcond) {
    // traj2setpoint = (W H^-1 Fr Is) \ (W H^-1 Fr), W = I with 1e6 on the first move (mpc2mpqp.jl:552-563);
    // the cost then sees the setpoint on every stage: Fr <- Fr Is
    Mat X;
    if (!solve_lu(H, Frp, X)) return cfail(LMPC_E_SETUP, "reference condensation: singular Hessian");
    for (int i = 0; i < nu && i < n; ++i) for (int j = 0; j < X.c; ++j) X(i, j) *= 1e6;
    Mat XI(n, ny), FI(n, ny);
    for (int i = 0; i < n; ++i) for (int k = 0; k < Np; ++k) for (int j = 0; j < ny; ++j) {
      XI(i, j) += X(i, k * ny + j);
      FI(i, j) += Frp(i, k * ny + j);
    }
    if (!lstsq_qr(XI, X, T2S)) return cfail(LMPC_E_SETUP, "reference condensation: rank-deficient least squares");
    add_block(fth, 0, nx, FI);
  }
  { Mat Hs(n, n); for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Hs(i, j) = 0.5 * (H(i, j) + H(j, i)); H = Hs; }
  // generalized-parameter cost (mpc2mpqp.jl:478-508): f += Umap' rep(eu) + Gx' rep(ex), and the p columns of
  // f_theta = Umap' rep(Eu) + Gx' rep(Ex); Umap picks the Nc moves out of the first Nc of the Np stages, Gx is
  // the plant-state part of the predictor rows x_1 .. x_Np
  // (stage_parameter_matrix, mpc2mpqp.jl:148: with parameter_preview stage k pairs u_k and x_{k+1} with p block k)
  if (md->ex || md->eu || npar > 0) {
    const Mat Exm = from_colmajor(md->Ex, nx, npb), Eum = from_colmajor(md->Eu, nu, npb);
    for (int b = 0; b < Nc; ++b)
      for (int i = 0; i < nu; ++i) {
        if (md->eu) fvec[b * nu + i] += md->eu[i];
        for (int j = 0; j < npb; ++j) fth(b * nu + i, opar + (pp ? b * npb : 0) + j) += Eum(i, j);
      }
    for (int k = 1; k <= Np; ++k)
      for (int i = 0; i < nx; ++i) {
        const double exi = md->ex ? md->ex[i] : 0.0;
        for (int c2 = 0; c2 < n; ++c2) {
          const double gx = Gam(k * nxe + i, c2);
          if (gx == 0.0) continue;
          fvec[c2] += gx * exi;
          for (int j = 0; j < npb; ++j) fth(c2, opar + (pp ? (k - 1) * npb : 0) + j) += gx * Exm(i, j);
        }
      }
  }

  // ---- constraints ------------------------------------------------------------------------------------
  std::vector<double> bu, bl;
  std::vector<int> soft, prio;
  Mat W(0, nth);
  int ms = 0;
  std::vector<std::vector<double>> Arows, Wrows;
  std::vector<double> gbu, gbl;
  std::vector<int> gsoft, gprio;
  if ((md->umax || md->umin) && hasK) {
    // lb <= v_k - K x_k <= ub: general rows A = I - K~ Gamma, W = K~ Phi (+ K~ Psi) (mpc2mpqp.jl:206-245)
    Mat row(nu, nxe);
    set_block(row, 0, 0, K);
    for (int k = 0; k < Nc; ++k) {
      Mat Gk(nxe, n), Pk(nxe, nxe);
      for (int i = 0; i < nxe; ++i) {
        for (int j = 0; j < n; ++j) Gk(i, j) = Gam(k * nxe + i, j);
        for (int j = 0; j < nxe; ++j) Pk(i, j) = Phi(k * nxe + i, j);
      }
      const Mat KG = mul(row, Gk), KP = mul(row, Pk);
      const Mat KPsi = dp ? mul(row, Psi[k]) : Mat(nu, 0);
      for (int i = 0; i < nu; ++i) {
        std::vector<double> ar(n), wr(nth, 0.0);
        for (int j = 0; j < n; ++j) ar[j] = -KG(i, j);
        ar[k * nu + i] += 1.0;
        for (int j = 0; j < nxe; ++j) wr[th_col(j)] = KP(i, j);
        for (int j = 0; j < KPsi.c; ++j) wr[nx + nrp + j] = KPsi(i, j);
        Arows.push_back(ar); Wrows.push_back(wr);
        gbu.push_back(md->umax ? md->umax[i] : 1e30); gbl.push_back(md->umin ? md->umin[i] : -1e30);
        gsoft.push_back(0); gprio.push_back(0);
      }
    }
  } else if (md->umax || md->umin) {
    ms = n;
    for (int b = 0; b < Nc; ++b)
      for (int i = 0; i < nu; ++i) {
        bu.push_back(md->umax ? md->umax[i] : 1e30);
        bl.push_back(md->umin ? md->umin[i] : -1e30);
        soft.push_back(0); prio.push_back(0);
      }
    W = Mat(ms, nth);
  }
  // general rows (mpc2mpqp.jl:249-354): lb <= [Ax Ar Ad Aup] x~_{k-1} + Au u_{k-1} <= ub, k in ks
  for (int ci = 0; ci < md->n_constraints; ++ci) {
    const lmpc_constraint &c = md->constraints[ci];
    const int mi = c.rows;
    if (mi <= 0) continue;
    const Mat Au = from_colmajor(c.Au, mi, nu);
    Mat Ax = from_colmajor(c.Ax, mi, nx);
    if (hasK) add_block(Ax, 0, 0, mul(Au, K), -1.0);
    const Mat Ar = from_colmajor(c.Ar, mi, nr), Ad = from_colmajor(c.Ad, mi, nd), Aupm = from_colmajor(c.Aup, mi, nup);
    if (rp && c.Ar) for (int i = 0; i < mi * ny; ++i) if (c.Ar[i] != 0.0)
      return cfail(LMPC_E_UNSUPPORTED, "constraints on the reference are not supported with reference_preview");
    const Mat Adp = dp ? from_colmajor(c.Ad, mi, ndm) : Mat(mi, 0);
    Mat row(mi, nxe);
    set_block(row, 0, 0, Ax); set_block(row, 0, nx, Ar); set_block(row, 0, nx + nr, Ad); set_block(row, 0, nx + nr + nd, Aupm);
    const int kmax = any_nonzero(Au) ? Np : Np + 1;
    for (int q = 0; q < c.nks; ++q) {
      const int k = c.ks[q];
      if (k < 1 || k > kmax) continue;
      // x~_{k-1} = Phi_{k-1} x0 + Gam_{k-1} U ; u_{k-1} only for k <= Nc (eyeU has zero rows beyond)
      Mat Gk(nxe, n), Pk(nxe, nxe);
      for (int i = 0; i < nxe; ++i) {
        for (int j = 0; j < n; ++j) Gk(i, j) = Gam((k - 1) * nxe + i, j);
        for (int j = 0; j < nxe; ++j) Pk(i, j) = Phi((k - 1) * nxe + i, j);
      }
      Mat Ak = mul(row, Gk);
      if (k <= Nc) add_block(Ak, 0, (k - 1) * nu, Au);
      const Mat Wk = mul(row, Pk);
      Mat Wdk(mi, dp ? ndm * Np : 0);
      if (dp) {   // -Ax~ Psi_{k-1} and the direct term -Ad d_{k-2} (disturbance_preview_direct, mpc2mpqp.jl:94-109)
        Wdk = mul(row, Psi[k - 1]);
        if (c.Ad && k >= 2) add_block(Wdk, 0, (k - 2) * ndm, Adp);
      }
      for (int i = 0; i < mi; ++i) {
        std::vector<double> ar(n), wr(nth, 0.0);
        for (int j = 0; j < n; ++j) ar[j] = Ak(i, j);
        for (int j = 0; j < nxe; ++j) wr[th_col(j)] = -Wk(i, j);
        for (int j = 0; j < Wdk.c; ++j) wr[nx + nrp + j] = -Wdk(i, j);
        // parameter_preview_direct (mpc2mpqp.jl:125-143): constraint stage k sees p block min(k, Np)
        if (c.Ap) for (int j = 0; j < npb; ++j) wr[opar + (pp ? (std::min(k, Np) - 1) * npb : 0) + j] = -c.Ap[(size_t)j * mi + i];
        Arows.push_back(ar); Wrows.push_back(wr);
        gbu.push_back(c.ub ? c.ub[i] : 1e30); gbl.push_back(c.lb ? c.lb[i] : -1e30);
        gsoft.push_back(c.soft ? 1 : 0); gprio.push_back(c.prio);
      }
    }
  }
  // ---- model offsets ------------------------------------------------------------------------------------
  if (noff) {   // collapse the constant state (value 1): mpc2mpqp.jl:395-399, 518-522
    const int sc = nth - 1;
    for (int i = 0; i < n; ++i) fvec[i] += fth(i, sc);
    for (size_t i = 0; i < Wrows.size(); ++i) { gbu[i] += Wrows[i][sc]; gbl[i] += Wrows[i][sc]; Wrows[i].pop_back(); }
    Mat f2(n, nth - 1);
    for (int i = 0; i < n; ++i) for (int j = 0; j < nth - 1; ++j) f2(i, j) = fth(i, j);
    fth = f2;
    W = Mat(ms, nth - 1);          // the simple bounds never depend on theta
    nth = nth_out;
  }
  if (nrp > 0 && md->h_offset) {   // the outputs track r - h_offset (mpc2mpqp.jl:523-530)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < nrp; ++j) fvec[i] -= fth(i, nx + j) * md->h_offset[j % ny];
  }
  // ---- move blocking (apply_move_block, mpc2mpqp.jl:830-857) -------------------------------------------
  // U_old = T z: the blocks of all inputs are interleaved pass by pass; the last block of every
  // input is clipped to one step (beyond Nc the control is held anyway); of the simple bounds only
  // the first step of each block survives.
  int nq = n;
  if (md->move_blocks && md->n_move_blocks) {
    std::vector<std::vector<int>> mb(nu);
    size_t maxlen = 0;
    const int32_t *pb = md->move_blocks;
    for (int iu = 0; iu < nu; ++iu) {
      for (int b = 0; b < md->n_move_blocks[iu]; ++b) mb[iu].push_back(*pb++);
      if (mb[iu].empty()) return cfail(LMPC_E_INVALID, "move blocking: an input has no blocks");
      maxlen = std::max(maxlen, mb[iu].size());
    }
    int nnew = 0;
    for (auto &v : mb) nnew += (int)v.size();
    Mat T(n, nnew);
    const int nbound = (md->umax || md->umin) ? n : 0;   // control-bound rows: simple, or the top of A with a feedback
    std::vector<int> keep, counter(nu);
    for (int iu = 0; iu < nu; ++iu) counter[iu] = iu;
    int new_id = 0;
    for (size_t pass = 0; pass < maxlen; ++pass)
      for (int iu = 0; iu < nu; ++iu) {
        if (mb[iu].size() <= pass) continue;
        const int block = (mb[iu].size() != pass + 1) ? mb[iu][pass] : 1;
        for (int b = 0; b < block; ++b) {
          const int row = counter[iu] + nu * b;
          if (row >= n) return cfail(LMPC_E_INVALID, "move blocking: blocks exceed the control horizon Nc");
          T(row, new_id) = 1.0;
        }
        if (counter[iu] < nbound) keep.push_back(counter[iu]);
        counter[iu] += nu * block;
        ++new_id;
      }
    const Mat Tt = tr(T);
    H = mul(Tt, mul(H, T));
    fth = mul(Tt, fth);
    {
      std::vector<double> f2(nnew, 0.0);
      for (int j = 0; j < n; ++j) for (int c2 = 0; c2 < nnew; ++c2) f2[c2] += T(j, c2) * fvec[j];
      fvec.swap(f2);
    }
    if (ms > 0) {
      std::vector<double> bu2, bl2;
      for (int r : keep) { bu2.push_back(bu[r]); bl2.push_back(bl[r]); }
      bu.swap(bu2); bl.swap(bl2);
      ms = (int)keep.size();
      soft.assign(ms, 0); prio.assign(ms, 0);
      W = Mat(ms, nth);
    }
    if (hasK && nbound > 0) {   // mpc2mpqp.jl:854: with a feedback the bounds are rows of A; keep one per block
      std::vector<int> rows = keep;
      for (int i = nbound; i < (int)Arows.size(); ++i) rows.push_back(i);
      std::vector<std::vector<double>> A2, W2; std::vector<double> u2, l2; std::vector<int> s2, p2;
      for (int i : rows) { A2.push_back(Arows[i]); W2.push_back(Wrows[i]); u2.push_back(gbu[i]); l2.push_back(gbl[i]); s2.push_back(gsoft[i]); p2.push_back(gprio[i]); }
      Arows.swap(A2); Wrows.swap(W2); gbu.swap(u2); gbl.swap(l2); gsoft.swap(s2); gprio.swap(p2);
    }
    for (auto &ar : Arows) {
      std::vector<double> r2(nnew, 0.0);
      for (int j = 0; j < n; ++j) if (ar[j] != 0.0) for (int c2 = 0; c2 < nnew; ++c2) r2[c2] += ar[j] * T(j, c2);
      ar.swap(r2);
    }
    nq = nnew;
  }
  // stable sort of the general rows by priority (mpc2mpqp.jl:859-866)
  std::vector<int> order(Arows.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return gprio[a] < gprio[b]; });
  {
    std::vector<std::vector<double>> A2, W2; std::vector<double> u2, l2; std::vector<int> s2, p2;
    for (int i : order) { A2.push_back(Arows[i]); W2.push_back(Wrows[i]); u2.push_back(gbu[i]); l2.push_back(gbl[i]); s2.push_back(gsoft[i]); p2.push_back(gprio[i]); }
    Arows.swap(A2); Wrows.swap(W2); gbu.swap(u2); gbl.swap(l2); gsoft.swap(s2); gprio.swap(p2);
  }
  if (md->preprocess) {
    // remove_redundant (mpc2mpqp.jl:733-773)
    std::vector<int> keep; std::vector<double> scale;
    for (size_t i = 0; i < Arows.size(); ++i) {
      std::vector<double> &a = Arows[i];
      double nf = 0.0; for (double v : a) nf += v * v; nf = std::sqrt(nf);
      if (!(nf > 1e-10)) continue;
      int nz = 0; while (nz < nq && !(std::fabs(a[nz]) > 1e-12)) ++nz;
      if (a[nz] < 0) {
        for (double &v : a) v = -v + 0.0;
        const double t = gbu[i]; gbu[i] = -gbl[i]; gbl[i] = -t;
        for (double &v : Wrows[i]) v = -v + 0.0;
      }
      int cnt = 0; for (double v : a) if (std::fabs(v) > 1e-12) ++cnt;
      if (cnt == 1 && nz < ms && prio[nz] == gprio[i] && soft[nz] == gsoft[i]) {
        bool same = true;
        for (int j = 0; j < nth; ++j) if (Wrows[i][j] - W(nz, j) != 0.0) same = false;
        if (same) { bu[nz] = std::min(bu[nz], gbu[i] / nf); bl[nz] = std::max(bl[nz], gbl[i] / nf); continue; }
      }
      keep.push_back((int)i); scale.push_back(1.0 / nf);
    }
    if (keep.size() < Arows.size()) {   // rows are normalised only when something was removed (reference behaviour)
      std::vector<std::vector<double>> A2, W2; std::vector<double> u2, l2; std::vector<int> s2, p2;
      for (size_t q = 0; q < keep.size(); ++q) {
        const int i = keep[q]; const double sc = scale[q];
        std::vector<double> a = Arows[i], w = Wrows[i];
        for (double &v : a) v *= sc; for (double &v : w) v *= sc;
        A2.push_back(a); W2.push_back(w); u2.push_back(gbu[i] * sc); l2.push_back(gbl[i] * sc); s2.push_back(gsoft[i]); p2.push_back(gprio[i]);
      }
      Arows.swap(A2); Wrows.swap(W2); gbu.swap(u2); gbl.swap(l2); gsoft.swap(s2); gprio.swap(p2);
    }
    // remove_duplicate (mpc2mpqp.jl:775-828): rows equal to 6 digits in [A W soft prio]
    std::map<std::vector<long long>, std::vector<int>> groups;
    std::vector<std::vector<long long>> first_seen;
    for (size_t i = 0; i < Arows.size(); ++i) {
      std::vector<long long> key;
      for (double v : Arows[i]) key.push_back(std::llround(v * 1e6));
      for (double v : Wrows[i]) key.push_back(std::llround(v * 1e6));
      key.push_back(gsoft[i]); key.push_back(gprio[i]);
      if (!groups.count(key)) first_seen.push_back(key);
      groups[key].push_back((int)i);
    }
    if (first_seen.size() < Arows.size()) {
      std::vector<std::vector<double>> A2, W2; std::vector<double> u2, l2; std::vector<int> s2, p2;
      for (auto &key : first_seen) {
        const std::vector<int> &g = groups[key];
        const int i0 = g[0];
        double mu = gbu[i0], ml = gbl[i0];
        for (int i : g) { mu = std::min(mu, gbu[i]); ml = std::max(ml, gbl[i]); }
        A2.push_back(Arows[i0]); W2.push_back(Wrows[i0]); u2.push_back(mu); l2.push_back(ml); s2.push_back(gsoft[i0]); p2.push_back(gprio[i0]);
      }
      Arows.swap(A2); Wrows.swap(W2); gbu.swap(u2); gbl.swap(l2); gsoft.swap(s2); gprio.swap(p2);
    }
  }
  lmpc_mpqp *q = new lmpc_mpqp();
  const int mg = (int)Arows.size(), m = ms + mg;
  q->n = nq; q->m = m; q->ms = ms; q->nth = nth; q->nx = nx; q->nr = nrp; q->nd = ndp; q->nuprev = nup; q->nu = nu;
  q->H = H; q->f_theta = fth; q->f = fvec; q->traj2setpoint = T2S;
  q->A = Mat(mg, nq); q->W = Mat(m, nth);
  for (int i = 0; i < ms; ++i) for (int j = 0; j < nth; ++j) q->W(i, j) = W(i, j);
  q->bu = bu; q->bl = bl;
  for (int i = 0; i < mg; ++i) {
    for (int j = 0; j < nq; ++j) q->A(i, j) = Arows[i][j];
    for (int j = 0; j < nth; ++j) q->W(ms + i, j) = Wrows[i][j];
    q->bu.push_back(gbu[i]); q->bl.push_back(gbl[i]);
    soft.push_back(gsoft[i]);
  }
  // senses + clamping (mpc2mpqp.jl:868-888)
  q->senses.assign(m, 0);
  for (int i = 0; i < m; ++i) {
    if (q->bu[i] > 1e20 && q->bl[i] < -1e20) q->senses[i] = LMPC_IMMUTABLE;
    else if (std::fabs(q->bu[i] - q->bl[i]) < 1e-12) q->senses[i] = LMPC_ACTIVE + LMPC_IMMUTABLE;
    if (soft[i]) q->senses[i] += LMPC_SOFT;
    q->bu[i] = std::min(std::max(q->bu[i], -1e30), 1e30);
    q->bl[i] = std::min(std::max(q->bl[i], -1e30), 1e30);
  }
  *out = q;
  return 0;
}

extern "C" int lmpc_gpu_mpqp_dims(const lmpc_mpqp *q, int32_t *dims9) {
  if (!q || !dims9) return cfail(LMPC_E_INVALID, "null argument");
  const int32_t d[9] = {q->n, q->m, q->ms, q->nth, q->nu, q->nx, q->nr, q->nd, q->nuprev};
  memcpy(dims9, d, sizeof(d));
  return 0;
}

extern "C" int lmpc_gpu_mpqp_traj2setpoint(const lmpc_mpqp *q, double *T, int32_t *rows, int32_t *cols) {
  if (!q) return cfail(LMPC_E_INVALID, "null argument");
  if (rows) *rows = q->traj2setpoint.r;
  if (cols) *cols = q->traj2setpoint.c;
  if (T) for (int j = 0; j < q->traj2setpoint.c; ++j) for (int i = 0; i < q->traj2setpoint.r; ++i)
    T[(size_t)j * q->traj2setpoint.r + i] = q->traj2setpoint(i, j);
  return 0;
}

extern "C" int lmpc_gpu_mpqp_get(const lmpc_mpqp *q, double *H, double *f, double *f_theta, double *A,
                                 double *bu, double *bl, double *W, int32_t *senses) {
  if (!q) return cfail(LMPC_E_INVALID, "null argument");
  const int n = q->n, m = q->m, mg = q->m - q->ms, nth = q->nth;
  if (H) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) H[(size_t)j * n + i] = q->H(i, j);
  if (f) memcpy(f, q->f.data(), sizeof(double) * n);
  if (f_theta) for (int j = 0; j < nth; ++j) for (int i = 0; i < n; ++i) f_theta[(size_t)j * n + i] = q->f_theta(i, j);
  if (A) for (int j = 0; j < n; ++j) for (int i = 0; i < mg; ++i) A[(size_t)j * mg + i] = q->A(i, j);
  if (bu) memcpy(bu, q->bu.data(), sizeof(double) * m);
  if (bl) memcpy(bl, q->bl.data(), sizeof(double) * m);
  if (W) for (int j = 0; j < nth; ++j) for (int i = 0; i < m; ++i) W[(size_t)j * m + i] = q->W(i, j);
  if (senses) memcpy(senses, q->senses.data(), sizeof(int32_t) * m);
  return 0;
}

extern "C" int lmpc_gpu_mpqp_free(lmpc_mpqp *q) { delete q; return 0; }
