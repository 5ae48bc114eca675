// Stepped closed loop: the pieces of the Simulation loop (/root/reference/src/simulation.jl:93-113)
// that sit around compute_control when the plant is a user function (scenario.dynamics, :39,110) or the
// observer has its own state (KalmanFilter / OffsetFreeObserver, /root/reference/src/observer.jl:1-148,203-222).
// One thread per closed loop, state in HBM between the kernels of a step:
//   step_pre  : measured disturbance of the step, measurement + correct! (observer.jl:114-124), theta
//   (solve)   : the controller's LDP kernel, warm-started from the previous step's working set
//   step_post : worst flag / iteration total / control trajectory, predict! (observer.jl:104-112), the affine
//               plant step unless the caller steps the plant itself
// Arithmetic: fma chains in index order, exactly the oracle's (oracle/ldp_oracle.c: step_chunk).
#include "ldp_device.cuh"

namespace {

constexpr int SS_BLOCK = 128;
constexpr int SS_MAXX = 64, SS_MAXY = 32, SS_MAXD = 32, SS_MAXU = 16;

LMPC_DEVINL double row_fma(const double *A, int i, int n, const double *v, double acc) {
  if (A) for (int c = 0; c < n; ++c) acc = fma(__ldg(A + i * n + c), v[c], acc);
  return acc;
}

__global__ void __launch_bounds__(SS_BLOCK) step_pre_kernel(StepDesc d, int t, int64_t N) {
  const int64_t q = (int64_t)blockIdx.x * SS_BLOCK + threadIdx.x;
  if (q >= N) return;
  const int nx = d.nx, nxo = d.nxo, ndm = d.ndm, ny = d.ny, ndo = nxo - nx;
  double dm[SS_MAXD], xh[SS_MAXX];
  for (int i = 0; i < ndm; ++i) {
    dm[i] = d.Dtraj ? d.Dtraj[(size_t)t * ndm + i] : (d.D ? d.D[q * ndm + i] : 0.0);
    d.Dm[q * ndm + i] = dm[i];
  }
  const double *xs = d.Xs + q * nx;
  if (d.has_obs) {
    double xsl[SS_MAXX], inn[SS_MAXY];
    for (int i = 0; i < nx; ++i) xsl[i] = xs[i];
    for (int i = 0; i < nxo; ++i) xh[i] = d.Xh[q * nxo + i];
    for (int i = 0; i < ny; ++i) {
      double y = d.Vtraj ? d.Vtraj[(size_t)t * ny + i] : 0.0;
      y = row_fma(d.Cm, i, nx, xsl, y); y = row_fma(d.Ddm, i, ndm, dm, y);
      if (d.hm) y += d.hm[i];
      double yh = row_fma(d.Co, i, nxo, xh, 0.0); yh = row_fma(d.Ddo, i, ndm, dm, yh);
      if (d.ho) yh += d.ho[i];
      inn[i] = y - yh;
    }
    for (int j = 0; j < nxo; ++j) d.Xh[q * nxo + j] = row_fma(d.Ko, j, ny, inn, xh[j]);
    for (int j = 0; j < nxo; ++j) xh[j] = d.Xh[q * nxo + j];
  } else {
    for (int j = 0; j < nx; ++j) { xh[j] = xs[j]; d.Xh[q * nxo + j] = xh[j]; }
  }
  double *th = d.Theta + q * d.nth;
  int o = 0;
  for (int i = 0; i < nx; ++i) th[o++] = xh[i];
  for (int i = 0; i < d.nr; ++i) th[o++] = d.Rtraj ? d.Rtraj[(size_t)t * d.nr + i] : (d.R ? d.R[q * d.nr + i] : 0.0);
  for (int i = 0; i < ndm; ++i) th[o++] = dm[i];
  for (int i = 0; i < ndo; ++i) th[o++] = xh[nx + i];
  for (int i = 0; i < d.nuprev; ++i) th[o++] = d.Uprev[q * d.nu + i];
}

__global__ void __launch_bounds__(SS_BLOCK) step_post_kernel(StepDesc d, int t, int64_t N) {
  const int64_t q = (int64_t)blockIdx.x * SS_BLOCK + threadIdx.x;
  if (q >= N) return;
  const int nx = d.nx, nxo = d.nxo, ndm = d.ndm, nu = d.nu;
  double u[SS_MAXU], dm[SS_MAXD];
  for (int i = 0; i < nu; ++i) { u[i] = d.U[q * nu + i]; d.Uprev[q * nu + i] = u[i]; }
  for (int i = 0; i < ndm; ++i) dm[i] = d.Dm[q * ndm + i];
  const int f = d.flag_step[q];
  int worst = t == 0 ? EXIT_OPTIMAL : d.flag_worst[q];
  if (f < 0) { if (worst > 0 || f < worst) worst = f; }
  else if (worst > 0 && f > worst) worst = f;
  d.flag_worst[q] = worst;
  if (d.iters_tot) d.iters_tot[q] = (t == 0 ? 0 : d.iters_tot[q]) + d.iters_step[q];
  if (d.Utraj) for (int i = 0; i < nu; ++i) d.Utraj[((size_t)q * d.T + t) * nu + i] = u[i];
  if (d.has_obs) {
    double xh[SS_MAXX];
    for (int i = 0; i < nxo; ++i) xh[i] = d.Xh[q * nxo + i];
    for (int i = 0; i < nxo; ++i) {
      double acc = row_fma(d.Fo, i, nxo, xh, 0.0);
      acc = row_fma(d.Go, i, nu, u, acc); acc = row_fma(d.Gdo, i, ndm, dm, acc);
      d.Xh[q * nxo + i] = d.fo ? acc + d.fo[i] : acc;
    }
  }
  if (!d.user_plant) {
    double xs[SS_MAXX];
    for (int i = 0; i < nx; ++i) xs[i] = d.Xs[q * nx + i];
    for (int i = 0; i < nx; ++i) {
      double acc = row_fma(d.Ft, i, nx, xs, 0.0);
      acc = row_fma(d.Gt, i, nu, u, acc); acc = row_fma(d.Gdt, i, ndm, dm, acc);
      d.Xn[q * nx + i] = d.ft ? acc + d.ft[i] : acc;
    }
  }
}

}  // namespace

bool step_dims_ok(const StepDesc &d) {
  return d.nx >= 1 && d.nx <= SS_MAXX && d.nxo >= d.nx && d.nxo <= SS_MAXX && d.ny >= 0 && d.ny <= SS_MAXY &&
         d.ndm >= 0 && d.ndm <= SS_MAXD && d.nu >= 1 && d.nu <= SS_MAXU;
}
int launch_step_pre(const StepDesc &d, int t, int64_t N, void *stream) {
  step_pre_kernel<<<(unsigned)((N + SS_BLOCK - 1) / SS_BLOCK), SS_BLOCK, 0, (cudaStream_t)stream>>>(d, t, N);
  return (int)cudaGetLastError();
}
int launch_step_post(const StepDesc &d, int t, int64_t N, void *stream) {
  step_post_kernel<<<(unsigned)((N + SS_BLOCK - 1) / SS_BLOCK), SS_BLOCK, 0, (cudaStream_t)stream>>>(d, t, N);
  return (int)cudaGetLastError();
}
