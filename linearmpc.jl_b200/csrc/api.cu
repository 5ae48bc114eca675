// C ABI of liblmpcgpu.so (declared in include/lmpc_gpu.h).  Host-side plumbing only:
// handle lifetime, constant upload, host<->device staging with a 3-slot copy/compute
// pipeline, kernel selection.  There is NO CPU fallback: without an sm_100 device every
// entry point fails with LMPC_E_NOGPU.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <functional>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <cmath>
#include <algorithm>
#include "lmpc_common.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}
void lmpc_count_kernel(int n) { g_launches.fetch_add(n); }
namespace {

int set_err(int code, const char *fmt, const char *detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}
#define CUDA_OK(call)                                                                  \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) return set_err(LMPC_E_CUDA, #call ": %s", cudaGetErrorString(e__)); \
  } while (0)

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return set_err(LMPC_E_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    cap = bytes;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

constexpr int kSlots = 3;
constexpr int64_t kChunk = 1 << 20;   // instances per pipeline stage for host buffers

// pinned staging for callers that hand us pageable arrays (a Julia Matrix, a numpy array)
struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return set_err(LMPC_E_CUDA, "cudaHostAlloc: %s", cudaGetErrorString(e));
    cap = bytes;
    return 0;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Slot {
  cudaStream_t stream = nullptr;
  DevBuf in[4], u, flag, iters, au, al, aux[3];
  PinBuf hin[4], hu, hflag, hiters, hau, hal;
  int64_t pend_off = 0, pend_cnt = 0;   // results sitting in the pinned buffers, not yet handed out
};

// Error paths of the host pipelines return in the middle of the stage loop while other slots may still
// have D2H copies into the CALLER's arrays in flight: wait for all of them before handing control back.
struct SlotGuard {
  Slot *slots;
  bool armed = true;
  explicit SlotGuard(Slot *s) : slots(s) {}
  ~SlotGuard() {
    if (!armed) return;
    for (int i = 0; i < kSlots; ++i) {
      if (slots[i].stream) cudaStreamSynchronize(slots[i].stream);
      slots[i].pend_cnt = 0;
    }
    cudaGetLastError();
  }
};

bool is_pageable(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

// The driver's own pageable path stages through one thread (about 8 GB/s measured); several
// host threads copying into / out of our pinned ring keep the PCIe link busy instead.
int copy_threads(int share) {
  static const int n = [] {
    if (const char *e = getenv("LMPC_GPU_COPY_THREADS")) return std::max(1, atoi(e));
    const int hc = (int)std::thread::hardware_concurrency();
    return std::max(1, std::min(8, hc / 2));
  }();
  return std::max(1, n / std::max(1, share));
}

// A few persistent copy workers per process (spawning threads per 2^18-instance stage cost a
// visible fraction of the stage).  One job at a time: callers (the per-device pipelines of a
// multi-device handle) take turns, which is fine because one job already saturates the workers.
class CopyPool {
 public:
  static CopyPool &get() { static CopyPool p; return p; }
  void run(int64_t rows, int nthreads, const std::function<void(int64_t, int64_t)> &f) {
    if (nthreads <= 1 || rows < 4096) { f(0, rows); return; }
    std::lock_guard<std::mutex> job(job_mu_);
    grow(nthreads - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      f_ = &f; rows_ = rows; parts_ = nthreads; next_ = 1; pending_ = nthreads - 1; ++gen_;
    }
    cv_.notify_all();
    f(0, rows / nthreads);                       // the caller is worker 0
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
    f_ = nullptr;
  }

 private:
  CopyPool() = default;
  ~CopyPool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  void grow(int n) {
    while ((int)th_.size() < n) th_.emplace_back([this] { work(); });
  }
  void work() {
    uint64_t seen = 0;
    for (;;) {
      int part;
      const std::function<void(int64_t, int64_t)> *f;
      int64_t rows; int parts;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || (gen_ != seen && next_ < parts_); });
        if (stop_) return;
        part = next_++;
        if (next_ >= parts_) seen = gen_;
        f = f_; rows = rows_; parts = parts_;
      }
      (*f)(rows * part / parts, rows * (part + 1) / parts);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::mutex job_mu_, mu_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> th_;
  const std::function<void(int64_t, int64_t)> *f_ = nullptr;
  int64_t rows_ = 0;
  int parts_ = 0, next_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

void parallel_rows(int64_t rows, int nthreads, const std::function<void(int64_t, int64_t)> &f) {
  CopyPool::get().run(rows, nthreads, f);
}

bool is_device_ptr(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int check_device(int device, int *sm_count) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt <= 0) {
    cudaGetLastError();
    return set_err(LMPC_E_NOGPU, "no CUDA device visible: liblmpcgpu has no CPU fallback%s");
  }
  if (device < 0 || device >= cnt) return set_err(LMPC_E_INVALID, "device index out of range%s");
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_err(LMPC_E_NOGPU, "device is not sm_100 (B200): kernels are built for sm_100a only%s");
  *sm_count = prop.multiProcessorCount;
  return 0;
}

}  // namespace

struct lmpc_handle {
  int device = 0, sm_count = 148, kernel = 0;
  int copy_share = 1;   // divisor of the copy-thread count (1: the shared pool runs one job at a time anyway)
  LdpLayout lay;
  LdpSettings st;
  double *h_blk = nullptr;
  double *d_blk = nullptr;
  cudaStream_t stream = nullptr;
  Slot slots[kSlots];
  DevBuf dyn;   // F, G, Gd for simulate
  DevBuf counter;   // failed-instance count of the last *_bcast call
  std::mutex mu;   // host-buffer calls share the staging slots / dyn: one at a time per handle
};

struct lmpc_bst {
  int device = 0, sm_count = 148;
  BstDev dev;
  double *d_blk = nullptr;
  cudaStream_t stream = nullptr;
  Slot slots[kSlots];
  std::mutex mu;
};

extern "C" {

void lmpc_gpu_default_settings(lmpc_settings *s) {
  if (!s) return;
  s->primal_tol = 1e-6; s->dual_tol = 1e-12; s->zero_tol = 1e-11; s->progress_tol = 1e-6;
  s->fval_bound = 1e30; s->cycle_tol = 10; s->iter_limit = 10000; s->kernel = LMPC_KERNEL_AUTO;
  s->certified_iterations = 0;
}

const char *lmpc_gpu_last_error(void) { return g_err; }

int lmpc_gpu_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) { cudaGetLastError(); return 0; }
  return cnt;
}

int64_t lmpc_gpu_launch_count(void) { return g_launches.load(); }

int lmpc_gpu_create(const lmpc_problem *p, const lmpc_settings *s, int device, lmpc_handle **out) {
  if (!p || !out) return set_err(LMPC_E_INVALID, "lmpc_gpu_create: null argument%s");
  *out = nullptr;
  LdpHost hs;
  char msg[256] = "";
  int rc = lmpc_host_setup(p, &hs, msg, sizeof(msg));
  if (rc != 0) return set_err(rc, "%s", msg);
  int sm = 0;
  rc = check_device(device, &sm);
  if (rc != 0) { free(hs.blk); return rc; }
  lmpc_settings def;
  lmpc_gpu_default_settings(&def);
  if (s) def = *s;
  lmpc_handle *h = new lmpc_handle();
  h->device = device; h->sm_count = sm; h->lay = hs.lay; h->h_blk = hs.blk;
  h->st.primal_tol = def.primal_tol; h->st.dual_tol = def.dual_tol; h->st.zero_tol = def.zero_tol;
  h->st.progress_tol = def.progress_tol; h->st.fval_bound = def.fval_bound;
  h->st.rho_soft = 1.0 / (p->soft_weight > 0 ? p->soft_weight : 1e6);
  h->st.cycle_tol = def.cycle_tol; h->st.iter_limit = def.iter_limit;
  if (def.certified_iterations > 0) {
    // certify(mpc; range).max_iterations (src/certify.jl:18-29): hard cap + static capacity sizing
    if (h->st.iter_limit > def.certified_iterations + 1) h->st.iter_limit = def.certified_iterations + 1;
    const int capc = h->lay.neq + def.certified_iterations;
    if (capc >= 1 && capc < h->lay.cap) h->lay.cap = capc;
  }
  int kern = def.kernel;
  if (kern == LMPC_KERNEL_AUTO) {
    if (ldp_thread_supports(h->lay)) kern = LMPC_KERNEL_THREAD;
    else if (ldp_warp_supports(h->lay)) kern = LMPC_KERNEL_WARP;
    else if (ldp_cta_supports(h->lay)) kern = LMPC_KERNEL_CTA;
    else kern = -1;
  } else {
    const bool ok = (kern == LMPC_KERNEL_THREAD && ldp_thread_supports(h->lay)) ||
                    (kern == LMPC_KERNEL_WARP && ldp_warp_supports(h->lay)) ||
                    (kern == LMPC_KERNEL_CTA && ldp_cta_supports(h->lay));
    if (!ok) kern = -1;
  }
  if (kern < 0) {
    free(hs.blk); delete h;
    return set_err(LMPC_E_UNSUPPORTED, "no kernel supports these dimensions (n, m, capacity)%s");
  }
  h->kernel = kern;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc((void **)&h->d_blk, (size_t)h->lay.total * sizeof(double)) != cudaSuccess ||
      cudaMemcpy(h->d_blk, h->h_blk, (size_t)h->lay.total * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    const char *es = cudaGetErrorString(cudaGetLastError());
    lmpc_gpu_destroy(h);
    return set_err(LMPC_E_CUDA, "lmpc_gpu_create: %s", es);
  }
  for (int i = 0; i < kSlots; ++i)
    if (cudaStreamCreateWithFlags(&h->slots[i].stream, cudaStreamNonBlocking) != cudaSuccess) {
      lmpc_gpu_destroy(h);
      return set_err(LMPC_E_CUDA, "cudaStreamCreate failed%s");
    }
  *out = h;
  return 0;
}

static void free_slots(Slot *slots) {
  for (int i = 0; i < kSlots; ++i) {
    Slot &s = slots[i];
    for (auto &b : s.in) b.release();
    for (auto &b : s.aux) b.release();
    s.u.release(); s.flag.release(); s.iters.release(); s.au.release(); s.al.release();
    for (auto &b : s.hin) b.release();
    s.hu.release(); s.hflag.release(); s.hiters.release(); s.hau.release(); s.hal.release();
    if (s.stream) cudaStreamDestroy(s.stream);
    s.stream = nullptr;
  }
}

int lmpc_gpu_destroy(lmpc_handle *h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  free_slots(h->slots);
  h->dyn.release();
  h->counter.release();
  if (h->d_blk) cudaFree(h->d_blk);
  if (h->stream) cudaStreamDestroy(h->stream);
  free(h->h_blk);
  delete h;
  return 0;
}

int lmpc_gpu_get_ldp(const lmpc_handle *h, double *Mbar, double *Gram, double *Dth, double *du,
                     double *dl, double *Rinv, double *Uth, double *uoff) {
  if (!h) return set_err(LMPC_E_INVALID, "null handle%s");
  const LdpLayout &l = h->lay;
  const double *b = h->h_blk;
  if (Mbar) memcpy(Mbar, b + l.oM, sizeof(double) * l.m * l.n);
  if (Gram) memcpy(Gram, b + l.oG, sizeof(double) * l.m * l.m);
  if (Dth) memcpy(Dth, b + l.oDth, sizeof(double) * l.m * l.nth);
  if (du) memcpy(du, b + l.odu, sizeof(double) * l.m);
  if (dl) memcpy(dl, b + l.odl, sizeof(double) * l.m);
  if (Rinv) memcpy(Rinv, b + l.oRinv, sizeof(double) * l.nu * l.n);
  if (Uth) memcpy(Uth, b + l.oUth, sizeof(double) * l.nu * l.nth);
  if (uoff) memcpy(uoff, b + l.ouoff, sizeof(double) * l.nu);
  return 0;
}

int lmpc_gpu_host_ldp(const lmpc_problem *p, double *Mbar, double *Gram, double *Dth, double *du,
                      double *dl, double *Rinv, double *Uth, double *uoff, int32_t *senses,
                      int32_t *cap) {
  if (!p) return set_err(LMPC_E_INVALID, "null problem%s");
  LdpHost hs;
  char msg[256] = "";
  int rc = lmpc_host_setup(p, &hs, msg, sizeof(msg));
  if (rc != 0) return set_err(rc, "%s", msg);
  lmpc_handle tmp;
  tmp.lay = hs.lay; tmp.h_blk = hs.blk;
  lmpc_gpu_get_ldp(&tmp, Mbar, Gram, Dth, du, dl, Rinv, Uth, uoff);
  if (senses) memcpy(senses, hs.blk + hs.lay.oSense, sizeof(int32_t) * hs.lay.m);
  if (cap) *cap = hs.lay.cap;
  free(hs.blk);
  return 0;
}

int lmpc_gpu_get_info(const lmpc_handle *h, int32_t *kernel, int32_t *cap, int32_t *senses) {
  if (!h) return set_err(LMPC_E_INVALID, "null handle%s");
  if (kernel) *kernel = h->kernel;
  if (cap) *cap = h->lay.cap;
  if (senses) memcpy(senses, h->h_blk + h->lay.oSense, sizeof(int32_t) * h->lay.m);
  return 0;
}

int lmpc_gpu_get_tiering(const lmpc_handle *h, int32_t *fast_cap, int64_t *redone) {
  if (!h) return set_err(LMPC_E_INVALID, "null handle%s");
  if (fast_cap) *fast_cap = (h->kernel == LMPC_KERNEL_CTA) ? ldp_cta_fast_cap(h->lay) : 0;
  if (redone) {
    cudaSetDevice(h->device);
    *redone = (h->kernel == LMPC_KERNEL_CTA) ? (int64_t)ldp_cta_redo_count(h->device) : 0;
  }
  return 0;
}

void *lmpc_gpu_stream(lmpc_handle *h) { return h ? (void *)h->stream : nullptr; }

static int run_kernel(lmpc_handle *h, const ThetaSegs &th, int64_t N, const SolveOut &o,
                      const SimDesc *sim, cudaStream_t s) {
  int e;
  switch (h->kernel) {
    case LMPC_KERNEL_THREAD:
      // box-only problems take the specialised fast path (same arithmetic, fewer instructions)
      e = (!sim && !o.ws_u && ldp_box_supports(h->lay))
              ? launch_ldp_box(h->lay, h->st, h->d_blk, h->h_blk, th, N, o, h->sm_count, s)
              : launch_ldp_thread(h->lay, h->st, h->d_blk, th, N, o, sim, h->sm_count, s);
      break;
    case LMPC_KERNEL_WARP: e = launch_ldp_warp(h->lay, h->st, h->d_blk, th, N, o, sim, h->sm_count, s); break;
    case LMPC_KERNEL_CTA: e = launch_ldp_cta(h->lay, h->st, h->d_blk, th, N, o, sim, h->sm_count, s); break;
    default: return set_err(LMPC_E_INVALID, "handle has no kernel%s");
  }
  if (e != 0) return set_err(LMPC_E_CUDA, "kernel launch: %s", cudaGetErrorString((cudaError_t)e));
  g_launches.fetch_add(1);
  return 0;
}

int lmpc_gpu_solve_batch_async(lmpc_handle *h, const double *Theta, int64_t N, double *U,
                               int32_t *exitflag, int32_t *iters, uint64_t *au, uint64_t *al,
                               void *stream) {
  if (!h || !Theta || !U || !exitflag || N < 0 || ((au == nullptr) != (al == nullptr)))
    return set_err(LMPC_E_INVALID, "lmpc_gpu_solve_batch_async: bad argument%s");
  CUDA_OK(cudaSetDevice(h->device));
  ThetaSegs th;
  memset(&th, 0, sizeof(th));
  th.nseg = 1; th.ptr[0] = Theta; th.len[0] = h->lay.nth; th.stride[0] = h->lay.nth;
  SolveOut o{U, exitflag, iters, au, al};
  return run_kernel(h, th, N, o, nullptr, stream ? (cudaStream_t)stream : h->stream);
}

// hand a finished stage's results from the pinned ring to the caller's pageable arrays
static int drain_slot(lmpc_handle *h, Slot &s, double *U, int ustride, int32_t *flag, int32_t *iters,
                      uint64_t *au, uint64_t *al) {
  CUDA_OK(cudaStreamSynchronize(s.stream));
  if (s.pend_cnt == 0) return 0;
  const LdpLayout &l = h->lay;
  const int words = (l.m + 63) / 64;
  const int64_t off = s.pend_off;
  parallel_rows(s.pend_cnt, copy_threads(h->copy_share), [&](int64_t a, int64_t b) {
    const double *su = (const double *)s.hu.p;
    if (ustride == l.nu) memcpy(U + (off + a) * l.nu, su + a * l.nu, (size_t)(b - a) * l.nu * 8);
    else for (int64_t i = a; i < b; ++i) memcpy(U + (off + i) * ustride, su + i * l.nu, (size_t)l.nu * 8);
    if (flag) memcpy(flag + off + a, (const int32_t *)s.hflag.p + a, (size_t)(b - a) * 4);
    if (iters) memcpy(iters + off + a, (const int32_t *)s.hiters.p + a, (size_t)(b - a) * 4);
    if (au) {
      memcpy(au + (off + a) * words, (const uint64_t *)s.hau.p + a * words, (size_t)(b - a) * words * 8);
      memcpy(al + (off + a) * words, (const uint64_t *)s.hal.p + a * words, (size_t)(b - a) * words * 8);
    }
  });
  s.pend_cnt = 0;
  return 0;
}

// host buffers: H2D -> kernel -> D2H in kChunk-sized stages round-robined over kSlots streams.
// Pinned caller memory is DMA'd in place; pageable memory goes through a pinned ring filled and
// emptied by several host threads while the other stages' DMA and kernels run.
static int solve_segments_host(lmpc_handle *h, const double *const *seg, const int *len,
                               const int *stride, int nseg, int64_t N, double *U, int ustride,
                               int32_t *flag, int32_t *iters, uint64_t *au, uint64_t *al,
                               bool warm = false, int64_t *n_failed = nullptr) {
  std::lock_guard<std::mutex> lock(h->mu);
  SlotGuard guard(h->slots);
  const LdpLayout &l = h->lay;
  const int words = (l.m + 63) / 64;
  bool stage = is_pageable(U) || (flag && is_pageable(flag));
  if (n_failed) {
    if (h->counter.reserve(8)) return LMPC_E_CUDA;
    CUDA_OK(cudaMemsetAsync(h->counter.p, 0, 8, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
  }
  for (int g = 0; g < nseg; ++g) stage = stage || is_pageable(seg[g]);
  const int64_t chunk = stage ? kChunk / 4 : kChunk;   // smaller stages: the ring fills sooner
  for (int i = 0; i < kSlots; ++i) h->slots[i].pend_cnt = 0;
  for (int64_t off = 0, c = 0; off < N; off += chunk, ++c) {
    const int64_t cnt = (N - off < chunk) ? (N - off) : chunk;
    Slot &s = h->slots[c % kSlots];
    if (stage) { int rc = drain_slot(h, s, U, ustride, flag, iters, au, al); if (rc) return rc; }
    ThetaSegs th;
    memset(&th, 0, sizeof(th));
    th.nseg = nseg;
    for (int g = 0; g < nseg; ++g) {
      // stride 0: one vector shared by every instance (broadcast segment), uploaded once per stage
      const size_t bytes = (stride[g] == 0 ? (size_t)len[g] : (size_t)cnt * stride[g]) * sizeof(double);
      if (s.in[g].reserve(bytes)) return LMPC_E_CUDA;
      const double *src = seg[g] + off * stride[g];
      if (stage && stride[g] == 0) {
        if (s.hin[g].reserve(bytes)) return LMPC_E_CUDA;
        memcpy(s.hin[g].p, src, bytes);
        src = (const double *)s.hin[g].p;
      } else if (stage) {
        if (s.hin[g].reserve(bytes)) return LMPC_E_CUDA;
        char *dst = (char *)s.hin[g].p;
        const size_t row = (size_t)stride[g] * 8;
        parallel_rows(cnt, copy_threads(h->copy_share), [&](int64_t a, int64_t b) {
          memcpy(dst + a * row, (const char *)src + a * row, (size_t)(b - a) * row);
        });
        src = (const double *)s.hin[g].p;
      }
      CUDA_OK(cudaMemcpyAsync(s.in[g].p, src, bytes, cudaMemcpyHostToDevice, s.stream));
      th.ptr[g] = (const double *)s.in[g].p; th.len[g] = len[g]; th.stride[g] = stride[g];
    }
    if (s.u.reserve((size_t)cnt * l.nu * sizeof(double)) || s.flag.reserve((size_t)cnt * 4)) return LMPC_E_CUDA;
    if (iters && s.iters.reserve((size_t)cnt * 4)) return LMPC_E_CUDA;
    if (au && (s.au.reserve((size_t)cnt * words * 8) || s.al.reserve((size_t)cnt * words * 8))) return LMPC_E_CUDA;
    SolveOut o{(double *)s.u.p, (int32_t *)s.flag.p, iters ? (int32_t *)s.iters.p : nullptr,
               au ? (uint64_t *)s.au.p : nullptr, au ? (uint64_t *)s.al.p : nullptr};
    if (warm) {   // the caller's working sets go in through the same device buffers the final ones come out of
      CUDA_OK(cudaMemcpyAsync(s.au.p, au + off * words, (size_t)cnt * words * 8, cudaMemcpyHostToDevice, s.stream));
      CUDA_OK(cudaMemcpyAsync(s.al.p, al + off * words, (size_t)cnt * words * 8, cudaMemcpyHostToDevice, s.stream));
      o.ws_u = o.act_u; o.ws_l = o.act_l;
    }
    int rc = run_kernel(h, th, cnt, o, nullptr, s.stream);
    if (rc) return rc;
    if (n_failed) {   // exit flags < 1 are counted on the device: the caller may then skip the flag array altogether
      int e = launch_count_failed((const int32_t *)s.flag.p, cnt, (unsigned long long *)h->counter.p, s.stream);
      if (e) return set_err(LMPC_E_CUDA, "count kernel launch: %s", cudaGetErrorString((cudaError_t)e));
      g_launches.fetch_add(1);
    }
    if (stage) {
      if (s.hu.reserve((size_t)cnt * l.nu * 8) || s.hflag.reserve((size_t)cnt * 4)) return LMPC_E_CUDA;
      if (iters && s.hiters.reserve((size_t)cnt * 4)) return LMPC_E_CUDA;
      if (au && (s.hau.reserve((size_t)cnt * words * 8) || s.hal.reserve((size_t)cnt * words * 8))) return LMPC_E_CUDA;
      CUDA_OK(cudaMemcpyAsync(s.hu.p, s.u.p, (size_t)cnt * l.nu * 8, cudaMemcpyDeviceToHost, s.stream));
      if (flag) CUDA_OK(cudaMemcpyAsync(s.hflag.p, s.flag.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.stream));
      if (iters) CUDA_OK(cudaMemcpyAsync(s.hiters.p, s.iters.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.stream));
      if (au) {
        CUDA_OK(cudaMemcpyAsync(s.hau.p, s.au.p, (size_t)cnt * words * 8, cudaMemcpyDeviceToHost, s.stream));
        CUDA_OK(cudaMemcpyAsync(s.hal.p, s.al.p, (size_t)cnt * words * 8, cudaMemcpyDeviceToHost, s.stream));
      }
      s.pend_off = off; s.pend_cnt = cnt;
      continue;
    }
    if (ustride == l.nu) {
      CUDA_OK(cudaMemcpyAsync(U + off * l.nu, s.u.p, (size_t)cnt * l.nu * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    } else {
      CUDA_OK(cudaMemcpy2DAsync(U + off * ustride, (size_t)ustride * 8, s.u.p, (size_t)l.nu * 8,
                                (size_t)l.nu * 8, (size_t)cnt, cudaMemcpyDeviceToHost, s.stream));
    }
    if (flag) CUDA_OK(cudaMemcpyAsync(flag + off, s.flag.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.stream));
    if (iters) CUDA_OK(cudaMemcpyAsync(iters + off, s.iters.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.stream));
    if (au) {
      CUDA_OK(cudaMemcpyAsync(au + off * words, s.au.p, (size_t)cnt * words * 8, cudaMemcpyDeviceToHost, s.stream));
      CUDA_OK(cudaMemcpyAsync(al + off * words, s.al.p, (size_t)cnt * words * 8, cudaMemcpyDeviceToHost, s.stream));
    }
  }
  for (int i = 0; i < kSlots; ++i) {
    int rc = drain_slot(h, h->slots[i], U, ustride, flag, iters, au, al);
    if (rc) return rc;
  }
  if (n_failed) {
    unsigned long long cnt = 0;
    CUDA_OK(cudaMemcpy(&cnt, h->counter.p, 8, cudaMemcpyDeviceToHost));
    *n_failed = (int64_t)cnt;
  }
  guard.armed = false;
  return 0;
}

int lmpc_gpu_solve_batch(lmpc_handle *h, const double *Theta, int64_t N, double *U,
                         int32_t *exitflag, int32_t *iters, uint64_t *au, uint64_t *al) {
  if (!h || !Theta || !U || !exitflag || N < 0 || ((au == nullptr) != (al == nullptr)))
    return set_err(LMPC_E_INVALID, "lmpc_gpu_solve_batch: bad argument%s");
  if (N == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  if (is_device_ptr(Theta)) {
    int rc = lmpc_gpu_solve_batch_async(h, Theta, N, U, exitflag, iters, au, al, h->stream);
    if (rc) return rc;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  }
  const double *seg[1] = {Theta};
  int len[1] = {h->lay.nth}, stride[1] = {h->lay.nth};
  return solve_segments_host(h, seg, len, stride, 1, N, U, h->lay.nu, exitflag, iters, au, al);
}

int lmpc_gpu_solve_batch_warm(lmpc_handle *h, const double *Theta, int64_t N, double *U,
                              int32_t *exitflag, int32_t *iters, uint64_t *au, uint64_t *al) {
  if (!h || !Theta || !U || !exitflag || !au || !al || N < 0)
    return set_err(LMPC_E_INVALID, "lmpc_gpu_solve_batch_warm: bad argument%s");
  if (N == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  if (is_device_ptr(Theta)) {
    ThetaSegs th;
    memset(&th, 0, sizeof(th));
    th.nseg = 1; th.ptr[0] = Theta; th.len[0] = h->lay.nth; th.stride[0] = h->lay.nth;
    SolveOut o{U, exitflag, iters, au, al, au, al};
    int rc = run_kernel(h, th, N, o, nullptr, h->stream);
    if (rc) return rc;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  }
  const double *seg[1] = {Theta};
  int len[1] = {h->lay.nth}, stride[1] = {h->lay.nth};
  return solve_segments_host(h, seg, len, stride, 1, N, U, h->lay.nu, exitflag, iters, au, al, true);
}

int lmpc_gpu_compute_control_batch(lmpc_handle *h, double *control, const double *state,
                                   const double *reference, const double *disturbance,
                                   int32_t nr, int32_t nd, int32_t nuprev, int64_t N,
                                   int32_t *exitflag) {
  if (!h || !control || !state || !exitflag || N < 0 || nr < 0 || nd < 0 || nuprev < 0)
    return set_err(LMPC_E_INVALID, "lmpc_gpu_compute_control_batch: bad argument%s");
  const LdpLayout &l = h->lay;
  const int nx = l.nth - nr - nd - nuprev;
  if (nx <= 0 || nuprev > l.nu || (nr > 0 && !reference) )
    return set_err(LMPC_E_INVALID, "compute_control_batch: nx + nr + nd + nuprev must equal nth%s");
  if (N == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  const double *seg[4]; int len[4], stride[4], ns = 0;
  seg[ns] = state; len[ns] = nx; stride[ns] = nx; ns++;
  if (nr > 0) { seg[ns] = reference; len[ns] = nr; stride[ns] = nr; ns++; }
  if (nd > 0) {
    if (!disturbance) return set_err(LMPC_E_INVALID, "disturbance is NULL but nd > 0%s");
    seg[ns] = disturbance; len[ns] = nd; stride[ns] = nd; ns++;
  }
  if (nuprev > 0) { seg[ns] = control; len[ns] = nuprev; stride[ns] = l.nu; ns++; }
  if (is_device_ptr(state)) {
    // in/out aliasing of `control`: each thread reads its own u_prev before it writes u
    ThetaSegs th;
    memset(&th, 0, sizeof(th));
    th.nseg = ns;
    for (int g = 0; g < ns; ++g) { th.ptr[g] = seg[g]; th.len[g] = len[g]; th.stride[g] = stride[g]; }
    SolveOut o{control, exitflag, nullptr, nullptr, nullptr};
    int rc = run_kernel(h, th, N, o, nullptr, h->stream);
    if (rc) return rc;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  }
  return solve_segments_host(h, seg, len, stride, ns, N, control, l.nu, exitflag, nullptr, nullptr, nullptr);
}

int lmpc_gpu_compute_control_batch_bcast(lmpc_handle *h, double *control, const double *state,
                                         const double *reference, const double *disturbance,
                                         const double *uprev, int32_t nr, int32_t nd, int32_t nuprev,
                                         int64_t N, int32_t bcast, int32_t *exitflag, int64_t *n_failed) {
  if (!h || !control || !state || N < 0 || nr < 0 || nd < 0 || nuprev < 0 || (!exitflag && !n_failed))
    return set_err(LMPC_E_INVALID, "lmpc_gpu_compute_control_batch_bcast: bad argument%s");
  const LdpLayout &l = h->lay;
  const int nx = l.nth - nr - nd - nuprev;
  if (nx <= 0 || nuprev > l.nu || (nr > 0 && !reference) || (nd > 0 && !disturbance) || (nuprev > 0 && !uprev))
    return set_err(LMPC_E_INVALID, "compute_control_batch_bcast: nx + nr + nd + nuprev must equal nth, no NULL block%s");
  if (n_failed) *n_failed = 0;
  if (N == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  const double *seg[4]; int len[4], stride[4], ns = 0;
  seg[ns] = state; len[ns] = nx; stride[ns] = nx; ns++;
  if (nr > 0) { seg[ns] = reference; len[ns] = nr; stride[ns] = (bcast & LMPC_BCAST_REFERENCE) ? 0 : nr; ns++; }
  if (nd > 0) { seg[ns] = disturbance; len[ns] = nd; stride[ns] = (bcast & LMPC_BCAST_DISTURBANCE) ? 0 : nd; ns++; }
  if (nuprev > 0) { seg[ns] = uprev; len[ns] = nuprev; stride[ns] = (bcast & LMPC_BCAST_UPREV) ? 0 : nuprev; ns++; }
  if (is_device_ptr(state)) {
    ThetaSegs th;
    memset(&th, 0, sizeof(th));
    th.nseg = ns;
    for (int g = 0; g < ns; ++g) { th.ptr[g] = seg[g]; th.len[g] = len[g]; th.stride[g] = stride[g]; }
    std::lock_guard<std::mutex> lock(h->mu);
    int32_t *fl = exitflag;
    if (!fl) {        // flags stay on the device: scratch in slot 0
      if (h->slots[0].flag.reserve((size_t)N * 4)) return LMPC_E_CUDA;
      fl = (int32_t *)h->slots[0].flag.p;
    }
    SolveOut o{control, fl, nullptr, nullptr, nullptr};
    int rc = run_kernel(h, th, N, o, nullptr, h->stream);
    if (rc) return rc;
    if (n_failed) {
      if (h->counter.reserve(8)) return LMPC_E_CUDA;
      CUDA_OK(cudaMemsetAsync(h->counter.p, 0, 8, h->stream));
      int e = launch_count_failed(fl, N, (unsigned long long *)h->counter.p, h->stream);
      if (e) return set_err(LMPC_E_CUDA, "count kernel launch: %s", cudaGetErrorString((cudaError_t)e));
      g_launches.fetch_add(1);
      unsigned long long cnt = 0;
      CUDA_OK(cudaMemcpyAsync(&cnt, h->counter.p, 8, cudaMemcpyDeviceToHost, h->stream));
      CUDA_OK(cudaStreamSynchronize(h->stream));
      *n_failed = (int64_t)cnt;
      return 0;
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  }
  return solve_segments_host(h, seg, len, stride, ns, N, control, l.nu, exitflag, nullptr, nullptr, nullptr, false,
                             n_failed);
}

int lmpc_gpu_simulate_batch(lmpc_handle *h, const double *F, const double *G, const double *Gd,
                            int32_t nx, int32_t nr, int32_t nd, int32_t nuprev,
                            const double *X0, const double *R, const double *D, const double *U0,
                            int64_t N, int32_t T, int32_t warm_start,
                            double *Xend, double *Ulast, double *Utraj,
                            int32_t *exitflag, int32_t *iters) {
  return lmpc_gpu_simulate_batch_traj(h, F, G, Gd, nx, nr, nd, nuprev, X0, R, D, U0, nullptr, nullptr,
                                      N, T, warm_start, Xend, Ulast, Utraj, exitflag, iters);
}

int lmpc_gpu_simulate_batch_traj(lmpc_handle *h, const double *F, const double *G, const double *Gd,
                                 int32_t nx, int32_t nr, int32_t nd, int32_t nuprev,
                                 const double *X0, const double *R, const double *D, const double *U0,
                                 const double *Rtraj, const double *Dtraj,
                                 int64_t N, int32_t T, int32_t warm_start,
                                 double *Xend, double *Ulast, double *Utraj,
                                 int32_t *exitflag, int32_t *iters) {
  lmpc_sim_ext e;
  memset(&e, 0, sizeof(e));
  e.Rtraj = Rtraj; e.Dtraj = Dtraj;
  return lmpc_gpu_simulate_batch_ex(h, F, G, Gd, nx, nr, nd, nuprev, X0, R, D, U0, &e, N, T, warm_start, Xend,
                                    Ulast, Utraj, exitflag, iters);
}

// column-major (rows x cols) host matrix -> row-major, appended to v; returns the offset
static size_t push_rowmajor(std::vector<double> &v, const double *A, int rows, int cols) {
  const size_t off = v.size();
  v.resize(off + (size_t)rows * cols);
  for (int i = 0; i < rows; ++i)
    for (int c = 0; c < cols; ++c) v[off + (size_t)i * cols + c] = A[(size_t)c * rows + i];
  return off;
}

int lmpc_gpu_simulate_batch_ex(lmpc_handle *h, const double *F, const double *G, const double *Gd,
                               int32_t nx, int32_t nr, int32_t nd, int32_t nuprev,
                               const double *X0, const double *R, const double *D, const double *U0,
                               const lmpc_sim_ext *ext,
                               int64_t N, int32_t T, int32_t warm_start,
                               double *Xend, double *Ulast, double *Utraj,
                               int32_t *exitflag, int32_t *iters) {
  if (!h || !F || !G || !X0 || !Xend || !Ulast || !exitflag || N < 0 || T < 0)
    return set_err(LMPC_E_INVALID, "lmpc_gpu_simulate_batch: bad argument%s");
  const LdpLayout &l = h->lay;
  if (nx + nr + nd + nuprev != l.nth || nuprev > l.nu || (nd > 0 && !Gd))
    return set_err(LMPC_E_INVALID, "simulate_batch: nx + nr + nd + nuprev must equal nth%s");
  const double *Rtraj = ext ? ext->Rtraj : nullptr, *Dtraj = ext ? ext->Dtraj : nullptr;
  const bool observer = ext && ext->C && ext->K;
  const bool extended = ext && (ext->Ft || ext->Gt || ext->Gdt || observer || ext->Xhat_end);
  if (ext && (ext->C != nullptr) != (ext->K != nullptr))
    return set_err(LMPC_E_INVALID, "simulate_batch_ex: an observer needs both C and K%s");
  if (observer && ext->ny <= 0) return set_err(LMPC_E_INVALID, "simulate_batch_ex: observer with ny <= 0%s");
  if (N == 0) return 0;
  std::lock_guard<std::mutex> lock(h->mu);
  CUDA_OK(cudaSetDevice(h->device));
  SlotGuard guard(h->slots);
  const int nu = l.nu, ny = observer ? ext->ny : 0;
  // everything the kernels read besides the per-trajectory inputs, row-major, in one device buffer
  std::vector<double> dynh;
  const size_t oF = push_rowmajor(dynh, F, nx, nx), oG = push_rowmajor(dynh, G, nx, nu);
  const size_t oGd = nd > 0 ? push_rowmajor(dynh, Gd, nx, nd) : dynh.size();
  size_t oR = 0, oD = 0, oFt = oF, oGt = oG, oGdt = oGd, oC = 0, oDd = 0, oK = 0, oV = 0;
  const int rstride = (ext && ext->rstride > 0) ? ext->rstride : nr;
  const int dstride = (ext && ext->dstride > 0) ? ext->dstride : nd;
  // a windowed trajectory is read up to (T - 1) * stride + n entries
  if (Rtraj && nr > 0) { oR = dynh.size(); dynh.insert(dynh.end(), Rtraj, Rtraj + ((size_t)(T > 0 ? T - 1 : 0) * rstride + nr)); }
  if (Dtraj && nd > 0) { oD = dynh.size(); dynh.insert(dynh.end(), Dtraj, Dtraj + ((size_t)(T > 0 ? T - 1 : 0) * dstride + nd)); }
  if (extended) {
    if (ext->Ft) oFt = push_rowmajor(dynh, ext->Ft, nx, nx);
    if (ext->Gt) oGt = push_rowmajor(dynh, ext->Gt, nx, nu);
    if (ext->Gdt && nd > 0) oGdt = push_rowmajor(dynh, ext->Gdt, nx, nd);
    if (observer) {
      oC = push_rowmajor(dynh, ext->C, ny, nx);
      if (ext->Dd && nd > 0) oDd = push_rowmajor(dynh, ext->Dd, ny, nd);
      oK = push_rowmajor(dynh, ext->K, nx, ny);
      if (ext->Vtraj) { oV = dynh.size(); dynh.insert(dynh.end(), ext->Vtraj, ext->Vtraj + (size_t)ny * T); }
    }
  }
  if (h->dyn.reserve(dynh.size() * sizeof(double))) return LMPC_E_CUDA;
  CUDA_OK(cudaMemcpyAsync(h->dyn.p, dynh.data(), dynh.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  const double *db = (const double *)h->dyn.p;
  SimDesc sd;
  memset(&sd, 0, sizeof(sd));
  sd.F = db + oF; sd.G = db + oG; sd.Gd = db + oGd;
  sd.nx = nx; sd.nr = nr; sd.nd = nd; sd.nuprev = nuprev; sd.T = T; sd.warm = warm_start;
  sd.Rtraj = (Rtraj && nr > 0) ? db + oR : nullptr;
  sd.Dtraj = (Dtraj && nd > 0) ? db + oD : nullptr;
  sd.rstride = rstride; sd.dstride = dstride;
  if (extended) {
    sd.ext = 1; sd.ny = ny;
    sd.Ft = db + oFt; sd.Gt = db + oGt; sd.Gdt = db + oGdt;
    if (observer) {
      sd.Cm = db + oC; sd.Kobs = db + oK;
      sd.Ddm = (ext->Dd && nd > 0) ? db + oDd : nullptr;
      sd.Vtraj = ext->Vtraj ? db + oV : nullptr;
    }
  }
  double *Xhat = extended ? ext->Xhat_end : nullptr;
  ThetaSegs th;
  memset(&th, 0, sizeof(th));
  if (is_device_ptr(X0)) {
    sd.X0 = X0; sd.R = R; sd.D = D; sd.U0 = U0; sd.Xend = Xend; sd.Ulast = Ulast; sd.Utraj = Utraj; sd.Xhat = Xhat;
    SolveOut o{nullptr, exitflag, iters, nullptr, nullptr};
    int rc = run_kernel(h, th, N, o, &sd, h->stream);
    if (rc) return rc;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
  }
  for (int64_t off = 0, c = 0; off < N; off += kChunk, ++c) {
    const int64_t cnt = (N - off < kChunk) ? (N - off) : kChunk;
    Slot &s = h->slots[c % kSlots];
    const double *src[4] = {X0, R, D, U0};
    const int ln[4] = {nx, nr, nd, nuprev};
    const double *dv[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int g = 0; g < 4; ++g) {
      if (!src[g] || ln[g] == 0) continue;
      const size_t bytes = (size_t)cnt * ln[g] * sizeof(double);
      if (s.in[g].reserve(bytes)) return LMPC_E_CUDA;
      CUDA_OK(cudaMemcpyAsync(s.in[g].p, src[g] + off * ln[g], bytes, cudaMemcpyHostToDevice, s.stream));
      dv[g] = (const double *)s.in[g].p;
    }
    if (s.aux[0].reserve((size_t)cnt * nx * 8) || s.u.reserve((size_t)cnt * nu * 8) ||
        s.flag.reserve((size_t)cnt * 4) || s.iters.reserve((size_t)cnt * 4)) return LMPC_E_CUDA;
    if (Utraj && s.aux[1].reserve((size_t)cnt * nu * T * 8)) return LMPC_E_CUDA;
    if (Xhat && s.aux[2].reserve((size_t)cnt * nx * 8)) return LMPC_E_CUDA;
    SimDesc d = sd;
    d.X0 = dv[0]; d.R = dv[1]; d.D = dv[2]; d.U0 = dv[3];
    d.Xend = (double *)s.aux[0].p; d.Ulast = (double *)s.u.p; d.Utraj = Utraj ? (double *)s.aux[1].p : nullptr;
    d.Xhat = Xhat ? (double *)s.aux[2].p : nullptr;
    SolveOut o{nullptr, (int32_t *)s.flag.p, (int32_t *)s.iters.p, nullptr, nullptr};
    int rc = run_kernel(h, th, cnt, o, &d, s.stream);
    if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(Xend + off * nx, d.Xend, (size_t)cnt * nx * 8, cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaMemcpyAsync(Ulast + off * nu, d.Ulast, (size_t)cnt * nu * 8, cudaMemcpyDeviceToHost, s.stream));
    if (Utraj) CUDA_OK(cudaMemcpyAsync(Utraj + off * nu * T, d.Utraj, (size_t)cnt * nu * T * 8, cudaMemcpyDeviceToHost, s.stream));
    if (Xhat) CUDA_OK(cudaMemcpyAsync(Xhat + off * nx, d.Xhat, (size_t)cnt * nx * 8, cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaMemcpyAsync(exitflag + off, s.flag.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.stream));
    if (iters) CUDA_OK(cudaMemcpyAsync(iters + off, s.iters.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.stream));
  }
  for (int i = 0; i < kSlots; ++i) CUDA_OK(cudaStreamSynchronize(h->slots[i].stream));
  guard.armed = false;
  return 0;
}

int lmpc_gpu_simulate_batch_stepped(lmpc_handle *h, const double *F, const double *G, const double *Gd,
                                    int32_t nx, int32_t nr, int32_t nuprev,
                                    const double *X0, const double *R, const double *D, const double *U0,
                                    const lmpc_sim_stepped *cfg, int64_t N, int32_t T, int32_t warm_start,
                                    double *Xend, double *Ulast, double *Utraj, int32_t *exitflag, int32_t *iters) {
  if (!h || !F || !G || !X0 || !Xend || !Ulast || !exitflag || !cfg || N < 0 || T < 0)
    return set_err(LMPC_E_INVALID, "lmpc_gpu_simulate_batch_stepped: bad argument%s");
  const LdpLayout &l = h->lay;
  const bool obs = cfg->Ko != nullptr;
  const int nu = l.nu, ndm = cfg->ndm, nxo = obs ? cfg->nxo : nx, ny = obs ? cfg->ny : 0, ndo = nxo - nx;
  if (nxo < nx || nx + nr + ndm + ndo + nuprev != l.nth || nuprev > nu || (ndm > 0 && !Gd && !cfg->Gdt && !cfg->dynamics))
    return set_err(LMPC_E_INVALID, "simulate_batch_stepped: nx + nr + ndm + (nxo - nx) + nuprev must equal nth%s");
  if (obs && (!cfg->Co || !cfg->Cm || !cfg->Fo || !cfg->Go || ny <= 0))
    return set_err(LMPC_E_INVALID, "simulate_batch_stepped: an observer needs Cm, Co, Fo, Go, Ko and ny > 0%s");
  if (is_device_ptr(X0)) return set_err(LMPC_E_UNSUPPORTED, "simulate_batch_stepped takes host arrays%s");
  if (N == 0 || T == 0) return 0;
  std::lock_guard<std::mutex> lock(h->mu);
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  // constants, row-major, one upload
  std::vector<double> ch;
  auto mat = [&](const double *A, int r, int c) -> long { return (A && r > 0 && c > 0) ? (long)push_rowmajor(ch, A, r, c) : -1; };
  auto vec = [&](const double *v, size_t n) -> long { if (!v || n == 0) return -1; const long o = (long)ch.size(); ch.insert(ch.end(), v, v + n); return o; };
  const long oFt = mat(cfg->Ft ? cfg->Ft : F, nx, nx), oGt = mat(cfg->Gt ? cfg->Gt : G, nx, nu);
  const long oGdt = mat(cfg->Gdt ? cfg->Gdt : Gd, nx, ndm), oft = vec(cfg->ft, nx);
  const long oCm = mat(cfg->Cm, ny, nx), oDdm = mat(cfg->Ddm, ny, ndm), ohm = vec(cfg->hm, ny), oV = vec(cfg->Vtraj, (size_t)ny * T);
  const long oFo = obs ? mat(cfg->Fo, nxo, nxo) : -1, oGo = obs ? mat(cfg->Go, nxo, nu) : -1, oGdo = obs ? mat(cfg->Gdo, nxo, ndm) : -1;
  const long ofo = obs ? vec(cfg->fo, nxo) : -1, oCo = obs ? mat(cfg->Co, ny, nxo) : -1, oDdo = obs ? mat(cfg->Ddo, ny, ndm) : -1;
  const long oho = obs ? vec(cfg->ho, ny) : -1, oKo = obs ? mat(cfg->Ko, nxo, ny) : -1;
  const long oRt = vec(cfg->Rtraj, (size_t)nr * T), oDt = vec(cfg->Dtraj, (size_t)ndm * T);
  if (ch.empty()) ch.push_back(0.0);
  if (h->dyn.reserve(ch.size() * sizeof(double))) return LMPC_E_CUDA;
  CUDA_OK(cudaMemcpyAsync(h->dyn.p, ch.data(), ch.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  const double *cb = (const double *)h->dyn.p;
  auto P = [&](long o) -> const double * { return o < 0 ? nullptr : cb + o; };
  // loop state, one slot's buffers (a closed loop is sequential in time: no chunk pipeline here)
  Slot &s = h->slots[0];
  const int words = (l.m + 63) / 64;
  DevBuf xs, xn, xh, up, dm, theta, wsu, wsl, fstep, istep;
  struct Rel { DevBuf *b[10]; ~Rel() { for (DevBuf *x : b) x->release(); } } rel{{&xs, &xn, &xh, &up, &dm, &theta, &wsu, &wsl, &fstep, &istep}};
  if (xs.reserve((size_t)N * nx * 8) || xn.reserve((size_t)N * nx * 8) || xh.reserve((size_t)N * nxo * 8) ||
      up.reserve((size_t)N * nu * 8) || dm.reserve((size_t)N * (ndm > 0 ? ndm : 1) * 8) || theta.reserve((size_t)N * l.nth * 8) ||
      wsu.reserve((size_t)N * words * 8) || wsl.reserve((size_t)N * words * 8) || fstep.reserve((size_t)N * 4) ||
      istep.reserve((size_t)N * 4) || s.u.reserve((size_t)N * nu * 8) || s.flag.reserve((size_t)N * 4) ||
      s.iters.reserve((size_t)N * 4))
    return LMPC_E_CUDA;
  if (nr > 0 && R && s.in[1].reserve((size_t)N * nr * 8)) return LMPC_E_CUDA;
  if (ndm > 0 && D && s.in[2].reserve((size_t)N * ndm * 8)) return LMPC_E_CUDA;
  if (Utraj && s.aux[1].reserve((size_t)N * nu * T * 8)) return LMPC_E_CUDA;
  CUDA_OK(cudaMemcpyAsync(xs.p, X0, (size_t)N * nx * 8, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemsetAsync(xh.p, 0, (size_t)N * nxo * 8, st));
  CUDA_OK(cudaMemcpy2DAsync(xh.p, (size_t)nxo * 8, X0, (size_t)nx * 8, (size_t)nx * 8, (size_t)N, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemsetAsync(up.p, 0, (size_t)N * nu * 8, st));
  if (U0 && nuprev > 0)
    CUDA_OK(cudaMemcpy2DAsync(up.p, (size_t)nu * 8, U0, (size_t)nuprev * 8, (size_t)nuprev * 8, (size_t)N, cudaMemcpyHostToDevice, st));
  if (nr > 0 && R) CUDA_OK(cudaMemcpyAsync(s.in[1].p, R, (size_t)N * nr * 8, cudaMemcpyHostToDevice, st));
  if (ndm > 0 && D) CUDA_OK(cudaMemcpyAsync(s.in[2].p, D, (size_t)N * ndm * 8, cudaMemcpyHostToDevice, st));
  StepDesc d;
  memset(&d, 0, sizeof(d));
  d.Ft = P(oFt); d.Gt = P(oGt); d.Gdt = P(oGdt); d.ft = P(oft);
  d.Cm = P(oCm); d.Ddm = P(oDdm); d.hm = P(ohm); d.Vtraj = P(oV);
  d.Fo = P(oFo); d.Go = P(oGo); d.Gdo = P(oGdo); d.fo = P(ofo); d.Co = P(oCo); d.Ddo = P(oDdo); d.ho = P(oho); d.Ko = P(oKo);
  d.Rtraj = P(oRt); d.Dtraj = P(oDt);
  d.R = (nr > 0 && R) ? (const double *)s.in[1].p : nullptr;
  d.D = (ndm > 0 && D) ? (const double *)s.in[2].p : nullptr;
  d.nx = nx; d.nu = nu; d.nr = nr; d.ndm = ndm; d.nxo = nxo; d.ny = ny; d.nuprev = nuprev; d.nth = l.nth; d.T = T;
  d.has_obs = obs; d.user_plant = cfg->dynamics != nullptr;
  d.Xs = (double *)xs.p; d.Xn = (double *)xn.p; d.Xh = (double *)xh.p; d.Uprev = (double *)up.p;
  d.Theta = (double *)theta.p; d.U = (double *)s.u.p; d.Dm = (double *)dm.p;
  d.Utraj = Utraj ? (double *)s.aux[1].p : nullptr;
  d.flag_step = (int32_t *)fstep.p; d.iters_step = (int32_t *)istep.p;
  d.flag_worst = (int32_t *)s.flag.p; d.iters_tot = (int32_t *)s.iters.p;
  if (!step_dims_ok(d)) return set_err(LMPC_E_UNSUPPORTED, "simulate_batch_stepped: dimension beyond the step kernels' limits%s");
  ThetaSegs th;
  memset(&th, 0, sizeof(th));
  th.nseg = 1; th.ptr[0] = d.Theta; th.len[0] = l.nth; th.stride[0] = l.nth;
  for (int t = 0; t < T; ++t) {
    int e = launch_step_pre(d, t, N, st);
    if (e) return set_err(LMPC_E_CUDA, "step kernel: %s", cudaGetErrorString((cudaError_t)e));
    const bool warm = warm_start && t > 0;
    SolveOut o{d.U, d.flag_step, d.iters_step, (uint64_t *)wsu.p, (uint64_t *)wsl.p,
               warm ? (const uint64_t *)wsu.p : nullptr, warm ? (const uint64_t *)wsl.p : nullptr};
    int rc = run_kernel(h, th, N, o, nullptr, st);
    if (rc) return rc;
    e = launch_step_post(d, t, N, st);
    if (e) return set_err(LMPC_E_CUDA, "step kernel: %s", cudaGetErrorString((cudaError_t)e));
    if (cfg->dynamics) {
      rc = cfg->dynamics(cfg->user, (void *)st, N, t, d.Xs, d.U, ndm > 0 ? d.Dm : nullptr, d.Xn);
      if (rc) { cudaStreamSynchronize(st); return set_err(LMPC_E_INVALID, "simulate_batch_stepped: the plant callback failed%s"); }
    }
    std::swap(d.Xs, d.Xn);
  }
  CUDA_OK(cudaMemcpyAsync(Xend, d.Xs, (size_t)N * nx * 8, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(Ulast, d.Uprev, (size_t)N * nu * 8, cudaMemcpyDeviceToHost, st));
  if (Utraj) CUDA_OK(cudaMemcpyAsync(Utraj, d.Utraj, (size_t)N * nu * T * 8, cudaMemcpyDeviceToHost, st));
  if (cfg->Xhat_end) CUDA_OK(cudaMemcpyAsync(cfg->Xhat_end, d.Xh, (size_t)N * nxo * 8, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(exitflag, d.flag_worst, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
  if (iters) CUDA_OK(cudaMemcpyAsync(iters, d.iters_tot, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------------------
// explicit MPC
// ---------------------------------------------------------------------------------------
int lmpc_gpu_create_bst_oriented(const double *halfplanes, int32_t n_hp, const double *feedbacks,
                                 int32_t n_fb, const int32_t *hp_list, const int32_t *jump_list,
                                 int32_t n_nodes, const double *clipping, int32_t nth, int32_t nu,
                                 int32_t le_child, int device, lmpc_bst **out) {
  if (!hp_list || !jump_list || n_nodes <= 0 || (le_child != 0 && le_child != 1))
    return set_err(LMPC_E_INVALID, "lmpc_gpu_create_bst_oriented: bad argument%s");
  if (le_child == 0)
    return lmpc_gpu_create_bst(halfplanes, n_hp, feedbacks, n_fb, hp_list, jump_list, n_nodes, clipping, nth, nu,
                               device, out);
  // The producer walks to next + 1 when a'theta <= b.  Normalise to this library's convention (<= goes to
  // `next`) by exchanging the two children of every internal node: they are adjacent entries, so the
  // exchange is a swap of (hp, jump) with the relative jumps corrected by the one slot each entry moved.
  std::vector<int32_t> hp(hp_list, hp_list + n_nodes), jp(jump_list, jump_list + n_nodes);
  std::vector<char> seen(n_nodes, 0);
  for (int i = 0; i < n_nodes; ++i) {
    const int j = jump_list[i];
    if (j == 0) continue;
    const int c0 = i + j, c1 = c0 + 1;
    if (j < 0 || c1 >= n_nodes || seen[c0] || seen[c1])
      return set_err(LMPC_E_INVALID, "bst: not a tree (child out of range or shared by two parents)%s");
    seen[c0] = seen[c1] = 1;
    if (jump_list[c0] == 1) return set_err(LMPC_E_INVALID, "bst: a node's children overlap its sibling%s");
    hp[c0] = hp_list[c1]; jp[c0] = jump_list[c1] ? jump_list[c1] + 1 : 0;
    hp[c1] = hp_list[c0]; jp[c1] = jump_list[c0] ? jump_list[c0] - 1 : 0;
  }
  return lmpc_gpu_create_bst(halfplanes, n_hp, feedbacks, n_fb, hp.data(), jp.data(), n_nodes, clipping, nth, nu,
                             device, out);
}

int lmpc_gpu_create_bst(const double *halfplanes, int32_t n_hp, const double *feedbacks,
                        int32_t n_fb, const int32_t *hp_list, const int32_t *jump_list,
                        int32_t n_nodes, const double *clipping, int32_t nth, int32_t nu,
                        int device, lmpc_bst **out) {
  if (!halfplanes || !feedbacks || !hp_list || !jump_list || !out || n_hp < 0 || n_fb <= 0 ||
      n_nodes <= 0 || nth <= 0 || nu <= 0)
    return set_err(LMPC_E_INVALID, "lmpc_gpu_create_bst: bad argument%s");
  *out = nullptr;
  // validate the walk: every internal node must jump forward inside the array and name a
  // valid half-plane; every leaf must name a valid feedback
  for (int i = 0; i < n_nodes; ++i) {
    const int j = jump_list[i];
    if (j == 0) { if (hp_list[i] < 0 || hp_list[i] >= n_fb) return set_err(LMPC_E_INVALID, "bst: leaf names a missing feedback%s"); }
    else if (j < 0 || i + j + 1 >= n_nodes || hp_list[i] < 0 || hp_list[i] >= n_hp)
      return set_err(LMPC_E_INVALID, "bst: internal node out of range%s");
  }
  int sm = 0;
  int rc = check_device(device, &sm);
  if (rc) return rc;
  const int ld = nth + 1;
  const int pairs = (ld + 1) / 2;          // 16-byte pairs per row [a | b | pad] / [F | g | pad]
  auto al2 = [](size_t x) { return (x + 1) & ~(size_t)1; };
  // node records fit one 32-bit word when every jump and index is below 65536
  bool packed = n_hp < 65536 && n_fb < 65536;
  for (int i = 0; i < n_nodes && packed; ++i) packed = jump_list[i] < 65536;
  const size_t node_d = packed ? al2(((size_t)n_nodes + 1) / 2) : al2((size_t)n_nodes);
  // replicate the half-plane and leaf-law tables so that lane l reads replica l mod rep
  // (conflict-free at 8); as many replicas (power of two <= 8) as fit in ~200 KB of shared memory
  const size_t rows = (size_t)n_hp + (size_t)n_fb * nu;
  // fp32 screening rows [a | b] (bst_eval.cu): 16-byte quads of four floats, nth <= 16 only
  const int quads = (nth <= 16) ? (ld + 3) / 4 : 0;
  int rep = 8;
  // what has to sit in shared memory: with the screen, the fp32 rows + the laws; without it, the fp64 rows + the laws
  auto smem_doubles = [&](int r) {
    return quads ? node_d + ((size_t)n_fb * nu * pairs * 2 + (size_t)n_hp * quads * 2) * r + 2 * nu + 2
                 : node_d + rows * pairs * 2 * r + 2 * nu;
  };
  // (the closed-loop kernel stages the whole block, fp64 rows included: keep that within reach too when possible)
  auto all_doubles = [&](int r) { return smem_doubles(r) + (quads ? (size_t)n_hp * pairs * 2 * r : 0); };
  while (rep > 1 && (smem_doubles(rep) * sizeof(double) > 160 * 1024 || all_doubles(rep) * sizeof(double) > 210 * 1024)) rep >>= 1;
  const size_t onodes = 0, ofb = node_d, oclip = ofb + (size_t)n_fb * nu * pairs * 2 * rep;
  const size_t ohp32 = oclip + al2((size_t)2 * nu), oscr = ohp32 + (size_t)n_hp * quads * 2 * rep;
  const size_t ohp = oscr + (quads ? 2 : 0), total = ohp + (size_t)n_hp * pairs * 2 * rep;
  std::vector<double> blk(total, 0.0);
  if (packed) {
    uint32_t *nodes = reinterpret_cast<uint32_t *>(blk.data() + onodes);
    for (int i = 0; i < n_nodes; ++i) nodes[i] = ((uint32_t)jump_list[i] << 16) | (uint32_t)hp_list[i];
  } else {
    int32_t *nodes = reinterpret_cast<int32_t *>(blk.data() + onodes);
    for (int i = 0; i < n_nodes; ++i) { nodes[2 * i] = jump_list[i]; nodes[2 * i + 1] = hp_list[i]; }
  }
  for (int hI = 0; hI < n_hp; ++hI)      // column-major (nth+1) x n_hp: column hI = [a; b]
    for (int c = 0; c < ld; ++c)
      for (int p2 = 0; p2 < rep; ++p2)
        blk[ohp + (((size_t)hI * pairs + c / 2) * rep + p2) * 2 + (c & 1)] = halfplanes[(size_t)hI * ld + c];
  for (int r = 0; r < n_fb; ++r)          // block r: (nth+1) x nu column-major; row (r, i) = [F_i | g_i]
    for (int i = 0; i < nu; ++i)
      for (int c = 0; c < ld; ++c)
        for (int p2 = 0; p2 < rep; ++p2)
          blk[ofb + ((((size_t)r * nu + i) * pairs + c / 2) * rep + p2) * 2 + (c & 1)] =
              feedbacks[(size_t)r * ld * nu + (size_t)i * ld + c];
  if (clipping)
    for (int i = 0; i < nu; ++i) { blk[oclip + 2 * i] = clipping[i]; blk[oclip + 2 * i + 1] = clipping[nu + i]; }
  if (quads) {
    float *h32 = reinterpret_cast<float *>(blk.data() + ohp32);
    double a1max = 0.0, bmax = 0.0;
    for (int hI = 0; hI < n_hp; ++hI) {
      double a1 = 0.0;
      for (int c = 0; c < ld; ++c) {
        const double v = halfplanes[(size_t)hI * ld + c];
        if (c < nth) a1 += std::fabs(v); else bmax = std::max(bmax, std::fabs(v));
        for (int p2 = 0; p2 < rep; ++p2) h32[(((size_t)hI * quads + c / 4) * rep + p2) * 4 + (c & 3)] = (float)v;
      }
      a1max = std::max(a1max, a1);
    }
    // bound constants of the screen: |d32 - (a'theta - b)| <= 2^-18 (a1max * max|theta_i| + bmax), see bst_eval.cu
    blk[oscr] = a1max * (1.0 + 1e-6); blk[oscr + 1] = bmax * (1.0 + 1e-6);
  }
  lmpc_bst *t = new lmpc_bst();
  t->device = device; t->sm_count = sm;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc((void **)&t->d_blk, total * sizeof(double)) != cudaSuccess ||
      cudaMemcpy(t->d_blk, blk.data(), total * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    const char *es = cudaGetErrorString(cudaGetLastError());
    lmpc_gpu_destroy_bst(t);
    return set_err(LMPC_E_CUDA, "lmpc_gpu_create_bst: %s", es);
  }
  for (int i = 0; i < kSlots; ++i)
    if (cudaStreamCreateWithFlags(&t->slots[i].stream, cudaStreamNonBlocking) != cudaSuccess) {
      lmpc_gpu_destroy_bst(t);
      return set_err(LMPC_E_CUDA, "cudaStreamCreate failed%s");
    }
  t->dev.blk = t->d_blk;
  t->dev.onodes = (int)onodes; t->dev.ohp = (int)ohp; t->dev.ofb = (int)ofb; t->dev.oclip = (int)oclip;
  t->dev.pairs = pairs; t->dev.rep = rep; t->dev.packed = packed ? 1 : 0;
  t->dev.total = (int)total; t->dev.has_clip = clipping ? 1 : 0;
  t->dev.ohp32 = (int)ohp32; t->dev.quads = quads; t->dev.oscr = (int)oscr; t->dev.staged = (int)(quads ? ohp : total);
  t->dev.nth = nth; t->dev.nu = nu; t->dev.n_hp = n_hp; t->dev.n_fb = n_fb; t->dev.n_nodes = n_nodes;
  *out = t;
  return 0;
}

int lmpc_gpu_destroy_bst(lmpc_bst *t) {
  if (!t) return 0;
  cudaSetDevice(t->device);
  free_slots(t->slots);
  if (t->d_blk) cudaFree(t->d_blk);
  if (t->stream) cudaStreamDestroy(t->stream);
  delete t;
  return 0;
}

int lmpc_gpu_bst_kernel(const lmpc_bst *t) { return t ? bst_eval_kind(t->dev) : -1; }

int lmpc_gpu_evaluate_batch_async(lmpc_bst *t, const double *Theta, int64_t N, double *U,
                                  int32_t *region, void *stream) {
  if (!t || !Theta || !U || N < 0) return set_err(LMPC_E_INVALID, "lmpc_gpu_evaluate_batch_async: bad argument%s");
  CUDA_OK(cudaSetDevice(t->device));
  int e = launch_bst_eval(t->dev, Theta, N, U, region, t->sm_count, stream ? stream : (void *)t->stream);
  if (e) return set_err(LMPC_E_CUDA, "bst kernel launch: %s", cudaGetErrorString((cudaError_t)e));
  g_launches.fetch_add(1);
  return 0;
}

int lmpc_gpu_evaluate_batch(lmpc_bst *t, const double *Theta, int64_t N, double *U, int32_t *region) {
  if (!t || !Theta || !U || N < 0) return set_err(LMPC_E_INVALID, "lmpc_gpu_evaluate_batch: bad argument%s");
  if (N == 0) return 0;
  CUDA_OK(cudaSetDevice(t->device));
  if (is_device_ptr(Theta)) {
    int rc = lmpc_gpu_evaluate_batch_async(t, Theta, N, U, region, t->stream);
    if (rc) return rc;
    CUDA_OK(cudaStreamSynchronize(t->stream));
    return 0;
  }
  std::lock_guard<std::mutex> lock(t->mu);
  const int nth = t->dev.nth, nu = t->dev.nu;
  // pageable caller arrays go through the pinned ring, as in solve_segments_host
  const bool stage = is_pageable(Theta) || is_pageable(U) || (region && is_pageable(region));
  const int64_t chunk = stage ? kChunk / 4 : kChunk;
  const int nthr = copy_threads(1);
  auto drain = [&](Slot &s) -> int {
    CUDA_OK(cudaStreamSynchronize(s.stream));
    if (s.pend_cnt == 0) return 0;
    const int64_t off = s.pend_off;
    parallel_rows(s.pend_cnt, nthr, [&](int64_t a, int64_t b) {
      memcpy(U + (off + a) * nu, (const double *)s.hu.p + a * nu, (size_t)(b - a) * nu * 8);
      if (region) memcpy(region + off + a, (const int32_t *)s.hflag.p + a, (size_t)(b - a) * 4);
    });
    s.pend_cnt = 0;
    return 0;
  };
  for (int i = 0; i < kSlots; ++i) t->slots[i].pend_cnt = 0;
  for (int64_t off = 0, c = 0; off < N; off += chunk, ++c) {
    const int64_t cnt = (N - off < chunk) ? (N - off) : chunk;
    Slot &s = t->slots[c % kSlots];
    if (stage) { int rc = drain(s); if (rc) return rc; }
    if (s.in[0].reserve((size_t)cnt * nth * 8) || s.u.reserve((size_t)cnt * nu * 8)) return LMPC_E_CUDA;
    if (region && s.flag.reserve((size_t)cnt * 4)) return LMPC_E_CUDA;
    const double *src = Theta + off * nth;
    if (stage) {
      if (s.hin[0].reserve((size_t)cnt * nth * 8) || s.hu.reserve((size_t)cnt * nu * 8)) return LMPC_E_CUDA;
      if (region && s.hflag.reserve((size_t)cnt * 4)) return LMPC_E_CUDA;
      char *dst = (char *)s.hin[0].p;
      const size_t row = (size_t)nth * 8;
      parallel_rows(cnt, nthr, [&](int64_t a, int64_t b) { memcpy(dst + a * row, (const char *)src + a * row, (size_t)(b - a) * row); });
      src = (const double *)s.hin[0].p;
    }
    CUDA_OK(cudaMemcpyAsync(s.in[0].p, src, (size_t)cnt * nth * 8, cudaMemcpyHostToDevice, s.stream));
    int rc = lmpc_gpu_evaluate_batch_async(t, (const double *)s.in[0].p, cnt, (double *)s.u.p,
                                           region ? (int32_t *)s.flag.p : nullptr, s.stream);
    if (rc) return rc;
    if (stage) {
      CUDA_OK(cudaMemcpyAsync(s.hu.p, s.u.p, (size_t)cnt * nu * 8, cudaMemcpyDeviceToHost, s.stream));
      if (region) CUDA_OK(cudaMemcpyAsync(s.hflag.p, s.flag.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.stream));
      s.pend_off = off; s.pend_cnt = cnt;
      continue;
    }
    CUDA_OK(cudaMemcpyAsync(U + off * nu, s.u.p, (size_t)cnt * nu * 8, cudaMemcpyDeviceToHost, s.stream));
    if (region) CUDA_OK(cudaMemcpyAsync(region + off, s.flag.p, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.stream));
  }
  for (int i = 0; i < kSlots; ++i) { int rc = drain(t->slots[i]); if (rc) return rc; }
  return 0;
}

int lmpc_gpu_simulate_bst_batch(lmpc_bst *t, const double *F, const double *G, const double *Gd,
                                int32_t nx, int32_t nr, int32_t nd, int32_t nuprev,
                                const double *X0, const double *R, const double *D, const double *U0,
                                const double *Rtraj, const double *Dtraj, int64_t N, int32_t T,
                                double *Xend, double *Ulast, double *Utraj) {
  if (!t || !F || !G || !X0 || !Xend || !Ulast || N < 0 || T < 0 || (nd > 0 && !Gd))
    return set_err(LMPC_E_INVALID, "lmpc_gpu_simulate_bst_batch: bad argument%s");
  const int nth = t->dev.nth, nu = t->dev.nu;
  if (nx + nr + nd + nuprev != nth || nuprev > nu)
    return set_err(LMPC_E_INVALID, "simulate_bst_batch: nx + nr + nd + nuprev must equal nth%s");
  if (nth > 16 || nu > 4)
    return set_err(LMPC_E_UNSUPPORTED, "simulate_bst_batch: supports nth <= 16 and nu <= 4%s");
  if (N == 0) return 0;
  std::lock_guard<std::mutex> lock(t->mu);
  CUDA_OK(cudaSetDevice(t->device));
  std::vector<double> dynh;
  push_rowmajor(dynh, F, nx, nx);
  push_rowmajor(dynh, G, nx, nu);
  if (nd > 0) push_rowmajor(dynh, Gd, nx, nd);
  size_t oR = 0, oD = 0;
  if (Rtraj && nr > 0) { oR = dynh.size(); dynh.insert(dynh.end(), Rtraj, Rtraj + (size_t)nr * T); }
  if (Dtraj && nd > 0) { oD = dynh.size(); dynh.insert(dynh.end(), Dtraj, Dtraj + (size_t)nd * T); }
  Slot &s0 = t->slots[0];
  if (s0.aux[2].reserve(dynh.size() * sizeof(double))) return LMPC_E_CUDA;
  CUDA_OK(cudaMemcpyAsync(s0.aux[2].p, dynh.data(), dynh.size() * sizeof(double), cudaMemcpyHostToDevice, t->stream));
  CUDA_OK(cudaStreamSynchronize(t->stream));
  const double *db = (const double *)s0.aux[2].p;
  const double *dR = (Rtraj && nr > 0) ? db + oR : nullptr, *dD = (Dtraj && nd > 0) ? db + oD : nullptr;
  auto launch = [&](const double *x0, const double *r, const double *d, const double *u0, int64_t cnt, double *xe,
                    double *ul, double *ut, cudaStream_t st) -> int {
    int e = launch_bst_sim(t->dev, db, nx, nr, nd, nuprev, x0, r, d, u0, dR, dD, cnt, T, xe, ul, ut, t->sm_count, st);
    if (e) return set_err(LMPC_E_CUDA, "bst closed-loop kernel launch: %s", cudaGetErrorString((cudaError_t)e));
    g_launches.fetch_add(1);
    return 0;
  };
  if (is_device_ptr(X0)) {
    int rc = launch(X0, R, D, U0, N, Xend, Ulast, Utraj, t->stream);
    if (rc) return rc;
    CUDA_OK(cudaStreamSynchronize(t->stream));
    return 0;
  }
  for (int64_t off = 0, c = 0; off < N; off += kChunk, ++c) {
    const int64_t cnt = (N - off < kChunk) ? (N - off) : kChunk;
    Slot &s = t->slots[c % kSlots];
    const double *src[4] = {X0, R, D, U0};
    const int ln[4] = {nx, nr, nd, nuprev};
    const double *dv[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int g = 0; g < 4; ++g) {
      if (!src[g] || ln[g] == 0) continue;
      const size_t bytes = (size_t)cnt * ln[g] * sizeof(double);
      if (s.in[g].reserve(bytes)) return LMPC_E_CUDA;
      CUDA_OK(cudaMemcpyAsync(s.in[g].p, src[g] + off * ln[g], bytes, cudaMemcpyHostToDevice, s.stream));
      dv[g] = (const double *)s.in[g].p;
    }
    if (s.aux[0].reserve((size_t)cnt * nx * 8) || s.u.reserve((size_t)cnt * nu * 8)) return LMPC_E_CUDA;
    if (Utraj && s.aux[1].reserve((size_t)cnt * nu * T * 8)) return LMPC_E_CUDA;
    int rc = launch(dv[0], dv[1], dv[2], dv[3], cnt, (double *)s.aux[0].p, (double *)s.u.p,
                    Utraj ? (double *)s.aux[1].p : nullptr, s.stream);
    if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(Xend + off * nx, s.aux[0].p, (size_t)cnt * nx * 8, cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaMemcpyAsync(Ulast + off * nu, s.u.p, (size_t)cnt * nu * 8, cudaMemcpyDeviceToHost, s.stream));
    if (Utraj) CUDA_OK(cudaMemcpyAsync(Utraj + off * nu * T, s.aux[1].p, (size_t)cnt * nu * T * 8, cudaMemcpyDeviceToHost, s.stream));
  }
  for (int i = 0; i < kSlots; ++i) CUDA_OK(cudaStreamSynchronize(t->slots[i].stream));
  return 0;
}

// ---------------------------------------------------------------------------------------
// one process, several GPUs: contiguous theta ranges, one worker thread per device
// ---------------------------------------------------------------------------------------
struct lmpc_multi {
  std::vector<lmpc_handle *> h;
};

int lmpc_gpu_create_multi(const lmpc_problem *p, const lmpc_settings *s, const int32_t *devices,
                          int32_t n_devices, lmpc_multi **out) {
  if (!p || !out || n_devices <= 0) return set_err(LMPC_E_INVALID, "lmpc_gpu_create_multi: bad argument%s");
  *out = nullptr;
  lmpc_multi *m = new lmpc_multi();
  for (int i = 0; i < n_devices; ++i) {
    lmpc_handle *h = nullptr;
    const int rc = lmpc_gpu_create(p, s, devices ? devices[i] : i, &h);
    if (rc != 0) {
      for (auto *x : m->h) lmpc_gpu_destroy(x);
      delete m;
      return rc;
    }
    m->h.push_back(h);
  }
  *out = m;
  return 0;
}

int lmpc_gpu_destroy_multi(lmpc_multi *m) {
  if (!m) return 0;
  for (auto *x : m->h) lmpc_gpu_destroy(x);
  delete m;
  return 0;
}

int lmpc_gpu_solve_batch_multi(lmpc_multi *m, const double *Theta, int64_t N, double *U,
                               int32_t *exitflag, int32_t *iters, uint64_t *au, uint64_t *al) {
  if (!m || m->h.empty() || !Theta || !U || !exitflag || N < 0)
    return set_err(LMPC_E_INVALID, "lmpc_gpu_solve_batch_multi: bad argument%s");
  const int G = (int)m->h.size();
  const LdpLayout &l = m->h[0]->lay;
  const int words = (l.m + 63) / 64;
  std::vector<int> rc(G, 0);
  std::vector<std::string> msg(G);
  std::vector<std::thread> th;
  for (int g = 0; g < G; ++g) {
    const int64_t lo = N * g / G, hi = N * (g + 1) / G;     // contiguous range [g N / G, (g+1) N / G)
    th.emplace_back([=, &rc, &msg]() {
      if (hi <= lo) return;
      rc[g] = lmpc_gpu_solve_batch(m->h[g], Theta + lo * l.nth, hi - lo, U + lo * l.nu, exitflag + lo,
                                   iters ? iters + lo : nullptr, au ? au + lo * words : nullptr,
                                   al ? al + lo * words : nullptr);
      if (rc[g]) msg[g] = g_err;      // thread-local error text of the worker
    });
  }
  for (auto &t : th) t.join();
  for (int g = 0; g < G; ++g)
    if (rc[g]) return set_err(rc[g], "device worker: %s", msg[g].c_str());
  return 0;
}

// run f(g, lo, hi) for the contiguous range of every device on its own thread; first error wins
extern "C++" {
template <class Fn>
static int for_each_device(int G, int64_t N, Fn f) {
  std::vector<int> rc(G, 0);
  std::vector<std::string> msg(G);
  std::vector<std::thread> th;
  for (int g = 0; g < G; ++g) {
    const int64_t lo = N * g / G, hi = N * (g + 1) / G;
    th.emplace_back([=, &rc, &msg]() {
      if (hi <= lo) return;
      rc[g] = f(g, lo, hi);
      if (rc[g]) msg[g] = g_err;
    });
  }
  for (auto &t : th) t.join();
  for (int g = 0; g < G; ++g)
    if (rc[g]) return set_err(rc[g], "device worker: %s", msg[g].c_str());
  return 0;
}
}  // extern "C++"

int lmpc_gpu_simulate_batch_multi(lmpc_multi *m, const double *F, const double *G, const double *Gd,
                                  int32_t nx, int32_t nr, int32_t nd, int32_t nuprev,
                                  const double *X0, const double *R, const double *D, const double *U0,
                                  const lmpc_sim_ext *ext, int64_t N, int32_t T, int32_t warm_start,
                                  double *Xend, double *Ulast, double *Utraj,
                                  int32_t *exitflag, int32_t *iters) {
  if (!m || m->h.empty() || !X0 || !Xend || !Ulast || !exitflag || N < 0)
    return set_err(LMPC_E_INVALID, "lmpc_gpu_simulate_batch_multi: bad argument%s");
  const int nu = m->h[0]->lay.nu;
  return for_each_device((int)m->h.size(), N, [=](int g, int64_t lo, int64_t hi) {
    lmpc_sim_ext e;
    memset(&e, 0, sizeof(e));
    if (ext) e = *ext;
    if (e.Xhat_end) e.Xhat_end += lo * nx;
    return lmpc_gpu_simulate_batch_ex(m->h[g], F, G, Gd, nx, nr, nd, nuprev, X0 + lo * nx, R ? R + lo * nr : nullptr,
                                      D ? D + lo * nd : nullptr, U0 ? U0 + lo * nuprev : nullptr, &e, hi - lo, T,
                                      warm_start, Xend + lo * nx, Ulast + lo * nu,
                                      Utraj ? Utraj + lo * nu * T : nullptr, exitflag + lo, iters ? iters + lo : nullptr);
  });
}

struct lmpc_bst_multi {
  std::vector<lmpc_bst *> t;
  int nth = 0, nu = 0;
};

int lmpc_gpu_create_bst_multi(const double *halfplanes, int32_t n_hp, const double *feedbacks,
                              int32_t n_fb, const int32_t *hp_list, const int32_t *jump_list,
                              int32_t n_nodes, const double *clipping, int32_t nth, int32_t nu,
                              const int32_t *devices, int32_t n_devices, lmpc_bst_multi **out) {
  if (!out || n_devices <= 0) return set_err(LMPC_E_INVALID, "lmpc_gpu_create_bst_multi: bad argument%s");
  *out = nullptr;
  lmpc_bst_multi *m = new lmpc_bst_multi();
  m->nth = nth; m->nu = nu;
  for (int i = 0; i < n_devices; ++i) {
    lmpc_bst *t = nullptr;
    const int rc = lmpc_gpu_create_bst(halfplanes, n_hp, feedbacks, n_fb, hp_list, jump_list, n_nodes, clipping, nth,
                                       nu, devices ? devices[i] : i, &t);
    if (rc != 0) { lmpc_gpu_destroy_bst_multi(m); return rc; }
    m->t.push_back(t);
  }
  *out = m;
  return 0;
}

int lmpc_gpu_destroy_bst_multi(lmpc_bst_multi *m) {
  if (!m) return 0;
  for (auto *x : m->t) lmpc_gpu_destroy_bst(x);
  delete m;
  return 0;
}

int lmpc_gpu_evaluate_batch_multi(lmpc_bst_multi *m, const double *Theta, int64_t N, double *U, int32_t *region) {
  if (!m || m->t.empty() || !Theta || !U || N < 0)
    return set_err(LMPC_E_INVALID, "lmpc_gpu_evaluate_batch_multi: bad argument%s");
  const int nth = m->nth, nu = m->nu;
  return for_each_device((int)m->t.size(), N, [=](int g, int64_t lo, int64_t hi) {
    return lmpc_gpu_evaluate_batch(m->t[g], Theta + lo * nth, hi - lo, U + lo * nu, region ? region + lo : nullptr);
  });
}

// ---------------------------------------------------------------------------------------
// result window: the final gather fused into the solve kernel's own stores (CUDA IPC + NVLink)
// ---------------------------------------------------------------------------------------
int lmpc_gpu_window_create(int device, size_t bytes, void **devptr, unsigned char handle[64]) {
  if (!devptr || !handle || bytes == 0) return set_err(LMPC_E_INVALID, "lmpc_gpu_window_create: bad argument%s");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  int sm = 0;
  int rc = check_device(device, &sm);
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(device));
  void *p = nullptr;
  CUDA_OK(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) { cudaFree(p); return set_err(LMPC_E_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
  memcpy(handle, &hd, 64);
  *devptr = p;
  return 0;
}

int lmpc_gpu_window_open(int device, const unsigned char handle[64], void **devptr) {
  if (!devptr || !handle) return set_err(LMPC_E_INVALID, "lmpc_gpu_window_open: bad argument%s");
  int sm = 0;
  int rc = check_device(device, &sm);
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(device));
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle, 64);
  void *p = nullptr;
  CUDA_OK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  *devptr = p;
  return 0;
}

int lmpc_gpu_window_read(void *dst_host, const void *devptr, size_t bytes) {
  if (!dst_host || !devptr) return set_err(LMPC_E_INVALID, "lmpc_gpu_window_read: bad argument%s");
  CUDA_OK(cudaMemcpy(dst_host, devptr, bytes, cudaMemcpyDeviceToHost));
  return 0;
}

int lmpc_gpu_window_close(void *devptr, int owner) {
  if (!devptr) return 0;
  if (owner) CUDA_OK(cudaFree(devptr));
  else CUDA_OK(cudaIpcCloseMemHandle(devptr));
  return 0;
}

// ---------------------------------------------------------------------------------------
// measurement helpers
// ---------------------------------------------------------------------------------------
int lmpc_gpu_fp64_peak(int device, double *tflops) {
  if (!tflops) return set_err(LMPC_E_INVALID, "null output%s");
  int sm = 0;
  int rc = check_device(device, &sm);
  if (rc) return rc;
  CUDA_OK(cudaSetDevice(device));
  double *d = nullptr;
  CUDA_OK(cudaMalloc((void **)&d, 8));
  cudaEvent_t a, b;
  CUDA_OK(cudaEventCreate(&a));
  CUDA_OK(cudaEventCreate(&b));
  const int iters = 1 << 16;
  launch_fp64_peak(1 << 10, d, sm, nullptr);   // warm-up
  CUDA_OK(cudaDeviceSynchronize());
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    CUDA_OK(cudaEventRecord(a, nullptr));
    int e = launch_fp64_peak(iters, d, sm, nullptr);
    CUDA_OK(cudaEventRecord(b, nullptr));
    CUDA_OK(cudaEventSynchronize(b));
    if (e) return set_err(LMPC_E_CUDA, "fp64 peak kernel: %s", cudaGetErrorString((cudaError_t)e));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, a, b));
    const double flops = 2.0 * 8.0 * (double)iters * 256.0 * 8.0 * sm;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  g_launches.fetch_add(4);
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
  *tflops = best;
  return 0;
}

void *lmpc_gpu_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void lmpc_gpu_host_free(void *p) { if (p) cudaFreeHost(p); }

}  // extern "C"
