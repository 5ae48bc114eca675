// How many L2 loads must a thread keep in flight before a dependent DFMA chain fed by them runs at the DFMA
// latency?  Shape of the CTA kernel's scan / primal-update loops: B independent loads (row stride 1440 B,
// coalesced across the warp), then B dependent DFMAs.  Also the latency of one small TMA bulk copy.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int B>
__global__ void __launch_bounds__(128, 2) k(const double *g, double *out, long long *cyc, int steps) {
  const int lane = threadIdx.x;
  double a = 1.0;
  const double *p = g + lane + (blockIdx.x % 7) * 180 * 180;
  long long t0 = clock64();
  for (int it = 0; it < steps; it += B) {
    double v[B];
#pragma unroll
    for (int j = 0; j < B; ++j) v[j] = __ldcg(p + (size_t)((it + j) % 180) * 180);
#pragma unroll
    for (int j = 0; j < B; ++j) a = fma(v[j], 1.0000001, a);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * 128 + threadIdx.x] = a;
}
// double-buffered: next batch's loads are issued before this batch is consumed
template <int B>
__global__ void __launch_bounds__(128, 2) k2(const double *g, double *out, long long *cyc, int steps) {
  const int lane = threadIdx.x;
  double a = 1.0;
  const double *p = g + lane + (blockIdx.x % 7) * 180 * 180;
  double v[B], w[B];
#pragma unroll
  for (int j = 0; j < B; ++j) v[j] = __ldcg(p + (size_t)(j % 180) * 180);
  long long t0 = clock64();
  for (int it = 0; it < steps; it += 2 * B) {
#pragma unroll
    for (int j = 0; j < B; ++j) w[j] = __ldcg(p + (size_t)((it + B + j) % 180) * 180);
#pragma unroll
    for (int j = 0; j < B; ++j) a = fma(v[j], 1.0000001, a);
#pragma unroll
    for (int j = 0; j < B; ++j) v[j] = __ldcg(p + (size_t)((it + 2 * B + j) % 180) * 180);
#pragma unroll
    for (int j = 0; j < B; ++j) a = fma(w[j], 1.0000001, a);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * 128 + threadIdx.x] = a;
}
__global__ void tma_lat(const double *g, long long *cyc, int bytes) {
  __shared__ __align__(16) double buf[4096];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0;
    uint32_t ph = 0;
    for (int rep = 0; rep < 64; ++rep) {
      long long t0 = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(d), "l"(g + (rep % 16) * 512), "r"(bytes), "r"(b) : "memory");
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b), "r"(ph) : "memory");
      ph ^= 1;
      tot += clock64() - t0;
    }
    cyc[0] = tot / 64;
  }
}
template <class K>
void run(const char *name, K kern, const double *g, double *out, long long *cyc, int grid) {
  const int steps = 180 * 16;
  kern<<<grid, 128>>>(g, out, cyc, steps);
  kern<<<grid, 128>>>(g, out, cyc, steps);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s grid %4d: %.1f cycles/step\n", name, grid, (double)c / steps);
}
int main() {
  double *g, *out; long long *cyc;
  cudaMalloc(&g, 8 * 180 * 180 * 8 + 4096); cudaMemset(g, 0, 8 * 180 * 180 * 8 + 4096);
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  for (int grid : {1, 296}) {
    run("batch 8", k<8>, g, out, cyc, grid);
    run("batch 16", k<16>, g, out, cyc, grid);
    run("batch 32", k<32>, g, out, cyc, grid);
    run("double-buffered 8", k2<8>, g, out, cyc, grid);
    run("double-buffered 16", k2<16>, g, out, cyc, grid);
    run("double-buffered 32", k2<32>, g, out, cyc, grid);
  }
  for (int bytes : {256, 1440, 5760}) {
    tma_lat<<<1, 32>>>(g, cyc, bytes);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("TMA bulk copy %5d B: %lld cycles issue -> landed\n", bytes, c);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
