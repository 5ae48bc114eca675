// Latency microbenchmarks that size the kernels' serial chains (one warp, clock64 around dependent ops):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, const double *g, long long *cyc, double *sm_init) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double a = 1.0 + lane * 1e-9, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // 1. dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) a = fma(a, b, c);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // 2. SHFL broadcast + DFMA per step (the triangular-sweep step)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) { const double z = __shfl_sync(0xffffffffu, a, i & 31); a = fma(-b, z, a); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // 3. LDS (dependent address) + DFMA per step
  int idx = lane;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) { const double z = sm[idx]; a = fma(-b, z, a); idx = (idx + (int)(a > 1e300) + 1) & 1023; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // 4. independent LDS feeding a dependent DFMA chain (chain bound by DFMA latency)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) a = fma(-b, sm[(lane + i) & 1023], a);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // 5. L2 hit latency: dependent pointer chase through a 1 MB buffer (warm)
  long long p = 0;
  const long long *gl = (const long long *)g;
  for (int i = 0; i < 64; ++i) p = gl[p];
  t0 = clock64();
  for (int i = 0; i < 256; ++i) p = __ldcg(gl + p);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // 6. 8 independent L2 loads then 8 dependent DFMAs (the scan loop's shape), x32
  const double *gd = g + 262144;
  t0 = clock64();
  for (int it = 0; it < 32; ++it) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldcg(gd + (it * 8 + j) * 256 + lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) a = fma(v[j], b, a);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // 7. DADD dependent, DMUL dependent
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) a = __dadd_rn(a, c);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  out[threadIdx.x] = a + (double)p;
}
int main() {
  double *out, *g; long long *cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&g, 4 << 20); cudaMalloc(&cyc, 64);
  long long *h = (long long *)malloc(2 << 20);
  const int n = (2 << 20) / 8;
  for (int i = 0; i < n; ++i) h[i] = (i + 4099) % n;      // stride > a line, stays in the first 2 MB
  cudaMemcpy(g, h, 2 << 20, cudaMemcpyHostToDevice);
  cudaMemset((char *)g + (2 << 20), 0, 2 << 20);
  for (int warps = 1; warps <= 8; warps *= 2) {
    k<<<1, 32 * warps>>>(out, g, cyc, nullptr);
    k<<<1, 32 * warps>>>(out, g, cyc, nullptr);
    cudaDeviceSynchronize();
    long long c[8];
    cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost);
    printf("warps=%d  DFMA dep %.1f  SHFL+DFMA %.1f  LDS(dep addr)+DFMA %.1f  LDS(indep)+DFMA %.1f  L2 chase %.1f  [8 L2 loads + 8 DFMA]/8 %.1f  DADD dep %.1f  (cycles per step)\n",
           warps, c[0] / 256.0, c[1] / 256.0, c[2] / 256.0, c[3] / 256.0, c[4] / 256.0, c[5] / 256.0, c[6] / 256.0);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
