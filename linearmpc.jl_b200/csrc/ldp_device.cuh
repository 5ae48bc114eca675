// Device helpers shared by the LDP kernels: TMA bulk staging of the constant block,
// typed views into it, theta gather from row-segments.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
#include "lmpc_common.h"

#define LMPC_DEVINL __device__ __forceinline__

// ---- TMA (cp.async.bulk) staging of the per-controller constant block ------------------
// One elected thread arms an mbarrier with the byte count and issues a single bulk copy
// global -> shared; everybody waits on the barrier's phase.  SASS: UBLKCP + SYNCS.
LMPC_DEVINL uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

LMPC_DEVINL void stage_constants_tma(double *dst_smem, const double *src_gmem, int total_doubles,
                                     uint64_t *mbar) {
  const uint32_t bar = smem_u32(mbar);
  const uint32_t dst = smem_u32(dst_smem);
  const uint32_t bytes = static_cast<uint32_t>(total_doubles) * 8u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
    // the copy size must be a multiple of 16 B; chunk to stay well under the tx-count limit
    uint32_t off = 0;
    while (off < bytes) {
      uint32_t chunk = bytes - off;
      if (chunk > 65536u) chunk = 65536u;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"(dst + off), "l"(reinterpret_cast<const char *>(src_gmem) + off), "r"(chunk), "r"(bar)
          : "memory");
      off += chunk;
    }
  }
  // wait for phase 0 to complete
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar)
        : "memory");
  }
}

// ---- per-warp TMA tiles (theta prefetch): one mbarrier per warp, one bulk copy per tile -------
LMPC_DEVINL void tile_issue(uint64_t *bar, double *dst, const double *src, uint32_t bytes) {
  const uint32_t b = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
}
LMPC_DEVINL void tile_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t b = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(b), "r"(parity) : "memory");
  }
}

// typed view of the constant block (pointers into shared memory)
struct ConstView {
  const double *M, *MT, *G, *Dth, *du, *dl, *Rinv, *Uth, *uoff;
  const int32_t *sense;
};
LMPC_DEVINL ConstView make_view(const double *blk, const LdpLayout &lay) {
  ConstView v;
  v.M = blk + lay.oM;
  v.MT = blk + lay.oMT;
  v.G = blk + lay.oG;
  v.Dth = blk + lay.oDth;
  v.du = blk + lay.odu;
  v.dl = blk + lay.odl;
  v.Rinv = blk + lay.oRinv;
  v.Uth = blk + lay.oUth;
  v.uoff = blk + lay.ouoff;
  v.sense = reinterpret_cast<const int32_t *>(blk + lay.oSense);
  return v;
}

// streaming (read-once) global load: keep theta out of L1
LMPC_DEVINL double ld_stream(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// ---- compile-time loops and the canonical tree sum (shared by the thread-per-QP kernels) ----
#define LIDX(i, c) ((i) * ((i)-1) / 2 + (c))

// Compile-time loops.  Every index into per-thread state is a constant at the front end, so
// the state is scalarised into registers before any optimisation pass can rewrite a chain of
// "slot == run-time position" selects into one dynamically indexed (local-memory) access.
template <int B, int E, class F>
LMPC_DEVINL void sfor(F &&f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    sfor<B + 1, E>(f);
  }
}
template <int B, int E, class F>   // B-1, B-2, ..., E
LMPC_DEVINL void sfor_down(F &&f) {
  if constexpr (B > E) {
    f(std::integral_constant<int, B - 1>{});
    sfor_down<B - 1, E>(f);
  }
}
#define SFOR(var, B, E) sfor<B, E>([&](auto var##_c) { constexpr int var = decltype(var##_c)::value;
#define SFOR_DOWN(var, B, E) sfor_down<B, E>([&](auto var##_c) { constexpr int var = decltype(var##_c)::value;
#define SEND });

// Canonical cross-element sum (see oracle/ldp_oracle.c: lanesum): element i lives on lane
// i mod 32 and the lanes are combined by an xor-butterfly, i.e. the value on lane 0 is
// S(base, step) = S(base, 2 step) + S(base + step, 2 step), S(base, 32) = p[base].  With
// LEN <= 32 every lane holds at most one product, so a thread can evaluate the same tree
// statically (sub-trees that only contain padding zeros are skipped: x + 0 == x).
template <int LEN, int BASE, int STEP>
LMPC_DEVINL double tree_sum(const double (&p)[LEN]) {
  if constexpr (STEP >= 32) return p[BASE];
  else if constexpr (BASE + STEP >= LEN) return tree_sum<LEN, BASE, 2 * STEP>(p);
  else return __dadd_rn(tree_sum<LEN, BASE, 2 * STEP>(p), tree_sum<LEN, BASE + STEP, 2 * STEP>(p));
}

