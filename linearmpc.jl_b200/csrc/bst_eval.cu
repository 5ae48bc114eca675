// K4: batched explicit-MPC evaluation: binary-search-tree point location + affine law.
//
// Replaces ParametricDAQP.evaluate(bst, theta) (call site /root/reference/src/utils.jl:57) and
// the generated {fname}_evaluate (/root/reference/src/codegen.jl:113-124) for N parameters.
// One thread per theta; theta lives in registers (NTH is a template parameter), the tree
// (half-planes, leaf laws, node lists) is staged once per CTA into shared memory with a TMA
// bulk copy.  The kernel is HBM-bound by design: 8(nth+nu) bytes per theta, everything else on chip.
#include <cstdlib>
#include "ldp_device.cuh"

namespace {

constexpr int K4_BLOCK = 1024;
#ifndef K4_SCREEN
#define K4_SCREEN 1        // fp32 screening of the branch decisions (0: fp64 rows from shared memory, as in r01)
#endif

// Tree block as packed by lmpc_gpu_create_bst (api.cu), all offsets in doubles:
//   nodes  n_nodes x {jump, index}: packed into ONE 32-bit word (16 + 16 bits) when every jump and
//          index fits (a warp's divergent node loads then cover 128 B instead of 256 B), else int2
//   hp     half-planes [a(nth) | b | pad] cut into 16-byte pairs and REPLICATED `rep` times:
//          pair c of row r, replica p sits at 16-byte slot (r * pairs + c) * rep + p, so that a lane
//          reading replica (lane mod rep) always hits bank group (lane mod rep) when rep == 8: the
//          divergent per-lane row gathers of a warp become conflict-free
//   fb     leaf laws [F(nth) | g | pad], same pair-replicated layout (row = law * nu + output)
//   clip   nu x 2
struct BstLayout {
  int onodes, ohp, ofb, oclip, pairs, rep, packed, total;
  int ohp32, quads, oscr, staged;   // fp32 screening rows (lmpc_common.h: BstDev)
};

// theta tiles: each warp streams its 32 parameters (one contiguous 32*nth*8-byte block of the
// column-major Theta) into its own shared-memory slot with ONE TMA bulk copy, issued a whole tree
// walk ahead; lanes then read their rows conflict-free.  This takes the theta traffic off the
// L1TEX path that the table gathers saturate.
// SCREEN (r02): the kernel is bound by the shared-memory bytes a visited node costs (64 B of fp64 half-plane per lane).
// The branch decision only needs the SIGN of a'theta - b, so it is first evaluated on an fp32 copy of the row (32 B per
// lane for nth = 7, FFMA on the idle FP32 pipe) together with a rigorous bound on how far the fp32 value can be from
// the exact one:  |d32 - (a'theta - b)| <= (n + 4) 2^-24 (sum |a_i theta_i| + |b|) <= 2^-18 (A1 max|theta_i| + B),
// A1 = max_rows sum|a_i|, B = max_rows |b| (n <= 16; rounding of a, theta, b to fp32, the fp32 fma chain and the
// final subtraction; the fp64 evaluation's own error is 2^-29 of that).  Outside the band the fp32 sign IS the fp64
// sign; inside it (a few theta per million, NaN / overflow included) the lane re-evaluates the row in fp64 from
// global memory, exactly as the unscreened kernel does.  Decisions, hence results, are bit-identical; the fp64 rows
// leave shared memory altogether (invpend: 158 KB -> 70 KB staged).  The leaf law stays fp64.
template <int NTH, bool SMEM, bool PACKED, bool SCREEN = false>
__global__ void __launch_bounds__(K4_BLOCK, 1)
bst_eval_kernel(const double *__restrict__ blk, BstLayout bl, int nth_rt, int nu, int has_clip,
                const double *__restrict__ Theta, int64_t N, double *__restrict__ U,
                int32_t *__restrict__ region, int use_tiles) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ __align__(8) uint64_t tbar[K4_BLOCK / 32];
  const double *base = blk;
  const int staged = SCREEN ? bl.staged : bl.total;
  if (SMEM) {
    stage_constants_tma(smem, blk, staged, &mbar);
    base = smem;
  }
  const int nth = NTH > 0 ? NTH : nth_rt;
  const int rep = bl.rep, pairs = bl.pairs;
  // replica choice for the half-plane rows: near the root the lanes of a warp sit on a handful of
  // nodes, so the replica is picked by NODE id there (lanes on the same node share one broadcast
  // read; breadth-first ids make the <= rep nodes of a level hit distinct bank groups); deeper
  // down the walks have scattered and the replica is picked by LANE (exactly 4 wavefronts per load)
  const double2 *hp0 = reinterpret_cast<const double2 *>((SCREEN ? blk : base) + bl.ohp);   // SCREEN: fallback rows in global memory
  const float4 *hq0 = reinterpret_cast<const float4 *>(base + bl.ohp32);
  const int quads = bl.quads;
  const float scrA = SCREEN ? (float)base[bl.oscr] * 1.0000002f : 0.f, scrB = SCREEN ? (float)base[bl.oscr + 1] * 1.0000002f : 0.f;
  const int lane_rep = threadIdx.x & (rep - 1);
#ifndef K4_NODE_REP_IDS
#define K4_NODE_REP_IDS 15          // node ids below this (levels 0..3 of a breadth-first tree)
#endif
  const double2 *fb2 = reinterpret_cast<const double2 *>(base + bl.ofb) + (threadIdx.x & (rep - 1));
  const double *clip = base + bl.oclip;
  const int2 *nodes2 = reinterpret_cast<const int2 *>(base + bl.onodes);
  const uint32_t *nodes1 = reinterpret_cast<const uint32_t *>(base + bl.onodes);
  auto load_node = [&](int i) -> int2 {
    if (PACKED) { const uint32_t w = nodes1[i]; return make_int2((int)(w >> 16), (int)(w & 0xffffu)); }
    return nodes2[i];
  };
  constexpr int TH = NTH > 0 ? NTH : 1;
  constexpr int P2 = (TH + 2) / 2;     // 16-byte pairs covering [a(nth) | b]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tiles = SMEM && NTH > 0 && use_tiles;
  double *tile = smem + staged + (size_t)warp * 32 * TH;      // this warp's theta slot
  uint32_t tphase = 0;
  const int64_t stride = (int64_t)gridDim.x * K4_BLOCK;
  const int64_t q0 = (int64_t)blockIdx.x * K4_BLOCK + warp * 32;   // first theta of the warp's tile
  if (tiles) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbar[warp])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // full tiles only (byte count a multiple of 16); a ragged last tile is loaded directly
    if (lane == 0 && q0 + 32 <= N) tile_issue(&tbar[warp], tile, Theta + q0 * TH, 32u * TH * 8u);
  }
  for (int64_t q = (int64_t)blockIdx.x * K4_BLOCK + threadIdx.x; q - lane < N; q += stride) {
    const int64_t qt = q - lane;              // tile base (warp-uniform)
    const bool live = q < N;
    const double *src = Theta + q * nth;
    double th[TH];
    if (NTH > 0) {
      if (tiles && qt + 32 <= N) {
        tile_wait(&tbar[warp], tphase);
        tphase ^= 1u;
#pragma unroll
        for (int c = 0; c < TH; ++c) th[c] = tile[lane * TH + c];
        __syncwarp();
        // prefetch the warp's next tile while this one is walked
        if (lane == 0 && qt + stride + 32 <= N)
          tile_issue(&tbar[warp], tile, Theta + (qt + stride) * TH, 32u * TH * 8u);
      } else {
#pragma unroll
        for (int c = 0; c < TH; ++c) th[c] = live ? ld_stream(src + c) : 0.0;
      }
    }
    if (!live) continue;
    int id = 0;
    int2 nd = load_node(0);
    float thf[TH];
    float scrE = 0.f;
    if (SCREEN) {
      float mx = 0.f;
#pragma unroll
      for (int c = 0; c < TH; ++c) { thf[c] = (float)th[c]; mx = fmaxf(mx, fabsf(thf[c])); }
      scrE = 3.8146973e-6f * fmaf(scrA, mx * 1.0000002f, scrB);      // 2^-18 (A1 max|theta| + B); NaN / inf -> fallback
      if (!(scrA * mx < 1e37f && scrB < 1e37f)) scrE = __int_as_float(0x7f800000);   // partial sums could overflow fp32: always fp64
    }
    while (nd.x != 0) {
      const int repsel = (id < K4_NODE_REP_IDS ? (id & (rep - 1)) : lane_rep);
      int go = -1;
      if (SCREEN) {
        constexpr int Q4 = (TH + 4) / 4;     // float4 quads covering [a(nth) | b]
        const float4 *h4 = hq0 + (size_t)nd.y * quads * rep + repsel;
        float v[4 * Q4];
#pragma unroll
        for (int c = 0; c < Q4; ++c) { const float4 t4 = h4[c * rep]; v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w; }
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < TH; ++c) acc = fmaf(v[c], thf[c], acc);
        const float d = acc - v[TH];
        if (fabsf(d) > scrE) go = d > 0.f ? 1 : 0;
      }
      if (go < 0) {
        const double2 *h2 = hp0 + (size_t)nd.y * pairs * rep + repsel;
        double acc = 0.0, rhs;
        if (NTH > 0) {
          double v[2 * P2];
#pragma unroll
          for (int c = 0; c < P2; ++c) { const double2 t2 = h2[c * rep]; v[2 * c] = t2.x; v[2 * c + 1] = t2.y; }
#pragma unroll
          for (int c = 0; c < TH; ++c) acc = fma(v[c], th[c], acc);
          rhs = v[TH];
        } else {
          rhs = 0.0;
          for (int c = 0; c <= nth; ++c) {
            const double2 t2 = h2[(c >> 1) * rep];
            const double hv = (c & 1) ? t2.y : t2.x;
            if (c < nth) acc = fma(hv, src[c], acc); else rhs = hv;
          }
        }
        go = acc <= rhs ? 0 : 1;
      }
      id = id + nd.x + go;
      nd = load_node(id);
    }
    const int reg = nd.y;
    for (int i = 0; i < nu; ++i) {
      const double2 *f2 = fb2 + ((size_t)reg * nu + i) * pairs * rep;
      double acc;
      if (NTH > 0) {
        double v[2 * P2];
#pragma unroll
        for (int c = 0; c < P2; ++c) { const double2 t2 = f2[c * rep]; v[2 * c] = t2.x; v[2 * c + 1] = t2.y; }
        acc = v[TH];
#pragma unroll
        for (int c = 0; c < TH; ++c) acc = fma(v[c], th[c], acc);
      } else {
        const double2 g2 = f2[(nth >> 1) * rep];
        acc = (nth & 1) ? g2.y : g2.x;
        for (int c = 0; c < nth; ++c) {
          const double2 t2 = f2[(c >> 1) * rep];
          acc = fma((c & 1) ? t2.y : t2.x, src[c], acc);
        }
      }
      if (has_clip) {
        if (acc < clip[2 * i]) acc = clip[2 * i];
        if (acc > clip[2 * i + 1]) acc = clip[2 * i + 1];
      }
      U[q * nu + i] = acc;
    }
    if (region) region[q] = reg;
  }
}

// ---- K4r (r02): the screened walk as a lane-refill ring ------------------------------------------------------------
// Two findings of the ncu capture of the kernel above drive this variant (profiles/r02_k4_bst_ncu_full.md):
//  (1) it is issue-bound (84 % of the slots) and a visited node costs 34 instructions, half of them address
//      arithmetic on run-time layout fields.  Here the replication factor (8), quads / pairs per row and NU are
//      compile-time, the tables are addressed with 32-bit shared-memory addresses, and a node costs ~22;
//  (2) 92 % of the invpend walks end after 4-5 nodes, 8 % take up to 20, so a warp that walks 32 theta together
//      waits 10.2 steps on average for 4.85 steps of work per lane (SIMT 19 / 32).  Here a lane that reaches its leaf
//      writes its result and takes the NEXT theta of the warp's stream while the stragglers keep walking:
//      each warp owns a two-tile shared-memory ring fed by TMA bulk copies (one 32-theta tile = one contiguous block of
//      Theta); a round = refill the free lanes from the ring (rank among the free lanes -> stream position), walk
//      until at most `park` lanes are still inside the tree, apply the leaf law for the lanes that arrived.
// Per theta the arithmetic is the kernel's above (same fp32 screen and error band, same fp64 fallback, same fp64
// law in the same fma order): results are bit-identical, only the order in which theta are visited changes.
#ifndef K4_RING
#define K4_RING 1
#endif
#ifndef K4R_REP_IDS
#define K4R_REP_IDS 0         // > 0: node ids below this pick the half-plane replica by node id (broadcast reads near the root)
#endif
#ifndef K4_PARK
#define K4_PARK 8            // a round's walk ends once this many lanes or fewer are still walking
#endif
LMPC_DEVINL float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
LMPC_DEVINL double2 lds_d2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
LMPC_DEVINL uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

template <int NTH, int NU, typename QI, int BLK>   // NU: outputs (0 = run-time count); QI: theta index type (int32_t while
                                                    // N < 2^31); BLK: threads per CTA (1024, or 512 when the ring of a wide theta needs room)
__global__ void __launch_bounds__(BLK, 1)
bst_eval_ring_kernel(const double *__restrict__ blk, BstLayout bl, int nu_rt, int has_clip,
                     const double *__restrict__ Theta, int64_t N, double *__restrict__ U,
                     int32_t *__restrict__ region, int park) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ __align__(8) uint64_t tbar[BLK / 32][2];
  stage_constants_tma(smem, blk, bl.staged, &mbar);
  constexpr int TH = NTH, Q4 = (TH + 4) / 4, P2 = (TH + 2) / 2, REP = 8;
  constexpr uint32_t HROW = Q4 * REP * 16, FROW = P2 * REP * 16;      // bytes per replicated row
  const int nu = NU > 0 ? NU : nu_rt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_nodes = s_base + (uint32_t)bl.onodes * 8u;
  const uint32_t s_hq = s_base + (uint32_t)bl.ohp32 * 8u;
  const uint32_t s_hq_lane = s_hq + (uint32_t)(lane & (REP - 1)) * 16u;
  const uint32_t s_fb_lane = s_base + (uint32_t)bl.ofb * 8u + (uint32_t)(lane & (REP - 1)) * 16u;
  const double2 *hp0 = reinterpret_cast<const double2 *>(blk + bl.ohp) + (lane & (REP - 1));   // fp64 rows: global memory
  const float scrA = (float)smem[bl.oscr] * 1.0000002f, scrB = (float)smem[bl.oscr + 1] * 1.0000002f;
  const bool scr_ovf = !(scrB < 1e37f);
  double clo[NU > 0 ? NU : 1], chi[NU > 0 ? NU : 1];
  if (NU > 0) {
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      clo[i] = has_clip ? smem[bl.oclip + 2 * i] : -__longlong_as_double(0x7ff0000000000000ll);
      chi[i] = has_clip ? smem[bl.oclip + 2 * i + 1] : __longlong_as_double(0x7ff0000000000000ll);
    }
  }
  const uint32_t root = lds_u32(s_nodes);

  // per-lane walk state
  double th[TH];
  float thf[TH];
  float scrE = 0.f;
  uint32_t w = 0;            // node word of the current node: jump << 16 | index; < 0x10000 = leaf (or no theta)
  int id = 0;
  QI q = -1;                 // theta index, -1 = the lane holds none

  auto screen = [&]() {
    float mx = 0.f;
#pragma unroll
    for (int c = 0; c < TH; ++c) { thf[c] = (float)th[c]; mx = fmaxf(mx, fabsf(thf[c])); }
    scrE = 3.8146973e-6f * fmaf(scrA, mx * 1.0000002f, scrB);      // 2^-18 (A1 max|theta| + B); NaN / inf -> fallback
    if (!(scrA * mx < 1e37f) || scr_ovf) scrE = __int_as_float(0x7f800000);   // partial sums could overflow fp32: always fp64
  };
  // one node on the fp32 screen.  A lane whose |d| falls inside the error band keeps walking on the fp32 sign but
  // is marked: at its leaf it repeats the whole walk in fp64 (walk64; a few theta per million), so the hot loop holds
  // no fp64 code at all.  Any root-to-leaf path is valid, so the provisional walk always ends at a leaf.
  float margin = 0.f;        // min over the visited nodes of |d| (NaN-propagating): the walk is exact iff margin > scrE
  auto step = [&]() {
#if K4R_REP_IDS > 0
    const uint32_t rb = id < K4R_REP_IDS ? s_hq + ((uint32_t)(id & (REP - 1)) << 4) : s_hq_lane;
#else
    const uint32_t rb = s_hq_lane;
#endif
    uint32_t a;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(w & 0xffffu), "r"(HROW), "r"(rb));
    float v[4 * Q4];
#pragma unroll
    for (int c = 0; c < Q4; ++c) { const float4 t4 = lds_f4(a + c * (REP * 16)); v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w; }
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < TH; ++c) acc = fmaf(v[c], thf[c], acc);
    const float d = acc - v[TH];
    asm("min.NaN.f32 %0, %0, %1;" : "+f"(margin) : "f"(fabsf(d)));
    int gm;                                    // -1 when a'theta > b
    asm("set.gt.s32.f32 %0, %1, %2;" : "=r"(gm) : "f"(acc), "f"(v[TH]));
    id = id + (int)(w >> 16) - gm;
    w = lds_u32(s_nodes + ((uint32_t)id << 2));
  };
  auto walk64 = [&]() {
    id = 0;
    w = root;
    while (w >= 0x10000u) {
      const double2 *h2 = hp0 + (size_t)(w & 0xffffu) * (P2 * REP);
      double vv[2 * P2];
#pragma unroll
      for (int c = 0; c < P2; ++c) { const double2 t2 = h2[c * REP]; vv[2 * c] = t2.x; vv[2 * c + 1] = t2.y; }
      double ad = 0.0;
#pragma unroll
      for (int c = 0; c < TH; ++c) ad = fma(vv[c], th[c], ad);
      id += (int)(w >> 16) + (ad <= vv[TH] ? 0 : 1);
      w = lds_u32(s_nodes + ((uint32_t)id << 2));
    }
  };
  auto leaf_out = [&]() {
    if (!(margin > scrE)) walk64();
    const uint32_t reg = w & 0xffffu;
    auto law = [&](int i) {
      const uint32_t a = s_fb_lane + (reg * (uint32_t)nu + (uint32_t)i) * FROW;
      double v[2 * P2];
#pragma unroll
      for (int c = 0; c < P2; ++c) { const double2 t2 = lds_d2(a + c * (REP * 16)); v[2 * c] = t2.x; v[2 * c + 1] = t2.y; }
      double acc = v[TH];
#pragma unroll
      for (int c = 0; c < TH; ++c) acc = fma(v[c], th[c], acc);
      return acc;
    };
    if (NU > 0) {
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        double acc = law(i);
        if (acc < clo[i]) acc = clo[i];
        if (acc > chi[i]) acc = chi[i];
        U[(int64_t)q * NU + i] = acc;
      }
    } else {
      for (int i = 0; i < nu; ++i) {
        double acc = law(i);
        if (has_clip) {
          const double lo = smem[bl.oclip + 2 * i], hi = smem[bl.oclip + 2 * i + 1];
          if (acc < lo) acc = lo;
          if (acc > hi) acc = hi;
        }
        U[(int64_t)q * nu + i] = acc;
      }
    }
    if (region) region[q] = (int32_t)reg;
  };

  // this warp's stream: the full tiles t0, t0 + ts, ... of Theta (tile = 32 theta = 32 * TH * 8 contiguous bytes)
  const int64_t nt_full = N >> 5;
  const int64_t t0 = (int64_t)blockIdx.x * (BLK / 32) + warp, ts = (int64_t)gridDim.x * (BLK / 32);
  const uint32_t ntw = t0 < nt_full ? (uint32_t)((nt_full - t0 + ts - 1) / ts) : 0u;
  const uint32_t total = ntw * 32u;
  double *ring = smem + bl.staged + (size_t)warp * 64 * TH;
  const uint32_t s_ring = smem_u32(ring);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbar[warp][0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbar[warp][1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0) {
    if (ntw > 0) tile_issue(&tbar[warp][0], ring, Theta + t0 * 32 * TH, 32u * TH * 8u);
    if (ntw > 1) tile_issue(&tbar[warp][1], ring + 32 * TH, Theta + (t0 + ts) * 32 * TH, 32u * TH * 8u);
  }
  const double *next_src = Theta + (t0 + 2 * ts) * 32 * TH;     // source of the next tile to request (third of the stream)
  uint32_t head = 0;          // theta of the stream handed to lanes so far (warp-uniform)
  uint32_t landed = 0;        // tiles of the stream waited for so far (warp-uniform)
  const unsigned lt = (1u << lane) - 1u;
  int pre = 0;
  while (true) {
    // refill: the lanes outside the tree take the next theta of the stream, in lane order
    const bool need = w < 0x10000u;
    const unsigned fm = __ballot_sync(0xffffffffu, need);
    const uint32_t want = head + (uint32_t)__popc(fm);
    const uint32_t nhead = want < total ? want : total;
    if (nhead > head) {
      const uint32_t jl = (nhead - 1u) >> 5;
      while (landed <= jl) { tile_wait(&tbar[warp][landed & 1u], (landed >> 1) & 1u); ++landed; }
      const uint32_t p = head + (uint32_t)__popc(fm & lt);
      if (need && p < total) {
        const uint32_t sa = s_ring + (p & 63u) * (TH * 8u);
#pragma unroll
        for (int c = 0; c < TH; ++c) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(th[c]) : "r"(sa + c * 8u));
        q = (QI)(((QI)t0 + (QI)(p >> 5) * (QI)ts) << 5) + (QI)(p & 31u);
        screen();
        id = 0;
        w = root;
        margin = __int_as_float(0x7f800000);
      }
      __syncwarp();
      // a tile whose 32 theta are all in registers frees its half of the ring for the tile two ahead
      // (a refill hands out at most 32 theta, so at most one tile boundary is crossed)
      const uint32_t j = head >> 5;
      if ((nhead >> 5) != j && j + 2u < ntw) {
        if (lane == 0) tile_issue(&tbar[warp][j & 1u], ring + (size_t)(j & 1u) * 32 * TH, next_src, 32u * TH * 8u);
        next_src += ts * (32 * TH);
      }
      head = nhead;
    }
    if (__ballot_sync(0xffffffffu, q >= 0) == 0u) break;
    // walk until at most `lim` lanes are still inside the tree (to the end once the stream is exhausted)
    // The "how many lanes are still walking" vote costs half a node step: the first `pre` steps of a round run
    // without it, pre = one less than what the previous round needed (adapts to the tree and to the theta).
    const int lim = head < total ? park : 0;
    int ns = 0;
#pragma unroll 1
    for (; ns + 2 <= pre; ns += 2) {
      if (w >= 0x10000u) step();
      if (w >= 0x10000u) step();
    }
    if (ns < pre) {
      if (w >= 0x10000u) step();
      ++ns;
    }
    while (true) {
      const bool in = w >= 0x10000u;
      if (__popc(__ballot_sync(0xffffffffu, in)) <= lim) break;
      if (in) step();
      ++ns;
    }
    pre = ns > 0 ? ns - 1 : 0;
    if (q >= 0 && w < 0x10000u) { leaf_out(); q = -1; }
  }
  // ragged tail (N mod 32 theta): first warp of the grid, straight from global memory
  if (blockIdx.x == 0 && warp == 0) {
    const int64_t qq = (nt_full << 5) + lane;
    if (qq < N) {
#pragma unroll
      for (int c = 0; c < TH; ++c) th[c] = ld_stream(Theta + qq * TH + c);
      q = (QI)qq;
      screen();
      id = 0;
      w = root;
      margin = __int_as_float(0x7f800000);
      while (w >= 0x10000u) step();
      leaf_out();
    }
  }
}

// threads per CTA of the ring kernel for this tree: 1024 when the tables and two theta tiles per warp fit, else 512, else 0
static int k4_ring_block(const BstDev &t) {
  if (!(K4_RING && t.rep == 8 && t.nth >= 1 && t.nth <= 16)) return 0;
  for (int blk = 1024; blk >= 512; blk >>= 1)
    if ((size_t)t.staged * sizeof(double) + (size_t)blk * 2 * t.nth * sizeof(double) <= 225 * 1024) return blk;
  return 0;
}

template <int NTH, int BLK>
int launch_k4_ring(const BstDev &t, const BstLayout &bl, const double *Theta, int64_t N, double *U,
                   int32_t *region, int sm_count, cudaStream_t s) {
  const size_t smem = (size_t)bl.staged * sizeof(double) + (size_t)BLK * 2 * NTH * sizeof(double);
  int64_t blocks = (N + BLK - 1) / BLK;
  if (blocks > sm_count) blocks = sm_count;
  int park = K4_PARK;
  if (const char *e = getenv("LMPC_K4_PARK")) park = atoi(e);
  const bool small = N < (1ll << 31) - 64 && !getenv("LMPC_K4_INDEX64");
#define K4R_GO(NUU)                                                                                             \
  {                                                                                                             \
    auto kern = small ? bst_eval_ring_kernel<NTH, NUU, int32_t, BLK> : bst_eval_ring_kernel<NTH, NUU, int64_t, BLK>; \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    if (e != cudaSuccess) return (int)e;                                                                        \
    kern<<<(unsigned)blocks, BLK, smem, s>>>(t.blk, bl, t.nu, t.has_clip, Theta, N, U, region, park);           \
    return (int)cudaGetLastError();                                                                             \
  }
  if (t.nu == 1) K4R_GO(1)
  if (t.nu == 2) K4R_GO(2)
  K4R_GO(0)
#undef K4R_GO
}

template <int NTH, bool SMEM, bool PACKED, bool SCREEN = false>
int launch_k4s(const BstDev &t, const BstLayout &bl, const double *Theta, int64_t N, double *U,
               int32_t *region, int sm_count, cudaStream_t s) {
  auto kern = bst_eval_kernel<NTH, SMEM, PACKED, SCREEN>;
  size_t smem = SMEM ? (size_t)(SCREEN ? bl.staged : bl.total) * sizeof(double) : 0;
  // room for one 32-theta tile per warp next to the tables?
  const size_t tile_bytes = (size_t)K4_BLOCK * (NTH > 0 ? NTH : 0) * sizeof(double);
  const int use_tiles = SMEM && NTH > 0 && smem + tile_bytes <= 225 * 1024 &&
                        (reinterpret_cast<uintptr_t>(Theta) & 15u) == 0;   // TMA needs 16-byte alignment
  if (use_tiles) smem += tile_bytes;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K4_BLOCK, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = (N + K4_BLOCK - 1) / K4_BLOCK;
  const int64_t cap = (int64_t)sm_count * per_sm;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, K4_BLOCK, smem, s>>>(t.blk, bl, t.nth, t.nu, t.has_clip, Theta, N, U, region,
                                                use_tiles);
  return (int)cudaGetLastError();
}

}  // namespace
// which evaluate kernel a tree gets (16-byte aligned Theta): 0 = fp64 rows (generic), 1 = fp32-screened walk, 32 theta per
// warp in lockstep, 2 = fp32-screened lane-refill ring (K4r)
int bst_eval_kind(const BstDev &t) {
  if (!(K4_SCREEN && t.nth >= 1 && t.nth <= 16 && t.quads > 0 && t.packed && (size_t)t.staged * sizeof(double) <= 165 * 1024))
    return 0;
  if (k4_ring_block(t) > 0)
    return 2;
  return 1;
}
namespace {

template <int NTH>
int launch_k4(const BstDev &t, const BstLayout &bl, const double *Theta, int64_t N, double *U,
              int32_t *region, int sm_count, cudaStream_t s) {
  // the fp32 screen needs compile-time nth, 32-bit node words and its staged prefix in shared memory
  if constexpr (NTH > 0) {
    const int kind = bst_eval_kind(t);
    if (kind >= 1) {
      // lane-refill ring: 8 replicas, room for two theta tiles per warp, TMA-aligned Theta
      if (kind == 2 && (reinterpret_cast<uintptr_t>(Theta) & 15u) == 0 && !getenv("LMPC_K4_NO_RING"))
        return k4_ring_block(t) == 1024 ? launch_k4_ring<NTH, 1024>(t, bl, Theta, N, U, region, sm_count, s)
                                        : launch_k4_ring<NTH, 512>(t, bl, Theta, N, U, region, sm_count, s);
      return launch_k4s<NTH, true, true, true>(t, bl, Theta, N, U, region, sm_count, s);
    }
  }
  const bool in_smem = (size_t)bl.total * sizeof(double) <= 220 * 1024;
  if (in_smem && bl.packed) return launch_k4s<NTH, true, true>(t, bl, Theta, N, U, region, sm_count, s);
  if (in_smem) return launch_k4s<NTH, true, false>(t, bl, Theta, N, U, region, sm_count, s);
  if (bl.packed) return launch_k4s<NTH, false, true>(t, bl, Theta, N, U, region, sm_count, s);
  return launch_k4s<NTH, false, false>(t, bl, Theta, N, U, region, sm_count, s);
}

// ---- closed loop with the explicit law: Simulation(empc, scenario) (src/simulation.jl:37-116) ----
// Thread per closed loop: theta = [x; r; d; u_prev] lives in registers, every step walks the tree
// (same tables, same arithmetic as bst_eval_kernel), applies the leaf law (+ clipping) and the
// linear dynamics x+ = F x + G u + Gd d (src/model.jl:70); F, G, Gd sit behind the tables in shared
// memory.  Run-time sizes nx, nr, nd, nu <= NTH are handled by fully unrolled, predicated loops so
// that theta never leaves the register file.
constexpr int K4S_BLOCK = 256;
constexpr int K4S_MAXU = 4;

template <int NTH>
__global__ void __launch_bounds__(K4S_BLOCK)
bst_sim_kernel(const double *__restrict__ blk, BstLayout bl, int nu, int has_clip, const double *__restrict__ dyn,
               int nx, int nr, int nd, int nup, const double *__restrict__ X0, const double *__restrict__ R,
               const double *__restrict__ D, const double *__restrict__ U0, const double *__restrict__ Rtraj,
               const double *__restrict__ Dtraj, int64_t N, int T, double *__restrict__ Xend,
               double *__restrict__ Ulast, double *__restrict__ Utraj) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t mbar;
  stage_constants_tma(smem, blk, bl.total, &mbar);
  const int ndyn = nx * (nx + nu + nd);
  double *dsm = smem + ((bl.total + 1) & ~1);
  for (int i = threadIdx.x; i < ndyn; i += K4S_BLOCK) dsm[i] = dyn[i];
  __syncthreads();
  const double *Fm = dsm, *Gm = dsm + nx * nx, *Gdm = Gm + nx * nu;
  const int rep = bl.rep, pairs = bl.pairs;
  const double2 *hp2 = reinterpret_cast<const double2 *>(smem + bl.ohp) + (threadIdx.x & (rep - 1));
  const double2 *fb2 = reinterpret_cast<const double2 *>(smem + bl.ofb) + (threadIdx.x & (rep - 1));
  const double *clip = smem + bl.oclip;
  const int2 *nodes2 = reinterpret_cast<const int2 *>(smem + bl.onodes);
  const uint32_t *nodes1 = reinterpret_cast<const uint32_t *>(smem + bl.onodes);
  const bool packed = bl.packed != 0;
  auto load_node = [&](int i) -> int2 {
    if (packed) { const uint32_t w = nodes1[i]; return make_int2((int)(w >> 16), (int)(w & 0xffffu)); }
    return nodes2[i];
  };
  constexpr int P2 = (NTH + 2) / 2;
  for (int64_t q = (int64_t)blockIdx.x * K4S_BLOCK + threadIdx.x; q < N; q += (int64_t)gridDim.x * K4S_BLOCK) {
    double th[NTH], u[K4S_MAXU];
#pragma unroll
    for (int c = 0; c < NTH; ++c) {
      double v = 0.0;
      if (c < nx) v = X0[q * nx + c];
      else if (c < nx + nr) v = R ? R[q * nr + (c - nx)] : 0.0;
      else if (c < nx + nr + nd) v = D ? D[q * nd + (c - nx - nr)] : 0.0;
      th[c] = v;
    }
#pragma unroll
    for (int i = 0; i < K4S_MAXU; ++i) u[i] = (U0 && i < nup) ? U0[q * nup + i] : 0.0;
    for (int t = 0; t < T; ++t) {
#pragma unroll
      for (int c = 0; c < NTH; ++c) {
        if (Rtraj && c >= nx && c < nx + nr) th[c] = Rtraj[(size_t)t * nr + (c - nx)];
        if (Dtraj && c >= nx + nr && c < nx + nr + nd) th[c] = Dtraj[(size_t)t * nd + (c - nx - nr)];
#pragma unroll
        for (int i = 0; i < K4S_MAXU; ++i) if (c == nx + nr + nd + i && i < nup) th[c] = u[i];
      }
      // tree walk (ParametricDAQP.evaluate, src/utils.jl:57)
      int id = 0;
      int2 nd2 = load_node(0);
      while (nd2.x != 0) {
        const double2 *h2 = hp2 + (size_t)nd2.y * pairs * rep;
        double v[2 * P2];
#pragma unroll
        for (int c = 0; c < P2; ++c) { const double2 t2 = h2[c * rep]; v[2 * c] = t2.x; v[2 * c + 1] = t2.y; }
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < NTH; ++c) acc = fma(v[c], th[c], acc);
        id = id + nd2.x + (acc <= v[NTH] ? 0 : 1);
        nd2 = load_node(id);
      }
      const int reg = nd2.y;
#pragma unroll
      for (int i = 0; i < K4S_MAXU; ++i) {
        if (i < nu) {
          const double2 *f2 = fb2 + ((size_t)reg * nu + i) * pairs * rep;
          double v[2 * P2];
#pragma unroll
          for (int c = 0; c < P2; ++c) { const double2 t2 = f2[c * rep]; v[2 * c] = t2.x; v[2 * c + 1] = t2.y; }
          double acc = v[NTH];
#pragma unroll
          for (int c = 0; c < NTH; ++c) acc = fma(v[c], th[c], acc);
          if (has_clip) {
            if (acc < clip[2 * i]) acc = clip[2 * i];
            if (acc > clip[2 * i + 1]) acc = clip[2 * i + 1];
          }
          u[i] = acc;
          if (Utraj) Utraj[((size_t)q * T + t) * nu + i] = acc;
        }
      }
      // x+ = F x + G u + Gd d, fma chains in index order (oracle: orc_bst_simulate_batch)
      double xn[NTH];
#pragma unroll
      for (int i = 0; i < NTH; ++i) {
        double acc = 0.0;
        if (i < nx) {
#pragma unroll
          for (int c = 0; c < NTH; ++c) if (c < nx) acc = fma(Fm[i * nx + c], th[c], acc);
#pragma unroll
          for (int c = 0; c < K4S_MAXU; ++c) if (c < nu) acc = fma(Gm[i * nu + c], u[c], acc);
#pragma unroll
          for (int c = 0; c < NTH; ++c) if (c >= nx + nr && c < nx + nr + nd) acc = fma(Gdm[i * nd + (c - nx - nr)], th[c], acc);
        }
        xn[i] = acc;
      }
#pragma unroll
      for (int i = 0; i < NTH; ++i) if (i < nx) th[i] = xn[i];
    }
#pragma unroll
    for (int c = 0; c < NTH; ++c) if (c < nx) Xend[q * nx + c] = th[c];
#pragma unroll
    for (int i = 0; i < K4S_MAXU; ++i) if (i < nu) Ulast[q * nu + i] = u[i];
  }
}

template <int NTH>
int launch_k4_sim(const BstDev &t, const BstLayout &bl, const double *dyn, int nx, int nr, int nd, int nup,
                  const double *X0, const double *R, const double *D, const double *U0, const double *Rtraj,
                  const double *Dtraj, int64_t N, int T, double *Xend, double *Ulast, double *Utraj,
                  int sm_count, cudaStream_t s) {
  auto kern = bst_sim_kernel<NTH>;
  const size_t smem = ((size_t)((bl.total + 1) & ~1) + (size_t)nx * (nx + t.nu + nd)) * sizeof(double);
  if (smem > 220 * 1024) return (int)cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K4S_BLOCK, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = (N + K4S_BLOCK - 1) / K4S_BLOCK;
  if (blocks > (int64_t)sm_count * per_sm) blocks = (int64_t)sm_count * per_sm;
  kern<<<(unsigned)blocks, K4S_BLOCK, smem, s>>>(t.blk, bl, t.nu, t.has_clip, dyn, nx, nr, nd, nup, X0, R, D, U0,
                                                 Rtraj, Dtraj, N, T, Xend, Ulast, Utraj);
  return (int)cudaGetLastError();
}

}  // namespace

int launch_bst_eval(const BstDev &t, const double *Theta, int64_t N, double *U, int32_t *region,
                    int sm_count, void *stream) {
  if (N <= 0) return 0;
  BstLayout bl;
  bl.onodes = t.onodes; bl.ohp = t.ohp; bl.ofb = t.ofb; bl.oclip = t.oclip;
  bl.pairs = t.pairs; bl.rep = t.rep; bl.packed = t.packed; bl.total = t.total;
  bl.ohp32 = t.ohp32; bl.quads = t.quads; bl.oscr = t.oscr; bl.staged = t.staged;
  cudaStream_t s = (cudaStream_t)stream;
#define K4_CASE(NN) case NN: return launch_k4<NN>(t, bl, Theta, N, U, region, sm_count, s);
  switch (t.nth) {
    K4_CASE(1) K4_CASE(2) K4_CASE(3) K4_CASE(4) K4_CASE(5) K4_CASE(6) K4_CASE(7) K4_CASE(8)
    K4_CASE(9) K4_CASE(10) K4_CASE(11) K4_CASE(12) K4_CASE(13) K4_CASE(14) K4_CASE(15) K4_CASE(16)
    default: return launch_k4<0>(t, bl, Theta, N, U, region, sm_count, s);
  }
}

int launch_bst_sim(const BstDev &t, const double *dyn, int nx, int nr, int nd, int nup, const double *X0,
                   const double *R, const double *D, const double *U0, const double *Rtraj, const double *Dtraj,
                   int64_t N, int T, double *Xend, double *Ulast, double *Utraj, int sm_count, void *stream) {
  if (N <= 0) return 0;
  if (t.nth < 1 || t.nth > 16 || t.nu > K4S_MAXU || nx + nr + nd + nup != t.nth) return (int)cudaErrorInvalidValue;
  BstLayout bl;
  bl.onodes = t.onodes; bl.ohp = t.ohp; bl.ofb = t.ofb; bl.oclip = t.oclip;
  bl.pairs = t.pairs; bl.rep = t.rep; bl.packed = t.packed; bl.total = t.total;
  bl.ohp32 = t.ohp32; bl.quads = t.quads; bl.oscr = t.oscr; bl.staged = t.staged;
  cudaStream_t s = (cudaStream_t)stream;
#define K4S_CASE(NN) case NN: return launch_k4_sim<NN>(t, bl, dyn, nx, nr, nd, nup, X0, R, D, U0, Rtraj, Dtraj, N, T, Xend, Ulast, Utraj, sm_count, s);
  switch (t.nth) {
    K4S_CASE(1) K4S_CASE(2) K4S_CASE(3) K4S_CASE(4) K4S_CASE(5) K4S_CASE(6) K4S_CASE(7) K4S_CASE(8)
    K4S_CASE(9) K4S_CASE(10) K4S_CASE(11) K4S_CASE(12) K4S_CASE(13) K4S_CASE(14) K4S_CASE(15) K4S_CASE(16)
    default: return (int)cudaErrorInvalidValue;
  }
}

// ---- FP64 FMA peak microbenchmark (roofline denominator measured in the same run) -------
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double *out) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 123.456) out[0] = r;  // keep the chain alive
}

int launch_fp64_peak(int iters, double *d_out, int sm_count, void *stream) {
  fp64_peak_kernel<<<sm_count * 8, 256, 0, (cudaStream_t)stream>>>(iters, d_out);
  return (int)cudaGetLastError();
}

// ---- failed-instance count: how many exit flags are < 1 (what `@assert exitflag >= 1`, src/utils.jl:46,
// would trip on).  Lets a Monte-Carlo caller leave the 4-byte-per-solve flag array on the device. ----
__global__ void __launch_bounds__(256) count_failed_kernel(const int32_t *__restrict__ flag, int64_t n,
                                                           unsigned long long *counter) {
  unsigned cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) cnt += flag[i] < 1;
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(counter, (unsigned long long)cnt);
}
int launch_count_failed(const int32_t *flag, int64_t n, unsigned long long *counter, void *stream) {
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  count_failed_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(flag, n, counter);
  return (int)cudaGetLastError();
}
