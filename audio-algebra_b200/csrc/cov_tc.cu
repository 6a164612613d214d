// VICReg covariance loss contractions on tcgen05 (aa_mixer.py:355-364 through the Gram identity of aa_ops.cu), fp32-accurate:
//
//   forward   G = Xc Xc^T                       [B][B],  Xc = z - column mean,  K = D      (gram_bb_tc_kernel, split over D)
//   backward  dX = k (G Xc - Xc diag(s))        [B][D],                         K = B      (cov_bwd_tc_kernel)
//
// Both are kind::tf32 GEMMs with every fp32 operand split into hi = its top 19 bits (exact in TF32) and lo = value - hi; three
// products a_lo b_hi + a_hi b_lo + a_hi b_hi are accumulated in fp32 in TMEM (the dropped lo lo term is ~2^-22 relative), chains
// are closed every kChainChunks K chunks (tcgen05 accumulates with truncation: profiles/ubench/tf32_accum_ubench) and summed in
// registers by the epilogue warps.  Loader warps read z (and the means / the Gram matrix) with coalesced 16-byte loads, centre,
// split and store the operands MMA-ready:
//   * K-major operands (both operands of the forward: rows of z are contiguous along D = K; the G operand of the backward) in
//     the no-swizzle panel layout [K/4][rows][16 B] of proj_tc.cu / gram_tc.cu;
//   * the Xc operand of the backward is contracted over the batch, i.e. over the SLOW dimension of z [B][D]: it is an MN-major
//     operand.  For tf32 the only MN-major shared-memory layout tcgen05 accepts is SWIZZLE_128B with 32-byte atoms (descriptor
//     layout type 1; CUTLASS: "for mn-major tf32 operands, SW128_32B is the only available smem layout"): 128-byte rows of 32
//     consecutive columns, one row per K index, the four 32-byte pieces of a row XOR-ed with (row & 3); K groups of 4 rows are
//     SBO = 512 B apart, 32-column groups LBO apart.  A loader thread's float4 of four consecutive columns is one 16-byte store.
//     (Round 1 probed MN-major through the no-swizzle layout and got wrong products: that layout does not exist for tf32.)
#include "aa_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace {

constexpr int CM = 128;                    // tile rows (UMMA M) and columns (UMMA N)
constexpr int CK = 32;                     // K per chunk
constexpr int CPS = CM * 16 + 16;          // K-major panel stride in bytes (128 rows x 16 B; +16 keeps the 8 panels of a row off the same banks)
constexpr int CPL = (CK / 4) * CPS;        // one K-major plane of a chunk: 8 panels
constexpr int CPL_MN = CK * 128 * 4;       // one MN-major plane of a chunk: 4 column groups x 32 K rows x 128 B
constexpr int kCovStages = 3;
constexpr int kChainChunks = 2;            // K chunks per accumulator chain (24 MMAs)
constexpr int kLoadWarps = 8, kEpiWarps = 8;
constexpr int kCovThreads = 32 * (1 + kLoadWarps + kEpiWarps);
constexpr int kGramStage = 4 * CPL;        // A hi, A lo, B hi, B lo
constexpr int kGramSmem = 1024 + kCovStages * kGramStage + 1024;
constexpr int kOffX = (2 * CPL + 1023) / 1024 * 1024;     // backward stage: G hi, G lo (K-major), then at the next 1024-byte boundary Xc hi, Xc lo (MN-major)
constexpr int kBwdStage = kOffX + 2 * CPL_MN;             // a multiple of 1024: every stage's swizzled planes keep their alignment
constexpr int kBwdSmem = 1024 + kCovStages * kBwdStage + 1024;

__device__ __forceinline__ uint32_t c_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void c_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void c_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void c_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "CV_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra CV_DONE;\n"
      "bra CV_WAIT;\n"
      "CV_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void c_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void c_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void c_umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void c_umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void c_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor: 8-row groups 128 B apart (SBO), K core matrices `lbo` bytes apart
__device__ __forceinline__ uint64_t c_desc_k(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// MN-major tf32 operand, SWIZZLE_128B with 32-byte atoms (layout type 1): 32-column groups `lbo` bytes apart, 4-row K groups `sbo` apart
__device__ __forceinline__ uint64_t c_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ float c_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ void c_sts4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void c_split_store(uint32_t hi_addr, uint32_t lo_addr, float4 v) {
  const float4 h = make_float4(c_hi(v.x), c_hi(v.y), c_hi(v.z), c_hi(v.w));
  c_sts4(hi_addr, h);
  c_sts4(lo_addr, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
}

// Barriers at base: full[s] (one arrival per loader warp), empty[s] (tcgen05.commit), accfull[q], accempty[q] (one arrival per epilogue warp)
struct CovBars {
  uint32_t base;
  __device__ uint32_t full(int s) const { return base + 8u * s; }
  __device__ uint32_t empty(int s) const { return base + 32u + 8u * s; }
  __device__ uint32_t accfull(int q) const { return base + 64u + 8u * q; }
  __device__ uint32_t accempty(int q) const { return base + 80u + 8u * q; }
  __device__ uint32_t tmem_slot() const { return base + 96u; }
};

__device__ __forceinline__ uint32_t cov_prologue(const CovBars& B, int warp) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < kCovStages; ++s) { c_mbar_init(B.full(s), kLoadWarps); c_mbar_init(B.empty(s), 1); }
    for (int q = 0; q < 2; ++q) { c_mbar_init(B.accfull(q), 1); c_mbar_init(B.accempty(q), kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(B.tmem_slot()), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  c_fence_before();
  __syncthreads();
  c_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(B.tmem_slot()));
  return tmem_base;
}
__device__ __forceinline__ void cov_epilogue_end(uint32_t tmem_base, int warp) {
  c_fence_before();
  __syncthreads();
  if (warp == 0) {
    c_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// The MMA issuer of both kernels: per chunk 3 terms x 4 K steps; chains of kChainChunks chunks alternate between two TMEM slots.
// MNB: the B operand is MN-major (backward); otherwise K-major panels.
template <bool MNB>
__device__ __forceinline__ void cov_issue(const CovBars& B, uint32_t tmem_base, uint32_t stage0, uint32_t stage_bytes, int n_chunks,
                                          uint32_t offA_hi, uint32_t offA_lo, uint32_t offB_hi, uint32_t offB_lo) {
  // instruction descriptor: D = F32, A = B = TF32 (format 2), N = 128, M = 128; bit 16 = B is MN-major
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (MNB ? (1u << 16) : 0u) | ((uint32_t)(CM >> 3) << 17) | ((uint32_t)(CM >> 4) << 24);
  for (int i = 0; i < n_chunks; ++i) {
    const int s = i % kCovStages;
    const int chain = i / kChainChunks, q = chain & 1;
    const bool first = (i % kChainChunks) == 0;
    if (first && chain >= 2) {   // the epilogue has drained this TMEM slot
      c_mbar_wait(B.accempty(q), (uint32_t)(((chain >> 1) - 1) & 1));
      c_fence_after();
    }
    c_mbar_wait(B.full(s), (uint32_t)((i / kCovStages) & 1));
    c_fence_after();
    const uint32_t st = stage0 + (uint32_t)s * stage_bytes;
#pragma unroll
    for (int term = 0; term < 3; ++term) {   // a_lo b_hi, a_hi b_lo, a_hi b_hi
      const uint32_t ab = st + (term == 0 ? offA_lo : offA_hi), bb = st + (term == 1 ? offB_lo : offB_hi);
#pragma unroll
      for (int kk = 0; kk < CK / 8; ++kk) {
        const uint64_t da = c_desc_k(ab + (uint32_t)(2 * kk) * CPS, CPS);
        const uint64_t db = MNB ? c_desc_mn(bb + (uint32_t)kk * 1024u, CK * 128u, 512u) : c_desc_k(bb + (uint32_t)(2 * kk) * CPS, CPS);
        c_umma_tf32(tmem_base + (uint32_t)q * 128u, da, db, idesc, (first && term == 0 && kk == 0) ? 0u : 1u);
      }
    }
    c_umma_commit(B.empty(s));
    if ((i % kChainChunks) == kChainChunks - 1 || i == n_chunks - 1) c_umma_commit(B.accfull(q));
  }
}

// Epilogue warps: thread = (TMEM lane = tile row, half of the 128 columns); sums the chains into acc[64].
__device__ __forceinline__ void cov_drain(const CovBars& B, uint32_t tmem_base, int n_chunks, int quarter, int half, int lane, float (&acc)[64]) {
#pragma unroll
  for (int j = 0; j < 64; ++j) acc[j] = 0.f;
  const int n_chains = (n_chunks + kChainChunks - 1) / kChainChunks;
  for (int c = 0; c < n_chains; ++c) {
    const int q = c & 1;
    c_mbar_wait(B.accfull(q), (uint32_t)((c >> 1) & 1));
    c_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)q * 128u + (uint32_t)half * 64u;
    uint32_t v[32];
    c_tmem_ld32(taddr, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
    c_tmem_ld32(taddr + 32u, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[32 + j] += __uint_as_float(v[j]);
    c_fence_before();
    __syncwarp();
    if (lane == 0) c_mbar_arrive(B.accempty(q));
  }
}

// ------------------------------------------------------------------------------------------------ forward: G = Xc Xc^T
struct GramBBArgs {
  const float* z;       // [b][d]
  const float* mean;    // [d]
  int b;
  long long d, d_per_split;   // d_per_split % 32 == 0
  float* parts;         // [splits][b][b] (only rows >= cols tiles are written; gram_reduce_kernel mirrors)
};

__global__ void __launch_bounds__(kCovThreads, 1) gram_bb_tc_kernel(const GramBBArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (c_smem_u32(smem_raw) + 1023u) & ~1023u;
  const CovBars B{base};
  const uint32_t stage0 = base + 1024u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile (ti >= tj) from the linear lower-triangle index
  int ti = 0, rem = (int)blockIdx.x;
  while (rem > ti) { rem -= ti + 1; ++ti; }
  const int tj = rem;
  const bool diag = ti == tj;
  const int split = blockIdx.y;
  const long long k0 = (long long)split * a.d_per_split, k1 = min(a.d, k0 + a.d_per_split);
  const int n_chunks = k1 > k0 ? (int)((k1 - k0 + CK - 1) / CK) : 0;
  const uint32_t tmem_base = cov_prologue(B, warp);

  // (an empty split -- possible when the chunk count does not divide -- runs no chunk and writes a zero tile)
  if (warp == 0) {
    if (lane == 0) cov_issue<false>(B, tmem_base, stage0, kGramStage, n_chunks, 0u, CPL, diag ? 0u : 2u * CPL, diag ? (uint32_t)CPL : 3u * CPL);
  } else if (warp <= kLoadWarps) {
    // ===================== loaders: 256 threads; float4 = 4 consecutive d of one row; 4 (A) + 4 (B) float4 per thread and chunk =====================
    const int lt = threadIdx.x - 32;
    const int pn = lt & 7;              // panel: d offset 4 pn inside the chunk
    const int r0 = lt >> 3;             // rows r0 + 32 u
    constexpr int kDepth = 2;
    float4 va[kDepth][4], vb[kDepth][4], mu[kDepth];
    auto issue = [&](int slot, int i) {
      const long long dd = k0 + (long long)i * CK + 4 * pn;
      const bool ok = i < n_chunks && dd < k1;       // d % 4 == 0 and d_per_split % 32 == 0: a float4 is inside or outside as a whole
      mu[slot] = ok ? __ldg(reinterpret_cast<const float4*>(a.mean + dd)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ra = ti * CM + r0 + 32 * u, rb = tj * CM + r0 + 32 * u;
        va[slot][u] = (ok && ra < a.b) ? __ldg(reinterpret_cast<const float4*>(a.z + (long long)ra * a.d + dd)) : mu[slot];
        if (!diag) vb[slot][u] = (ok && rb < a.b) ? __ldg(reinterpret_cast<const float4*>(a.z + (long long)rb * a.d + dd)) : mu[slot];
      }
    };
    auto process = [&](int slot, int i) {
      const int s = i % kCovStages;
      if (i >= kCovStages) c_mbar_wait(B.empty(s), (uint32_t)(((i / kCovStages) - 1) & 1));
      const uint32_t st = stage0 + (uint32_t)s * kGramStage + (uint32_t)pn * CPS;
      const float4 m = mu[slot];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t row = (uint32_t)(r0 + 32 * u) * 16u;
        const float4 x = va[slot][u];
        c_split_store(st + row, st + CPL + row, make_float4(x.x - m.x, x.y - m.y, x.z - m.z, x.w - m.w));
        if (!diag) {
          const float4 y = vb[slot][u];
          c_split_store(st + 2u * CPL + row, st + 3u * CPL + row, make_float4(y.x - m.x, y.y - m.y, y.z - m.z, y.w - m.w));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the operands are read by the tensor core
      __syncwarp();
      if (lane == 0) c_mbar_arrive(B.full(s));
    };
    issue(0, 0);
    for (int i = 0; i < n_chunks; i += 2) {
      issue(1, i + 1);
      process(0, i);
      if (i + 1 < n_chunks) { issue(0, i + 2); process(1, i + 1); }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 1 - kLoadWarps;          // 0..7
    const int quarter = warp & 3, half = ew >> 2;  // consecutive warps cover the four TMEM lane quarters
    float acc[64];
    cov_drain(B, tmem_base, n_chunks, quarter, half, lane, acc);
    const int r = ti * CM + quarter * 32 + lane, c0 = tj * CM + half * 64;
    if (r < a.b) {
      float* dst = a.parts + ((long long)split * a.b + r) * a.b + c0;
#pragma unroll
      for (int j = 0; j < 64; ++j)
        if (c0 + j < a.b) dst[j] = acc[j];
    }
  }
  cov_epilogue_end(tmem_base, warp);
}

// ------------------------------------------------------------------------------------------------ backward: dX = k (G Xc - Xc diag(s))
struct CovBwdArgs {
  const float* z;       // [b][d]
  const float* stats;   // mean [d], unbiased variance [d]
  const float* gram;    // [b][b]
  int b;
  long long d;
  const float* gloss;   // scalar or NULL
  float gscale;
  float* gz;            // [b][d]
  int accumulate;
  uint32_t mn_lbo, mn_sbo;   // (descriptor fields are fixed in cov_issue; kept for the probe build)
};

__global__ void __launch_bounds__(kCovThreads, 1) cov_bwd_tc_kernel(const CovBwdArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (c_smem_u32(smem_raw) + 1023u) & ~1023u;
  const CovBars B{base};
  const uint32_t stage0 = base + 1024u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long c0 = (long long)blockIdx.x * CM;   // column tile
  const int tr = blockIdx.y;                         // row tile
  const int n_chunks = (a.b + CK - 1) / CK;
  constexpr uint32_t OFF_XHI = kOffX;
  static_assert(kBwdStage % 1024 == 0, "stage layout");
  const uint32_t tmem_base = cov_prologue(B, warp);

  if (warp == 0) {
    if (lane == 0) cov_issue<true>(B, tmem_base, stage0, kBwdStage, n_chunks, 0u, CPL, OFF_XHI, OFF_XHI + CPL_MN);
  } else if (warp <= kLoadWarps) {
    // ===================== loaders: per chunk (32 batch rows j) 4 float4 of G (rows of the tile, 4 consecutive j) and 4 float4 of Xc =====================
    const int lt = threadIdx.x - 32;
    // G: panel pn = 4 consecutive j, rows r0 + 32 u of the row tile
    const int pn = lt & 7, r0 = lt >> 3;
    // Xc: thread = (batch row j = jq + 8 u, float4 q of the 128-column tile): 32 lanes = one 512-byte row segment
    const int q = lt & 31, jq = lt >> 5;
    const long long col = c0 + 4 * q;
    const bool col_ok = col < a.d;
    const float4 mu = col_ok ? __ldg(reinterpret_cast<const float4*>(a.stats + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int kDepth = 2;
    float4 vg[kDepth][4], vx[kDepth][4];
    auto issue = [&](int slot, int i) {
      const int j4 = i * CK + 4 * pn;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = tr * CM + r0 + 32 * u;
        vg[slot][u] = (i < n_chunks && r < a.b && j4 < a.b) ? __ldg(reinterpret_cast<const float4*>(a.gram + (long long)r * a.b + j4))
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);   // b % 4 == 0
        const int j = i * CK + jq + 8 * u;
        vx[slot][u] = (i < n_chunks && col_ok && j < a.b) ? __ldg(reinterpret_cast<const float4*>(a.z + (long long)j * a.d + col)) : mu;
      }
    };
    auto process = [&](int slot, int i) {
      const int s = i % kCovStages;
      if (i >= kCovStages) c_mbar_wait(B.empty(s), (uint32_t)(((i / kCovStages) - 1) & 1));
      const uint32_t st = stage0 + (uint32_t)s * kBwdStage;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t ga = st + (uint32_t)pn * CPS + (uint32_t)(r0 + 32 * u) * 16u;
        c_split_store(ga, ga + CPL, vg[slot][u]);
        // MN-major SW128 / 32-byte atoms: column group q >> 3, K row j, 16-byte piece q & 7: 32-byte piece ((q & 7) >> 1) ^ (j & 3)
        const uint32_t j = (uint32_t)(jq + 8 * u);
        const uint32_t xa = st + OFF_XHI + (uint32_t)(q >> 3) * (CK * 128u) + j * 128u + (((((uint32_t)q & 7u) >> 1) ^ (j & 3u)) << 5) + (((uint32_t)q & 1u) << 4);
        const float4 x = vx[slot][u];
        c_split_store(xa, xa + CPL_MN, make_float4(x.x - mu.x, x.y - mu.y, x.z - mu.z, x.w - mu.w));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) c_mbar_arrive(B.full(s));
    };
    issue(0, 0);
    for (int i = 0; i < n_chunks; i += 2) {
      issue(1, i + 1);
      process(0, i);
      if (i + 1 < n_chunks) { issue(0, i + 2); process(1, i + 1); }
    }
  } else {
    // ===================== epilogue: thread = (row of the tile, 64 columns) =====================
    const int ew = warp - 1 - kLoadWarps;
    const int quarter = warp & 3, half = ew >> 2;
    float acc[64];
    cov_drain(B, tmem_base, n_chunks, quarter, half, lane, acc);
    const int r = tr * CM + quarter * 32 + lane;
    const long long cc = c0 + half * 64;
    if (r < a.b) {
      const float kf = 4.0f * a.gscale * (a.gloss ? a.gloss[0] : 1.0f) / ((float)(a.b - 1) * (float)(a.b - 1) * (float)a.d);
      const float bm1 = (float)(a.b - 1);
      const float* zr = a.z + (long long)r * a.d + cc;
      float* gr = a.gz + (long long)r * a.d + cc;
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        if (cc + j < a.d) {   // d % 4 == 0
          const float4 zz = __ldg(reinterpret_cast<const float4*>(zr + j));
          const float4 m = __ldg(reinterpret_cast<const float4*>(a.stats + cc + j));
          const float4 v = __ldg(reinterpret_cast<const float4*>(a.stats + a.d + cc + j));
          float4 o;
          o.x = kf * (acc[j] - bm1 * v.x * (zz.x - m.x));
          o.y = kf * (acc[j + 1] - bm1 * v.y * (zz.y - m.y));
          o.z = kf * (acc[j + 2] - bm1 * v.z * (zz.z - m.z));
          o.w = kf * (acc[j + 3] - bm1 * v.w * (zz.w - m.w));
          if (a.accumulate) {
            const float4 p = *reinterpret_cast<const float4*>(gr + j);
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
          }
          *reinterpret_cast<float4*>(gr + j) = o;
        }
      }
    }
  }
  cov_epilogue_end(tmem_base, warp);
}

}  // namespace

namespace aa {

static bool cov_tc_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("AA_COV_TC") ? atoi(getenv("AA_COV_TC")) : 1;
  return v != 0;
}

// Shapes the tcgen05 kernels take: float4-aligned rows, enough work to fill the tiles.
bool cov_tc_eligible(const float* z, const float* stats, int64_t b, int64_t d) {
  return cov_tc_enabled() && b >= 64 && d >= 1024 && d % 4 == 0 && ((uintptr_t)z & 15) == 0 && ((uintptr_t)stats & 15) == 0;
}

int cov_tc_splits(int64_t b, int64_t d) {
  const int64_t nt = (b + CM - 1) / CM, tiles = nt * (nt + 1) / 2;
  const int64_t chunks = (d + CK - 1) / CK;
  int64_t s = std::max<int64_t>(1, num_sms() / tiles);
  s = std::min<int64_t>(s, std::max<int64_t>(1, chunks / 4));
  return (int)std::min<int64_t>(s, 65535);
}

// parts[split][b][b] partial Gram tiles (lower-triangle tiles), to be summed by gram_reduce_kernel
int gram_bb_tc(const float* z, const float* mean, int64_t b, int64_t d, float* parts, int splits, cudaStream_t stream) {
  AA_CUDA(aa::ensure_dyn_smem(gram_bb_tc_kernel, kGramSmem));
  GramBBArgs a;
  a.z = z; a.mean = mean; a.b = (int)b; a.d = d; a.parts = parts;
  const long long chunks = (d + CK - 1) / CK;
  a.d_per_split = (chunks + splits - 1) / splits * CK;
  const int nt = (int)((b + CM - 1) / CM);
  gram_bb_tc_kernel<<<dim3(nt * (nt + 1) / 2, splits), kCovThreads, kGramSmem, stream>>>(a);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int cov_bwd_tc(const float* z, const float* stats, const float* gram, int64_t b, int64_t d, const float* gloss, float gscale, float* gz,
               int accumulate, cudaStream_t stream) {
  AA_CUDA(aa::ensure_dyn_smem(cov_bwd_tc_kernel, kBwdSmem));
  CovBwdArgs a;
  a.z = z; a.stats = stats; a.gram = gram; a.b = (int)b; a.d = d; a.gloss = gloss; a.gscale = gscale; a.gz = gz; a.accumulate = accumulate;
  a.mn_lbo = CK * 128; a.mn_sbo = 512;
  const long long ct = (d + CM - 1) / CM;
  AA_REQUIRE(ct < (1LL << 31), "d too large");
  cov_bwd_tc_kernel<<<dim3((unsigned)ct, (unsigned)((b + CM - 1) / CM)), kCovThreads, kBwdSmem, stream>>>(a);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

}  // namespace aa
