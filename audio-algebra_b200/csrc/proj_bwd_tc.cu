// AudioAlgebra projector half (4 EmbedBlocks, aa_mixer.py:205-260) BACKWARD on tcgen05, fp32-accurate, for the standard
// 64 -> 64 -> 64 -> 64 -> 64 residual configuration (the forward is proj_tc.cu):
//
//   recompute  u_l = h_l W_l^T + b_l,  h_{l+1} = h_l + act_l(u_l)            (l = 0..2; act = GELU(erf); layer 3 has no activation)
//   backward   du_l = g_{l+1} * act_l'(u_l),  dW_l += du_l^T h_l,  db_l += sum_tok du_l,  g_l = g_{l+1} + du_l W_l,  gx = g_0 + gout
//
// Three GEMM shapes per layer and 128-token tile, all kind::f16 with bf16 operands and fp32 accumulation in TMEM:
//   (a) forward recompute   [128 tok x 64] x W_l^T      A = h_l   K-major (K = in features),  B = W_l  K-major  (rows = out)
//   (b) data gradient       [128 tok x 64] x W_l        A = du_l  K-major (K = out features), B = W_l  MN-major (rows = out = K)
//   (c) weight gradient     du_l^T [64 x 128 tok] x h_l A = du_l  MN-major (rows = tok = K),  B = h_l  MN-major (rows = tok = K)
// Every fp32 value is split into THREE bf16 pieces v = p0 + p1 + p2 (24 mantissa bits) and the six products p_i q_j with
// i + j <= 2 are accumulated (the dropped ones are <= 2^-24 relative): the fp32 parity gates hold (measured against float64 in the
// tests) at the tensor-pipe cost of a 3-term TF32 split.  Why bf16 pieces and not TF32 hi / lo as in the forward: for 16-bit types
// the K-major and the MN-major SWIZZLE_128B layouts are the SAME bytes (128-byte rows, 16-byte pieces XOR-ed with row & 7), so ONE
// shared-memory copy of du_l serves (b) as a K-major and (c) as an MN-major operand, and one copy of W_l serves (a) and (b): all
// four layers' weights stay resident (96 KB) next to du (4 planes: 3 pieces + a zero plane) and h (3 planes) -- 213 KB.  With TF32 the
// only MN-major layout is the 32-byte-atom swizzle (cov_tc.cu), which no K-major layout matches: du and W would be needed twice, 260 KB.
// (c) runs with M = 128: A = [du_p0 ; du_p1] (two 64-feature MN groups LBO apart) against h_p0, h_p1, h_p2, then [du_p2 ; 0] against h_p0;
// accumulator rows 0..63 + rows 64..127 = dW_l, kept in TMEM (4 x 64 columns) across tiles and drained every kDrainTiles tiles.
// One CTA per SM: warp 0 issues the MMAs; 16 warps = 4 TMEM lane quarters x 4 feature quarters (thread = one token x 16 features)
// keep h and g in registers and park u_0..2 in free TMEM columns of their own lane (h_l is rebuilt on the way back as
// h_{l+1} - GELU(u_l)), write the split operands, apply
// bias / GELU / GELU' to the accumulator and reduce the bias gradients with warp shuffles in a fixed order.
#include "aa_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace {

constexpr int BM = 128;                  // tokens per tile (UMMA M of (a), (b); K of (c))
constexpr int BD = 64;                   // features
constexpr int BSPLIT = 4;                // threads per token
constexpr int BCP = BD / BSPLIT;         // features per thread
constexpr int kBwdTcThreads = 32 + BM * BSPLIT;
constexpr int PLANE_T = BM * 128;        // a token-row plane: 128 rows x 128 B (64 bf16)
constexpr int PLANE_W = BD * 128;        // a weight plane: 64 rows (out) x 128 B (64 in)
constexpr int OFF_BIAS = 64;             // b[4][64] fp32
constexpr int OFF_DB = 2048;             // bias-gradient slots [16 warps][4 layers][16] fp32
constexpr int OFF_W = 6144;              // [4 layers][3 pieces] weight planes (1024-byte aligned)
constexpr int OFF_DU = OFF_W + 4 * 3 * PLANE_W;   // du pieces 0..2 + a zero plane
constexpr int OFF_H = OFF_DU + 4 * PLANE_T;       // h pieces 0..2
constexpr int kBwdTcSmem = OFF_H + 3 * PLANE_T + 1024;
constexpr int kDrainTiles = 4;           // weight-gradient accumulators are drained every kDrainTiles tiles: tcgen05 accumulates with truncation, so the
                                         // error of a chain grows with its length (one chain over 14 tiles: 8e-6 against float64, every 2 tiles 1.2e-6 at +0.08 ms per call, every 4: see DESIGN.md 4.3)
constexpr int kPartFloats = 4 * BM * BD + 4 * 4 * BD;   // per CTA: dW accumulators [4][128][64], bias partials [4][4 quarters][64]
static_assert(OFF_W % 1024 == 0 && OFF_DU % 1024 == 0 && OFF_H % 1024 == 0, "swizzle atoms need 1024-byte alignment");

struct ProjBwdTcArgs {
  const float* w[4];
  const float* b[4];
  const float* x;
  const float* gout;
  float* gx;            // may be NULL
  int accumulate_gx;
  float* partials;      // [grid][kPartFloats]
  long long n_tiles;
  int t, tiles_t;
};

__device__ __forceinline__ uint32_t b_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void b_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void b_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void b_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra PB_DONE;\n"
      "bra PB_WAIT;\n"
      "PB_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void b_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void b_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void b_umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void b_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void b_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void b_tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// SWIZZLE_128B descriptors (layout type 2), 8-row groups 1024 B apart (SBO).  K-major: rows = M / N index, K along the 128-byte row.
__device__ __forceinline__ uint64_t b_desc_k(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major: rows = K index, 64 M / N elements along the 128-byte row; further 64-element M / N groups `lbo` bytes apart.
__device__ __forceinline__ uint64_t b_desc_mn(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t b_pack(float lo, float hi) {   // {hi : bits 31..16, lo : bits 15..0}, round to nearest even
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// v[0..7] -> three 16-byte pieces (8 bf16 each): v = p0 + p1 + p2 up to 2^-24 relative
__device__ __forceinline__ void b_split8(const float* v, uint4& p0, uint4& p1, uint4& p2) {
  uint32_t a[4], b[4], c[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float x0 = v[2 * q], x1 = v[2 * q + 1];
    a[q] = b_pack(x0, x1);
    const float r0 = x0 - __uint_as_float(a[q] << 16), r1 = x1 - __uint_as_float(a[q] & 0xffff0000u);
    b[q] = b_pack(r0, r1);
    const float s0 = r0 - __uint_as_float(b[q] << 16), s1 = r1 - __uint_as_float(b[q] & 0xffff0000u);
    c[q] = b_pack(s0, s1);
  }
  p0 = make_uint4(a[0], a[1], a[2], a[3]);
  p1 = make_uint4(b[0], b[1], b[2], b[3]);
  p2 = make_uint4(c[0], c[1], c[2], c[3]);
}
__device__ __forceinline__ void b_sts16(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// this thread's 16 features of token row r -> the three piece planes of `buf` (pieces 2 split, 2 split + 1 of the row, swizzled)
__device__ __forceinline__ void b_write_planes(uint32_t buf, int r, int split, const float (&v)[BCP]) {
#pragma unroll
  for (int hq = 0; hq < 2; ++hq) {
    uint4 p0, p1, p2;
    b_split8(v + 8 * hq, p0, p1, p2);
    const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)(2 * split + hq)) ^ ((uint32_t)r & 7u)) << 4);
    b_sts16(buf + off, p0);
    b_sts16(buf + PLANE_T + off, p1);
    b_sts16(buf + 2 * PLANE_T + off, p2);
  }
}
// Column sums of v[16] over the 32 lanes in a fixed order: the result for feature f = 8 b4 + 4 b3 + 2 b2 + b1 (b_k = bit k of the lane)
// is returned in every lane (lanes differing in bit 0 hold the same value).
__device__ __forceinline__ float b_colsum16(const float (&v)[BCP], int lane) {
  float a8[8], a4[4], a2[2];
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float send = u16 ? v[j] : v[j + 8], keep = u16 ? v[j + 8] : v[j];
    a8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = u8 ? a8[j] : a8[j + 4], keep = u8 ? a8[j + 4] : a8[j];
    a4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = u4 ? a4[j] : a4[j + 2], keep = u4 ? a4[j + 2] : a4[j];
    a2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = u2 ? a2[0] : a2[1], keep = u2 ? a2[1] : a2[0];
  const float a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  return a1 + __shfl_xor_sync(0xffffffffu, a1, 1);
}

__global__ void __launch_bounds__(kBwdTcThreads, 1) proj_bwd_tc_kernel(const ProjBwdTcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (b_smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* gbase = smem_raw + (base - b_smem_u32(smem_raw));
  const uint32_t afull = base, accfull = base + 8, wdone = base + 24, tmem_slot = base + 16;
  const uint32_t sbias = base + OFF_BIAS, sdb = base + OFF_DB, sW = base + OFF_W, sDU = base + OFF_DU, sH = base + OFF_H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    b_mbar_init(afull, 4 * BSPLIT);
    b_mbar_init(accfull, 1);
    b_mbar_init(wdone, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident weights: W_l[o][i] -> piece planes [l][piece][o][64 i] (SWIZZLE_128B rows); biases; zero plane; bias-gradient slots
  for (int idx = threadIdx.x; idx < 4 * BD * 8; idx += kBwdTcThreads) {
    const int l = idx / (BD * 8), rem = idx - l * (BD * 8);
    const int o = rem >> 3, c = rem & 7;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.w[l] + o * BD + 8 * c));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.w[l] + o * BD + 8 * c) + 1);
    const float v[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    uint4 p0, p1, p2;
    b_split8(v, p0, p1, p2);
    const uint32_t off = (uint32_t)(l * 3) * PLANE_W + (uint32_t)o * 128u + ((((uint32_t)c) ^ ((uint32_t)o & 7u)) << 4);
    b_sts16(sW + off, p0);
    b_sts16(sW + PLANE_W + off, p1);
    b_sts16(sW + 2 * PLANE_W + off, p2);
  }
  for (int i = threadIdx.x; i < 4 * BD; i += kBwdTcThreads)
    reinterpret_cast<float*>(gbase + OFF_BIAS)[i] = __ldg(a.b[i / BD] + (i % BD));
  for (int i = threadIdx.x; i < PLANE_T / 16; i += kBwdTcThreads) b_sts16(sDU + 3 * PLANE_T + 16u * i, make_uint4(0u, 0u, 0u, 0u));
  for (int i = threadIdx.x; i < 16 * 4 * 16; i += kBwdTcThreads) reinterpret_cast<float*>(gbase + OFF_DB)[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  b_fence_before();
  __syncthreads();
  b_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = BF16, N = 64, M = 128; bit 15 / 16: A / B is MN-major
      const uint32_t id0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BD >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t id_kk = id0, id_kmn = id0 | (1u << 16), id_mnmn = id0 | (1u << 15) | (1u << 16);
      // the six piece products, smallest first: (a piece, b piece)
      const int ta[6] = {2, 1, 0, 1, 0, 0}, tb[6] = {0, 1, 2, 0, 1, 0};
      uint32_t n = 0;
      int it = 0;
      for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const bool first_tile = (it % kDrainTiles) == 0;   // the epilogue drained the weight-gradient accumulators after the previous tile
        for (int l = 0; l < 3; ++l, ++n) {   // (a) forward recompute: acc = h_l W_l^T
          b_mbar_wait(afull, n & 1u);
          b_fence_after();
          const uint32_t wl = sW + (uint32_t)(l * 3) * PLANE_W;
#pragma unroll
          for (int term = 0; term < 6; ++term)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              b_umma(tmem_base, b_desc_k(sH + (uint32_t)ta[term] * PLANE_T + 32u * kk), b_desc_k(wl + (uint32_t)tb[term] * PLANE_W + 32u * kk), id_kk,
                     (term | kk) != 0 ? 1u : 0u);
          b_commit(accfull);
        }
        for (int l = 3; l >= 0; --l, ++n) {
          b_mbar_wait(afull, n & 1u);
          b_fence_after();
          const uint32_t wl = sW + (uint32_t)(l * 3) * PLANE_W;
          // (b) data gradient: acc = du_l W_l  (K = out features: 16 weight rows per step)
#pragma unroll
          for (int term = 0; term < 6; ++term)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              b_umma(tmem_base, b_desc_k(sDU + (uint32_t)ta[term] * PLANE_T + 32u * kk),
                     b_desc_mn(wl + (uint32_t)tb[term] * PLANE_W + 2048u * kk, PLANE_W), id_kmn, (term | kk) != 0 ? 1u : 0u);
          b_commit(accfull);   // the token warps only need g_l to go on; the weight-gradient MMAs below run under their GELU' arithmetic
          // (c) weight gradient: accW_l += [du_p0 ; du_p1]^T (h_p0 + h_p1 + h_p2) + [du_p2 ; 0]^T h_p0   (K = 128 tokens: 16 rows per step)
          const uint32_t tw = tmem_base + 64u + 64u * (uint32_t)l;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t a01 = b_desc_mn(sDU + 2048u * kk, PLANE_T), a2z = b_desc_mn(sDU + 2u * PLANE_T + 2048u * kk, PLANE_T);
            b_umma(tw, a2z, b_desc_mn(sH + 2048u * kk, PLANE_T), id_mnmn, (first_tile && kk == 0) ? 0u : 1u);
            b_umma(tw, a01, b_desc_mn(sH + 2u * PLANE_T + 2048u * kk, PLANE_T), id_mnmn, 1u);
            b_umma(tw, a01, b_desc_mn(sH + PLANE_T + 2048u * kk, PLANE_T), id_mnmn, 1u);
            b_umma(tw, a01, b_desc_mn(sH + 2048u * kk, PLANE_T), id_mnmn, 1u);
          }
          b_commit(wdone);     // du_l / h_l planes may be overwritten, accW_l may be drained
        }
      }
    }
  } else {
    // ===================== token rows: thread = one token of the tile x 16 features =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int split = (warp - 1) >> 2;            // feature quarter
    const int r = quarter * 32 + lane, c_lo = split * BCP;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c_lo;
    const float* bias = reinterpret_cast<const float*>(gbase + OFF_BIAS);
    float* dbs = reinterpret_cast<float*>(gbase + OFF_DB) + (warp - 1) * (4 * BCP);
    const int fdb = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);   // feature b_colsum16 leaves in this lane
    uint32_t n = 0, nw = 0;
    float* part = a.partials + (size_t)blockIdx.x * kPartFloats;
    auto hand_over = [&]() {   // operands written -> tensor core; returns when the accumulator of this phase is complete
      b_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) b_mbar_arrive(afull);
      b_mbar_wait(accfull, n & 1u);
      b_fence_after();
      ++n;
    };
    int it = 0;
    for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const long long bi = tile / a.tiles_t;
      const int t0 = (int)(tile - bi * a.tiles_t) * BM;
      const bool valid = t0 + r < a.t;
      const long long goff = (bi * BD + c_lo) * (long long)a.t + t0 + r;
      float h[BCP], g[BCP];
#pragma unroll
      for (int c = 0; c < BCP; ++c) h[c] = valid ? __ldg(a.x + goff + (long long)c * a.t) : 0.f;
#pragma unroll 1
      for (int l = 0; l < 3; ++l) {
        b_write_planes(sH, r, split, h);
        hand_over();
        uint32_t v[BCP];
        b_tmem_ld16(taddr, v);
#pragma unroll
        for (int j = 0; j < BCP; ++j) {
          const float uu = __uint_as_float(v[j]) + bias[l * BD + c_lo + j];
          v[j] = __float_as_uint(uu);
          h[j] += 0.5f * uu * (1.0f + erff(uu * 0.70710678118654752440f));
        }
        b_tmem_st16(taddr + 320u + 64u * (uint32_t)l, v);   // u_l parked in this token's TMEM lane (columns 320.. are free) until the way back
      }
      // h = h_3 (input of the last block); g = dL/dh_4 = gout
#pragma unroll
      for (int c = 0; c < BCP; ++c) g[c] = valid ? __ldg(a.gout + goff + (long long)c * a.t) : 0.f;
#pragma unroll 1
      for (int l = 3; l >= 0; --l) {
        float du[BCP];
        if (l == 3) {
#pragma unroll
          for (int j = 0; j < BCP; ++j) du[j] = g[j];
        } else {
          uint32_t uv[BCP];
          b_tmem_ld16(taddr + 320u + 64u * (uint32_t)l, uv);
#pragma unroll
          for (int j = 0; j < BCP; ++j) {
            const float uu = __uint_as_float(uv[j]);
            const float cdf = 0.5f * (1.0f + erff(uu * 0.70710678118654752440f));
            const float pdf = 0.39894228040143267794f * __expf(-0.5f * uu * uu);
            du[j] = g[j] * (cdf + uu * pdf);
            h[j] -= uu * cdf;   // h_l = h_{l+1} - GELU(u_l)
          }
        }
        const float cs = b_colsum16(du, lane);
        if ((lane & 1) == 0) dbs[l * BCP + fdb] += cs;   // this warp's 32 tokens, accumulated over the CTA's tiles in tile order
        if (l < 3) { b_mbar_wait(wdone, nw & 1u); ++nw; }   // the previous layer's weight-gradient MMAs have read the planes
        b_write_planes(sDU, r, split, du);
        b_write_planes(sH, r, split, h);
        hand_over();
        uint32_t v[BCP];
        b_tmem_ld16(taddr, v);
#pragma unroll
        for (int j = 0; j < BCP; ++j) g[j] += __uint_as_float(v[j]);
      }
      if (a.gx && valid) {   // dL/dx = dL/dh_0 + gout (outer residual)
#pragma unroll
        for (int c = 0; c < BCP; ++c) {
          const long long o = goff + (long long)c * a.t;
          float gv = g[c] + __ldg(a.gout + o);
          if (a.accumulate_gx) gv += a.gx[o];
          a.gx[o] = gv;
        }
      }
      b_mbar_wait(wdone, nw & 1u);   // layer 0's weight-gradient MMAs: planes free for the next tile, accumulators complete
      ++nw;
      b_fence_after();
      // every MMA issued so far is complete: every kDrainTiles tiles (and after the CTA's last tile) the four
      // weight-gradient accumulators are added into this CTA's partial sums (same thread, same order every run)
      if ((it % kDrainTiles) == kDrainTiles - 1 || tile + gridDim.x >= a.n_tiles) {
        const bool first_drain = it < kDrainTiles;
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
          uint32_t v[BCP];
          b_tmem_ld16(taddr + 64u + 64u * (uint32_t)l, v);
          float4* dst = reinterpret_cast<float4*>(part + ((size_t)l * BM + r) * BD + c_lo);
#pragma unroll
          for (int j = 0; j < BCP; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (!first_drain) { const float4 p = dst[j / 4]; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
            dst[j / 4] = o;
          }
        }
      }
    }
    __syncwarp();
    for (int i = lane; i < 4 * BCP; i += 32) {
      const int l = i / BCP, f = i % BCP;
      part[4 * BM * BD + (l * 4 + quarter) * BD + c_lo + f] = dbs[l * BCP + f];
    }
  }
  b_fence_before();
  __syncthreads();
  if (warp == 0) {
    b_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

struct ProjBwdTcReduceArgs {
  float* gw[4];
  float* gb[4];
  const float* partials;
  int n_parts;
  float scale;
  int accumulate;
};
// gw_l[o][i] = scale * sum_cta (acc_l[o][i] + acc_l[64 + o][i]);  gb_l[o] = scale * sum_cta sum_quarter.  A block owns 32 outputs
// (lane) and its 8 warps split the CTAs' partials; the 8 partial sums are added in warp order: bitwise reproducible.
__global__ void __launch_bounds__(256) proj_bwd_tc_reduce_kernel(const ProjBwdTcReduceArgs r) {
  __shared__ float sh[8][32];
  const int per = BD * BD + BD;
  const int lane = threadIdx.x & 31, wg = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + lane;
  float s = 0.f;
  int l = 0, e = 0;
  if (idx < 4 * per) {
    l = idx / per; e = idx % per;
    if (e < BD * BD) {
      const int o = e / BD, i = e % BD;
      for (int p = wg; p < r.n_parts; p += 8) {
        const float* q = r.partials + (size_t)p * kPartFloats + (size_t)l * BM * BD;
        s += q[o * BD + i] + q[(BD + o) * BD + i];
      }
    } else {
      const int o = e - BD * BD;
      for (int p = wg; p < r.n_parts; p += 8) {
        const float* q = r.partials + (size_t)p * kPartFloats + 4 * BM * BD + (size_t)l * 4 * BD;
        s += (q[o] + q[BD + o]) + (q[2 * BD + o] + q[3 * BD + o]);
      }
    }
  }
  sh[wg][lane] = s;
  __syncthreads();
  if (wg != 0 || idx >= 4 * per) return;
  s = ((sh[0][lane] + sh[1][lane]) + (sh[2][lane] + sh[3][lane])) + ((sh[4][lane] + sh[5][lane]) + (sh[6][lane] + sh[7][lane]));
  s *= r.scale;
  if (e < BD * BD) {
    if (r.gw[l]) r.gw[l][e] = r.accumulate ? r.gw[l][e] + s : s;
  } else {
    const int o = e - BD * BD;
    if (r.gb[l]) r.gb[l][o] = r.accumulate ? r.gb[l][o] + s : s;
  }
}

}  // namespace

namespace aa {

int64_t proj_bwd_tc_workspace_floats() { return (int64_t)num_sms() * kPartFloats; }

bool proj_bwd_tc_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("AA_PROJ_BWD_TC") ? atoi(getenv("AA_PROJ_BWD_TC")) : 1;
  return v != 0;
}

// Backward of one projector half on the tensor core; requires dims = hidden = 64 with residuals, 16-byte aligned weights.
int proj_bwd_tc(const float* const* w, const float* const* b, const float* x, const float* gout, int64_t batch, int64_t t, float* gx,
                int accumulate_gx, float* const* gw, float* const* gb, int accumulate_gw, float gscale, float* workspace,
                cudaStream_t stream) {
  AA_CUDA(aa::ensure_dyn_smem(proj_bwd_tc_kernel, kBwdTcSmem));
  ProjBwdTcArgs a;
  for (int l = 0; l < 4; ++l) { a.w[l] = w[l]; a.b[l] = b[l]; }
  a.x = x; a.gout = gout; a.gx = gx; a.accumulate_gx = accumulate_gx; a.partials = workspace; a.t = (int)t;
  a.tiles_t = (int)((t + BM - 1) / BM);
  a.n_tiles = batch * a.tiles_t;
  if (a.n_tiles == 0) return AA_OK;
  const int grid = (int)std::min<long long>(a.n_tiles, (long long)num_sms());
  proj_bwd_tc_kernel<<<grid, kBwdTcThreads, kBwdTcSmem, stream>>>(a);
  AA_LAUNCH_CHECK();
  ProjBwdTcReduceArgs r;
  for (int l = 0; l < 4; ++l) { r.gw[l] = gw[l]; r.gb[l] = gb[l]; }
  r.partials = workspace; r.n_parts = grid; r.scale = gscale; r.accumulate = accumulate_gw;
  proj_bwd_tc_reduce_kernel<<<(4 * (BD * BD + BD) + 31) / 32, 256, 0, stream>>>(r);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

}  // namespace aa
