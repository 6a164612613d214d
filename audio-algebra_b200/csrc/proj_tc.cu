// AudioAlgebra projector half (4 EmbedBlocks, aa_mixer.py:205-260) FORWARD on tcgen05, fp32-accurate:
//
//     h0 = x^T;  h_{l+1} = h_l + act_l(h_l W_l^T + b_l)  (GELU(erf) for l < 3, identity for l = 3);  out = x + h_4^T
//
// for the standard 64 -> 64 -> 64 -> 64 -> 64 residual configuration.  Each Linear is a [128 tokens x 64] x [64 x 64] GEMM on
// the tensor core with kind::tf32; to keep the fp32 parity gates (1e-4) every operand is split into hi = its top 19 bits
// (exactly representable in TF32) and lo = value - hi, and three products are accumulated in fp32 in TMEM:
// a_lo w_hi + a_hi w_lo + a_hi w_hi (the dropped a_lo w_lo term is ~2^-22 relative).
// Operands sit in shared memory in the no-swizzle K-major panel layout [K/4][rows][4 x fp32 = 16 B] (conv_ru.cuh); the
// hi / lo weights of all four layers stay resident (130 KB).  One CTA per SM: warp 0 issues the MMAs, warps 1..4 own the 128
// token rows of the tile: warps 1..16 = 4 TMEM lane quarters x 4 feature quarters (thread = one token x 16 features): they keep
// h (and x for the outer residual) in registers, write the split A operand, and apply bias + GELU + the inner residual to the
// accumulator they read back from TMEM.
#include "aa_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace {

constexpr int PM = 128;                  // tokens per tile (UMMA M)
constexpr int PD = 64;                   // features
constexpr int PSA = PM * 16 + 16;        // A panel stride in bytes (+16: 16 panels written per row do not share banks)
constexpr int PSW = PD * 16 + 16;        // W panel stride
constexpr int A_BYTES = (PD / 4) * PSA;  // one A operand (hi or lo): 16 panels
constexpr int W_BYTES = (PD / 4) * PSW;  // one layer's weights (hi or lo)
constexpr int OFF_BIAS = 64;
constexpr int OFF_WHI = OFF_BIAS + 4 * PD * 4;
constexpr int OFF_WLO = OFF_WHI + 4 * W_BYTES;
constexpr int OFF_AHI = OFF_WLO + 4 * W_BYTES;
constexpr int OFF_ALO = OFF_AHI + A_BYTES;
constexpr int kProjTcSmem = OFF_ALO + A_BYTES + 128;
constexpr int NSPLIT = 4;                // threads per token: each owns PD / NSPLIT = 16 features (16 epilogue warps hide the erf latency)
constexpr int CP = PD / NSPLIT;
constexpr int kProjTcThreads = 32 + PM * NSPLIT;

struct ProjTcArgs {
  const float* w[4];
  const float* b[4];
  const float* x;
  float* out;
  long long batch, n_tiles;
  int t, tiles_t;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PJ_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra PJ_DONE;\n"
      "bra PJ_WAIT;\n"
      "PJ_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, fp32 containers read as TF32, fp32 accumulate, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor: 8-row groups 128 B apart (SBO), K core matrices `lbo` bytes apart
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ float gelu_erf(float u) { return 0.5f * u * (1.0f + erff(u * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(kProjTcThreads, 1) proj_fwd_tc_kernel(const ProjTcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t afull = base, accfull = base + 8, tmem_slot = base + 16;
  const uint32_t sbias = base + OFF_BIAS, sWhi = base + OFF_WHI, sWlo = base + OFF_WLO, sAhi = base + OFF_AHI, sAlo = base + OFF_ALO;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(afull, 4 * NSPLIT);
    mbar_init(accfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident weights: W_l[o][i] -> panels [i/4][o][4 i] as hi (top 19 bits) and lo (remainder); biases
  for (int idx = threadIdx.x; idx < 4 * PD * (PD / 4); idx += kProjTcThreads) {
    const int l = idx / (PD * (PD / 4)), rem = idx - l * (PD * (PD / 4));
    const int o = rem / (PD / 4), pn = rem - o * (PD / 4);
    const float4 w = __ldg(reinterpret_cast<const float4*>(a.w[l] + o * PD) + pn);
    const float4 hi = make_float4(tf32_hi(w.x), tf32_hi(w.y), tf32_hi(w.z), tf32_hi(w.w));
    const uint32_t off = (uint32_t)l * W_BYTES + (uint32_t)pn * PSW + (uint32_t)o * 16u;
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sWhi + off), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sWlo + off), "f"(w.x - hi.x), "f"(w.y - hi.y), "f"(w.z - hi.z),
                 "f"(w.w - hi.w)
                 : "memory");
  }
  for (int i = threadIdx.x; i < 4 * PD; i += kProjTcThreads)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + 4u * i), "f"(__ldg(a.b[i / PD] + (i % PD))) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32 (format 2), both K-major, N = 64, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(PD >> 3) << 17) | ((uint32_t)(PM >> 4) << 24);
      uint32_t n = 0;
      for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        for (int l = 0; l < 4; ++l, ++n) {
          mbar_wait(afull, n & 1u);
          tc_fence_after();
          const uint32_t whi = sWhi + (uint32_t)l * W_BYTES, wlo = sWlo + (uint32_t)l * W_BYTES;
#pragma unroll
          for (int term = 0; term < 3; ++term) {   // a_lo w_hi, a_hi w_lo, a_hi w_hi
            const uint32_t ab = (term == 0) ? sAlo : sAhi, wb = (term == 1) ? wlo : whi;
#pragma unroll
            for (int kk = 0; kk < PD / 8; ++kk)
              umma_tf32(tmem_base, make_desc_ns(ab + (uint32_t)(2 * kk) * PSA, PSA), make_desc_ns(wb + (uint32_t)(2 * kk) * PSW, PSW), idesc,
                        (term | kk) != 0 ? 1u : 0u);
          }
          umma_commit(accfull);
        }
      }
    }
  } else {
    // ===================== token rows: thread = one token of the tile x CP features =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int split = (warp - 1) >> 2;            // feature quarter: features [CP split, CP split + CP)
    const int r = quarter * 32 + lane, c_lo = split * CP;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c_lo;
    uint32_t n = 0;
    for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const long long bi = tile / a.tiles_t;
      const int t0 = (int)(tile - bi * a.tiles_t) * PM;
      const bool valid = t0 + r < a.t;
      const float* xp = a.x + (bi * PD + c_lo) * (long long)a.t + t0 + r;
      float x[CP], h[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) { x[c] = valid ? __ldg(xp + (long long)c * a.t) : 0.f; h[c] = x[c]; }
#pragma unroll 1
      for (int l = 0; l < 4; ++l, ++n) {
#pragma unroll
        for (int pn = 0; pn < CP / 4; ++pn) {     // this thread's panels of the split A operand
          const float4 hi = make_float4(tf32_hi(h[4 * pn]), tf32_hi(h[4 * pn + 1]), tf32_hi(h[4 * pn + 2]), tf32_hi(h[4 * pn + 3]));
          const uint32_t off = (uint32_t)(c_lo / 4 + pn) * PSA + (uint32_t)r * 16u;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sAhi + off), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sAlo + off), "f"(h[4 * pn] - hi.x), "f"(h[4 * pn + 1] - hi.y),
                       "f"(h[4 * pn + 2] - hi.z), "f"(h[4 * pn + 3] - hi.w)
                       : "memory");
        }
        tc_fence_before();                                              // accumulator reads of the previous layer are done
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the operand is read by the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(afull);
        mbar_wait(accfull, n & 1u);
        tc_fence_after();
        const uint32_t sb = sbias + (uint32_t)(l * PD + c_lo) * 4u;
        uint32_t v[CP];
        tmem_ld16(taddr, v);
#pragma unroll
        for (int j = 0; j < CP; ++j) {
          float bj;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bj) : "r"(sb + 4u * (uint32_t)j));
          const float u = __uint_as_float(v[j]) + bj;
          h[j] += (l < 3) ? gelu_erf(u) : u;
        }
      }
      if (valid) {
        float* op = a.out + (bi * PD + c_lo) * (long long)a.t + t0 + r;
#pragma unroll
        for (int c = 0; c < CP; ++c) op[(long long)c * a.t] = x[c] + h[c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

}  // namespace

namespace aa {

// Forward of one projector half on the tensor core; requires dims = hidden = 64 with residuals.  Returns AA_OK or an error code.
int proj_fwd_tc(const float* const* w, const float* const* b, const float* x, int64_t batch, int64_t t, float* out, cudaStream_t stream) {
  AA_CUDA(aa::ensure_dyn_smem(proj_fwd_tc_kernel, kProjTcSmem));   // per (kernel, device)
  ProjTcArgs a;
  for (int l = 0; l < 4; ++l) { a.w[l] = w[l]; a.b[l] = b[l]; }
  a.x = x; a.out = out; a.batch = batch; a.t = (int)t;
  a.tiles_t = (int)((t + PM - 1) / PM);
  a.n_tiles = batch * a.tiles_t;
  if (a.n_tiles == 0) return AA_OK;
  const int grid = (int)std::min<long long>(a.n_tiles, (long long)aa::num_sms());
  proj_fwd_tc_kernel<<<grid, kProjTcThreads, kProjTcSmem, stream>>>(a);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

}  // namespace aa
