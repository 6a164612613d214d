// bf16 tensor-core path of the conv encoder: Conv1d as an implicit GEMM on tcgen05 (sm_100a).
//
// Activations are kept channels-last in bf16 ([B][L][C], C contiguous) between layers, so that
//   * an input tile for filter tap j is ONE 3-D TMA box (channels x 128 positions x 1 batch element) whose row
//     coordinate is simply shifted by the tap offset; conv zero padding = TMA out-of-bounds zero fill; the stride-1 k7 layers
//     load one [128 + 6 d rows] halo box per channel chunk instead and address tap j as a row-shifted descriptor view of it;
//   * strided down-convs (k = 2s, stride s, pad s/2) become a 3-tap stride-1 conv on the view
//     [B][L/s][s*C] (row offsets -1, 0, +1 with partial column ranges), i.e. the same kernel;
//   * both MMA operands are K-major: A = activations [128 rows x BK channels], B = weights
//     [BN out-channels x BK] repacked once as W2[cout][tap-major K] bf16.
// One persistent CTA per SM, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer
// (accumulators in TMEM, double buffered), warps 2..5 = epilogue (tcgen05.ld -> bias (+ residual) -> ELU
// -> bf16 channels-last store, or fp32 channel-major (+ tanh) for the last layer).
// The first layer (2 input channels, K = 14) has no tensor-core shape and is memory bound: a CUDA-core
// kernel fuses the fader-scaled stem sum, the conv, ELU and the fp32 [B][2][N] -> bf16 [B][N][32] layout change.
#include "aa_common.cuh"
#include "encoder.cuh"
#include "packed_f32x2.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

constexpr int BM = 128;             // output positions per tile (UMMA M)
constexpr int kMaxChunks = 128;     // K chunks per tile (C = 1024 with 7 taps: 112)
constexpr int kMaxEpi = 4;          // epilogue warp groups (4 warps each); warp 0: TMA, warp 1: MMA
constexpr int kTcMaxThreads = 64 + 128 * kMaxEpi;
constexpr int kMaxAcc = 8;

struct TcArgs {
  int n_chunks;                     // K chunks of BK channels per tile
  short chunk_off[kMaxChunks];      // row offset of the A box for chunk q
  short chunk_col[kMaxChunks];      // column (channel) coordinate of the A box for chunk q
  int bn, n_tiles_n, m_tiles;       // N tile width, tiles along cout, tiles along rows (per batch element)
  long long tiles;                  // batch * m_tiles * n_tiles_n
  int lout, lpad, cout;             // valid output rows, rows to zero-fill up to, output channels
  int stages;
  int n_acc, n_epi;                 // TMEM accumulator buffers, epilogue warp groups (tile i -> buffer i % n_acc, group i % n_epi)
  const float* bias;
  const __nv_bfloat16* res;         // residual [B][lout_stride][cout] or NULL
  __nv_bfloat16* out;               // [B][lout_stride][cout] bf16 channels-last, or NULL
  float* out_f32;                   // [B][cout][lout] fp32 channel-major (last layer), or NULL
  long long out_row_stride;         // rows per batch element in `out` / `res`
  int elu, tanh_out;
  int res_tma;                      // RES layers with bn % 64 == 0: residual tile in / result tile out through TMA (see the epilogue)
  // halo mode (stride-1 k-tap layers, non-RES): ONE [halo_rows x 64 channels] activation box per channel chunk serves all taps
  // (tap j = the same box with the descriptor start advanced by j * dil rows), only the weight boxes stream per (chunk, tap)
  int halo_rows, halo_cc, halo_taps, halo_dil, halo_row0, a_stages;   // rows of the box, channel chunks, taps, dilation, row offset of tap 0, A slots
  int halo_pair;                    // halo mode, one N tile (bn = cout): every weight box serves TWO of the CTA's tiles (two halo boxes, two accumulators)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Blocking wait: try_wait with a suspend-time hint lets the hardware park the thread until the phase completes (or the
// hint expires) instead of returning at once -- without it the waiting warps (producer, MMA issuer, idle epilogue groups)
// burn more than half of the SM's issue slots on SYNCS / BRA / YIELD polling, slots the epilogue warps need.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra TC_DONE;\n"
      "bra TC_WAIT;\n"
      "TC_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
// shared -> global tile store (bulk async group of the issuing thread); rows / columns outside the tensor are clipped
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1),
               "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, issued by one thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One lane of a converged warp.  The MMA warp runs its loop with all 32 lanes (so that addresses, descriptors and loop state are
// provably warp-uniform and live in uniform registers) and elects a lane only for the tcgen05 instructions themselves.  With the
// whole loop under `if (lane == 0)` the compiler wraps EVERY tcgen05.mma in an ELECT / 5 x R2UR.BROADCAST / BRA.U.ANY sequence:
// ~130 clocks per MMA where the tensor pipe needs 64 (profiles/ubench/umma_rate_ubench.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
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

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100): dense rows of BK bf16,
// 8-row swizzle atoms stacked along M/N.  BK = 64 -> 128-byte rows, SWIZZLE_128B, SBO = 1024 B;
// BK = 32 -> 64-byte rows, SWIZZLE_64B, SBO = 512 B.  LBO is unused for swizzled K-major (set to 1).
template <int BK>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  constexpr uint64_t sbo = (BK == 64) ? (1024 >> 4) : (512 >> 4);
  constexpr uint64_t layout = (BK == 64) ? 2 : 4;   // SWIZZLE_128B = 2, SWIZZLE_64B = 4
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                           // leading byte offset (>>4), bits [16,30)
  d |= sbo << 32;                                   // stride byte offset (>>4), bits [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version 1 (Blackwell), bits [46,48)
  d |= layout << 61;                                // layout type, bits [61,64)
  return d;
}

// ELU(alpha = 1), branch free; ex2.approx error (~2^-22 relative on exp) is far below the bf16 output precision
__device__ __forceinline__ float elu1(float v) { return fmaxf(v, 0.f) + (__expf(fminf(v, 0.f)) - 1.0f); }

#include "conv_ru.cuh"

// RES: the layer adds a residual (ROLE_RES_SECOND); compile-time so that the epilogue carries no per-element branches
template <int BK, bool RES>
__global__ void __launch_bounds__(kTcMaxThreads, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ CUtensorMap tmR,
                                                                const __grid_constant__ CUtensorMap tmO, const TcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // swizzle atoms need 1024-byte alignment
  constexpr uint32_t A_BYTES = BM * BK * 2;
  const uint32_t B_BYTES = (uint32_t)a.bn * BK * 2;
  const bool halo = !RES && a.halo_rows > 0;
  const uint32_t HALO_SLOT = halo ? (((uint32_t)a.halo_rows * (uint32_t)(BK * 2) + 1023u) & ~1023u) : 0u;   // A slots come first
  const uint32_t STAGE = (halo ? 0u : A_BYTES) + ((B_BYTES + 1023u) & ~1023u);
  const uint32_t ring = base + (uint32_t)a.a_stages * HALO_SLOT;       // operand ring (halo mode: weight boxes only)
  const uint32_t bars = ring + (uint32_t)a.stages * STAGE;              // full[S], empty[S], tfull[8], tempty[8], tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (a.stages + s); };
  auto tfull_bar = [&](int i) { return bars + 8u * (2 * a.stages + i); };
  auto tempty_bar = [&](int i) { return bars + 8u * (2 * a.stages + kMaxAcc + i); };
  const uint32_t tmem_slot = bars + 8u * (2 * a.stages + 2 * kMaxAcc);
  // res_tma: per epilogue group one [bn / 64][128 rows][128 B] SWIZZLE_128B tile buffer (residual in, result out, in place) at
  // bars + 1024 + g * (BM * bn * 2 + 1024), followed by the group's bias copy; rfull[g] = residual landed, rempty[g] = the
  // group's previous result has been read out of the buffer by its TMA store
  auto rfull_bar = [&](int g) { return bars + 384u + 8u * g; };
  auto rempty_bar = [&](int g) { return bars + 416u + 8u * g; };
  const uint32_t RBUF = (uint32_t)BM * (uint32_t)a.bn * 2u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int acc_cols = a.n_acc * a.bn;
  const uint32_t tmem_cols = (acc_cols <= 32) ? 32 : (acc_cols <= 64 ? 64 : (acc_cols <= 128 ? 128 : (acc_cols <= 256 ? 256 : 512)));

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < a.n_acc; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 4); }
    for (int g = 0; g < kMaxEpi; ++g) { mbar_init(rfull_bar(g), 1); mbar_init(rempty_bar(g), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(bars + 448u + 8u * i, 1); mbar_init(bars + 480u + 8u * i, 1); }   // afull[4], aempty[4] (halo mode)
    if (halo) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");   // tmR = the halo box map in this mode
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (RES && a.res_tma) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
    }
  }
  if (warp == 1) {   // TMEM allocation (whole warp), address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int tiles_per_b = a.m_tiles * a.n_tiles_n;

  if (warp >= 2 + 4 * a.n_epi) {
    // unused epilogue warps of this launch configuration
  } else if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues: see elect_one) =====================
    {
      const uint32_t ubase = __shfl_sync(0xffffffffu, base, 0), ubars = __shfl_sync(0xffffffffu, bars, 0);
      const uint32_t uring = __shfl_sync(0xffffffffu, ring, 0);
      int s = 0, has = 0;
      uint32_t ph = 0, hph = 0;
      int it = 0;
      if (halo && a.halo_pair) {
        // two tiles of this CTA per pass over the weights (n_tiles_n = 1: every tile uses the same weight boxes)
        for (long long t = blockIdx.x; t < a.tiles; t += 2LL * gridDim.x) {
          const long long tB = t + gridDim.x;
          const int nb = tB < a.tiles ? 2 : 1;
          for (int cc = 0; cc < a.halo_cc; ++cc) {
            for (int w = 0; w < nb; ++w) {
              const long long tt = w ? tB : t;
              const int b = (int)(tt / tiles_per_b), m0 = (int)(tt % tiles_per_b) * BM;
              const uint32_t af = ubars + 448u + 8u * has;
              mbar_wait(ubars + 480u + 8u * has, hph ^ 1u);
              if (elect_one()) {
                mbar_expect_tx(af, (uint32_t)a.halo_rows * (uint32_t)(BK * 2));
                tma_load_3d(ubase + (uint32_t)has * HALO_SLOT, &tmR, af, cc * BK, m0 + a.halo_row0, b);
              }
              if (++has == a.a_stages) { has = 0; hph ^= 1u; }
            }
            for (int j = 0; j < a.halo_taps; ++j) {
              const uint32_t fb = ubars + 8u * s;
              mbar_wait(ubars + 8u * (a.stages + s), ph ^ 1u);
              if (elect_one()) {
                mbar_expect_tx(fb, B_BYTES);
                tma_load_2d(uring + (uint32_t)s * STAGE, &tmB, fb, (j * a.halo_cc + cc) * BK, 0);
              }
              if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
          }
        }
      } else
      for (long long t = blockIdx.x; t < a.tiles; t += gridDim.x, ++it) {
        const int b = (int)(t / tiles_per_b);
        const int r = (int)(t % tiles_per_b);
        const int nt = r % a.n_tiles_n, mt = r / a.n_tiles_n;
        const int m0 = mt * BM, n0 = nt * a.bn;
        if (halo) {
          for (int cc = 0; cc < a.halo_cc; ++cc) {
            const uint32_t af = ubars + 448u + 8u * has;
            mbar_wait(ubars + 480u + 8u * has, hph ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(af, (uint32_t)a.halo_rows * (uint32_t)(BK * 2));
              tma_load_3d(ubase + (uint32_t)has * HALO_SLOT, &tmR, af, cc * BK, m0 + a.halo_row0, b);
            }
            if (++has == a.a_stages) { has = 0; hph ^= 1u; }
            for (int j = 0; j < a.halo_taps; ++j) {
              const uint32_t fb = ubars + 8u * s;
              mbar_wait(ubars + 8u * (a.stages + s), ph ^ 1u);
              if (elect_one()) {
                mbar_expect_tx(fb, B_BYTES);
                tma_load_2d(uring + (uint32_t)s * STAGE, &tmB, fb, (j * a.halo_cc + cc) * BK, n0);
              }
              if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
          }
          continue;
        }
        for (int q = 0; q < a.n_chunks; ++q) {
          const uint32_t fb = ubars + 8u * s;
          mbar_wait(ubars + 8u * (a.stages + s), ph ^ 1u);
          const uint32_t sa = uring + (uint32_t)s * STAGE;
          if (elect_one()) {
            mbar_expect_tx(fb, A_BYTES + B_BYTES);
            tma_load_3d(sa, &tmA, fb, a.chunk_col[q], m0 + a.chunk_off[q], b);
            tma_load_2d(sa + A_BYTES, &tmB, fb, q * BK, n0);
          }
          if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
        if (RES && a.res_tma) {
          // this tile's residual -> its epilogue group's buffer: ONE DRAM round trip, requested before the tile's MMAs have run.
          // (Measured alternative: the group's leader thread requesting the group's next residual right after its store --
          // no coupling to this warp -- was slower at C = 128: 78 vs 67 us.)
          const int g = it % a.n_epi, k = it / a.n_epi;
          if (k > 0) mbar_wait(ubars + 416u + 8u * g, (uint32_t)(k - 1) & 1u);
          const uint32_t rb = ubars + 1024u + (uint32_t)g * (RBUF + 1024u), rf = ubars + 384u + 8u * g;
          if (elect_one()) {
            mbar_expect_tx(rf, RBUF);
            for (int bx = 0; bx < a.bn / 64; ++bx) tma_load_3d(rb + (uint32_t)bx * 16384u, &tmR, rf, n0 + 64 * bx, m0, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues: see elect_one) =====================
    {
      const uint32_t ubase = __shfl_sync(0xffffffffu, base, 0), ubars = __shfl_sync(0xffffffffu, bars, 0);
      const uint32_t utmem = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t uring = __shfl_sync(0xffffffffu, ring, 0);
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major, N, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int s = 0, has = 0;
      uint32_t ph = 0, hph = 0;
      int it = 0;
      if (halo && a.halo_pair) {
        for (long long t = blockIdx.x; t < a.tiles; t += 2LL * gridDim.x, it += 2) {
          const int nb = t + gridDim.x < a.tiles ? 2 : 1;
          uint32_t tmem_d[2];
          for (int w = 0; w < nb; ++w) {
            const int acc = (it + w) % a.n_acc;
            mbar_wait(ubars + 8u * (2 * a.stages + kMaxAcc + acc), (((uint32_t)((it + w) / a.n_acc)) & 1u) ^ 1u);
            tmem_d[w] = utmem + (uint32_t)(acc * a.bn);
          }
          tc_fence_after();
          for (int cc = 0; cc < a.halo_cc; ++cc) {
            uint32_t xa[2];
            int slot[2];
            for (int w = 0; w < nb; ++w) {
              mbar_wait(ubars + 448u + 8u * has, hph);
              slot[w] = has;
              xa[w] = ubase + (uint32_t)has * HALO_SLOT;
              if (++has == a.a_stages) { has = 0; hph ^= 1u; }
            }
            for (int j = 0; j < a.halo_taps; ++j) {
              mbar_wait(ubars + 8u * s, ph);
              tc_fence_after();
              const uint64_t db = make_desc<BK>(uring + (uint32_t)s * STAGE);
              for (int w = 0; w < nb; ++w) {
                const uint64_t da = make_desc<BK>(xa[w] + (uint32_t)(j * a.halo_dil) * (uint32_t)(BK * 2));
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  if (elect_one()) umma_bf16(tmem_d[w], da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (cc | j | k) != 0 ? 1u : 0u);
              }
              if (elect_one()) umma_commit(ubars + 8u * (a.stages + s));   // frees the weight stage
              if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
            for (int w = 0; w < nb; ++w)
              if (elect_one()) umma_commit(ubars + 480u + 8u * slot[w]);   // frees the halo slots
          }
          for (int w = 0; w < nb; ++w)
            if (elect_one()) umma_commit(ubars + 8u * (2 * a.stages + (it + w) % a.n_acc));
        }
      } else
      for (long long t = blockIdx.x; t < a.tiles; t += gridDim.x, ++it) {
        const int acc = it % a.n_acc;
        mbar_wait(ubars + 8u * (2 * a.stages + kMaxAcc + acc), (((uint32_t)(it / a.n_acc)) & 1u) ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = utmem + (uint32_t)(acc * a.bn);
        if (halo) {
          for (int cc = 0; cc < a.halo_cc; ++cc) {
            mbar_wait(ubars + 448u + 8u * has, hph);
            const uint32_t xa = ubase + (uint32_t)has * HALO_SLOT;
            for (int j = 0; j < a.halo_taps; ++j) {
              mbar_wait(ubars + 8u * s, ph);
              tc_fence_after();
              // tap j = the halo box seen from row j * dil on (SWIZZLE_128B is a function of the shared-memory address, so a
              // descriptor may start at any 128-byte row of a TMA-written tile: profiles/ubench/sw128_rowshift_ubench)
              const uint64_t da = make_desc<BK>(xa + (uint32_t)(j * a.halo_dil) * (uint32_t)(BK * 2)), db = make_desc<BK>(uring + (uint32_t)s * STAGE);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                if (elect_one()) umma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (cc | j | k) != 0 ? 1u : 0u);
              if (elect_one()) umma_commit(ubars + 8u * (a.stages + s));   // frees the weight stage
              if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
            if (elect_one()) umma_commit(ubars + 480u + 8u * has);         // frees the halo slot when all its taps have been read
            if (++has == a.a_stages) { has = 0; hph ^= 1u; }
          }
          if (elect_one()) umma_commit(ubars + 8u * (2 * a.stages + acc));
          continue;
        }
        for (int q = 0; q < a.n_chunks; ++q) {
          mbar_wait(ubars + 8u * s, ph);
          tc_fence_after();
          const uint32_t sa = uring + (uint32_t)s * STAGE;
          const uint64_t da = make_desc<BK>(sa), db = make_desc<BK>(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)   // 16 bf16 = 32 bytes along K inside the swizzle atom
            if (elect_one()) umma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (q | k) != 0 ? 1u : 0u);
          if (elect_one()) umma_commit(ubars + 8u * (a.stages + s));   // frees the smem stage when these MMAs retire
          if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) umma_commit(ubars + 8u * (2 * a.stages + acc));   // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue: 4 warps, warp w owns TMEM lanes 32*(w%4) .. +31 =====================
    // The [128 x bn] bf16 output tile is staged in shared memory (rows padded by 16 B -> conflict-free) so that
    // the residual is read and the result written with fully coalesced 16-byte accesses.
    const int quarter = warp & 3;                          // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;                       // epilogue group
    const int row_in_tile = quarter * 32 + lane;
    const int et = (threadIdx.x - 64) & 127;               // 0..127 within the group
    const uint32_t RS = (uint32_t)a.bn * 2u + 16u;         // row stride in bytes
    const uint32_t grp_bytes = (a.out_f32 == nullptr ? (uint32_t)BM * RS : 0u) + (uint32_t)a.bn * 4u;
    const bool rtma = RES && a.res_tma != 0;
    const uint32_t stg = rtma ? bars + 1024u + (uint32_t)grp * (RBUF + 1024u)
                              : bars + 512u + (uint32_t)grp * grp_bytes;   // this group's staging area, then its bias copy
    const int vec_per_row = a.bn / 8;                      // 16-byte vectors per row (power of two)
    const int vpr_shift = 31 - __clz(vec_per_row);
    const uint32_t sbias = rtma ? stg + RBUF : stg + (a.out_f32 == nullptr ? (uint32_t)BM * RS : 0u);   // bn floats
    const int bar_id = 1 + grp;
    int bias_n0 = -1;
    int it = 0;
    for (long long t = blockIdx.x; t < a.tiles; t += gridDim.x, ++it) {
      if (it % a.n_epi != grp) continue;
      const int acc = it % a.n_acc;
      const int b = (int)(t / tiles_per_b);
      const int r = (int)(t % tiles_per_b);
      const int nt = r % a.n_tiles_n, mt = r / a.n_tiles_n;
      const int m0 = mt * BM, n0 = nt * a.bn;
      const int m = m0 + row_in_tile;
      const bool valid = m < a.lout;
      const long long tile_off = ((long long)b * a.out_row_stride + m0) * a.cout + n0;   // element offset of (row 0, col 0)
      if (n0 != bias_n0) {                                  // (re)load this N tile's bias into shared memory
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        for (int i = et; i < a.bn; i += 128) {
          const float bv = (n0 + i < a.cout) ? __ldg(a.bias + n0 + i) : 0.f;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + 4u * i), "f"(bv) : "memory");
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        bias_n0 = n0;
      }
      if (RES && a.out_f32 == nullptr && !rtma) {           // residual tile -> staging (overlaps with the MMAs)
        // loads in batches of 4 before their shared-memory stores: one L2 round trip per batch, not per 16-byte piece
        const int n_vec = BM * vec_per_row;                 // multiple of 512 (bn >= 32)
        const __nv_bfloat16* rbase = a.res + tile_off;
        for (int i0 = et; i0 < n_vec; i0 += 512) {
          uint4 rv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int idx = i0 + 128 * u;
            const int rr = idx >> vpr_shift, cv = idx & (vec_per_row - 1);
            rv[u] = make_uint4(0u, 0u, 0u, 0u);
            if (m0 + rr < a.lout) rv[u] = __ldg(reinterpret_cast<const uint4*>(rbase + (size_t)rr * a.cout) + cv);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int idx = i0 + 128 * u;
            const int rr = idx >> vpr_shift, cv = idx & (vec_per_row - 1);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)rr * RS + (uint32_t)cv * 16u), "r"(rv[u].x),
                         "r"(rv[u].y), "r"(rv[u].z), "r"(rv[u].w)
                         : "memory");
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      }
      mbar_wait(tfull_bar(acc), ((uint32_t)(it / a.n_acc)) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * a.bn);
      if (rtma) {
        // The residual tile was laid down by TMA (SWIZZLE_128B: 128-byte rows of 64 channels, 16-byte piece p of row r at
        // ((p ^ (r & 7)) << 4) -- conflict free for lanes = consecutive rows); the result overwrites it IN PLACE (a thread touches
        // only its own row's pieces) and leaves through ONE TMA store per 64-channel box: no per-thread global loads (their four
        // dependent DRAM round trips per tile were the critical path of the 1x1 layers: ncu long_scoreboard 3.3 cycles per issued
        // instruction at 58 % of the HBM roofline), no copy-out loop.
        mbar_wait(rfull_bar(grp), ((uint32_t)(it / a.n_epi)) & 1u);
        const uint32_t rrow = stg + (uint32_t)row_in_tile * 128u;
        const uint32_t rx = (uint32_t)(row_in_tile & 7);
        for (int c0 = 0; c0 < a.bn; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)c0, v);
          float bia[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bia[j]), "=f"(bia[j + 1]), "=f"(bia[j + 2]), "=f"(bia[j + 3])
                         : "r"(sbias + 4u * (uint32_t)(c0 + j)));
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t col = (uint32_t)c0 + 8u * g;
            const uint32_t addr = rrow + (col >> 6) * 16384u + ((((col & 63u) >> 3) ^ rx) << 4);
            float2 y0 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1])), make_float2(bia[8 * g + 0], bia[8 * g + 1]));
            float2 y1 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])), make_float2(bia[8 * g + 2], bia[8 * g + 3]));
            float2 y2 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])), make_float2(bia[8 * g + 4], bia[8 * g + 5]));
            float2 y3 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])), make_float2(bia[8 * g + 6], bia[8 * g + 7]));
            uint4 rres;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rres.x), "=r"(rres.y), "=r"(rres.z), "=r"(rres.w) : "r"(addr));
            y0 = __fadd2_rn(y0, unpack_bf16(rres.x));
            y1 = __fadd2_rn(y1, unpack_bf16(rres.y));
            y2 = __fadd2_rn(y2, unpack_bf16(rres.z));
            y3 = __fadd2_rn(y3, unpack_bf16(rres.w));
            uint4 o = make_uint4(pack_bf16(elu2(y0)), pack_bf16(elu2(y1)), pack_bf16(elu2(y2)), pack_bf16(elu2(y3)));
            if (!valid) o = make_uint4(0u, 0u, 0u, 0u);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
          }
        }
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile is read by the TMA store (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));          // accumulator drained: the MMA warp may reuse it
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // tile complete in the buffer
        if (et == 0) {
          for (int bx = 0; bx < a.bn / 64; ++bx) tma_store_3d(&tmO, stg + (uint32_t)bx * 16384u, n0 + 64 * bx, m0, b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the store has read the buffer: the next residual may land
          mbar_arrive(rempty_bar(grp));
        }
        continue;
      }
      for (int c0 = 0; c0 < a.bn; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);   // warp-collective: every lane participates
        float bia[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bia[j]), "=f"(bia[j + 1]), "=f"(bia[j + 2]), "=f"(bia[j + 3])
                       : "r"(sbias + 4u * (uint32_t)(c0 + j)));
        if (a.out_f32 != nullptr) {
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + c0 + j;
              if (n < a.cout) {
                float x = __uint_as_float(v[j]) + bia[j];
                if (a.elu) x = elu1(x);
                if (a.tanh_out) x = tanhf(x);
                a.out_f32[((long long)b * a.cout + n) * a.lout + m] = x;
              }
            }
          }
        } else {
          // bf16 channels-last output: +bias (+ residual from staging) -> ELU (every bf16 layer of the table has one) -> staging
          const uint32_t srow = stg + (uint32_t)row_in_tile * RS + (uint32_t)c0 * 2u;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float2 y0 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1])), make_float2(bia[8 * g + 0], bia[8 * g + 1]));
            float2 y1 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])), make_float2(bia[8 * g + 2], bia[8 * g + 3]));
            float2 y2 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])), make_float2(bia[8 * g + 4], bia[8 * g + 5]));
            float2 y3 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])), make_float2(bia[8 * g + 6], bia[8 * g + 7]));
            if (RES) {
              uint4 rres;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rres.x), "=r"(rres.y), "=r"(rres.z), "=r"(rres.w)
                           : "r"(srow + (uint32_t)g * 16u));
              y0 = __fadd2_rn(y0, unpack_bf16(rres.x));
              y1 = __fadd2_rn(y1, unpack_bf16(rres.y));
              y2 = __fadd2_rn(y2, unpack_bf16(rres.z));
              y3 = __fadd2_rn(y3, unpack_bf16(rres.w));
            }
            uint4 o = make_uint4(pack_bf16(elu2(y0)), pack_bf16(elu2(y1)), pack_bf16(elu2(y2)), pack_bf16(elu2(y3)));
            if (!valid) o = make_uint4(0u, 0u, 0u, 0u);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)g * 16u), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                         : "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));          // accumulator drained: the MMA warp may reuse it
      if (a.out_f32 == nullptr) {
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // tile complete in staging
        for (int idx = et; idx < BM * vec_per_row; idx += 128) {
          const int rr = idx >> vpr_shift, cv = idx & (vec_per_row - 1);
          if (m0 + rr < a.lpad) {                           // rows in [lout, lpad) carry zeros (tail of a strided view)
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(stg + (uint32_t)rr * RS + (uint32_t)cv * 16u));
            reinterpret_cast<uint4*>(a.out + tile_off + (long long)rr * a.cout)[cv] = v;
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // staging free for the next tile
      }
    }
    if (rtma && et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this thread's tile stores are complete
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

#include "conv_cg2.cuh"

// ---- layer 0: fp32 stems [B][cin<=4][N] (fader-scaled sum) -> conv k (<=7), cout <= 32 -> ELU -> bf16 [B][N][cout] ----
struct L0Args {
  const float* x[4];
  float fader[4];
  int n_in, cin, cout, k, pad, n, lpad;
  const float* w;      // [cout][cin][k]
  const float* bias;
  __nv_bfloat16* out;  // [B][row_stride][cout]
  long long row_stride;
};
__global__ void __launch_bounds__(256) conv_l0_kernel(const L0Args a) {
  __shared__ float Xs[4][256 + 8];
  __shared__ float Ws[32 * 4 * 7];
  __shared__ float Bs[32];
  const int b = blockIdx.y, l0 = blockIdx.x * 256;
  for (int e = threadIdx.x; e < a.cout * a.cin * a.k; e += 256) Ws[e] = a.w[e];
  if (threadIdx.x < a.cout) Bs[threadIdx.x] = a.bias[threadIdx.x];
  const int span = 256 + a.k - 1;
  for (int e = threadIdx.x; e < a.cin * span; e += 256) {
    const int c = e / span, j = e % span, pos = l0 - a.pad + j;
    float v = 0.f;
    if (pos >= 0 && pos < a.n) {
      const long long off = ((long long)b * a.cin + c) * a.n + pos;
      v = a.fader[0] * a.x[0][off];
      for (int s = 1; s < a.n_in; ++s) v = fmaf(a.fader[s], a.x[s][off], v);
    }
    Xs[c][j] = v;
  }
  __syncthreads();
  // results are staged in shared memory (row stride cout*2 + 16 bytes: conflict-free) and written as one
  // contiguous, fully coalesced block: 256 positions x cout channels are adjacent in the channels-last layout
  __shared__ __align__(16) unsigned char Os[256 * (32 * 2 + 16)];
  const int l = l0 + threadIdx.x;
  const int RS = a.cout * 2 + 16;
  for (int co0 = 0; co0 < a.cout; co0 += 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = Bs[co0 + i];
    for (int c = 0; c < a.cin; ++c)
      for (int kk = 0; kk < a.k; ++kk) {
        const float xv = Xs[c][threadIdx.x + kk];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(Ws[((co0 + i) * a.cin + c) * a.k + kk], xv, acc[i]);
      }
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float x0 = (l < a.n) ? elu1(acc[2 * i]) : 0.f, x1 = (l < a.n) ? elu1(acc[2 * i + 1]) : 0.f;
      __nv_bfloat162 p = __floats2bfloat162_rn(x0, x1);
      w[i] = *reinterpret_cast<uint32_t*>(&p);
    }
    *reinterpret_cast<uint4*>(Os + threadIdx.x * RS + co0 * 2) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __syncthreads();
  const int vpr = a.cout / 8;
  uint4* o = reinterpret_cast<uint4*>(a.out + ((long long)b * a.row_stride + l0) * a.cout);
  for (int idx = threadIdx.x; idx < 256 * vpr; idx += 256) {
    const int rr = idx / vpr, cv = idx - rr * vpr;
    if (l0 + rr < a.lpad) o[idx] = *reinterpret_cast<const uint4*>(Os + rr * RS + cv * 16);
  }
}

// The same for wider first layers (cout <= 64, e.g. the capacity-64 first stage of StackedDiffAEWrapper): 128 positions per block.
__global__ void __launch_bounds__(128) conv_l0_wide_kernel(const L0Args a) {
  __shared__ float Xs[4][128 + 8];
  __shared__ float Ws[64 * 4 * 7];
  __shared__ float Bs[64];
  __shared__ __align__(16) unsigned char Os[128 * (64 * 2 + 16)];
  const int b = blockIdx.y, l0 = blockIdx.x * 128;
  for (int e = threadIdx.x; e < a.cout * a.cin * a.k; e += 128) Ws[e] = a.w[e];
  if (threadIdx.x < a.cout) Bs[threadIdx.x] = a.bias[threadIdx.x];
  const int span = 128 + a.k - 1;
  for (int e = threadIdx.x; e < a.cin * span; e += 128) {
    const int c = e / span, j = e % span, pos = l0 - a.pad + j;
    float v = 0.f;
    if (pos >= 0 && pos < a.n) {
      const long long off = ((long long)b * a.cin + c) * a.n + pos;
      v = a.fader[0] * a.x[0][off];
      for (int s = 1; s < a.n_in; ++s) v = fmaf(a.fader[s], a.x[s][off], v);
    }
    Xs[c][j] = v;
  }
  __syncthreads();
  const int l = l0 + threadIdx.x;
  const int RS = a.cout * 2 + 16;
  for (int co0 = 0; co0 < a.cout; co0 += 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = Bs[co0 + i];
    for (int c = 0; c < a.cin; ++c)
      for (int kk = 0; kk < a.k; ++kk) {
        const float xv = Xs[c][threadIdx.x + kk];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(Ws[((co0 + i) * a.cin + c) * a.k + kk], xv, acc[i]);
      }
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float x0 = (l < a.n) ? elu1(acc[2 * i]) : 0.f, x1 = (l < a.n) ? elu1(acc[2 * i + 1]) : 0.f;
      __nv_bfloat162 p = __floats2bfloat162_rn(x0, x1);
      w[i] = *reinterpret_cast<uint32_t*>(&p);
    }
    *reinterpret_cast<uint4*>(Os + threadIdx.x * RS + co0 * 2) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __syncthreads();
  const int vpr = a.cout / 8;
  uint4* o = reinterpret_cast<uint4*>(a.out + ((long long)b * a.row_stride + l0) * a.cout);
  for (int idx = threadIdx.x; idx < 128 * vpr; idx += 128) {
    const int rr = idx / vpr, cv = idx - rr * vpr;
    if (l0 + rr < a.lpad) o[idx] = *reinterpret_cast<const uint4*>(Os + rr * RS + cv * 16);
  }
}

// Layer 0, common shape (cout = 32, k = 7, stride 1, 'same' padding): 2 positions per thread, all 32 output channels in
// packed fp32x2 accumulators (channel pairs); weights are read as broadcast float4 (4 channels of one (cin, tap)), so one
// LDS.128 feeds 4 FFMA2 -- the kernel is FMA-pipe bound (448 FMA per position) instead of shared-memory bound.
constexpr int kL0Pos = 512;                                   // positions per CTA
constexpr int kL0Smem = 4 * (kL0Pos + 8) * 4 + 4 * 7 * 32 * 4 + 32 * 4 + kL0Pos * 80;   // (also covers 2 buffers x CIN <= 2 of the register kernel)
__global__ void __launch_bounds__(256, 2) conv_l0_c32k7_kernel(const L0Args a) {
  extern __shared__ __align__(16) unsigned char l0smem[];
  float* Xs = reinterpret_cast<float*>(l0smem);                              // [cin][kL0Pos + 8]
  float4* Ws4 = reinterpret_cast<float4*>(Xs + 4 * (kL0Pos + 8));            // [(c*7 + j)][8] float4 = channels 4q..4q+3
  float* Bs = reinterpret_cast<float*>(Ws4 + 4 * 7 * 8);
  unsigned char* Os = reinterpret_cast<unsigned char*>(Bs + 32);             // [kL0Pos][80 B]
  const int b = blockIdx.y, l0 = blockIdx.x * kL0Pos, tid = threadIdx.x;
  for (int e = tid; e < a.cin * 7 * 32; e += 256) {
    const int co = e & 31, cj = e >> 5;                                      // cj = c*7 + j
    reinterpret_cast<float*>(Ws4)[cj * 32 + co] = a.w[(co * a.cin + cj / 7) * 7 + cj % 7];
  }
  if (tid < 32) Bs[tid] = a.bias[tid];
  constexpr int span = kL0Pos + 6;
  for (int e = tid; e < a.cin * span; e += 256) {
    const int c = e / span, j = e - c * span, pos = l0 - 3 + j;
    float v = 0.f;
    if (pos >= 0 && pos < a.n) {
      const long long off = ((long long)b * a.cin + c) * a.n + pos;
      v = a.fader[0] * __ldg(a.x[0] + off);
      for (int s = 1; s < a.n_in; ++s) v = fmaf(a.fader[s], __ldg(a.x[s] + off), v);
    }
    Xs[c * (kL0Pos + 8) + j] = v;
  }
  __syncthreads();
  float2 acc0[16], acc1[16];                                                 // positions 2 tid and 2 tid + 1
#pragma unroll
  for (int i = 0; i < 16; ++i) acc0[i] = acc1[i] = make_float2(Bs[2 * i], Bs[2 * i + 1]);
  for (int c = 0; c < a.cin; ++c) {
    float xw[8];                                                             // x[2 tid - 3 .. 2 tid + 4]
    const float2* xp = reinterpret_cast<const float2*>(Xs + c * (kL0Pos + 8) + 2 * tid);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = xp[i]; xw[2 * i] = t.x; xw[2 * i + 1] = t.y; }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const float4* wp = Ws4 + (c * 7 + j) * 8;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 w = wp[q];
        acc0[2 * q] = pfma2(make_float2(w.x, w.y), make_float2(xw[j], xw[j]), acc0[2 * q]);
        acc0[2 * q + 1] = pfma2(make_float2(w.z, w.w), make_float2(xw[j], xw[j]), acc0[2 * q + 1]);
        acc1[2 * q] = pfma2(make_float2(w.x, w.y), make_float2(xw[j + 1], xw[j + 1]), acc1[2 * q]);
        acc1[2 * q + 1] = pfma2(make_float2(w.z, w.w), make_float2(xw[j + 1], xw[j + 1]), acc1[2 * q + 1]);
      }
    }
  }
  const bool ok0 = l0 + 2 * tid < a.n, ok1 = l0 + 2 * tid + 1 < a.n;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t w0[4], w1[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      w0[h] = ok0 ? pack_bf16(elu2(acc0[4 * g + h])) : 0u;
      w1[h] = ok1 ? pack_bf16(elu2(acc1[4 * g + h])) : 0u;
    }
    *reinterpret_cast<uint4*>(Os + (2 * tid) * 80 + g * 16) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
    *reinterpret_cast<uint4*>(Os + (2 * tid + 1) * 80 + g * 16) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
  }
  __syncthreads();
  uint4* o = reinterpret_cast<uint4*>(a.out + ((long long)b * a.row_stride + l0) * 32);
#pragma unroll
  for (int i = 0; i < kL0Pos * 4 / 256; ++i) {
    const int idx = tid + 256 * i, rr = idx >> 2, cv = idx & 3;
    if (l0 + rr < a.lpad) o[idx] = *reinterpret_cast<const uint4*>(Os + rr * 80 + cv * 16);
  }
}

// Layer 0, stereo / mono input (CIN <= 2, cout = 32, k = 7): a thread owns 4 output channels (its 4 x CIN x 7 weights stay in
// registers) and a strip of 16 consecutive positions (the 22-sample input window per channel stays in registers too), so the
// inner loop is pure packed FFMA2: 28 x CIN per position and thread.  Lanes = 8 channel groups x 4 strips; a block of 256
// threads covers 512 positions.  Shared memory is touched for 14 x CIN weight float4s + 11 x CIN input float2s per thread and
// block (broadcast reads) and for the bf16 output staging.
template <int CIN>
__global__ void __launch_bounds__(256, 2) conv_l0_reg_kernel(const L0Args a, int tiles_per_row, int n_tiles) {
  extern __shared__ __align__(16) unsigned char l0smem[];
  float* Xs = reinterpret_cast<float*>(l0smem);                              // 2 buffers x [CIN][kL0Pos + 8], index 0 = position l0 - 3
  float4* Ws4 = reinterpret_cast<float4*>(Xs + 2 * CIN * (kL0Pos + 8));      // [(c*7 + j)][8] float4 = channels 4q..4q+3
  float* Bs = reinterpret_cast<float*>(Ws4 + CIN * 7 * 8);
  unsigned char* Os = reinterpret_cast<unsigned char*>(Bs + 32);             // [kL0Pos][80 B]
  const int tid = threadIdx.x;
  for (int e = tid; e < CIN * 7 * 32; e += 256) {
    const int co = e & 31, cj = e >> 5;                                      // cj = c*7 + j
    reinterpret_cast<float*>(Ws4)[cj * 32 + co] = a.w[(co * CIN + cj / 7) * 7 + cj % 7];
  }
  if (tid < 32) Bs[tid] = a.bias[tid];
  constexpr int span = kL0Pos + 6;
  constexpr int NLD = (CIN * span + 255) / 256;                              // input samples per thread and tile
  // element e of tile t: offset into a stem tensor, or -1 outside the signal (zero = conv padding) / past the last tile
  auto elem_off = [&](int t, int e) -> long long {
    if (t >= n_tiles || e >= CIN * span) return -1;
    const int b = t / tiles_per_row, l0 = (t - b * tiles_per_row) * kL0Pos;
    const int c = e / span, pos = l0 - 3 + (e - c * span);
    if (pos < 0 || pos >= a.n) return -1;
    return ((long long)b * CIN + c) * a.n + pos;
  };
  // the raw samples of the first two stems are loaded early (in flight during the FMAs of the current tile); the fader
  // scaling and the stem sum (aa_mixer.py:303,309) are applied when they are parked in shared memory
  auto load_raw = [&](int t, int e, float& r0, float& r1) {
    const long long off = elem_off(t, e);
    r0 = r1 = 0.f;
    if (off >= 0) {
      r0 = __ldg(a.x[0] + off);
      if (a.n_in > 1) r1 = __ldg(a.x[1] + off);
    }
  };
  auto stash = [&](float* X, int t, int e, float r0, float r1) {
    if (e < CIN * span) {
      float v = a.fader[0] * r0;
      if (a.n_in > 1) v = fmaf(a.fader[1], r1, v);
      if (a.n_in > 2) {
        const long long off = elem_off(t, e);
        if (off >= 0)
          for (int s2 = 2; s2 < a.n_in; ++s2) v = fmaf(a.fader[s2], __ldg(a.x[s2] + off), v);
      }
      const int c = e / span;
      X[c * (kL0Pos + 8) + (e - c * span)] = v;
    }
  };
#pragma unroll
  for (int i = 0; i < NLD; ++i) {
    float r0, r1;
    load_raw(blockIdx.x, tid + 256 * i, r0, r1);
    stash(Xs, blockIdx.x, tid + 256 * i, r0, r1);
  }
  __syncthreads();
  const int lane = tid & 31, cg = lane & 7, strip = (tid >> 5) * 4 + (lane >> 3);   // strip: positions 16 strip .. +15
  float2 w01[CIN * 7], w23[CIN * 7];                                        // channel pairs (4cg, 4cg+1), (4cg+2, 4cg+3)
#pragma unroll
  for (int cj = 0; cj < CIN * 7; ++cj) {
    const float4 w = Ws4[cj * 8 + cg];
    w01[cj] = make_float2(w.x, w.y);
    w23[cj] = make_float2(w.z, w.w);
  }
  const float2 b01 = make_float2(Bs[4 * cg], Bs[4 * cg + 1]), b23 = make_float2(Bs[4 * cg + 2], Bs[4 * cg + 3]);
  int buf = 0;
#pragma unroll 1
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, buf ^= 1) {
    const int b = t / tiles_per_row, l0 = (t - b * tiles_per_row) * kL0Pos;
    float nxt0[NLD], nxt1[NLD];                                             // next tile's raw inputs: in flight during the FMAs below
#pragma unroll
    for (int i = 0; i < NLD; ++i) load_raw(t + gridDim.x, tid + 256 * i, nxt0[i], nxt1[i]);
    const float* X = Xs + buf * CIN * (kL0Pos + 8);
    float xw[CIN][22];
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float2* xp = reinterpret_cast<const float2*>(X + c * (kL0Pos + 8) + 16 * strip);
#pragma unroll
      for (int i = 0; i < 11; ++i) { const float2 v = xp[i]; xw[c][2 * i] = v.x; xw[c][2 * i + 1] = v.y; }
    }
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      float2 a01 = b01, a23 = b23;
#pragma unroll
      for (int c = 0; c < CIN; ++c)
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const float2 xx = make_float2(xw[c][p + j], xw[c][p + j]);
          a01 = pfma2(w01[c * 7 + j], xx, a01);
          a23 = pfma2(w23[c * 7 + j], xx, a23);
        }
      const bool ok = l0 + 16 * strip + p < a.n;
      const uint32_t o0 = ok ? pack_bf16(elu2(a01)) : 0u, o1 = ok ? pack_bf16(elu2(a23)) : 0u;
      *reinterpret_cast<uint2*>(Os + (16 * strip + p) * 80 + cg * 8) = make_uint2(o0, o1);
    }
    __syncthreads();
    uint4* o = reinterpret_cast<uint4*>(a.out + ((long long)b * a.row_stride + l0) * 32);
#pragma unroll
    for (int i = 0; i < kL0Pos * 4 / 256; ++i) {
      const int idx = tid + 256 * i, rr = idx >> 2, cv = idx & 3;
      if (l0 + rr < a.lpad) o[idx] = *reinterpret_cast<const uint4*>(Os + rr * 80 + cv * 16);
    }
    float* Xn = Xs + (buf ^ 1) * CIN * (kL0Pos + 8);
#pragma unroll
    for (int i = 0; i < NLD; ++i) stash(Xn, t + gridDim.x, tid + 256 * i, nxt0[i], nxt1[i]);
    __syncthreads();
  }
}

// W [cout][cin][k] fp32 -> W2 [cout][K_total] bf16 in chunk order: column (tap j, c) <- W[co][c % cin][c / cin + ktap_base[j]]
struct PackArgs {
  int n_taps, cin, k, cout, k_total;
  int tap_col0[9], tap_width[9], tap_kbase[9], tap_kcol[9];
};
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w2, const PackArgs p) {
  const long long n = (long long)p.cout * p.k_total;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / p.k_total), kc = (int)(i % p.k_total);
    int j = 0;
    while (j + 1 < p.n_taps && kc >= p.tap_kcol[j + 1]) ++j;
    const int c = p.tap_col0[j] + (kc - p.tap_kcol[j]);
    const int ci = c % p.cin, kt = c / p.cin + p.tap_kbase[j];
    w2[i] = __float2bfloat16((kt >= 0 && kt < p.k) ? w[((long long)co * p.cin + ci) * p.k + kt] : 0.f);
  }
}

#include "conv_tf32.cuh"

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct LayerPlan {
  int bk = 64;                 // K chunk (channels per TMA box)
  int n_taps = 0;
  int tap_off[9], tap_col0[9], tap_width[9], tap_kbase[9];
  int view_c = 0;              // channels per row of the input view (s * cin)
  int view_s = 1;              // positions per row of the input view
  int k_total = 0;
  __nv_bfloat16* w2 = nullptr;
  float* w2f = nullptr;        // 3xTF32 path: planes [2][cout][k_total]
};

}  // namespace

namespace aa {

struct TcState {
  std::vector<LayerPlan> plans;
  bool weights_valid = false;
  int max_smem = 0;
  bool fuse_ru = true;          // ResidualUnits with C = 32 / 64 / 128 run as one fused kernel (conv_ru.cuh)
  bool fuse_ru128 = false;
  int ru_ctas_per_sm[2] = {1, 1};
  int bn_1x1_wide = 0;
  int nepi_k7_128 = 0, nacc_min = 0;
  int cg2 = 1;                 // wide non-residual layers as cta_group::2 MMAs on CTA pairs (conv_cg2.cuh; AA_TC_CG2=0: one CTA per tile)
  int halo = 1, halo_slots = 2;   // halo mode of the k-tap layers (AA_TC_HALO=0: one activation box per tap, 2: single-CTA kernel only, 3: also N = 128 pairs -- measured slower; AA_TC_HALO_SLOTS: halo boxes in flight)
  int halo_pair = 0, halo_pair_slots = 4;   // conv_tc_kernel halo mode, dev knob AA_TC_HALO_PAIR=1: two tiles per pass over the weights (correct, measured +-0 / slower)
  int cg2_nepi = 0;            // epilogue groups of the non-residual pair kernel (0 = default: 2; AA_TC_CG2_NEPI)
  int cg2r = 1;                // ResidualUnit 1x1 layers at C >= 256 on CTA pairs (conv_tc2_kernel<256, true>; AA_TC_CG2R=0: one CTA per tile, 2: + L2 prefetch of the residual)
  int res_tma = 1;             // ResidualUnit 1x1 layers: residual / result tiles through TMA (AA_RES_TMA=0: per-thread loads, staged stores)
};

static int plan_layer(const ConvLayer& l, LayerPlan& p, bool tf32 = false) {
  p = LayerPlan();
  if (l.stride == 1) {
    p.view_s = 1; p.view_c = l.cin; p.n_taps = l.k;
    AA_REQUIRE(l.k <= 9, "kernel size %d not supported on the tensor-core path", l.k);
    for (int j = 0; j < l.k; ++j) { p.tap_off[j] = j * l.dil - l.pad; p.tap_col0[j] = 0; p.tap_width[j] = l.cin; p.tap_kbase[j] = j; }
  } else {
    const int s = l.stride, pd = l.pad;
    AA_REQUIRE(l.k == 2 * s && pd == (s + 1) / 2 && l.dil == 1 && s % 2 == 0, "strided conv shape not supported on the tensor-core path");
    p.view_s = s; p.view_c = s * l.cin; p.n_taps = 3;
    p.tap_off[0] = -1; p.tap_col0[0] = (s - pd) * l.cin; p.tap_width[0] = pd * l.cin;       p.tap_kbase[0] = -(s - pd);
    p.tap_off[1] = 0;  p.tap_col0[1] = 0;               p.tap_width[1] = s * l.cin;        p.tap_kbase[1] = pd;
    p.tap_off[2] = 1;  p.tap_col0[2] = 0;               p.tap_width[2] = (s - pd) * l.cin; p.tap_kbase[2] = pd + s;
  }
  p.bk = tf32 ? 32 : 64;       // channels per K chunk: 128 bytes of bf16 (64 bytes where a tap is 32 channels wide) or of fp32
  for (int j = 0; j < p.n_taps; ++j) if (p.tap_width[j] % 64 != 0) p.bk = 32;
  p.k_total = 0;
  for (int j = 0; j < p.n_taps; ++j) {
    AA_REQUIRE(p.tap_width[j] % p.bk == 0 && p.tap_col0[j] % p.bk == 0, "channel count %d not a multiple of 32", l.cin);
    p.k_total += p.tap_width[j];
  }
  AA_REQUIRE(p.k_total / p.bk <= (tf32 ? kTfMaxChunks : kMaxChunks), "too many K chunks (%d)", p.k_total / p.bk);
  AA_REQUIRE(l.cout % 32 == 0, "cout=%d must be a multiple of 32 on the tensor-core path", l.cout);
  return AA_OK;
}

int tc_create(TcState** out, const std::vector<ConvLayer>& layers) {
  AA_REQUIRE(get_encode_fn() != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  AA_REQUIRE(layers.size() >= 2 && layers[0].stride == 1 && layers[0].cout <= 64 && layers[0].cout % 8 == 0 && layers[0].cin <= 4 &&
                 layers[0].k <= 7,
             "first layer shape not supported on the tensor-core path");
  TcState* st = new TcState();
  st->plans.resize(layers.size());
  for (size_t i = 1; i < layers.size(); ++i) {
    int rc = plan_layer(layers[i], st->plans[i]);
    if (rc != AA_OK) { delete st; return rc; }
    AA_CUDA(cudaMalloc(&st->plans[i].w2, sizeof(__nv_bfloat16) * (size_t)layers[i].cout * st->plans[i].k_total));
  }
  int dev = 0;
  AA_CUDA(cudaGetDevice(&dev));
  AA_CUDA(cudaDeviceGetAttribute(&st->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  AA_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
  AA_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
  AA_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
  AA_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
  AA_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
  AA_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
  AA_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
  AA_CUDA(cudaFuncSetAttribute(conv_l0_c32k7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kL0Smem));
  AA_CUDA(cudaFuncSetAttribute(conv_l0_reg_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kL0Smem));
  AA_CUDA(cudaFuncSetAttribute(conv_l0_reg_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kL0Smem));
  AA_CUDA(cudaFuncSetAttribute(ru_fused_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, RuCfg<32>::SMEM));
  AA_CUDA(cudaFuncSetAttribute(ru_fused_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, RuCfg<64>::SMEM));
  AA_CUDA(cudaFuncSetAttribute(ru_fused_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, RuCfg<128>::SMEM));
  // C = 128 units: the fused kernel is correct but measured SLOWER than the two layer-wise kernels (420 vs 232 us at B=64): 229 KB of
  // k7 weights per 128-position tile must stream through a 2-stage ring (all the shared memory left), too shallow for the L2
  // latency.  Off by default; AA_RU128=1 enables it (tests, and the 2-CTA multicast follow-up).
  st->fuse_ru128 = getenv("AA_RU128") != nullptr;
  AA_CUDA(cudaFuncSetAttribute(ru_fused_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  AA_CUDA(cudaFuncSetAttribute(ru_fused_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  AA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&st->ru_ctas_per_sm[0], ru_fused_kernel<32>, RuCfg<32>::THREADS, RuCfg<32>::SMEM));
  AA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&st->ru_ctas_per_sm[1], ru_fused_kernel<64>, RuCfg<64>::THREADS, RuCfg<64>::SMEM));
  st->fuse_ru = getenv("AA_NO_RU_FUSION") == nullptr;
  // the C = 32 kernel is sized for two CTAs per SM (105 KB shared memory, 80 registers x 320 threads, 128 TMEM columns each);
  // the persistent tile loop is correct for any grid, so a conservative occupancy answer only costs a second wave
  st->ru_ctas_per_sm[0] = std::max(st->ru_ctas_per_sm[0], getenv("AA_RU_CTAS32") ? atoi(getenv("AA_RU_CTAS32")) : 2);
  if (getenv("AA_RES_TMA")) st->res_tma = atoi(getenv("AA_RES_TMA"));
  if (getenv("AA_1X1_BN256")) st->bn_1x1_wide = atoi(getenv("AA_1X1_BN256"));
  if (getenv("AA_TC_NEPI128")) st->nepi_k7_128 = atoi(getenv("AA_TC_NEPI128"));
  if (getenv("AA_TC_NACC")) st->nacc_min = atoi(getenv("AA_TC_NACC"));
  if (getenv("AA_TC_CG2")) st->cg2 = atoi(getenv("AA_TC_CG2"));
  if (getenv("AA_TC_CG2R")) st->cg2r = atoi(getenv("AA_TC_CG2R"));
  if (getenv("AA_TC_CG2_NEPI")) st->cg2_nepi = atoi(getenv("AA_TC_CG2_NEPI"));
  if (getenv("AA_TC_HALO_PAIR")) st->halo_pair = atoi(getenv("AA_TC_HALO_PAIR"));
  if (getenv("AA_TC_HALO_PAIR_SLOTS")) st->halo_pair_slots = std::max(2, std::min(4, atoi(getenv("AA_TC_HALO_PAIR_SLOTS"))));
  if (getenv("AA_TC_HALO")) st->halo = atoi(getenv("AA_TC_HALO"));
  if (getenv("AA_TC_HALO_SLOTS")) st->halo_slots = std::max(1, std::min(4, atoi(getenv("AA_TC_HALO_SLOTS"))));
  if (getenv("AA_DEBUG")) fprintf(stderr, "[aa] ru_fused CTAs/SM: C=32 -> %d, C=64 -> %d\n", st->ru_ctas_per_sm[0], st->ru_ctas_per_sm[1]);
  *out = st;
  return AA_OK;
}

void tc_destroy(TcState* st) {
  if (!st) return;
  for (auto& p : st->plans) cudaFree(p.w2);
  delete st;
}

void tc_invalidate_weights(TcState* st) { if (st) st->weights_valid = false; }

static int64_t rows_padded(int64_t l) { return (l + 3) / 4 * 4; }

static int64_t tc_max_act_elems(const std::vector<ConvLayer>& layers, int64_t batch, int64_t n) {
  int64_t l = n, mx = 0;
  for (const auto& ly : layers) {
    const int64_t lout = (l + 2 * ly.pad - (int64_t)ly.dil * (ly.k - 1) - 1) / ly.stride + 1;
    mx = std::max(mx, batch * rows_padded(lout) * ly.cout);
    l = lout;
  }
  return mx;
}

int64_t tc_workspace_bytes(const std::vector<ConvLayer>& layers, int64_t batch, int64_t n) {
  return 3 * (tc_max_act_elems(layers, batch, n) * 2 + 1024) + 1024;
}

int tc_forward(TcState* st, const std::vector<ConvLayer>& layers, const std::vector<float*>& w, const std::vector<float*>& bvec,
               const float* const* stems_host, const float* faders_host, int n_stems, int64_t batch, int64_t n, int apply_tanh,
               float* y, void* workspace, cudaStream_t stream) {
  AA_REQUIRE(batch < (1LL << 24) && n < (1LL << 30), "problem too large");
  EncodeTiledFn encode = get_encode_fn();
  if (!st->weights_valid) {
    for (size_t i = 1; i < layers.size(); ++i) {
      const auto& ly = layers[i];
      const LayerPlan& p = st->plans[i];
      PackArgs pa{};
      pa.n_taps = p.n_taps; pa.cin = ly.cin; pa.k = ly.k; pa.cout = ly.cout; pa.k_total = p.k_total;
      int kc = 0;
      for (int j = 0; j < p.n_taps; ++j) {
        pa.tap_col0[j] = p.tap_col0[j]; pa.tap_width[j] = p.tap_width[j]; pa.tap_kbase[j] = p.tap_kbase[j]; pa.tap_kcol[j] = kc;
        kc += p.tap_width[j];
      }
      const long long tot = (long long)ly.cout * p.k_total;
      pack_weights_kernel<<<(unsigned)std::min<long long>((tot + 255) / 256, 4096), 256, 0, stream>>>(w[i], p.w2, pa);
      AA_LAUNCH_CHECK();
    }
    st->weights_valid = true;
  }
  const int64_t buf_bytes = tc_max_act_elems(layers, batch, n) * 2 + 1024;
  unsigned char* wsb = reinterpret_cast<unsigned char*>(workspace);
  wsb += (1024 - (reinterpret_cast<uintptr_t>(wsb) & 1023)) & 1023;
  __nv_bfloat16* buf[3] = {reinterpret_cast<__nv_bfloat16*>(wsb), reinterpret_cast<__nv_bfloat16*>(wsb + buf_bytes),
                           reinterpret_cast<__nv_bfloat16*>(wsb + 2 * buf_bytes)};
  int cur = 0, res_buf = -1;
  int64_t l = n;
  {  // ---- layer 0 on CUDA cores ----
    const auto& ly = layers[0];
    const int64_t lout = (l + 2 * ly.pad - (int64_t)ly.dil * (ly.k - 1) - 1) / ly.stride + 1;
    AA_REQUIRE(lout == l && ly.dil == 1, "first layer must be a 'same' convolution");
    L0Args a{};
    a.n_in = n_stems;
    for (int s = 0; s < n_stems; ++s) { a.x[s] = stems_host[s]; a.fader[s] = faders_host ? faders_host[s] : 1.0f; }
    a.cin = ly.cin; a.cout = ly.cout; a.k = ly.k; a.pad = ly.pad; a.n = (int)n; a.lpad = (int)rows_padded(lout);
    a.w = w[0]; a.bias = bvec[0]; a.out = buf[0]; a.row_stride = rows_padded(lout);
    const int l0_tpr = (a.lpad + kL0Pos - 1) / kL0Pos;
    const long long l0_tiles = (long long)l0_tpr * batch;
    AA_REQUIRE(l0_tiles < (1LL << 31) - 4096, "problem too large");
    if (ly.cout == 32 && ly.k == 7 && ly.pad == 3 && ly.cin == 2)
      conv_l0_reg_kernel<2><<<(unsigned)std::min<long long>(l0_tiles, 2LL * aa::num_sms()), 256, kL0Smem, stream>>>(a, l0_tpr, (int)l0_tiles);
    else if (ly.cout == 32 && ly.k == 7 && ly.pad == 3 && ly.cin == 1)
      conv_l0_reg_kernel<1><<<(unsigned)std::min<long long>(l0_tiles, 2LL * aa::num_sms()), 256, kL0Smem, stream>>>(a, l0_tpr, (int)l0_tiles);
    else if (ly.cout == 32 && ly.k == 7 && ly.pad == 3)
      conv_l0_c32k7_kernel<<<dim3((unsigned)((a.lpad + kL0Pos - 1) / kL0Pos), (unsigned)batch), 256, kL0Smem, stream>>>(a);
    else if (ly.cout <= 32)
      conv_l0_kernel<<<dim3((unsigned)((a.lpad + 255) / 256), (unsigned)batch), 256, 0, stream>>>(a);
    else
      conv_l0_wide_kernel<<<dim3((unsigned)((a.lpad + 127) / 128), (unsigned)batch), 128, 0, stream>>>(a);
    AA_LAUNCH_CHECK();
    l = lout;
  }
  for (size_t i = 1; i < layers.size(); ++i) {
    const auto& ly = layers[i];
    const LayerPlan& p = st->plans[i];
    const int64_t lout = (l + 2 * ly.pad - (int64_t)ly.dil * (ly.k - 1) - 1) / ly.stride + 1;
    AA_REQUIRE(lout >= 1, "input too short for layer %zu", i);
    const bool last = (i + 1 == layers.size());
    int dst = 0;
    while (dst == cur || dst == res_buf) ++dst;
    // ---- fused ResidualUnit: k7 (dilated) conv -> ELU -> 1x1 conv -> + x -> ELU in one kernel (conv_ru.cuh) ----
    if (st->fuse_ru && ly.role == ROLE_RES_FIRST && i + 2 < layers.size() && layers[i + 1].role == ROLE_RES_SECOND && ly.k == 7 &&
        ly.stride == 1 && ly.cin == ly.cout && (ly.cin == 32 || ly.cin == 64 || (ly.cin == 128 && st->fuse_ru128)) && ly.pad == 3 * ly.dil && ly.dil <= 9 && ly.elu &&
        layers[i + 1].k == 1 && layers[i + 1].stride == 1 && layers[i + 1].cin == ly.cin && layers[i + 1].cout == ly.cin &&
        layers[i + 1].elu && lout == l) {
      RuArgs ra{};
      ra.x = buf[cur]; ra.out = buf[dst];
      ra.w7 = p.w2; ra.w1 = st->plans[i + 1].w2; ra.b7 = bvec[i]; ra.b1 = bvec[i + 1];
      ra.lout = (int)lout; ra.lpad = (int)rows_padded(lout); ra.dil = ly.dil;
      ra.m_tiles = (int)((rows_padded(lout) + BM - 1) / BM);
      ra.rows_alloc = rows_padded(l); ra.tiles = batch * ra.m_tiles;
      CUtensorMap tmX;
      {   // x [B][rows_alloc][C] bf16 as (8 channels, row, panel, batch): box = (8, 128 + 6d rows, C/8 panels, 1)
        cuuint64_t dims[4] = {8, (cuuint64_t)lout, (cuuint64_t)(ly.cin / 8), (cuuint64_t)batch};
        cuuint64_t strides[3] = {(cuuint64_t)ly.cin * 2, 16, (cuuint64_t)rows_padded(l) * ly.cin * 2};
        cuuint32_t box[4] = {8, (cuuint32_t)(BM + 6 * ly.dil), (cuuint32_t)(ly.cin / 8), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf[cur], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        AA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(x tile, fused unit) failed for layer %zu: %d", i, (int)r);
      }
      if (ly.cin == 32) {
        const int grid = (int)std::min<long long>(ra.tiles, (long long)aa::num_sms() * st->ru_ctas_per_sm[0]);
        ru_fused_kernel<32><<<grid, RuCfg<32>::THREADS, RuCfg<32>::SMEM, stream>>>(tmX, ra);
      } else if (ly.cin == 64) {
        const int grid = (int)std::min<long long>(ra.tiles, (long long)aa::num_sms() * st->ru_ctas_per_sm[1]);
        ru_fused_kernel<64><<<grid, RuCfg<64>::THREADS, RuCfg<64>::SMEM, stream>>>(tmX, ra);
      } else {
        const int grid = (int)std::min<long long>(ra.tiles, (long long)aa::num_sms());
        ru_fused_kernel<128><<<grid, RuCfg<128>::THREADS, RuCfg<128>::SMEM, stream>>>(tmX, ra);
      }
      AA_LAUNCH_CHECK();
      cur = dst;
      l = lout;
      ++i;   // the 1x1 layer is part of the fused kernel
      continue;
    }
    if (ly.role == ROLE_RES_FIRST) res_buf = cur;
    const int64_t in_rows_alloc = rows_padded(l);                 // rows per batch element of the input buffer
    const int64_t view_rows = (l + p.view_s - 1) / p.view_s;      // rows of the (possibly strided) view
    // ---- tensor maps ----
    CUtensorMap tmA, tmB;
    {
      cuuint64_t dims[3] = {(cuuint64_t)p.view_c, (cuuint64_t)view_rows, (cuuint64_t)batch};
      cuuint64_t strides[2] = {(cuuint64_t)p.view_c * 2, (cuuint64_t)in_rows_alloc * ly.cin * 2};
      cuuint32_t box[3] = {(cuuint32_t)p.bk, (cuuint32_t)BM, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf[cur], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          p.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      AA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed for layer %zu: %d", i, (int)r);
    }
    const int n_chunks_total = p.k_total / p.bk;
    // K-light layers (1x1 convs) are epilogue bound: narrower N tiles and more epilogue groups in flight
    int bn = std::min(ly.cout, (n_chunks_total <= 8 && !last) ? 128 : 256);
    // 8-chunk layers without a residual (the 128 -> 256 down-conv): one 256-wide tile reads the activations once: 97 -> 86 us.
    // (The C = 512 1x1 layers measured 113 us at bn = 128, 110 us at bn = 256 with res_tma, 147 us at bn = 256 without: unchanged
    //  unless the dev knob AA_1X1_BN256=1 is set.)
    if (n_chunks_total == 8 && !last && ly.cout >= 256 && (ly.role != ROLE_RES_SECOND || st->bn_1x1_wide)) bn = 256;
    // ResidualUnit 1x1 layers at C >= 256: 256-wide tiles on CTA pairs (conv_tc2_kernel<256, true>); measured at B = 64: C = 512
    // 110 -> 90 us, C = 256 78 -> 72 us (403 MB of activations = 61 us at the HBM roofline); AA_TC_CG2R=0: conv_tc_kernel
    const bool pair_res = st->cg2 && st->cg2r && ly.role == ROLE_RES_SECOND && !last && p.bk == 64 && p.n_taps == 1 && ly.cout % 256 == 0 &&
                          (ly.cout >= 256 || st->cg2r >= 3) && ly.elu && rows_padded(lout) >= 2 * BM && aa::num_sms() >= 2;
    if (pair_res) bn = 256;
    {
      cuuint64_t dims[2] = {(cuuint64_t)p.k_total, (cuuint64_t)ly.cout};
      cuuint64_t strides[1] = {(cuuint64_t)p.k_total * 2};
      cuuint32_t box[2] = {(cuuint32_t)p.bk, (cuuint32_t)bn};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.w2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          p.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      AA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed for layer %zu: %d", i, (int)r);
    }
    TcArgs a{};
    a.n_chunks = 0;
    for (int j = 0; j < p.n_taps; ++j)
      for (int c = 0; c < p.tap_width[j]; c += p.bk) {
        a.chunk_off[a.n_chunks] = (short)p.tap_off[j];
        a.chunk_col[a.n_chunks] = (short)(p.tap_col0[j] + c);
        ++a.n_chunks;
      }
    a.bn = bn; a.n_tiles_n = ly.cout / bn; a.m_tiles = (int)((rows_padded(lout) + BM - 1) / BM);
    a.tiles = batch * a.m_tiles * a.n_tiles_n;
    a.lout = (int)lout; a.lpad = (int)rows_padded(lout); a.cout = ly.cout;
    a.bias = bvec[i];
    a.res = (ly.role == ROLE_RES_SECOND) ? buf[res_buf] : nullptr;
    a.out = last ? nullptr : buf[dst];
    a.out_f32 = last ? y : nullptr;
    a.out_row_stride = rows_padded(lout);
    a.elu = ly.elu; a.tanh_out = (last && apply_tanh) ? 1 : 0;
    const int stage_bytes = BM * p.bk * 2 + ((bn * p.bk * 2 + 1023) & ~1023);
    a.n_epi = last ? 1 : (bn <= 64 ? 4 : (bn <= 128 ? (n_chunks_total <= 8 ? 3 : 2) : 1));
    // (dev knob AA_TC_NEPI128, measured on the C = 128 k7 layers: 1 group / 5 stages 124 us, 2 groups / 4 stages 120 us (default),
    //  3 groups / 3 stages 135 us: the layer is not bound by the depth of the operand ring)
    if (st->nepi_k7_128 > 0 && !last && bn == 128 && n_chunks_total > 8) a.n_epi = st->nepi_k7_128;
    a.n_acc = std::min(kMaxAcc, std::min(512 / bn, 2 * a.n_epi));
    if (st->nacc_min > 0) a.n_acc = std::max(a.n_acc, std::min(512 / bn, st->nacc_min));
    // ResidualUnit 1x1 layers with 64-channel boxes: residual in / result out through TMA, in place in one swizzled tile buffer per group
    // (measured per layer at B = 64: C = 128 95 -> 67 us, C = 256 98 -> 79 us; C = 512 -- 8 K chunks, bound by re-streaming the
    //  weights from L2 for every tile -- 110 -> 120 us: those layers keep the per-thread path unless AA_RES_TMA=2)
    a.res_tma = (a.res != nullptr && !last && bn % 64 == 0 && (st->res_tma > 1 || (st->res_tma == 1 && n_chunks_total <= 4))) ? 1 : 0;
    if (pair_res) a.res_tma = st->cg2r;   // the pair kernel always takes the TMA route (2: with the L2 prefetch)
    const int staging = a.res_tma ? 512 + a.n_epi * (BM * bn * 2 + 1024) : a.n_epi * ((last ? 0 : BM * (bn * 2 + 16)) + bn * 4);
    a.stages = std::max(2, std::min(8, (st->max_smem - 2048 - 512 - staging) / stage_bytes));
    int smem = a.stages * stage_bytes + 1024 + 512 + staging;
    const bool halo_shape = a.res == nullptr && !last && ly.stride == 1 && p.n_taps >= 3 && p.bk == 64 && ly.cin % 64 == 0 &&
                            BM + (p.n_taps - 1) * ly.dil <= 256;
    const bool pair_plain = st->cg2 && a.res == nullptr && !last && p.bk == 64 &&
                            (bn == 256 || (bn == 128 && (st->cg2 > 1 || (st->halo >= 3 && halo_shape)))) && ly.cout % bn == 0 &&
                            a.m_tiles >= 2 && aa::num_sms() >= 2;
    // halo mode (conv_tc_kernel, the k7 layers that stay on one CTA per tile, i.e. C = 128): one activation box per channel chunk for all taps
    CUtensorMap tmH = tmA;
    if (halo_shape && !pair_res && (pair_plain ? st->halo != 2 && st->halo >= 1 : st->halo >= 1)) {
      a.halo_rows = BM + (p.n_taps - 1) * ly.dil; a.halo_cc = ly.cin / 64; a.halo_taps = p.n_taps; a.halo_dil = ly.dil; a.halo_row0 = p.tap_off[0];
      const int slot = (a.halo_rows * 128 + 1023) & ~1023, bstage = (bn * 128 + 1023) & ~1023;
      a.a_stages = st->halo_slots;
      if (!pair_plain && st->halo_pair && a.n_tiles_n == 1 && 4 * bn <= 512 && a.tiles >= 2LL * aa::num_sms()) {
        a.halo_pair = 1;                     // two tiles per weight pass: four accumulators, one halo slot per tile and chunk in flight
        a.n_acc = 4;
        a.a_stages = std::max(a.a_stages, st->halo_pair_slots);
      }
      a.stages = std::max(2, std::min(8, (st->max_smem - 2048 - 512 - staging - a.a_stages * slot) / bstage));
      smem = a.a_stages * slot + a.stages * bstage + 1024 + 512 + staging;
      cuuint64_t dims[3] = {(cuuint64_t)p.view_c, (cuuint64_t)view_rows, (cuuint64_t)batch};
      cuuint64_t strides[2] = {(cuuint64_t)p.view_c * 2, (cuuint64_t)in_rows_alloc * ly.cin * 2};
      cuuint32_t box[3] = {64, (cuuint32_t)a.halo_rows, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = encode(&tmH, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf[cur], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      AA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(halo box) failed for layer %zu: %d", i, (int)r);
    }
    CUtensorMap tmR = tmH, tmO = tmA;   // placeholders unless res_tma (tmR = the halo box map in halo mode)
    if (a.res_tma) {
      cuuint64_t dims[3] = {(cuuint64_t)ly.cout, (cuuint64_t)rows_padded(lout), (cuuint64_t)batch};
      cuuint64_t strides[2] = {(cuuint64_t)ly.cout * 2, (cuuint64_t)rows_padded(lout) * ly.cout * 2};
      cuuint32_t box[3] = {64, (cuuint32_t)BM, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = encode(&tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf[res_buf], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      AA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(residual) failed for layer %zu: %d", i, (int)r);
      r = encode(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf[dst], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      AA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(result) failed for layer %zu: %d", i, (int)r);
    }
    AA_REQUIRE(pair_res || pair_plain || smem <= st->max_smem, "layer %zu does not fit in shared memory (%d bytes)", i, smem);
    const int grid = (int)std::min<long long>(a.tiles, aa::num_sms());
    const int threads = 64 + 128 * a.n_epi;
    AA_REQUIRE(last || ly.elu, "bf16 layers without ELU are not supported on the tensor-core path (layer %zu)", i);
    const bool res = a.res != nullptr;
    // (bn = 128 pairs -- AA_TC_CG2=2 -- measured SLOWER: C = 128 k7 120 -> 133 us, 64 -> 128 down-conv 102 -> 110 us; off)
    if (pair_res || pair_plain) {
      const int kCg2BN = bn;
      // wide non-residual layer: CTA pairs, M = 256 cta_group::2 MMAs, each CTA loads half of the weight box (conv_cg2.cuh)
      CUtensorMap tmBh;
      cuuint64_t dims[2] = {(cuuint64_t)p.k_total, (cuuint64_t)ly.cout};
      cuuint64_t strides[1] = {(cuuint64_t)p.k_total * 2};
      cuuint32_t box[2] = {(cuuint32_t)p.bk, (cuuint32_t)(kCg2BN / 2)};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = encode(&tmBh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.w2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      AA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B half) failed for layer %zu: %d", i, (int)r);
      const int stage2 = BM * 64 * 2 + (kCg2BN / 2) * 64 * 2;
      if (!pair_res) a.n_epi = (st->cg2_nepi > 0) ? std::min(2, st->cg2_nepi) : 2;   // column-split epilogue groups (measured: 2 groups -2 ... -4 us per pair layer)
      const int fixed = pair_res ? 1024 + 1024 + kCg2ResGroups * (BM * (kCg2BN / kCg2ResGroups) * 2 + 1024)
                                 : 1024 + 512 + BM * (kCg2BN * 2 + 16 * a.n_epi) + kCg2BN * 4 + 256;
      a.stages = std::max(2, std::min(8, (st->max_smem - fixed) / stage2));
      int smem2 = a.stages * stage2 + fixed;
      if (a.halo_rows > 0) {   // halo mode: A slots + a ring of weight halves
        const int slot = (a.halo_rows * 128 + 1023) & ~1023, bh = (kCg2BN / 2) * 64 * 2;
        a.stages = std::max(2, std::min(8, (st->max_smem - fixed - a.a_stages * slot) / bh));
        smem2 = a.a_stages * slot + a.stages * bh + fixed;
      }
      AA_REQUIRE(smem2 <= st->max_smem, "pair layer %zu does not fit in shared memory (%d bytes)", i, smem2);
      const long long units = batch * ((a.m_tiles + 1) / 2) * a.n_tiles_n;
      const int grid2 = (int)std::min<long long>(2 * units, (long long)(aa::num_sms() & ~1));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)grid2); cfg.blockDim = dim3((unsigned)(pair_res ? kCg2ResThreads : 64 + 128 * a.n_epi)); cfg.dynamicSmemBytes = (size_t)smem2; cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      cudaError_t le = pair_res   ? cudaLaunchKernelEx(&cfg, conv_tc2_kernel<256, true>, tmA, tmBh, tmR, tmO, a)
                       : bn == 256 ? cudaLaunchKernelEx(&cfg, conv_tc2_kernel<256, false>, tmA, tmBh, tmR, tmO, a)
                                   : cudaLaunchKernelEx(&cfg, conv_tc2_kernel<128, false>, tmA, tmBh, tmR, tmO, a);
      AA_REQUIRE(le == cudaSuccess, "conv_tc2_kernel launch failed for layer %zu: %s", i, cudaGetErrorString(le));
      AA_LAUNCH_CHECK();
      if (ly.role == ROLE_RES_SECOND) res_buf = -1;
      cur = dst;
      l = lout;
      continue;
    }
    if (p.bk == 64) {
      if (res) conv_tc_kernel<64, true><<<grid, threads, smem, stream>>>(tmA, tmB, tmR, tmO, a);
      else conv_tc_kernel<64, false><<<grid, threads, smem, stream>>>(tmA, tmB, tmR, tmO, a);
    } else {
      if (res) conv_tc_kernel<32, true><<<grid, threads, smem, stream>>>(tmA, tmB, tmR, tmO, a);
      else conv_tc_kernel<32, false><<<grid, threads, smem, stream>>>(tmA, tmB, tmR, tmO, a);
    }
    AA_LAUNCH_CHECK();
    if (ly.role == ROLE_RES_SECOND) res_buf = -1;
    cur = dst;
    l = lout;
  }
  return AA_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// 3xTF32 path (conv_tf32.cuh): layer-wise, fp32 channels-last activations as hi / lo planes
// ------------------------------------------------------------------------------------------------------------------
struct TfState {
  std::vector<LayerPlan> plans;
  bool weights_valid = false;
  int max_smem = 0;
  int chain = 2;               // K chunks (of 32 channels) per accumulator chain
};

static int tf_id_cols(const ConvLayer& l) { return l.role == ROLE_RES_SECOND ? l.cout : 0; }

// Quiet shape check (no error message): can this layer table run on the 3xTF32 path?  Mirrors tf_create / plan_layer / tf_forward.
bool tf_eligible(const std::vector<ConvLayer>& layers) {
  if (layers.size() < 2) return false;
  const ConvLayer& l0 = layers[0];
  if (l0.stride != 1 || l0.cout != 32 || l0.cin > 4 || l0.k > 7 || l0.dil != 1 || 2 * l0.pad != l0.k - 1 || !l0.elu) return false;
  for (size_t i = 1; i < layers.size(); ++i) {
    const ConvLayer& l = layers[i];
    const bool last = i + 1 == layers.size();
    if (l.cin % 32 != 0 || l.cout % 32 != 0 || (l.cout > 128 && l.cout % 128 != 0)) return false;
    if (!last && !l.elu) return false;
    if (l.stride == 1) {
      if (l.k > 9) return false;
    } else if (!(l.k == 2 * l.stride && l.pad == (l.stride + 1) / 2 && l.dil == 1 && l.stride % 2 == 0)) {
      return false;
    }
    const int k_total = (l.stride == 1 ? l.k : 2 * l.stride) * l.cin;
    if (k_total / 32 + (l.role == ROLE_RES_SECOND ? 4 : 0) > kTfMaxChunks) return false;
    if (l.role == ROLE_RES_SECOND && (l.cin != l.cout || l.stride != 1 || l.k != 1)) return false;
  }
  return true;
}

int tf_create(TfState** out, const std::vector<ConvLayer>& layers) {
  AA_REQUIRE(get_encode_fn() != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  AA_REQUIRE(layers.size() >= 2 && layers[0].stride == 1 && layers[0].cout == 32 && layers[0].cin <= 4 && layers[0].k <= 7,
             "first layer shape not supported on the 3xTF32 path (needs capacity 32, <= 4 input channels, k <= 7)");
  TfState* st = new TfState();
  st->plans.resize(layers.size());
  for (size_t i = 1; i < layers.size(); ++i) {
    int rc = plan_layer(layers[i], st->plans[i], true);
    if (rc != AA_OK) { delete st; return rc; }
    if (layers[i].role == ROLE_RES_SECOND)
      AA_REQUIRE(layers[i].cin == layers[i].cout && layers[i].stride == 1, "residual layer %zu must keep its shape", i);
    AA_CUDA(cudaMalloc(&st->plans[i].w2f, sizeof(float) * 2 * (size_t)layers[i].cout * (st->plans[i].k_total + tf_id_cols(layers[i]))));
  }
  int dev = 0;
  AA_CUDA(cudaGetDevice(&dev));
  AA_CUDA(cudaDeviceGetAttribute(&st->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  AA_CUDA(cudaFuncSetAttribute(conv_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
  if (getenv("AA_TF32_CHAIN")) st->chain = std::max(1, atoi(getenv("AA_TF32_CHAIN")));
  *out = st;
  return AA_OK;
}

void tf_destroy(TfState* st) {
  if (!st) return;
  for (auto& p : st->plans) cudaFree(p.w2f);
  delete st;
}

void tf_invalidate_weights(TfState* st) { if (st) st->weights_valid = false; }

// one activation buffer = two planes of max_act_elems floats (+ alignment slack)
int64_t tf_workspace_bytes(const std::vector<ConvLayer>& layers, int64_t batch, int64_t n) {
  return 3 * (tc_max_act_elems(layers, batch, n) * 8 + 1024) + 1024;
}

static CUresult tf_act_map(EncodeTiledFn encode, CUtensorMap* tm, float* buf, int64_t view_c, int64_t view_rows, int64_t batch,
                           int64_t batch_stride_elems, int64_t plane_elems) {
  cuuint64_t dims[4] = {(cuuint64_t)view_c, (cuuint64_t)view_rows, (cuuint64_t)batch, 2};
  cuuint64_t strides[3] = {(cuuint64_t)view_c * 4, (cuuint64_t)batch_stride_elems * 4, (cuuint64_t)plane_elems * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)BM, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int tf_forward(TfState* st, const std::vector<ConvLayer>& layers, const std::vector<float*>& w, const std::vector<float*>& bvec,
               const float* const* stems_host, const float* faders_host, int n_stems, int64_t batch, int64_t n, int apply_tanh,
               float* y, void* workspace, cudaStream_t stream) {
  AA_REQUIRE(batch < 65536 && n < (1LL << 30), "problem too large");
  EncodeTiledFn encode = get_encode_fn();
  if (!st->weights_valid) {
    for (size_t i = 1; i < layers.size(); ++i) {
      const auto& ly = layers[i];
      const LayerPlan& p = st->plans[i];
      PackArgs pa{};
      pa.n_taps = p.n_taps; pa.cin = ly.cin; pa.k = ly.k; pa.cout = ly.cout; pa.k_total = p.k_total;
      int kc = 0;
      for (int j = 0; j < p.n_taps; ++j) {
        pa.tap_col0[j] = p.tap_col0[j]; pa.tap_width[j] = p.tap_width[j]; pa.tap_kbase[j] = p.tap_kbase[j]; pa.tap_kcol[j] = kc;
        kc += p.tap_width[j];
      }
      const long long tot = (long long)ly.cout * (p.k_total + tf_id_cols(ly));
      pack_weights_tf32_kernel<<<(unsigned)std::min<long long>((tot + 255) / 256, 4096), 256, 0, stream>>>(w[i], p.w2f, pa, tf_id_cols(ly));
      AA_LAUNCH_CHECK();
    }
    st->weights_valid = true;
  }
  const int64_t plane_elems = tc_max_act_elems(layers, batch, n);
  const int64_t buf_bytes = plane_elems * 8 + 1024;
  unsigned char* wsb = reinterpret_cast<unsigned char*>(workspace);
  wsb += (1024 - (reinterpret_cast<uintptr_t>(wsb) & 1023)) & 1023;
  float* buf[3] = {reinterpret_cast<float*>(wsb), reinterpret_cast<float*>(wsb + buf_bytes), reinterpret_cast<float*>(wsb + 2 * buf_bytes)};
  int cur = 0, res_buf = -1;
  int64_t l = n;
  {  // ---- layer 0 on CUDA cores ----
    const auto& ly = layers[0];
    const int64_t lout = (l + 2 * ly.pad - (int64_t)ly.dil * (ly.k - 1) - 1) / ly.stride + 1;
    AA_REQUIRE(lout == l && ly.dil == 1, "first layer must be a 'same' convolution");
    L0fArgs a{};
    a.n_in = n_stems;
    for (int s = 0; s < n_stems; ++s) { a.x[s] = stems_host[s]; a.fader[s] = faders_host ? faders_host[s] : 1.0f; }
    a.cin = ly.cin; a.k = ly.k; a.pad = ly.pad; a.n = (int)n; a.lpad = (int)rows_padded(lout);
    a.w = w[0]; a.bias = bvec[0]; a.out_hi = buf[0]; a.out_lo = buf[0] + plane_elems; a.row_stride = rows_padded(lout);
    conv_l0_tf32_kernel<<<dim3((unsigned)((a.lpad + kL0fPos - 1) / kL0fPos), (unsigned)batch), 256, 0, stream>>>(a);
    AA_LAUNCH_CHECK();
    l = lout;
  }
  for (size_t i = 1; i < layers.size(); ++i) {
    const auto& ly = layers[i];
    const LayerPlan& p = st->plans[i];
    const int64_t lout = (l + 2 * ly.pad - (int64_t)ly.dil * (ly.k - 1) - 1) / ly.stride + 1;
    AA_REQUIRE(lout >= 1, "input too short for layer %zu", i);
    const bool last = (i + 1 == layers.size());
    int dst = 0;
    while (dst == cur || dst == res_buf) ++dst;
    if (ly.role == ROLE_RES_FIRST) res_buf = cur;
    const bool res = ly.role == ROLE_RES_SECOND;
    AA_REQUIRE(!res || (res_buf >= 0 && lout == l), "residual layer %zu without a saved input", i);
    const int64_t in_rows_alloc = rows_padded(l);
    const int64_t view_rows = (l + p.view_s - 1) / p.view_s;
    CUtensorMap tmA, tmB, tmR;
    CUresult r = tf_act_map(encode, &tmA, buf[cur], p.view_c, view_rows, batch, in_rows_alloc * ly.cin, plane_elems);
    AA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A, tf32) failed for layer %zu: %d", i, (int)r);
    tmR = tmA;
    if (res) {   // the ResidualUnit's input x: [B][rows_padded(l)][cout], rows >= l read as zeros
      r = tf_act_map(encode, &tmR, buf[res_buf], ly.cout, l, batch, in_rows_alloc * ly.cout, plane_elems);
      AA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(R, tf32) failed for layer %zu: %d", i, (int)r);
    }
    const int bn = std::min(ly.cout, 128);
    AA_REQUIRE(ly.cout % bn == 0 && bn % 32 == 0, "cout=%d must be a multiple of 32 (and of 128 above 128) on the 3xTF32 path", ly.cout);
    const int kt_ext = p.k_total + tf_id_cols(ly);
    {
      cuuint64_t dims[3] = {(cuuint64_t)kt_ext, (cuuint64_t)ly.cout, 2};
      cuuint64_t strides[2] = {(cuuint64_t)kt_ext * 4, (cuuint64_t)ly.cout * kt_ext * 4};
      cuuint32_t box[3] = {32, (cuuint32_t)bn, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p.w2f, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      AA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B, tf32) failed for layer %zu: %d", i, (int)r);
    }
    TfArgs a{};
    a.n_chunks = 0;
    for (int j = 0; j < p.n_taps; ++j)
      for (int c = 0; c < p.tap_width[j]; c += 32) {
        a.chunk_off[a.n_chunks] = (short)p.tap_off[j];
        a.chunk_col[a.n_chunks] = (short)(p.tap_col0[j] + c);
        ++a.n_chunks;
      }
    a.n_main = a.n_chunks;
    if (res)
      for (int c = 0; c < bn; c += 32) { a.chunk_off[a.n_chunks] = 0; a.chunk_col[a.n_chunks] = (short)c; ++a.n_chunks; }
    AA_REQUIRE(a.n_chunks <= kTfMaxChunks, "too many K chunks (%d)", a.n_chunks);
    a.chain = st->chain;
    a.bn = bn; a.n_tiles_n = ly.cout / bn; a.m_tiles = (int)((rows_padded(lout) + BM - 1) / BM);
    a.tiles = batch * a.m_tiles * a.n_tiles_n;
    a.lout = (int)lout; a.lpad = (int)rows_padded(lout); a.cout = ly.cout;
    a.bias = bvec[i];
    a.out_hi = last ? nullptr : buf[dst];
    a.out_lo = last ? nullptr : buf[dst] + plane_elems;
    a.out_f32 = last ? y : nullptr;
    a.out_row_stride = rows_padded(lout);
    a.elu = ly.elu; a.tanh_out = (last && apply_tanh) ? 1 : 0;
    AA_REQUIRE(last || ly.elu, "hidden layers without ELU are not supported on the 3xTF32 path (layer %zu)", i);
    const int fixed = 1024 + 512 + ly.cout * 4;            // alignment slack, barriers, bias
    const int stage_bytes = 2 * BM * 128 + 2 * bn * 128;
    a.stages = std::max(2, std::min(8, (st->max_smem - fixed) / stage_bytes));
    const int smem = a.stages * stage_bytes + fixed;
    AA_REQUIRE(smem <= st->max_smem, "layer %zu does not fit in shared memory (%d bytes)", i, smem);
    const int grid = (int)std::min<long long>(a.tiles, aa::num_sms());
    conv_tf32_kernel<<<grid, kTfThreads, smem, stream>>>(tmA, tmB, tmR, a);
    AA_LAUNCH_CHECK();
    if (res) res_buf = -1;
    cur = dst;
    l = lout;
  }
  return AA_OK;
}

}  // namespace aa
