// PCA scatter of 64-channel latents on tcgen05 (calc_effects_pca.py:81-89: rearrange 'b d n -> d (b n)', torch.cov * (n - 1)):
//
//     S = sum_p (y_p - c)(y_p - c)^T   (64 x 64),   s = sum_p (y_p - c),   cov_numerator += S - s s^T / n
//
// over all n = B T' points p of y [B][64][T'] in ONE pass over the latents (c = the first point: a pivot that keeps the raw
// moments well conditioned; the result does not depend on it).  The 64 x 64 rank-n update is a GEMM with K = points: every fp32
// value is split into hi = its top 19 bits (exact in TF32) and lo = value - hi, and ONE tcgen05.mma.kind::tf32 per 8 points
// computes both products the fp32 parity needs: A = [hi rows 0..63 ; lo rows 64..127] (M = 128), B = hi (N = 64, the same shared
// memory rows as A's upper half), so that the accumulator holds HH = hi hi^T in TMEM lanes 0..63 and LH = lo hi^T in lanes
// 64..127; S = HH + LH + LH^T (the dropped lo lo^T term is 2^-22 relative).  Operands sit in the no-swizzle K-major panel
// layout [points / 4][128 rows][16 B] (proj_tc.cu): a loader thread's float4 of four consecutive points of one channel is one
// 16-byte store per plane.  tcgen05 accumulates into TMEM with truncation (profiles/ubench/tf32_accum_ubench), so a chain is
// closed every kChain tiles: the epilogue warps (thread = TMEM lane) add the finished accumulator into registers while the
// tensor core continues in the other TMEM slot.  Persistent CTA per SM: warp 0 issues, warps 1..4 load / split / sum the channel
// totals, warps 5..8 drain.  Partials of all CTAs are summed in CTA order by a coalesced second kernel; a third (one block) forms HH + LH + LH^T - s s^T / n.
#include "aa_common.cuh"

#include <algorithm>

namespace {

constexpr int GC = 64;                    // channels
constexpr int GP = 64;                    // points per tile
constexpr int GPS = 128 * 16 + 16;        // panel stride in bytes (128 rows x 16 B; +16 keeps the 16 panels off the same banks)
constexpr int GA_BYTES = (GP / 4) * GPS;  // one operand buffer: 16 panels
constexpr int kGramStages = 2;
constexpr int kChain = 2;                 // tiles per accumulator chain (16 MMAs)
constexpr int kGramThreads = 32 + 128 + 128;
constexpr int kGramSmem = 256 + kGramStages * GA_BYTES + 128;

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "GR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra GR_DONE;\n"
      "bra GR_WAIT;\n"
      "GR_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void g_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void g_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void g_umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void g_umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void g_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
__device__ __forceinline__ uint64_t g_desc_ns(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ float g_tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

struct GramArgs {
  const float* y;         // [b][64][t]
  long long n_tiles;      // b * tiles_t
  int t, tiles_t;
  float* parts;           // [grid][128 * 64 + 64]: HH rows, LH rows, channel sums
  unsigned int* counter;  // arrival counter of the reduction kernel (zeroed here)
};

__global__ void __launch_bounds__(kGramThreads, 1) gram_tc_kernel(const GramArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (g_smem_u32(smem_raw) + 127u) & ~127u;
  // barriers: full[s] (loaders -> MMA), empty[s] (MMA done reading stage s), accfull[q], accempty[q] (TMEM slot q)
  const uint32_t full0 = base, empty0 = base + 16, accfull0 = base + 32, accempty0 = base + 48, tmem_slot = base + 64;
  const uint32_t sA = base + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kGramStages; ++s) {
      g_mbar_init(full0 + 8 * s, 4);      // one arrival per loader warp
      g_mbar_init(empty0 + 8 * s, 1);     // tcgen05.commit
    }
    for (int q = 0; q < 2; ++q) {
      g_mbar_init(accfull0 + 8 * q, 1);   // tcgen05.commit
      g_mbar_init(accempty0 + 8 * q, 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  g_fence_before();
  __syncthreads();
  g_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.counter = 0u;
  const long long my_tiles = a.n_tiles > blockIdx.x ? (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = TF32 (format 2), both K-major, N = 64, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (long long i = 0; i < my_tiles; ++i) {
        const int s = (int)(i % kGramStages);
        const long long chain = i / kChain;
        const int q = (int)(chain & 1);
        const bool first = (i % kChain) == 0;
        if (first && chain >= 2) {   // the epilogue has drained this TMEM slot
          g_mbar_wait(accempty0 + 8 * q, (uint32_t)(((chain >> 1) - 1) & 1));
          g_fence_after();
        }
        g_mbar_wait(full0 + 8 * s, (uint32_t)((i / kGramStages) & 1));
        g_fence_after();
        const uint32_t ab = sA + (uint32_t)s * GA_BYTES;
#pragma unroll
        for (int kk = 0; kk < GP / 8; ++kk) {
          const uint64_t d = g_desc_ns(ab + (uint32_t)(2 * kk) * GPS, GPS);   // A = all 128 rows, B = rows 0..63 of the same panels
          g_umma_tf32(tmem_base + (uint32_t)q * 64u, d, d, idesc, (first && kk == 0) ? 0u : 1u);
        }
        g_umma_commit(empty0 + 8 * s);
        if ((i % kChain) == kChain - 1 || i == my_tiles - 1) g_umma_commit(accfull0 + 8 * q);
      }
    }
  } else if (warp <= 4) {
    // ===================== loaders: thread = 8 float4 (4 consecutive points of one channel) per tile =====================
    const int lt = threadIdx.x - 32;               // 0..127
    const int pq = lt & 15;                        // panel (points 4 pq .. 4 pq + 3 of the tile)
    const int ch0 = lt >> 4;                       // channels ch0 + 8 j
    float piv[8], csum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      piv[j] = __ldg(a.y + (long long)(ch0 + 8 * j) * a.t);   // the first point of the channel
      csum[j] = 0.f;
    }
    // three tiles of loads are in flight per thread (48 KB per SM: what the HBM latency-bandwidth product asks for)
    constexpr int kDepth = 3;
    float4 v[kDepth][8];
    bool vok[kDepth];
    auto issue = [&](int slot, long long i) {
      const long long tile = blockIdx.x + i * gridDim.x;
      const long long bi = tile / a.tiles_t;
      const int t0 = (int)(tile - bi * a.tiles_t) * GP + 4 * pq;
      vok[slot] = i < my_tiles && t0 < a.t;            // t % 4 == 0: a float4 is inside or outside as a whole
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[slot][j] = vok[slot] ? __ldg(reinterpret_cast<const float4*>(a.y + (bi * GC + ch0 + 8 * j) * (long long)a.t + t0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto process = [&](int slot, long long i) {
      const int s = (int)(i % kGramStages);
      if (i >= kGramStages) g_mbar_wait(empty0 + 8 * s, (uint32_t)(((i / kGramStages) - 1) & 1));   // the tensor core has read this stage
      const uint32_t ab = sA + (uint32_t)s * GA_BYTES + (uint32_t)pq * GPS;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 d = make_float4(v[slot][j].x - piv[j], v[slot][j].y - piv[j], v[slot][j].z - piv[j], v[slot][j].w - piv[j]);
        if (!vok[slot]) d = make_float4(0.f, 0.f, 0.f, 0.f);
        csum[j] += (d.x + d.y) + (d.z + d.w);
        const float4 hi = make_float4(g_tf32_hi(d.x), g_tf32_hi(d.y), g_tf32_hi(d.z), g_tf32_hi(d.w));
        const uint32_t row = (uint32_t)(ch0 + 8 * j);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ab + row * 16u), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ab + (row + 64u) * 16u), "f"(d.x - hi.x), "f"(d.y - hi.y),
                     "f"(d.z - hi.z), "f"(d.w - hi.w)
                     : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the operand is read by the tensor core
      __syncwarp();
      if (lane == 0) g_mbar_arrive(full0 + 8 * s);
    };
    issue(0, 0);
    issue(1, 1);
    for (long long i = 0; i < my_tiles; i += kDepth) {
      issue(2, i + 2);
      process(0, i);
      if (i + 1 < my_tiles) { issue(0, i + 3); process(1, i + 1); }
      if (i + 2 < my_tiles) { issue(1, i + 4); process(2, i + 2); }
    }
    // channel sums: the 16 threads of a half warp share ch0 (lt >> 4): reduce over pq, fixed order
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float sv = csum[j];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
      if (pq == 0) a.parts[(size_t)blockIdx.x * (128 * 64 + 64) + 128 * 64 + ch0 + 8 * j] = sv;
    }
  } else {
    // ===================== epilogue: thread = TMEM lane (row of [HH ; LH]), 64 running sums in registers =====================
    const int quarter = warp & 3;                  // warps 5..8 -> TMEM lane quarters 1, 2, 3, 0
    const int r = quarter * 32 + lane;
    float acc[GC];
#pragma unroll
    for (int j = 0; j < GC; ++j) acc[j] = 0.f;
    const long long n_chains = (my_tiles + kChain - 1) / kChain;
    for (long long c = 0; c < n_chains; ++c) {
      const int q = (int)(c & 1);
      g_mbar_wait(accfull0 + 8 * q, (uint32_t)((c >> 1) & 1));
      g_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)q * 64u;
      uint32_t v[32];
      g_tmem_ld32(taddr, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
      g_tmem_ld32(taddr + 32u, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[32 + j] += __uint_as_float(v[j]);
      g_fence_before();
      __syncwarp();
      if (lane == 0) g_mbar_arrive(accempty0 + 8 * q);
    }
    float* dst = a.parts + (size_t)blockIdx.x * (128 * 64 + 64) + (size_t)r * 64;
#pragma unroll
    for (int j = 0; j < GC; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }
  g_fence_before();
  __syncthreads();
  if (warp == 0) {
    g_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// Partials of all CTAs summed in CTA order (bitwise reproducible), coalesced: R[e] = sum_p parts[p][e]
// as above for e < 128 * 64 + 64, then the LAST block to finish (arrival counter; the arithmetic order does not depend on which
// block that is) forms cov_num[i][j] += HH[i][j] + LH[i][j] + LH[j][i] - s_i s_j / n
__global__ void __launch_bounds__(256) gram_tc_sum_final_kernel(const float* __restrict__ parts, int n_parts, float* __restrict__ R,
                                                                unsigned int* __restrict__ counter, double n_points, float* __restrict__ cov_num,
                                                                double* __restrict__ count) {
  __shared__ float sLH[GC][GC + 1];
  __shared__ float ssum[GC];
  __shared__ bool last;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = 128 * 64 + 64;
  if (e < n) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int p = 0;
    for (; p + 3 < n_parts; p += 4) {
      s0 += __ldg(parts + (size_t)p * n + e);
      s1 += __ldg(parts + (size_t)(p + 1) * n + e);
      s2 += __ldg(parts + (size_t)(p + 2) * n + e);
      s3 += __ldg(parts + (size_t)(p + 3) * n + e);
    }
    for (; p < n_parts; ++p) s0 += __ldg(parts + (size_t)p * n + e);
    R[e] = (s0 + s1) + (s2 + s3);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  const volatile float* Rv = R;
  for (int q = threadIdx.x; q < GC * GC; q += blockDim.x) sLH[q / GC][q % GC] = Rv[64 * 64 + q];
  if (threadIdx.x < GC) ssum[threadIdx.x] = Rv[128 * 64 + threadIdx.x];
  __syncthreads();
  for (int q = threadIdx.x; q < GC * GC; q += blockDim.x) {
    const int i = q / GC, j = q % GC;
    const float g = (Rv[q] + sLH[i][j]) + sLH[j][i];
    cov_num[q] += g - (float)((double)ssum[i] * (double)ssum[j] / n_points);
  }
  if (threadIdx.x == 0) {
    if (count) count[0] += n_points;
    *counter = 0u;
  }
}
}  // namespace

namespace aa {

int64_t gram_tc_workspace_floats() { return (int64_t)(num_sms() + 1) * (128 * 64 + 64) + 16; }

// PCA scatter of y [b][64][t] (t % 4 == 0, y 16-byte aligned) accumulated into cov_num [64][64] and count; see the file header.
int gram_tc(const float* y, int64_t b, int64_t t, float* cov_num, double* count, float* workspace, cudaStream_t stream) {
  AA_CUDA(aa::ensure_dyn_smem(gram_tc_kernel, kGramSmem));
  GramArgs a;
  a.y = y; a.t = (int)t; a.tiles_t = (int)((t + GP - 1) / GP); a.n_tiles = b * a.tiles_t; a.parts = workspace;
  const int grid = (int)std::min<long long>(a.n_tiles, (long long)num_sms());
  a.counter = reinterpret_cast<unsigned int*>(workspace + (size_t)grid * (128 * 64 + 64) + 128 * 64 + 64);
  gram_tc_kernel<<<grid, kGramThreads, kGramSmem, stream>>>(a);
  AA_LAUNCH_CHECK();
  float* R = workspace + (size_t)grid * (128 * 64 + 64);
  unsigned int* counter = reinterpret_cast<unsigned int*>(R + 128 * 64 + 64);   // zeroed by gram_tc_kernel, reset by its last reader
  gram_tc_sum_final_kernel<<<(128 * 64 + 64 + 255) / 256, 256, 0, stream>>>(workspace, grid, R, counter, (double)b * (double)t, cov_num, count);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

}  // namespace aa
