// conv_tc2_kernel -- the wide (bn = 256) non-residual bf16 conv layers as `cta_group::2` MMAs (included by conv_tc.cu).
//
// Why: the k7 layers at C >= 256 are bound by operand streaming INTO the SM (~64 B/clk/SM through TMA): with one
// CTA per 128 x 256 output tile every K chunk brings 16 KB of activations + 32 KB of weights for 512 clocks of MMAs = 96 B/clk.
// A CTA PAIR (2-CTA cluster) computes a 256 x 256 tile with M = 256 MMAs issued by the leader CTA: each CTA loads its own 128
// activation rows and only HALF of the weight box (N / 2 = 128 rows of it), 32 KB per 512 clocks = 64 B/clk, and keeps its own
// 128 x 256 accumulator (double buffered: 2 x 256 TMEM columns) and its own epilogue.
//
// Protocol (rank = %cluster_ctarank, leader = rank 0; both CTAs walk the same unit sequence):
//   producers (warp 0 of both CTAs): wait own empty[s]; the leader arms ITS full[s] with the bytes of both CTAs; each CTA's TMA
//       loads (A rows of its row tile, its weight half) complete on the LEADER's full[s] (cluster address of rank 0);
//   MMA (warp 1 of the leader): wait full[s]; 4 x tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) -- the hardware reads the A
//       rows and the B half of each CTA at the same shared-memory offsets; tcgen05.commit.cta_group::2 multicast frees empty[s] in
//       both CTAs and, after the last chunk, completes tfull[acc] in both;
//   epilogues (4 warps per CTA): wait own tfull[acc]; TMEM -> +bias -> ELU -> bf16 -> staging -> coalesced stores; arrive on the
//       LEADER's tempty[acc] (8 arrivals: 4 warps x 2 CTAs) before the leader reuses the accumulator.
#pragma once

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {   // shared::cluster address of `saddr` in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads of a CTA pair: `.cta_group::2` lets the load complete on the LEADER's mbarrier (shared::cluster address with the CTA
// rank bit cleared, as CUTLASS' SM100_TMA_2SM_LOAD does) while the box lands in the issuing CTA's own shared memory
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

constexpr int kCg2Threads = 64 + 128;   // TMA warp, MMA warp, one epilogue group

template <int kCg2BN>   // N of the pair's tile: 256 (C >= 256 layers) or 128 (the C = 128 k7 layers)
__global__ void __launch_bounds__(kCg2Threads, 1) conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh,
                                                                  const TcArgs a) {
  constexpr int BK = 64;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = BM * BK * 2, BH_BYTES = (kCg2BN / 2) * BK * 2, STAGE = A_BYTES + BH_BYTES;
  const uint32_t bars = base + (uint32_t)a.stages * STAGE;   // full[S], empty[S], tfull[2], tempty[2], tmem slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (a.stages + s); };
  auto tfull_bar = [&](int i) { return bars + 8u * (2 * a.stages + i); };
  auto tempty_bar = [&](int i) { return bars + 8u * (2 * a.stages + 2 + i); };
  const uint32_t tmem_slot = bars + 8u * (2 * a.stages + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
  }
  if (warp == 1) {   // both CTAs of the pair allocate together (same warp id, same destination offset)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * kCg2BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers and TMEM exist before any cross-CTA signal
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // unit u of this cluster: (batch, row-tile pair, N tile); CTA `rank` owns row tile 2 p + rank (an odd row-tile count leaves a
  // ghost tile: zero-filled loads, clipped stores)
  const int mpairs = (a.m_tiles + 1) >> 1;
  const long long units = (a.tiles / ((long long)a.m_tiles * a.n_tiles_n)) * mpairs * a.n_tiles_n;
  const long long ufirst = (long long)(blockIdx.x >> 1), ustep = (long long)(gridDim.x >> 1);
  const int my_units = units > ufirst ? (int)((units - ufirst + ustep - 1) / ustep) : 0;
  auto unit_coords = [&](int i, int& b, int& m0, int& n0) {
    const long long u = ufirst + (long long)i * ustep;
    const int per_b = mpairs * a.n_tiles_n;
    b = (int)(u / per_b);
    const int r = (int)(u % per_b);
    n0 = (r % a.n_tiles_n) * kCg2BN;
    m0 = (2 * (r / a.n_tiles_n) + (int)rank) * BM;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_units; ++it) {
        int b, m0, n0;
        unit_coords(it, b, m0, n0);
        for (int q = 0; q < a.n_chunks; ++q) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (rank == 0) mbar_expect_tx(full_bar(s), 2u * STAGE);           // the leader's barrier counts both CTAs' bytes
          const uint32_t sa = base + (uint32_t)s * STAGE;
          tma2_load_3d(sa, &tmA, full_bar(s), a.chunk_col[q], m0 + a.chunk_off[q], b);
          tma2_load_2d(sa + A_BYTES, &tmBh, full_bar(s), q * BK, n0 + (int)rank * (kCg2BN / 2));
          if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      // D = F32, A = B = BF16, K-major, N = 256, M = 256 (cta_group::2)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kCg2BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_units; ++it) {
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), (((uint32_t)(it >> 1)) & 1u) ^ 1u);   // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kCg2BN);
        for (int q = 0; q < a.n_chunks; ++q) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)s * STAGE;
          const uint64_t da = make_desc<BK>(sa), db = make_desc<BK>(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma2_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (q | k) != 0 ? 1u : 0u);
          umma2_commit_mc(empty_bar(s), (uint16_t)3);   // frees the stage in both CTAs when these MMAs retire
          if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
        umma2_commit_mc(tfull_bar(acc), (uint16_t)3);   // accumulators complete in both CTAs
      }
    }
  } else {
    // ===================== epilogue (both CTAs): warp w owns TMEM lanes 32 (w % 4) .. + 31 of its CTA's 128 rows =====================
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const int et = (threadIdx.x - 64) & 127;
    const uint32_t RS = (uint32_t)kCg2BN * 2u + 16u;
    const uint32_t stg = bars + 512u;
    const uint32_t sbias = stg + (uint32_t)BM * RS;
    const uint32_t tempty0 = map_to_cta(tempty_bar(0), 0u), tempty1 = map_to_cta(tempty_bar(1), 0u);
    int bias_n0 = -1;
    for (int it = 0; it < my_units; ++it) {
      const int acc = it & 1;
      int b, m0, n0;
      unit_coords(it, b, m0, n0);
      const bool valid = m0 + row_in_tile < a.lout;
      const long long tile_off = ((long long)b * a.out_row_stride + m0) * a.cout + n0;
      if (n0 != bias_n0) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = et; i < kCg2BN; i += 128) {
          const float bv = (n0 + i < a.cout) ? __ldg(a.bias + n0 + i) : 0.f;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + 4u * i), "f"(bv) : "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        bias_n0 = n0;
      }
      mbar_wait(tfull_bar(acc), ((uint32_t)(it >> 1)) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kCg2BN);
      for (int c0 = 0; c0 < kCg2BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        float bia[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bia[j]), "=f"(bia[j + 1]), "=f"(bia[j + 2]), "=f"(bia[j + 3])
                       : "r"(sbias + 4u * (uint32_t)(c0 + j)));
        const uint32_t srow = stg + (uint32_t)row_in_tile * RS + (uint32_t)c0 * 2u;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float2 y0 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1])), make_float2(bia[8 * g + 0], bia[8 * g + 1]));
          const float2 y1 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])), make_float2(bia[8 * g + 2], bia[8 * g + 3]));
          const float2 y2 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])), make_float2(bia[8 * g + 4], bia[8 * g + 5]));
          const float2 y3 = __fadd2_rn(make_float2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])), make_float2(bia[8 * g + 6], bia[8 * g + 7]));
          uint4 o = make_uint4(pack_bf16(elu2(y0)), pack_bf16(elu2(y1)), pack_bf16(elu2(y2)), pack_bf16(elu2(y3)));
          if (!valid) o = make_uint4(0u, 0u, 0u, 0u);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (uint32_t)g * 16u), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc ? tempty1 : tempty0);   // the leader may reuse the accumulator once all 8 warps arrived
      asm volatile("bar.sync 1, 128;" ::: "memory");      // tile complete in staging
      for (int idx = et; idx < BM * (kCg2BN / 8); idx += 128) {
        const int rr = idx / (kCg2BN / 8), cv = idx % (kCg2BN / 8);
        if (m0 + rr < a.lpad) {
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(stg + (uint32_t)rr * RS + (uint32_t)cv * 16u));
          reinterpret_cast<uint4*>(a.out + tile_off + (long long)rr * a.cout)[cv] = v;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");      // staging free for the next tile
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody leaves while the pair may still read its shared memory / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * kCg2BN)) : "memory");
  }
}
