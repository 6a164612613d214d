// conv_tc2_kernel -- the wide (bn = 256) bf16 conv layers as `cta_group::2` MMAs (included by conv_tc.cu).
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
//   epilogues (n_epi groups of 4 warps per CTA, each owning a column slice of the tile): wait own tfull[acc]; TMEM -> +bias -> ELU
//       -> bf16 -> staging -> coalesced stores; arrive on the LEADER's tempty[acc] (8 n_epi arrivals: groups x 4 warps x 2 CTAs)
//       before the leader reuses the accumulator.
// Halo mode (stride-1 k-tap layers): the activations of a channel chunk arrive ONCE per CTA as a [128 + (k - 1) d rows x 64] box in
// a ring of their own (afull / aempty, both CTAs' boxes on the leader's barrier) and every tap is a row-shifted descriptor view of
// it; the operand ring then carries the weight halves only.  RES = true: the ResidualUnit 1x1 layers (see the kernel's comment).
// Producer and MMA warps run their loops with all 32 lanes and elect one per instruction (see elect_one in conv_tc.cu).
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

// tcgen05.ld without the wait, and the wait tied to the destination registers (so that no use can be scheduled above it):
// several loads in flight before one wait
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                 "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                 "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

constexpr int kCg2Threads = 64 + 128;   // TMA warp, MMA warp, one epilogue group
constexpr int kCg2Threads2 = 64 + 256;  // ... or two (a.n_epi = 2: the groups split the tile's columns)
constexpr int kCg2ResGroups = 4;           // RES: four epilogue groups, each owns a quarter of the tile's columns
constexpr int kCg2ResThreads = 64 + 128 * kCg2ResGroups;

__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {   // box -> L2 only
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// RES = true: the ResidualUnit 1x1 layers at C >= 256 (K = C is short, so one CTA per 128 x 128 tile re-streams 2 x 16 KB per 257
// clocks of MMAs -- twice what the SM can ingest; a pair at N = 256 brings 32 KB per 512 clocks).  Four epilogue groups per CTA split
// the tile's COLUMNS (all drain the same accumulator: an epilogue warp alone on its scheduler is latency bound -- ~10 k clocks per
// 128 x 128 tile -- so the per-tile epilogue time falls with the number of groups and comes under the short K loop); the
// residual tile comes in and the result leaves through TMA, in place in the group's SWIZZLE_128B buffer, as in conv_tc_kernel.
template <int kCg2BN, bool RES>   // N of the pair's tile: 256 (C >= 256 layers) or 128 (the C = 128 k7 layers)
__global__ void __launch_bounds__(RES ? kCg2ResThreads : kCg2Threads2, 1) conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmBh,
                                                                  const __grid_constant__ CUtensorMap tmR,
                                                                  const __grid_constant__ CUtensorMap tmO, const TcArgs a) {
  constexpr int BK = 64;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = BM * BK * 2, BH_BYTES = (kCg2BN / 2) * BK * 2;
  // halo mode (non-RES k-tap layers, as in conv_tc_kernel): one [halo_rows x 64] activation box per channel chunk in its own ring
  // (afull / aempty at bars + 448 / 480), the operand ring then carries the weight halves only
  const bool halo = !RES && a.halo_rows > 0;
  const uint32_t HALO_BYTES = (uint32_t)a.halo_rows * (uint32_t)(BK * 2), HALO_SLOT = halo ? ((HALO_BYTES + 1023u) & ~1023u) : 0u;
  const uint32_t STAGE = halo ? BH_BYTES : A_BYTES + BH_BYTES;
  const uint32_t ring = base + (uint32_t)a.a_stages * HALO_SLOT;
  const uint32_t bars = ring + (uint32_t)a.stages * STAGE;   // full[S], empty[S], tfull[2], tempty[2], tmem slot
  auto afull_bar = [&](int i) { return bars + 448u + 8u * i; };
  auto aempty_bar = [&](int i) { return bars + 480u + 8u * i; };
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (a.stages + s); };
  auto tfull_bar = [&](int i) { return bars + 8u * (2 * a.stages + i); };
  auto tempty_bar = [&](int i) { return bars + 8u * (2 * a.stages + 2 + i); };
  const uint32_t tmem_slot = bars + 8u * (2 * a.stages + 4);
  // RES: group g's [CW / 64][128 rows][128 B] SWIZZLE_128B tile buffer (residual in, result out) at bars + 1024 + g (RBUF + 1024),
  // then its bias copy; rfull[g] = residual landed (this CTA's own barrier), rempty[g] = the previous result has left the buffer
  constexpr int NG = RES ? kCg2ResGroups : 1;      // epilogue groups
  constexpr int CW = kCg2BN / NG;                  // columns per group
  constexpr uint32_t RBUF = (uint32_t)BM * CW * 2u;
  auto rfull_bar = [&](int g) { return bars + 384u + 8u * g; };
  auto rempty_bar = [&](int g) { return bars + 416u + 8u * g; };
  auto rbuf = [&](int g) { return bars + 1024u + (uint32_t)g * (RBUF + 1024u); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), RES ? 8 * NG : 8 * a.n_epi); }
    for (int g = 0; g < NG; ++g) { mbar_init(rfull_bar(g), 1); mbar_init(rempty_bar(g), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(afull_bar(i), 1); mbar_init(aempty_bar(i), 1); }
    if (halo) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");   // tmR = the halo box map in this mode
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
    if (RES) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
    }
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
    // ===================== TMA producer (both CTAs; whole warp runs the loop, one elected lane issues: see elect_one) =====================
    {
      const uint32_t ubase = __shfl_sync(0xffffffffu, base, 0), ubars = __shfl_sync(0xffffffffu, bars, 0);
      const uint32_t uring = __shfl_sync(0xffffffffu, ring, 0), urank = __shfl_sync(0xffffffffu, rank, 0);
      auto ufull = [&](int s) { return ubars + 8u * s; };
      auto uempty = [&](int s) { return ubars + 8u * (a.stages + s); };
      int s = 0, has = 0;
      uint32_t ph = 0, hph = 0;
      for (int it = 0; it < my_units; ++it) {
        int b, m0, n0;
        unit_coords(it, b, m0, n0);
        if (halo) {
          for (int cc = 0; cc < a.halo_cc; ++cc) {
            mbar_wait(ubars + 480u + 8u * has, hph ^ 1u);
            if (elect_one()) {
              if (urank == 0) mbar_expect_tx(ubars + 448u + 8u * has, 2u * HALO_BYTES);
              tma2_load_3d(ubase + (uint32_t)has * HALO_SLOT, &tmR, ubars + 448u + 8u * has, cc * BK, m0 + a.halo_row0, b);
            }
            if (++has == a.a_stages) { has = 0; hph ^= 1u; }
            for (int j = 0; j < a.halo_taps; ++j) {
              mbar_wait(uempty(s), ph ^ 1u);
              if (elect_one()) {
                if (urank == 0) mbar_expect_tx(ufull(s), 2u * BH_BYTES);
                tma2_load_2d(uring + (uint32_t)s * STAGE, &tmBh, ufull(s), (j * a.halo_cc + cc) * BK, n0 + (int)urank * (kCg2BN / 2));
              }
              if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
          }
          continue;
        }
        if (RES && a.res_tma > 1 && elect_one())   // this tile's residual -> L2 now, so that the load below (issued once the buffers are free) is short
          for (int bx = 0; bx < kCg2BN / 64; ++bx) tma_prefetch_3d(&tmR, n0 + 64 * bx, m0, b);
        for (int q = 0; q < a.n_chunks; ++q) {
          mbar_wait(uempty(s), ph ^ 1u);
          if (elect_one()) {
            if (urank == 0) mbar_expect_tx(ufull(s), 2u * STAGE);           // the leader's barrier counts both CTAs' bytes
            const uint32_t sa = uring + (uint32_t)s * STAGE;
            tma2_load_3d(sa, &tmA, ufull(s), a.chunk_col[q], m0 + a.chunk_off[q], b);
            tma2_load_2d(sa + A_BYTES, &tmBh, ufull(s), q * BK, n0 + (int)urank * (kCg2BN / 2));
          }
          if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
        if (RES) {   // this CTA's residual tile, one column slice per epilogue group, on the CTA's OWN barriers
          for (int g = 0; g < NG; ++g) {
            if (it > 0) mbar_wait(ubars + 416u + 8u * g, (uint32_t)(it - 1) & 1u);
            if (elect_one()) {
              const uint32_t rb = ubars + 1024u + (uint32_t)g * (RBUF + 1024u), rf = ubars + 384u + 8u * g;
              mbar_expect_tx(rf, RBUF);
              for (int bx = 0; bx < CW / 64; ++bx) tma_load_3d(rb + (uint32_t)bx * 16384u, &tmR, rf, n0 + g * CW + 64 * bx, m0, b);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; whole warp runs the loop, one elected lane issues) =====================
    // (with the loop under `lane == 0` every tcgen05.mma sits in an ELECT / R2UR.BROADCAST / BRA.U.ANY sequence of ~130 clocks:
    //  more than the 128 clocks an M = 256, N = 256, K = 16 MMA needs, twice what N = 128 needs)
    if (rank == 0) {
      const uint32_t ubase = __shfl_sync(0xffffffffu, base, 0), ubars = __shfl_sync(0xffffffffu, bars, 0);
      const uint32_t uring = __shfl_sync(0xffffffffu, ring, 0), utmem = __shfl_sync(0xffffffffu, tmem_base, 0);
      auto ufull = [&](int s) { return ubars + 8u * s; };
      auto uempty = [&](int s) { return ubars + 8u * (a.stages + s); };
      auto utfull = [&](int i) { return ubars + 8u * (2 * a.stages + i); };
      auto utempty = [&](int i) { return ubars + 8u * (2 * a.stages + 2 + i); };
      // D = F32, A = B = BF16, K-major, N = 256, M = 256 (cta_group::2)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kCg2BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int s = 0, has = 0;
      uint32_t ph = 0, hph = 0;
      for (int it = 0; it < my_units; ++it) {
        const int acc = it & 1;
        mbar_wait(utempty(acc), (((uint32_t)(it >> 1)) & 1u) ^ 1u);   // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = utmem + (uint32_t)(acc * kCg2BN);
        if (halo) {
          for (int cc = 0; cc < a.halo_cc; ++cc) {
            mbar_wait(ubars + 448u + 8u * has, hph);
            const uint32_t xa = ubase + (uint32_t)has * HALO_SLOT;
            for (int j = 0; j < a.halo_taps; ++j) {
              mbar_wait(ufull(s), ph);
              tc_fence_after();
              const uint64_t da = make_desc<BK>(xa + (uint32_t)(j * a.halo_dil) * (uint32_t)(BK * 2)), db = make_desc<BK>(uring + (uint32_t)s * STAGE);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                if (elect_one()) umma2_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (cc | j | k) != 0 ? 1u : 0u);
              if (elect_one()) umma2_commit_mc(uempty(s), (uint16_t)3);
              if (++s == a.stages) { s = 0; ph ^= 1u; }
            }
            if (elect_one()) umma2_commit_mc(ubars + 480u + 8u * has, (uint16_t)3);   // the halo slots of both CTAs are free once all taps have read them
            if (++has == a.a_stages) { has = 0; hph ^= 1u; }
          }
          if (elect_one()) umma2_commit_mc(utfull(acc), (uint16_t)3);
          continue;
        }
        for (int q = 0; q < a.n_chunks; ++q) {
          mbar_wait(ufull(s), ph);
          tc_fence_after();
          const uint32_t sa = uring + (uint32_t)s * STAGE;
          const uint64_t da = make_desc<BK>(sa), db = make_desc<BK>(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            if (elect_one()) umma2_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (q | k) != 0 ? 1u : 0u);
          if (elect_one()) umma2_commit_mc(uempty(s), (uint16_t)3);   // frees the stage in both CTAs when these MMAs retire
          if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) umma2_commit_mc(utfull(acc), (uint16_t)3);   // accumulators complete in both CTAs
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
    if constexpr (RES) {
      const int grp = (warp - 2) >> 2;                     // column slice of this group
      const int bar_id = 1 + grp;
      const uint32_t buf = rbuf(grp), gbias = buf + RBUF;
      const uint32_t rrow = buf + (uint32_t)row_in_tile * 128u;
      const uint32_t rx = (uint32_t)(row_in_tile & 7);
      for (int it = 0; it < my_units; ++it) {
        const int acc = it & 1;
        int b, m0, n0;
        unit_coords(it, b, m0, n0);
        const bool valid = m0 + row_in_tile < a.lout;
        const int nc = n0 + grp * CW;
        if (n0 != bias_n0) {
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          for (int i = et; i < CW; i += 128) {
            const float bv = (nc + i < a.cout) ? __ldg(a.bias + nc + i) : 0.f;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(gbias + 4u * i), "f"(bv) : "memory");
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          bias_n0 = n0;
        }
        mbar_wait(tfull_bar(acc), ((uint32_t)(it >> 1)) & 1u);
        tc_fence_after();
        mbar_wait(rfull_bar(grp), (uint32_t)it & 1u);
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kCg2BN + grp * CW);
        uint32_t vv[CW / 32][32];
#pragma unroll
        for (int h = 0; h < CW / 32; ++h) tmem_ld32_nowait(taddr + 32u * h, vv[h]);
#pragma unroll
        for (int h = 0; h < CW / 32; ++h) tmem_ld_wait32(vv[h]);
#pragma unroll
        for (int h = 0; h < CW / 32; ++h) {
          const int c0 = 32 * h;
          uint32_t(&v)[32] = vv[h];
          float bia[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bia[j]), "=f"(bia[j + 1]), "=f"(bia[j + 2]), "=f"(bia[j + 3])
                         : "r"(gbias + 4u * (uint32_t)(c0 + j)));
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
        if (lane == 0) mbar_arrive_cluster(acc ? tempty1 : tempty0);   // 8 NG arrivals: NG groups x 4 warps x 2 CTAs
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // the group's half tile is complete in the buffer
        if (et == 0) {
          for (int bx = 0; bx < CW / 64; ++bx) tma_store_3d(&tmO, buf + (uint32_t)bx * 16384u, nc + 64 * bx, m0, b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the store has read the buffer: the next residual may land
          mbar_arrive(rempty_bar(grp));
        }
      }
      if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this thread's tile stores are complete
    } else {
      // non-residual layers: a.n_epi (1 or 2) groups split the tile's columns (an epilogue warp alone on its scheduler is latency
      // bound; two groups halve the per-tile epilogue time, which matters where the K loop is short relative to the tile: N = 128)
      const int grp = (warp - 2) >> 2;
      if (grp < a.n_epi) {
        const int CWn = kCg2BN / a.n_epi;                    // columns of this group
        const uint32_t RSn = (uint32_t)CWn * 2u + 16u;
        const uint32_t gstg = bars + 512u + (uint32_t)grp * ((uint32_t)BM * RSn + (uint32_t)CWn * 4u);
        const uint32_t gbias = gstg + (uint32_t)BM * RSn;
        const int vpr = CWn / 8;
        for (int it = 0; it < my_units; ++it) {
          const int acc = it & 1;
          int b, m0, n0;
          unit_coords(it, b, m0, n0);
          const int nc = n0 + grp * CWn;
          const bool valid = m0 + row_in_tile < a.lout;
          const long long tile_off = ((long long)b * a.out_row_stride + m0) * a.cout + nc;
          if (n0 != bias_n0) {
            group_sync(grp);
            for (int i = et; i < CWn; i += 128) {
              const float bv = (nc + i < a.cout) ? __ldg(a.bias + nc + i) : 0.f;
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(gbias + 4u * i), "f"(bv) : "memory");
            }
            group_sync(grp);
            bias_n0 = n0;
          }
          mbar_wait(tfull_bar(acc), ((uint32_t)(it >> 1)) & 1u);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kCg2BN + grp * CWn);
          for (int c0 = 0; c0 < CWn; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            float bia[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bia[j]), "=f"(bia[j + 1]), "=f"(bia[j + 2]), "=f"(bia[j + 3])
                           : "r"(gbias + 4u * (uint32_t)(c0 + j)));
            const uint32_t srow = gstg + (uint32_t)row_in_tile * RSn + (uint32_t)c0 * 2u;
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
          if (lane == 0) mbar_arrive_cluster(acc ? tempty1 : tempty0);   // 8 n_epi arrivals: n_epi groups x 4 warps x 2 CTAs
          group_sync(grp);      // the group's column slice is complete in its staging area
          for (int idx = et; idx < BM * vpr; idx += 128) {
            const int rr = idx / vpr, cv = idx % vpr;
            if (m0 + rr < a.lpad) {
              uint4 v;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                           : "r"(gstg + (uint32_t)rr * RSn + (uint32_t)cv * 16u));
              reinterpret_cast<uint4*>(a.out + tile_off + (long long)rr * a.cout)[cv] = v;
            }
          }
          group_sync(grp);      // staging free for the next tile
        }
      }
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
