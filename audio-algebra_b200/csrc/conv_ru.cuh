// Fused ResidualUnit of the conv encoder on tcgen05 (included by conv_tc.cu, inside its anonymous namespace):
//
//     y = ELU( x + Conv1d_k1( ELU( Conv1d_k7,dilation d (x) + b7 ) ) + b1 )          (C -> C -> C channels, C = 32 / 64 / 128)
//
// i.e. the two layers ROLE_RES_FIRST / ROLE_RES_SECOND of the layer table (ResidualUnit) in ONE
// kernel: the intermediate never leaves the SM, x is read from HBM once (it also serves as the residual) and the
// result is written once -- 2 activation passes over HBM instead of 5.
//
// Shared-memory operand layout: K-major, NO swizzle, "panel" form  [C/8 panels][rows][8 channels = 16 B]
// (UMMA canonical layout ((8,m),(T,2)):((1T,SBO),(1,LBO)) with SBO = 128 B, LBO = panel stride).  Rows are uniformly
// 16 B apart inside a panel, so the A operand of filter tap j is the SAME x tile with the descriptor start address
// advanced by j*d rows: every x tile (128 + 6 d rows) is loaded once for all 7 taps.  The k7 and k1 weights stay
// resident in shared memory for the whole (persistent) kernel (C = 32 / 64); at C = 128 the 229 KB of k7 weights stream
// through a two-stage ring filled by a third producer warp (one [C][64 K] chunk per stage).
//
// Warp roles (64 + 128 NG threads): warp 0 = x-tile producer (16-byte cp.async straight into the panel layout, zero fill =
// conv padding), warp 1 = tcgen05.mma issuer (k7 of tile i+1 is issued before k1 of tile i: the tensor pipe works
// while the epilogue turns earlier tiles' first accumulators into k1 operands), warps 2.. = NG epilogue groups that
// take tiles round-robin: TMEM -> +b7 -> ELU -> bf16 panel tile (the k1 A operand) ; TMEM -> +b1 + x (from the resident x
// tile) -> ELU -> bf16 -> staged -> coalesced 16-byte stores.
#pragma once


struct RuArgs {
  const __nv_bfloat16* x;    // [B][rows_alloc][C] channels-last
  __nv_bfloat16* out;        // same geometry
  const __nv_bfloat16* w7;   // [C][7 C] tap-major K (pack_weights_kernel)
  const __nv_bfloat16* w1;   // [C][C]
  const float* b7;
  const float* b1;
  int lout, lpad, dil, m_tiles;
  long long rows_alloc, tiles;
};

template <int C>
struct RuCfg {
  static constexpr int P = C / 8;                        // 16-byte channel panels per row
  static constexpr bool STREAM = (C == 128);             // k7 weights (229 KB at C = 128) streamed through a ring instead of resident
  static constexpr int RA = (C == 32) ? 186 : 185;       // rows allocated per x tile (>= 128 + 6*9); = 2 mod 8 / odd:
                                                         // the producer's 16-byte pieces then spread over all banks
  static constexpr int XS = RA * 16;                     // x panel stride (bytes)
  static constexpr int XBYTES = P * XS;
  static constexpr int WS = C * 16 + (STREAM ? 16 : 0);  // weight panel stride (+16: the streaming producer writes 8 panels per row at once)
  static constexpr int NW = 2;                           // ring stages of streamed k7 weights
  static constexpr int WCH = 8 * WS;                     // one streamed chunk: [C out-channels][64 K] = 8 panels
  static constexpr int W7BYTES = STREAM ? NW * WCH : 7 * P * WS;
  static constexpr int W1BYTES = P * WS;
  static constexpr int TS = 2048 + (C == 32 ? 32 : 16);  // intermediate / staging panel stride (copy-out conflict free)
  static constexpr int TBYTES = P * TS;
  static constexpr int OFF_BIAS = 256;
  static constexpr int OFF_W7 = 2048;
  static constexpr int OFF_W1 = OFF_W7 + W7BYTES;
  static constexpr int OFF_X = OFF_W1 + W1BYTES;
#ifndef AA_RU_NG32
#define AA_RU_NG32 2
#endif
#ifndef AA_RU_NG64
#define AA_RU_NG64 2
#endif
#ifndef AA_RU_NX32
#define AA_RU_NX32 6
#endif
#ifndef AA_RU_NX64
#define AA_RU_NX64 4
#endif
  static constexpr int NG = (C == 32) ? AA_RU_NG32 : (C == 64) ? AA_RU_NG64 : 2;   // epilogue groups (4 warps each) = tiles in flight per CTA; C = 32 also runs two CTAs per SM
  static constexpr int EPI0 = STREAM ? 3 : 2;            // first epilogue warp (warp 0: x producer, 1: MMA issuer, 2: weight producer if STREAM)
  static constexpr int THREADS = 32 * EPI0 + 128 * NG;
  static constexpr int NX = (C == 32) ? AA_RU_NX32 : (C == 64) ? AA_RU_NX64 : 2;   // x tiles in flight (a stage is held until its tile's phase 2 has read the residual)
  // the 1x1 conv of tile i is issued after the k7 conv of tile i + LAG; LAG <= NX - 1, otherwise the x stage that tile
  // i + LAG needs would only be released by a phase 2 that waits for that very 1x1 conv
  static constexpr int LAG = (NG - 1 < NX - 1) ? NG - 1 : NX - 1;
  static constexpr int OFF_T = OFF_X + NX * XBYTES;
  static constexpr int SMEM = OFF_T + NG * TBYTES + 128; // + alignment slack
  static constexpr int TMEM_USED = 2 * C * NG;           // {acc1, acc2} per epilogue group
  static constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;   // power of two
};

// K-major, no-swizzle shared-memory matrix descriptor: 8-row groups 128 B apart (SBO), K core matrices `lbo` bytes apart
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell); layout type 0 = SWIZZLE_NONE
  return d;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// ELU(alpha = 1) of a pair: max(t, 2^(-|t| log2 e) - 1).  For t <= 0 this is exp(t) - 1 >= t; for t > 0 the second term lies in
// (-1, 0) < t.  The |t| and the sign ride on the FMUL operand modifiers: no clamp instructions.
__device__ __forceinline__ float2 elu2(float2 t) {
  float2 e = make_float2(__fmul_rn(-fabsf(t.x), 1.4426950408889634f), __fmul_rn(-fabsf(t.y), 1.4426950408889634f));
  asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(e.x));
  asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(e.y));
  e = __fadd2_rn(e, make_float2(-1.f, -1.f));
  return make_float2(fmaxf(t.x, e.x), fmaxf(t.y, e.y));
}
__device__ __forceinline__ uint32_t pack_bf16(float2 v) {
  __nv_bfloat162 o = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&o);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// named barrier of one epilogue group (immediate ids: a register id would reserve all 16 hardware barriers)
__device__ __forceinline__ void group_sync(int g) {
  if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else if (g == 1) asm volatile("bar.sync 2, 128;" ::: "memory");
  else if (g == 2) asm volatile("bar.sync 3, 128;" ::: "memory");
  else asm volatile("bar.sync 4, 128;" ::: "memory");
}

template <int C>
__global__ void __launch_bounds__(RuCfg<C>::THREADS, (C == 32) ? 2 : 1) ru_fused_kernel(const __grid_constant__ CUtensorMap tmX, const RuArgs a) {
  using Cfg = RuCfg<C>;
  constexpr int P = Cfg::P, NX = Cfg::NX, NG = Cfg::NG, LAG = Cfg::LAG, kRuThreads = Cfg::THREADS;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  // barriers: xfull[NX] xempty[NX] acc1full[NG] acc1empty[NG] tfull[NG] acc2full[NG] acc2empty[NG], tmem slot
  auto xfull = [&](int s) { return base + 8u * s; };
  auto xempty = [&](int s) { return base + 8u * (NX + s); };
  auto acc1full = [&](int g) { return base + 8u * (2 * NX + g); };
  auto acc1empty = [&](int g) { return base + 8u * (2 * NX + NG + g); };
  auto tfull = [&](int g) { return base + 8u * (2 * NX + 2 * NG + g); };
  auto acc2full = [&](int g) { return base + 8u * (2 * NX + 3 * NG + g); };
  auto acc2empty = [&](int g) { return base + 8u * (2 * NX + 4 * NG + g); };
  auto wfull = [&](int s) { return base + 8u * (2 * NX + 5 * NG + s); };
  auto wempty = [&](int s) { return base + 8u * (2 * NX + 5 * NG + Cfg::NW + s); };
  const uint32_t tmem_slot = base + 8u * (2 * NX + 5 * NG + 2 * Cfg::NW);
  const uint32_t sbias = base + Cfg::OFF_BIAS;   // b7[C], b1[C]
  const uint32_t sW7 = base + Cfg::OFF_W7, sW1 = base + Cfg::OFF_W1, sX = base + Cfg::OFF_X, sT = base + Cfg::OFF_T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NX; ++s) { mbar_init(xfull(s), 1); mbar_init(xempty(s), 4); }
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    for (int g = 0; g < NG; ++g) {
      mbar_init(acc1full(g), 1); mbar_init(acc1empty(g), 4); mbar_init(tfull(g), 4);
      mbar_init(acc2full(g), 1); mbar_init(acc2empty(g), 4);
    }
    for (int s2 = 0; s2 < Cfg::NW; ++s2) { mbar_init(wfull(s2), 32); mbar_init(wempty(s2), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident weights -> panel layout [K/8][cout][16 B]; biases
  for (int idx = threadIdx.x; idx < (Cfg::STREAM ? 0 : C * 7 * P); idx += kRuThreads) {
    const int co = idx / (7 * P), pn = idx - co * (7 * P);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.w7 + (size_t)co * 7 * C) + pn);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sW7 + (uint32_t)pn * Cfg::WS + (uint32_t)co * 16u), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
  }
  for (int idx = threadIdx.x; idx < C * P; idx += kRuThreads) {
    const int co = idx / P, pn = idx - co * P;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.w1 + (size_t)co * C) + pn);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sW1 + (uint32_t)pn * Cfg::WS + (uint32_t)co * 16u), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
  }
  for (int i = threadIdx.x; i < 2 * C; i += kRuThreads) {
    const float bv = (i < C) ? __ldg(a.b7 + i) : __ldg(a.b1 + i - C);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + 4u * i), "f"(bv) : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // weights are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int d = a.dil;
  const int R = BM + 6 * d;   // rows of an x tile
  const uint32_t xs = (uint32_t)R * 16u;   // panel stride of an x tile in shared memory = what one TMA box lays down: [C/8 panels][R rows][16 B]

  if (warp == 0) {
    // ===================== producer: x tiles (rows m0 - 3d .. m0 + 127 + 3d) =====================
    // ONE TMA box per tile: the tensor map views x [B][rows][C] as (8 channels = 16 B, row, panel, batch) with the panel as the
    // outer box dimension, so the box lands directly in the no-swizzle panel layout [C/8][R][16 B]; rows outside [0, lout) are
    // zero-filled = the conv padding.  (The first version copied 16-byte pieces with cp.async: ncu showed 3 of every 4 LDGSTS
    // serialised into 32 shared-memory wavefronts -- 22 M of the kernel's 52 M wavefronts, the LSU data pipe at 61 %.)
    if (lane == 0) {
      int it = 0;
      for (long long t = blockIdx.x; t < a.tiles; t += gridDim.x, ++it) {
        const int s = it % NX;
        if (it >= NX) mbar_wait(xempty(s), (uint32_t)((it / NX) - 1) & 1u);
        const int b = (int)(t / a.m_tiles);
        const int m0 = (int)(t - (long long)b * a.m_tiles) * BM;
        mbar_expect_tx(xfull(s), xs * (uint32_t)P);
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                         sX + (uint32_t)s * Cfg::XBYTES),
                     "l"(&tmX), "r"(xfull(s)), "r"(0), "r"(m0 - 3 * d), "r"(0), "r"(b)
                     : "memory");
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      auto issue_k1 = [&](int j) {
        const int g = j % NG;
        const uint32_t n = (uint32_t)(j / NG);
        mbar_wait(tfull(g), n & 1u);
        mbar_wait(acc2empty(g), (n & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(g * 2 * C + C);
        const uint32_t ta = sT + (uint32_t)g * Cfg::TBYTES;
#pragma unroll
        for (int kk = 0; kk < C / 16; ++kk)
          umma_bf16(tmem_d, make_desc_ns(ta + (uint32_t)(2 * kk) * Cfg::TS, Cfg::TS), make_desc_ns(sW1 + (uint32_t)(2 * kk) * Cfg::WS, Cfg::WS),
                    idesc, kk != 0 ? 1u : 0u);
        umma_commit(acc2full(g));
      };
      int it = 0;
      uint32_t wcount = 0;   // streamed weight chunks consumed so far
      for (long long t = blockIdx.x; t < a.tiles; t += gridDim.x, ++it) {
        const int g = it % NG, s = it % NX;
        const uint32_t n = (uint32_t)(it / NG);
        mbar_wait(xfull(s), (uint32_t)(it / NX) & 1u);
        mbar_wait(acc1empty(g), (n & 1u) ^ 1u);
        // (the x tile was written by TMA, i.e. by the async proxy the tensor core reads through: no proxy fence needed here)
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(g * 2 * C);
        const uint32_t xa = sX + (uint32_t)s * Cfg::XBYTES;
        if constexpr (Cfg::STREAM) {
          // k7 weights arrive chunk by chunk ([C][64 K] = tap j, K half h) through the ring filled by the weight producer warp
#pragma unroll 1
          for (int q = 0; q < 7 * (C / 64); ++q, ++wcount) {
            const int j = q / (C / 64), h = q - j * (C / 64), ws = (int)(wcount % Cfg::NW);
            mbar_wait(wfull(ws), (wcount / Cfg::NW) & 1u);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_after();
            const uint32_t wb = sW7 + (uint32_t)ws * Cfg::WCH;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(tmem_d, make_desc_ns(xa + (uint32_t)(8 * h + 2 * kk) * xs + (uint32_t)(j * d) * 16u, xs),
                        make_desc_ns(wb + (uint32_t)(2 * kk) * Cfg::WS, Cfg::WS), idesc, (q | kk) != 0 ? 1u : 0u);
            umma_commit(wempty(ws));           // frees the ring stage when these MMAs retire
          }
        } else {
#pragma unroll 1
          for (int j = 0; j < 7; ++j) {
#pragma unroll
            for (int kk = 0; kk < C / 16; ++kk)
              umma_bf16(tmem_d, make_desc_ns(xa + (uint32_t)(2 * kk) * xs + (uint32_t)(j * d) * 16u, xs),
                        make_desc_ns(sW7 + (uint32_t)(j * P + 2 * kk) * Cfg::WS, Cfg::WS), idesc, (j | kk) != 0 ? 1u : 0u);
          }
        }
        umma_commit(acc1full(g));
        if (it >= LAG) issue_k1(it - LAG);   // the oldest tile still waiting for its 1x1 conv: its operand is (nearly) ready by now
      }
      for (int j = (it >= LAG ? it - LAG : 0); j < it; ++j) issue_k1(j);
    }
  } else if (Cfg::STREAM && warp == 2) {
    // ===================== weight producer (C = 128): k7 weights, one [C][64 K] chunk per ring stage =====================
    // lanes = 4 out-channels x 8 panels: 128 contiguous bytes of a weight row per 8 lanes; panel stride WS = 2 KB + 16 B
    const int pn = lane & 7, co0 = lane >> 3;
    uint32_t wcount = 0;
    for (long long t = blockIdx.x; t < a.tiles; t += gridDim.x) {
#pragma unroll 1
      for (int q = 0; q < 7 * (C / 64); ++q, ++wcount) {
        const int ws = (int)(wcount % Cfg::NW);
        if (wcount >= (uint32_t)Cfg::NW) mbar_wait(wempty(ws), ((wcount / Cfg::NW) - 1u) & 1u);
        const __nv_bfloat16* src = a.w7 + (size_t)q * 64 + pn * 8;              // K offset of chunk q = (tap j) * C + h * 64 = 64 q
        const uint32_t dst = sW7 + (uint32_t)ws * Cfg::WCH + (uint32_t)pn * Cfg::WS;
#pragma unroll 4
        for (int co = co0; co < C; co += 4) cp_async16(dst + (uint32_t)co * 16u, src + (size_t)co * 7 * C, 16u);
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(wfull(ws)) : "memory");
      }
    }
  } else {
    // ===================== epilogue: group g = tiles with it % NG == g; thread = one row of the tile =====================
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    const int g = (warp - Cfg::EPI0) >> 2;
    const int r = quarter * 32 + lane;          // row in the tile
    const int et = (threadIdx.x - 32 * Cfg::EPI0) & 127;
    const uint32_t tg = sT + (uint32_t)g * Cfg::TBYTES;
    int it = 0;
    for (long long t = blockIdx.x; t < a.tiles; t += gridDim.x, ++it) {
      if (it % NG != g) continue;
      const uint32_t n = (uint32_t)(it / NG);
      const int s = it % NX;
      const int b = (int)(t / a.m_tiles);
      const int m0 = (int)(t - (long long)b * a.m_tiles) * BM;
      const bool valid = m0 + r < a.lout;
      const uint32_t tq = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * 2 * C);
      // ---- phase 1: h = ELU(acc1 + b7) -> bf16 panel tile (A operand of the 1x1 conv) ----
      mbar_wait(acc1full(g), n & 1u);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tq + (uint32_t)c0, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 ba, bb;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ba.x), "=f"(ba.y), "=f"(ba.z), "=f"(ba.w)
                       : "r"(sbias + 4u * (uint32_t)(c0 + 8 * q)));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w)
                       : "r"(sbias + 4u * (uint32_t)(c0 + 8 * q + 4)));
          const float2 h0 = elu2(__fadd2_rn(make_float2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1])), make_float2(ba.x, ba.y)));
          const float2 h1 = elu2(__fadd2_rn(make_float2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3])), make_float2(ba.z, ba.w)));
          const float2 h2 = elu2(__fadd2_rn(make_float2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5])), make_float2(bb.x, bb.y)));
          const float2 h3 = elu2(__fadd2_rn(make_float2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])), make_float2(bb.z, bb.w)));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tg + (uint32_t)(c0 / 8 + q) * Cfg::TS + (uint32_t)r * 16u),
                       "r"(pack_bf16(h0)), "r"(pack_bf16(h1)), "r"(pack_bf16(h2)), "r"(pack_bf16(h3))
                       : "memory");
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the panel tile is read by the tensor core
      __syncwarp();
      if (lane == 0) { mbar_arrive(acc1empty(g)); mbar_arrive(tfull(g)); }
      // ---- phase 2: y = ELU(acc2 + b1 + x) -> staged in the (now consumed) panel tile -> coalesced store ----
      mbar_wait(acc2full(g), n & 1u);
      tc_fence_after();
      const uint32_t xres = sX + (uint32_t)s * Cfg::XBYTES + (uint32_t)(3 * d + r) * 16u;
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tq + (uint32_t)(C + c0), v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 ba, bb;
          uint4 xr;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ba.x), "=f"(ba.y), "=f"(ba.z), "=f"(ba.w)
                       : "r"(sbias + 4u * (uint32_t)(C + c0 + 8 * q)));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w)
                       : "r"(sbias + 4u * (uint32_t)(C + c0 + 8 * q + 4)));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xr.x), "=r"(xr.y), "=r"(xr.z), "=r"(xr.w)
                       : "r"(xres + (uint32_t)(c0 / 8 + q) * xs));
          float2 y0 = __fadd2_rn(make_float2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1])), make_float2(ba.x, ba.y));
          float2 y1 = __fadd2_rn(make_float2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3])), make_float2(ba.z, ba.w));
          float2 y2 = __fadd2_rn(make_float2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5])), make_float2(bb.x, bb.y));
          float2 y3 = __fadd2_rn(make_float2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])), make_float2(bb.z, bb.w));
          y0 = elu2(__fadd2_rn(y0, unpack_bf16(xr.x)));
          y1 = elu2(__fadd2_rn(y1, unpack_bf16(xr.y)));
          y2 = elu2(__fadd2_rn(y2, unpack_bf16(xr.z)));
          y3 = elu2(__fadd2_rn(y3, unpack_bf16(xr.w)));
          uint4 o = make_uint4(pack_bf16(y0), pack_bf16(y1), pack_bf16(y2), pack_bf16(y3));
          if (!valid) o = make_uint4(0u, 0u, 0u, 0u);   // rows in [lout, lpad) carry zeros (tail of a strided view)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tg + (uint32_t)(c0 / 8 + q) * Cfg::TS + (uint32_t)r * 16u), "r"(o.x),
                       "r"(o.y), "r"(o.z), "r"(o.w)
                       : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(acc2empty(g)); mbar_arrive(xempty(s)); }
      group_sync(g);                                                    // tile complete in staging
      {
        uint4* o = reinterpret_cast<uint4*>(a.out + ((size_t)b * (size_t)a.rows_alloc + m0) * C);
#pragma unroll
        for (int i = 0; i < P; ++i) {
          const int idx = et + 128 * i;
          const int rr = idx / P, pn = idx % P;
          if (m0 + rr < a.lpad) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(tg + (uint32_t)pn * Cfg::TS + (uint32_t)rr * 16u));
            o[idx] = v;
          }
        }
      }
      group_sync(g);                                                    // staging free: the next phase 1 may overwrite it
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}
