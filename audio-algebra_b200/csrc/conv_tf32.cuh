// fp32-grade tensor-core path of the conv encoder ("3xTF32"): the same implicit-GEMM Conv1d as conv_tc.cu, with every
// fp32 operand split into two TF32-representable parts,  v = hi + lo,  hi = v rounded to TF32 (10 mantissa bits),
// lo = (v - hi) rounded to TF32, and three MMAs per K step accumulated in fp32 in
// TMEM:  a_lo * w_hi + a_hi * w_lo + a_hi * w_hi  (the dropped a_lo * w_lo term is ~2^-24 relative).
//
// Accumulator chains are kept SHORT.  tcgen05.mma adds into its fp32 accumulator with truncation (round toward zero), measured
// in profiles/ubench/tf32_accum_ubench.cu: a chain of S accumulate steps loses ~2^-24.2 * S of the sum, systematically (K = 4096
// in one chain: 4.3e-5 relative error, all of it bias; 38 such layers gave 4.4e-5 on the embeddings against 1.4e-6 for plain
// fp32).  So the MMA warp closes a chain every `chain` K chunks (2 chunks = 24 MMAs by default), hands that TMEM slot to the
// epilogue warps and continues in the next slot; the epilogue warps sum the partial accumulators in registers with ordinary
// round-to-nearest fp32 adds.  Each of the two epilogue groups owns 256 TMEM columns = 2 .. 8 accumulator slots, so the draining overlaps the MMAs.
//
// The residual of a ResidualUnit's 1x1 conv rides on the tensor core as well: x (hi and lo planes) times an identity block
// appended to the weight matrix -- two extra MMAs per K step of 8 instead of per-thread global loads in the epilogue.
//
// Activations are fp32 channels-last, stored as two planes [2][B][rows][C] (plane 0 = hi, plane 1 = lo) so that the
// consumer's TMA boxes land in shared memory ready for the MMA (there is no register pass in which to split them);
// the producing epilogue does the split.  A K chunk is 32 channels = 128 bytes = one SWIZZLE_128B row, i.e. byte-for-byte
// the smem tile and descriptor geometry of the bf16 kernel (kind::tf32 consumes 8 elements = 32 bytes per MMA).
// Included by conv_tc.cu (inside its anonymous namespace; uses its mbarrier / TMA / TMEM helpers).
#pragma once

constexpr int kTfMaxChunks = 144;      // 128 conv chunks (8 x 512 channels of a strided view) + 16 residual chunks
constexpr int kTfSlots = 16;           // TMEM accumulator slots at most: 8 per epilogue group (256 columns per group / bn, capped)
constexpr int kTfThreads = 64 + 256;   // warp 0: TMA, warp 1: MMA, 2 epilogue groups of 4 warps; 320 threads -> 200 registers each

struct TfArgs {
  int n_chunks;                     // K chunks of 32 channels per tile, residual chunks included
  int n_main;                       // chunks [0, n_main) read the layer input (tmA); [n_main, n_chunks) the residual tensor (tmR)
  int chain;                        // chunks per accumulator chain
  short chunk_off[kTfMaxChunks];    // row offset of the A box for chunk q
  short chunk_col[kTfMaxChunks];    // column (channel) coordinate of the A box for chunk q
  int bn, n_tiles_n, m_tiles;
  long long tiles;
  int lout, lpad, cout;
  int stages;
  const float* bias;
  float* out_hi;                    // output planes [B][out_row_stride][cout], or NULL for the last layer
  float* out_lo;
  float* out_f32;                   // [B][cout][lout] fp32 channel-major (last layer), or NULL
  long long out_row_stride;
  int elu, tanh_out;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
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
// Same MMA with the two shared-memory descriptors given as their low words only: the high word of a SWIZZLE_128B K-major
// descriptor (SBO = 1024 B, version 1, layout 2) is the constant 0x40004040, and the low word is (address >> 4) | LBO field, so a
// K step of 32 bytes is "+ 2" on a 32-bit value -- half the uniform-register traffic per issued MMA of the 64-bit form.
__device__ __forceinline__ void umma_tf32_lo(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(0x40004040u)
      : "memory");
}
__device__ __forceinline__ void stg_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
// Round-to-nearest split: hi = rna_tf32(v), lo = rna_tf32(v - hi).  Both parts are exactly TF32-representable, so the tensor core
// reads them unchanged, and the split errors (|lo| <= 2^-12 |v|, signed) are unbiased.
__device__ __forceinline__ float tf32_hi(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ float elu_exact(float v) { return v > 0.f ? v : expm1f(v); }

__global__ void __launch_bounds__(kTfThreads, 1) conv_tf32_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const __grid_constant__ CUtensorMap tmR, const TfArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = BM * 128;                                 // one plane: 128 rows x 32 fp32
  const uint32_t B_BYTES = (uint32_t)a.bn * 128u;                        // one plane: bn rows x 32 fp32 (multiple of 1024)
  const uint32_t STAGE = 2u * A_BYTES + 2u * B_BYTES;                    // A hi | A lo | W hi | W lo
  const uint32_t bars = base + (uint32_t)a.stages * STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (a.stages + s); };
  auto pfull_bar = [&](int i) { return bars + 8u * (2 * a.stages + i); };                // chain complete in TMEM slot i
  auto pempty_bar = [&](int i) { return bars + 8u * (2 * a.stages + kTfSlots + i); };    // slot i drained
  const uint32_t tmem_slot = bars + 8u * (2 * a.stages + 2 * kTfSlots);
  const uint32_t sbias = bars + 512u;                                                    // cout floats
  const uint32_t nsl = min(8u, 256u / (uint32_t)a.bn);                                   // slots per epilogue group (power of two)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (2u * nsl * a.bn <= 256u) ? 256u : 512u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2 * (int)nsl; ++i) { mbar_init(pfull_bar(i), 1); mbar_init(pempty_bar(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < a.cout; i += kTfThreads)
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + 4u * i), "f"(__ldg(a.bias + i)) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int tiles_per_b = a.m_tiles * a.n_tiles_n;
  const int chains_per_tile = (a.n_chunks + a.chain - 1) / a.chain;

  if (warp == 0) {
    // ===================== TMA producer: both planes of the A box and of the weight box per K chunk =====================
    // (whole warp runs the loop so that addresses and coordinates stay in uniform registers; one elected lane issues)
    {
      const uint32_t ubase = __shfl_sync(0xffffffffu, base, 0), ubars = __shfl_sync(0xffffffffu, bars, 0);
      int s = 0;
      uint32_t ph = 0;
      for (long long t = blockIdx.x; t < a.tiles; t += gridDim.x) {
        const int b = (int)(t / tiles_per_b);
        const int r = (int)(t % tiles_per_b);
        const int nt = r % a.n_tiles_n, mt = r / a.n_tiles_n;
        const int m0 = mt * BM, n0 = nt * a.bn;
        for (int q = 0; q < a.n_chunks; ++q) {
          const uint32_t fb = ubars + 8u * s;
          mbar_wait(ubars + 8u * (a.stages + s), ph ^ 1u);
          const uint32_t sa = ubase + (uint32_t)s * STAGE;
          const bool main = q < a.n_main;
          // a residual chunk multiplies the identity block: of x's channels only [n0, n0 + bn) reach this N tile
          const int col = main ? a.chunk_col[q] : n0 + a.chunk_col[q], row = m0 + a.chunk_off[q];
          const int wcol = main ? q * 32 : a.n_main * 32 + n0 + a.chunk_col[q];
          if (elect_one()) {
            mbar_expect_tx(fb, STAGE);
            if (main) { tma_load_4d(sa, &tmA, fb, col, row, b, 0); tma_load_4d(sa + A_BYTES, &tmA, fb, col, row, b, 1); }
            else      { tma_load_4d(sa, &tmR, fb, col, row, b, 0); tma_load_4d(sa + A_BYTES, &tmR, fb, col, row, b, 1); }
            tma_load_3d(sa + 2u * A_BYTES, &tmB, fb, wcol, n0, 0);
            tma_load_3d(sa + 2u * A_BYTES + B_BYTES, &tmB, fb, wcol, n0, 1);
          }
          if (++s == a.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 3 x 4 MMAs (K = 8 each) per chunk, a new accumulator slot every `chain` chunks =====================
    {
      // warp-uniform copies the compiler can keep in uniform registers (a shuffle from lane 0 is uniform by construction)
      const uint32_t ubase = __shfl_sync(0xffffffffu, base, 0);
      const uint32_t utmem = __shfl_sync(0xffffffffu, tmem_base, 0), ubars = __shfl_sync(0xffffffffu, bars, 0);
      // instruction descriptor: D = F32, A = B = TF32 (format 2), both K-major, N = bn, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      // Each epilogue group owns half of the slots (nsl = 256 columns / bn of them, at most 8) and rotates through them chain by chain, so a slot's
      // barriers are only ever waited on by one group, phase after phase.  (With slots shared across the groups a group skips the
      // phases the other one handles, and a parity wait two phases ahead can pass on the stale phase.)
      uint32_t cg0 = 0u, cg1 = 0u;                          // chains issued so far for the tiles of each group
      int it = 0;
      for (long long t = blockIdx.x; t < a.tiles; t += gridDim.x, ++it) {
        const int g = it & 1;
        for (int q0 = 0; q0 < a.n_chunks; q0 += a.chain) {
          const uint32_t c = g ? cg1 : cg0;
          if (g) ++cg1; else ++cg0;
          const uint32_t slot = nsl * g + (c & (nsl - 1u));
          mbar_wait((ubars + 8u * (2 * a.stages + kTfSlots + slot)), ((c / nsl) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = utmem + slot * (uint32_t)a.bn;
          const int q1 = min(q0 + a.chain, a.n_chunks);
          for (int q = q0; q < q1; ++q) {
            mbar_wait((ubars + 8u * s), ph);
            tc_fence_after();
            const uint32_t sa = ubase + (uint32_t)s * STAGE;
            const uint32_t a_hi = ((sa & 0x3FFFFu) >> 4) | 0x10000u, a_lo = a_hi + (A_BYTES >> 4);
            const uint32_t w_hi = a_hi + ((2u * A_BYTES) >> 4), w_lo = w_hi + (B_BYTES >> 4);
            const bool lead = elect_one();
            if (lead) {
              // small terms first, the hi x hi term last
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_tf32_lo(tmem_d, a_lo + 2u * k, w_hi + 2u * k, idesc, (q != q0 || k != 0) ? 1u : 0u);
              if (q < a.n_main) {                             // (the identity block of a residual chunk has no lo part)
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32_lo(tmem_d, a_hi + 2u * k, w_lo + 2u * k, idesc, 1u);
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_tf32_lo(tmem_d, a_hi + 2u * k, w_hi + 2u * k, idesc, 1u);
            }
            if (elect_one()) umma_commit((ubars + 8u * (a.stages + s)));
            if (++s == a.stages) { s = 0; ph ^= 1u; }
          }
          if (elect_one()) umma_commit((ubars + 8u * (2 * a.stages + slot)));
        }
      }
    }
  } else {
    // ===================== epilogue: 2 groups of 4 warps, alternate tiles; thread = one output row (position) =====================
    // Sums the chains of its tile in registers (round-to-nearest), then bias -> ELU -> hi / lo split.  A row's 32-channel piece
    // is 128 contiguous bytes in each output plane: written with 256-bit stores, whole 32-byte sectors per lane.
    const int quarter = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    int it = 0;
    for (long long t = blockIdx.x; t < a.tiles; t += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int b = (int)(t / tiles_per_b);
      const int r = (int)(t % tiles_per_b);
      const int nt = r % a.n_tiles_n, mt = r / a.n_tiles_n;
      const int m0 = mt * BM, n0 = nt * a.bn;
      const int m = m0 + row_in_tile;
      const bool valid = m < a.lout;
      const long long row_off = ((long long)b * a.out_row_stride + m) * a.cout + n0;   // element offset of (this row, col 0 of the tile)
      float sum[128];
      uint32_t c = (uint32_t)(it >> 1) * (uint32_t)chains_per_tile;   // this group's chain counter (same sequence as the MMA warp's)
      for (int j = 0; j < chains_per_tile; ++j, ++c) {
        const uint32_t slot = nsl * grp + (c & (nsl - 1u));
        mbar_wait(pfull_bar(slot), (c / nsl) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * (uint32_t)a.bn;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          if (c0 < a.bn) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            if (j == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) sum[c0 + i] = __uint_as_float(v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) sum[c0 + i] += __uint_as_float(v[i]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pempty_bar(slot));
      }
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        if (c0 < a.bn) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float b0, b1, b2, b3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(sbias + 4u * (uint32_t)(n0 + c0 + i)));
            sum[c0 + i] += b0; sum[c0 + i + 1] += b1; sum[c0 + i + 2] += b2; sum[c0 + i + 3] += b3;
          }
          if (a.out_f32 != nullptr) {
            if (valid) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float y = sum[c0 + i];
                if (a.elu) y = elu_exact(y);
                if (a.tanh_out) y = tanhf(y);
                a.out_f32[((long long)b * a.cout + n0 + c0 + i) * a.lout + m] = y;
              }
            }
          } else if (m < a.lpad) {                            // rows in [lout, lpad) carry zeros
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float hi[8], lo[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float y = valid ? elu_exact(sum[c0 + 8 * g + i]) : 0.f;
                hi[i] = tf32_hi(y);
                lo[i] = tf32_hi(y - hi[i]);
              }
              stg_v8(a.out_hi + row_off + c0 + 8 * g, hi);
              stg_v8(a.out_lo + row_off + c0 + 8 * g, lo);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- layer 0 of the 3xTF32 path: fp32 stems [B][cin<=4][N] (fader-scaled sum) -> conv k<=7, 32 channels -> ELU -> hi / lo planes ----
// Memory bound (256 output bytes per position).  Thread = 8 channels x 4 positions; the 4 lanes of a position write one
// contiguous 128-byte row per plane.
struct L0fArgs {
  const float* x[4];
  float fader[4];
  int n_in, cin, k, pad, n, lpad;
  const float* w;      // [32][cin][k]
  const float* bias;
  float* out_hi;       // [B][row_stride][32]
  float* out_lo;
  long long row_stride;
};
constexpr int kL0fPos = 256;
__global__ void __launch_bounds__(256) conv_l0_tf32_kernel(const L0fArgs a) {
  __shared__ float Xs[4][kL0fPos + 8];
  __shared__ __align__(16) float Ws[4 * 7][32];     // [cin * k][cout]
  __shared__ __align__(16) float Bs[32];
  const int b = blockIdx.y, l0 = blockIdx.x * kL0fPos;
  for (int e = threadIdx.x; e < 32 * a.cin * a.k; e += 256) {
    const int co = e / (a.cin * a.k), ck = e % (a.cin * a.k);
    Ws[ck][co] = a.w[e];
  }
  if (threadIdx.x < 32) Bs[threadIdx.x] = a.bias[threadIdx.x];
  const int span = kL0fPos + a.k - 1;
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
  const int cg = threadIdx.x & 3, p0 = threadIdx.x >> 2;   // channels 8 cg .. 8 cg + 7, positions p0 + 64 i
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = Bs[8 * cg + j];
  for (int c = 0; c < a.cin; ++c)
    for (int kk = 0; kk < a.k; ++kk) {
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[c * a.k + kk][8 * cg]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[c * a.k + kk][8 * cg + 4]);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float xv = Xs[c][p0 + 64 * i + kk];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wv[j], xv, acc[i][j]);
      }
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int l = l0 + p0 + 64 * i;
    if (l < a.lpad) {
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = (l < a.n) ? elu_exact(acc[i][j]) : 0.f;
        hi[j] = tf32_hi(y);
        lo[j] = tf32_hi(y - hi[j]);
      }
      const long long off = ((long long)b * a.row_stride + l) * 32 + 8 * cg;
      stg_v8(a.out_hi + off, hi);
      stg_v8(a.out_lo + off, lo);
    }
  }
}

// W [cout][cin][k] fp32 -> planes [2][cout][K_total + id_cols] (hi, lo) in chunk order (same column mapping as
// pack_weights_kernel); the last id_cols (= cout for a ResidualUnit's second conv, else 0) columns are the identity block
// that adds the residual on the tensor core
__global__ void pack_weights_tf32_kernel(const float* __restrict__ w, float* __restrict__ w2, const PackArgs p, int id_cols) {
  const int kt_ext = p.k_total + id_cols;
  const long long n = (long long)p.cout * kt_ext;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / kt_ext), kc = (int)(i % kt_ext);
    if (kc >= p.k_total) {
      w2[i] = (kc - p.k_total == co) ? 1.f : 0.f;
      w2[n + i] = 0.f;
      continue;
    }
    int j = 0;
    while (j + 1 < p.n_taps && kc >= p.tap_kcol[j + 1]) ++j;
    const int c = p.tap_col0[j] + (kc - p.tap_kcol[j]);
    const int ci = c % p.cin, kt = c / p.cin + p.tap_kbase[j];
    const float v = (kt >= 0 && kt < p.k) ? w[((long long)co * p.cin + ci) * p.k + kt] : 0.f;
    const float hi = tf32_hi(v);
    w2[i] = hi;
    w2[n + i] = tf32_hi(v - hi);
  }
}
