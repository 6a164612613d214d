// stft2048_v2_kernel -- n_fft = 2048 STFT / power / mel front-end, round-2 design (included by stft.cu).
//
// Same arithmetic as stft2048_kernel (one warp = one frame of a ROW PAIR in packed fp32x2 registers, 1024-point
// complex FFT = 32 x 32 with one transpose through a warp-private buffer, even/odd split against the partner lane),
// but the 12 warps of the (single) CTA of an SM are DECOUPLED: there is no bar.sync in the tile loop.
//   * tile = 12 consecutive frames of a row pair; samples arrive by TMA bulk copies into a ring of NBUF (2 when it
//     fits, else 1) tile buffers; a warp that has its samples in registers bumps the buffer's counter and the LAST
//     consumer issues the copy of the tile NBUF iterations ahead (nobody waits for the issue, and with two buffers
//     the copy has a whole tile period to land);
//     frames that touch the chunk edges (reflect padding, zero_pad_po2 tail) or rows that cannot use TMA are
//     gathered from global memory with the reference's index math into the warp's own buffer -- only that warp
//     takes the slow path;
//   * 8456 bytes per warp hold the [32][33] float2 transposes, the float4 partner exchange and finally the frame's
//     1025-bin power line; to make room for the second sample buffer only HALF of the W_1024^(lane k1) table sits
//     in shared memory: W^(lane (32 - j)) = W_32^lane conj(W^(lane j)) costs four scalar FP ops instead of a load;
//   * complex / power outputs (frequency-minor layout = torch.stft's own memory layout) are stored straight from
//     registers: lanes hold 32 consecutive bins per slot, so every store instruction writes a 128/256-byte run;
//     these two modes have no inter-warp dependency besides the sample ring;
//   * mel: the filterbank is applied as a BANDED matrix (<= 2 adjacent taps per bin, which is what HTK/Slaney
//     triangles are): lanes = (frame of the tile, row); every lane walks the bins of a run of the band once, two FMAs
//     per power value (tap of filter m and of filter m + 1), weights come from constant memory (kernel parameter
//     space, warp-uniform LDC) -- one conflict-free shared-memory wavefront per 24 power values where the v1
//     epilogue needed 12 per 96 -- and the runs are dealt to the warps by bin count.  The only rendezvous of a tile
//     is "all P lines written" (mbarrier); "all warps done reading my P line" is waited for one FFT later.
#pragma once

constexpr int kV2W = 12;                 // warps per CTA = frames per tile
constexpr int kV2Xb = 8456;              // per-warp buffer: 2114 words = 2 (mod 32): lanes (frame, row) read conflict free
constexpr int kV2Tables = 17 * 256 + 512 + 256 + 256;   // tw1 rows 0..16, Hann phases, W_2048^lane, W_32^lane

struct Stft2Args {
  const float* wav;          // [rows][n_in]
  float* out;
  int rows, n_in, n_pad, n_frames, hop, center_off, tiles_per_pair, n_tiles;
  int n_freq, n_mels, wav_aligned16, nbuf, diag;   // diag (timing experiments only, wrong results): 1 skip the mel walk, 2 skip the P-ready wait
  const float2* tw1;         // [32][32] W_1024^(k1 n2)
  const float* lane_consts;  // float4[32] Hann phases + float2[32] W_2048^lane
  const int4* mel_groups;    // [4] {first record of warps 0..3}, {4..7}, {8..11}, pad; then 48-byte records (3 x int4):
                             //   {byte offset of the first bin in a P line, flags (1: last group of a run, 2: last record of the warp),
                             //    filter to store at the end of the run or -1, 0}, then 4 x (tap of filter m_lo, tap of filter m_lo + 1),
                             //   taps beyond the end of the run are zero (x 0.25: P holds 4|X|^2)
  int mel_groups_len;        // number of int4 entries
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_acq_rel_shared(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// Branch-free variant of fetch_sample (reflect about 0 and n_pad - 1, zeros beyond n_in): the load is unconditional on a clamped
// address, so that the compiler can keep a whole batch of gathers in flight.
__device__ __forceinline__ float fetch_sample_nb(const float* __restrict__ p, int i, int n_in, int n_pad, bool center) {
  if (center) {
    i = abs(i);
    i = (i >= n_pad) ? 2 * (n_pad - 1) - i : i;
  }
  const bool ok = (unsigned)i < (unsigned)n_in;
  const float v = __ldg(p + (ok ? i : 0));
  return ok ? v : 0.f;
}

__device__ __forceinline__ TileGeom tile_geom2(const Stft2Args& a, int t, int span) {
  TileGeom g;
  const int pair = t / a.tiles_per_pair;
  g.f0 = (t - pair * a.tiles_per_pair) * kV2W;
  g.rowA = 2 * pair;
  g.hasB = g.rowA + 1 < a.rows;
  g.s0 = g.f0 * a.hop - a.center_off;
  g.lo = g.s0 < 0 ? -g.s0 : 0;
  g.hi = min(span, a.n_in - g.s0);
  g.tma = g.hasB && a.wav_aligned16 && (a.n_in & 3) == 0 && g.hi - g.lo >= 4 && g.lo < span;
  return g;
}

__device__ __forceinline__ void v2_issue(const Stft2Args& a, int t, int span, float* buf, uint64_t* bar) {
  const TileGeom g = tile_geom2(a, t, span);
  if (g.tma) issue_tma(buf, buf + span, a.wav + (size_t)g.rowA * (size_t)a.n_in + (g.s0 + g.lo), g.lo, g.hi - g.lo, a.n_in, bar);
}

template <int MODE>
__global__ void __launch_bounds__(kV2W * 32, 1) stft2048_v2_kernel(const __grid_constant__ Stft2Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hop = a.hop;
  const int span = (kV2W - 1) * hop + 2048;
  const int nbuf = a.nbuf;
  float* samples = reinterpret_cast<float*>(smem);                       // [nbuf][2 rows][span]
  unsigned char* stage = smem + (size_t)nbuf * (size_t)span * 8;
  unsigned char* xb_raw = stage + warp * kV2Xb;
  float2* XB = reinterpret_cast<float2*>(xb_raw);                         // transposes; later this frame's P line
  const uint32_t xb4 = (smem_u32(xb_raw) + 15u) & ~15u;                   // 16-byte aligned window for the float4 exchange
  unsigned char* tab = stage + kV2W * kV2Xb;
  float2* s_tw1 = reinterpret_cast<float2*>(tab);                          // [17][32] W_1024^(k1 lane), k1 = 0..16
  float4* s_lane = reinterpret_cast<float4*>(s_tw1 + 17 * 32);
  float2* s_tw2l = reinterpret_cast<float2*>(s_lane + 32);
  float2* s_w32 = s_tw2l + 32;                                              // [32] W_32^lane
  int4* s_groups = reinterpret_cast<int4*>(s_w32 + 32);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_groups + a.mel_groups_len);   // [0..1] samples full, [2] P ready, [3] mel done
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(mbar + 4);                      // [0..1] consumers of sample buffer b

  for (int i = tid; i < 17 * 32; i += kV2W * 32) s_tw1[i] = __ldg(a.tw1 + i);
  if (tid < 48) reinterpret_cast<float4*>(s_lane)[tid] = __ldg(reinterpret_cast<const float4*>(a.lane_consts) + tid);
  if (tid < 32) s_w32[tid] = __ldg(a.tw1 + 32 * 32 + tid);   // W_1024^(32 lane): row 32 of the extended table
  for (int i = tid; i < a.mel_groups_len; i += kV2W * 32) s_groups[i] = __ldg(a.mel_groups + i);
  if (tid == 0) {
    mbar_init(mbar + 0, 1);
    mbar_init(mbar + 1, 1);
    mbar_init(mbar + 2, kV2W);
    mbar_init(mbar + 3, kV2W);
    s_cnt[0] = s_cnt[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int n_tiles = a.n_tiles;
  if ((int)blockIdx.x >= n_tiles) return;
  if (tid == 0) {
    for (int b = 0; b < nbuf; ++b) {
      const int t = blockIdx.x + b * (int)gridDim.x;
      if (t < n_tiles) v2_issue(a, t, span, samples + (size_t)b * 2 * span, mbar + b);
    }
  }

  uint32_t ph_pr = 0, ph_md = 0, uses0 = 0, uses1 = 0;   // uses_b: TMA fills of sample buffer b waited for so far
  const bool center = a.center_off != 0;
  const int plane = (32 - lane) & 31;
  const int pshift = (lane == 0) ? 16 : 15;

#pragma unroll 1
  for (int t = blockIdx.x, it = 0; t < n_tiles; t += gridDim.x, ++it) {
    const TileGeom g = tile_geom2(a, t, span);
    const int f0 = g.f0;
    const long long rowA = g.rowA;
    const int frame = f0 + warp;
    const bool fvalid = frame < a.n_frames;
    const int b = (nbuf == 2) ? (it & 1) : 0;
    float* SA = samples + (size_t)b * 2 * span;
    float* SB = SA + span;
    if (g.tma) {   // tiles that cannot use TMA neither issue nor wait: the parity follows the number of fills, not the iteration
      if (b == 0) { mbar_wait(mbar + 0, uses0 & 1u); ++uses0; }
      else { mbar_wait(mbar + 1, uses1 & 1u); ++uses1; }
    }
    bool md_waited = false;
    float2 re[32], im[32];
    const int s_rel = warp * hop;
    const bool fast = fvalid && g.tma && s_rel >= g.lo && s_rel + 2048 <= g.hi;   // this frame lies inside the TMA-filled part
    // Done with sample buffer b: the last warp out refills it with the tile nbuf iterations ahead.  Warps that do not read the
    // buffer (no frame, or a frame gathered from global memory) say so BEFORE their slow work, so they never hold up the ring.
    auto release_samples = [&]() {
      __syncwarp();
      if (lane == 0) {
        const uint32_t old = atom_add_acq_rel_shared(s_cnt + b, 1u);
        if (old % kV2W == kV2W - 1) {
          const int tn = t + nbuf * (int)gridDim.x;
          if (tn < n_tiles) v2_issue(a, tn, span, SA, mbar + b);
        }
      }
      __syncwarp();
    };
    if (!fast) release_samples();
    if (fvalid) {
      const float4 ph = s_lane[lane];   // (cos phi0, sin phi0, cos phi1, sin phi1), phi = 2 pi (2 lane + {0,1}) / 2048
      if (fast) {
        const float2* FA = reinterpret_cast<const float2*>(SA + s_rel);
        const float2* FB = reinterpret_cast<const float2*>(SB + s_rel);
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
          const int n = 32 * n1 + lane;
          const float2 xa = FA[n], xb = FB[n];
          const float w0 = fmaf(0.5f * aa_consts::kSin32[n1], ph.y, fmaf(-0.5f * aa_consts::kCos32[n1], ph.x, 0.5f));
          const float w1 = fmaf(0.5f * aa_consts::kSin32[n1], ph.w, fmaf(-0.5f * aa_consts::kCos32[n1], ph.z, 0.5f));
          re[bitrev5(n1)] = make_float2(xa.x * w0, xb.x * w0);
          im[bitrev5(n1)] = make_float2(xa.y * w1, xb.y * w1);
        }
        release_samples();
      } else {
        // chunk edge / no TMA: gather the frame with the reference's reflect + zero-pad index math, half a frame (1024 samples of
        // both rows) at a time, through this warp's own buffer.  In mel mode that buffer still holds the previous tile's P line.
        if (MODE == MODE_MEL && it > 0) {
          mbar_wait(mbar + 3, ph_md);
          md_waited = true;
        }
        const float* __restrict__ pa = a.wav + (size_t)g.rowA * (size_t)a.n_in;
        const float* __restrict__ pb = g.hasB ? pa + a.n_in : pa;
        const float bmask = g.hasB ? 1.f : 0.f;
        const int sf = frame * hop - a.center_off;
        float* GA = reinterpret_cast<float*>(xb_raw);   // [1024] row A, then [1024] row B
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          __syncwarp();
#pragma unroll 8
          for (int j = lane; j < 1024; j += 32) {
            GA[j] = fetch_sample_nb(pa, sf + half * 1024 + j, a.n_in, a.n_pad, center);
            GA[1024 + j] = bmask * fetch_sample_nb(pb, sf + half * 1024 + j, a.n_in, a.n_pad, center);
          }
          __syncwarp();
          const float2* FA = reinterpret_cast<const float2*>(GA);
          const float2* FB = reinterpret_cast<const float2*>(GA + 1024);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int n1 = half * 16 + q;
            const float2 xa = FA[32 * q + lane], xb = FB[32 * q + lane];
            const float w0 = fmaf(0.5f * aa_consts::kSin32[n1], ph.y, fmaf(-0.5f * aa_consts::kCos32[n1], ph.x, 0.5f));
            const float w1 = fmaf(0.5f * aa_consts::kSin32[n1], ph.w, fmaf(-0.5f * aa_consts::kCos32[n1], ph.z, 0.5f));
            re[bitrev5(n1)] = make_float2(xa.x * w0, xb.x * w0);
            im[bitrev5(n1)] = make_float2(xa.y * w1, xb.y * w1);
          }
        }
        __syncwarp();
      }
    }

    if (fvalid) {
      fft32_dit(re, im);
      const float2 c32 = s_w32[lane];
#pragma unroll
      for (int j = 1; j <= 16; ++j) {
        const float2 tw = s_tw1[j * 32 + lane];
        {
          const float2 r = re[j], i = im[j];
          re[j] = pfma(i, -tw.y, pmuls(r, tw.x));
          im[j] = pfma(i, tw.x, pmuls(r, tw.y));
        }
        if (j < 16) {   // W^(lane (32 - j)) = W_32^lane * conj(W^(lane j))
          const float ux = fmaf(c32.y, tw.y, c32.x * tw.x), uy = fmaf(c32.y, tw.x, -c32.x * tw.y);
          const float2 r = re[32 - j], i = im[32 - j];
          re[32 - j] = pfma(i, -uy, pmuls(r, ux));
          im[32 - j] = pfma(i, ux, pmuls(r, uy));
        }
      }
    }
    if (MODE == MODE_MEL && it > 0) {   // the other warps have finished reading this warp's previous P line
      if (!md_waited) mbar_wait(mbar + 3, ph_md);
      ph_md ^= 1;
    }
    if (fvalid) {
#pragma unroll
      for (int k1 = 0; k1 < 32; ++k1) XB[k1 * 33 + lane] = re[k1];
      __syncwarp();
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) re[bitrev5(n2)] = XB[lane * 33 + n2];
      __syncwarp();
#pragma unroll
      for (int k1 = 0; k1 < 32; ++k1) XB[k1 * 33 + lane] = im[k1];
      __syncwarp();
#pragma unroll
      for (int n2 = 0; n2 < 32; ++n2) im[bitrev5(n2)] = XB[lane * 33 + n2];
      __syncwarp();
      fft32_dit(re, im);   // slot k2 of lane k1: Z[k1 + 32 k2]
      // publish the upper half (k2 >= 16) for the partner lane
#pragma unroll
      for (int k2 = 16; k2 < 32; ++k2)
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(xb4 + (uint32_t)((k2 - 16) * 32 + lane) * 16u), "f"(re[k2].x),
                     "f"(re[k2].y), "f"(im[k2].x), "f"(im[k2].y)
                     : "memory");
      __syncwarp();

      // Even/odd split for the pair (k, 1024-k), k = lane + 32 i (i < 16): own Z[k] in slot i, partner's Z[1024-k] in lane
      // (32-lane)&31 slot 31-i (lane 0: slot 32-i).  E2 = a + conj(b), O2 = (a - conj(b))/i, T = W_2048^k O2:
      // 2 X[k] = E2 + T, 2 X[1024-k] = conj(E2 - T).
      const float2 cl = s_tw2l[lane];
      const float2 z0r = re[0], z0i = im[0], z16r = re[16], z16i = im[16];
      if constexpr (MODE == MODE_MEL) {
        // P line (float2 = (rowA,rowB) per bin, 4|X|^2, bins 0..1024) written in place into this warp's own buffer:
        // 4|X[1024-k]|^2 right away (those addresses are already consumed), 4|X[k]|^2 after the last partner read.  Partner rows
        // sit at xb4 = buffer + {0, 8} bytes: row r = bytes [512 r, 512 r + 512) + shift.  After iteration i the unread rows are
        // < 15 - i plus lane 0's 16 bytes of row 15 - i, i.e. bytes < 7704 - 512 i (shift included); the upper-bin stores of
        // iteration i start at byte 7944 - 256 i and pk[j] (bytes [256 j, 256 j + 256)) goes out once 256 j >= 8192 - 512 i.
        float2* pl = XB;
        float2 pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          int ps = pshift - i;
          ps = ps > 15 ? 15 : ps;
          const float4 bq = lds_f4(xb4 + (uint32_t)(ps * 32 + plane) * 16u);
          __syncwarp();   // all lanes have read slot row `ps` before anyone overwrites it below
          const float2 br = make_float2(bq.x, bq.y), bi = make_float2(bq.z, bq.w);
          const float2 ar = re[i], ai = im[i];
          const float2 tw = make_float2(fmaf(cl.y, aa_consts::kSin64[i], cl.x * aa_consts::kCos64[i]),
                                        fmaf(cl.y, aa_consts::kCos64[i], -cl.x * aa_consts::kSin64[i]));
          const float2 e_r = padd(ar, br), e_i = psub(ai, bi);
          const float2 o_r = padd(ai, bi), o_i = psub(br, ar);
          const float2 xr = pfma(o_i, -tw.y, pfma(o_r, tw.x, e_r));
          const float2 xi = pfma(o_r, tw.y, pfma(o_i, tw.x, e_i));
          const float2 yr = pfma(o_i, tw.y, pfma(o_r, -tw.x, e_r));
          const float2 yi = pfma(o_r, -tw.y, pfma(o_i, -tw.x, e_i));
          pk[i] = pfma2(xi, xi, pmul(xr, xr));
          const float2 pq = pfma2(yi, yi, pmul(yr, yr));
          if (!(lane == 0 && i == 0)) pl[1024 - (lane + 32 * i)] = pq;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j <= i && j >= 32 - 2 * i && (j == i || j < 34 - 2 * i)) pl[lane + 32 * j] = pk[j];
        }
        __syncwarp();
        pl[lane] = pk[0];
        pl[lane + 32] = pk[1];
        __syncwarp();
        if (lane == 0) {
          const float2 dc = padd(z0r, z0i), ny = psub(z0r, z0i);
          pl[0] = pmuls(pmul(dc, dc), 4.f);
          pl[1024] = pmuls(pmul(ny, ny), 4.f);
          pl[512] = pmuls(pfma2(z16i, z16i, pmul(z16r, z16r)), 4.f);
        }
      } else {
        // complex / power, frequency-minor output [row][frame][1025]: stored straight from registers
        const bool hasB = rowA + 1 < a.rows;
        const long long e0 = (rowA * (long long)a.n_frames + frame) * a.n_freq;
        const long long e1 = e0 + (long long)a.n_frames * a.n_freq;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          int ps = pshift - i;
          ps = ps > 15 ? 15 : ps;
          const float4 bq = lds_f4(xb4 + (uint32_t)(ps * 32 + plane) * 16u);
          const float2 br = make_float2(bq.x, bq.y), bi = make_float2(bq.z, bq.w);
          const float2 ar = re[i], ai = im[i];
          const float2 tw = make_float2(fmaf(cl.y, aa_consts::kSin64[i], cl.x * aa_consts::kCos64[i]),
                                        fmaf(cl.y, aa_consts::kCos64[i], -cl.x * aa_consts::kSin64[i]));
          const float2 e_r = padd(ar, br), e_i = psub(ai, bi);
          const float2 o_r = padd(ai, bi), o_i = psub(br, ar);
          const float2 xr = pfma(o_i, -tw.y, pfma(o_r, tw.x, e_r));   // 2 Re X[k]
          const float2 xi = pfma(o_r, tw.y, pfma(o_i, tw.x, e_i));    // 2 Im X[k]
          const float2 yr = pfma(o_i, tw.y, pfma(o_r, -tw.x, e_r));   // 2 Re X[1024-k]
          const float2 yi = pfma(o_r, -tw.y, pfma(o_i, -tw.x, e_i));  // -2 Im X[1024-k]
          const int k = lane + 32 * i;
          const bool skip = (lane == 0 && i == 0);   // DC / Nyquist come from lane 0 below
          if constexpr (MODE == MODE_POWER) {
            const float2 pa_ = pmuls(pfma2(xi, xi, pmul(xr, xr)), 0.25f);
            const float2 pq = pmuls(pfma2(yi, yi, pmul(yr, yr)), 0.25f);
            if (!skip) {
              a.out[e0 + k] = pa_.x;
              a.out[e0 + 1024 - k] = pq.x;
              if (hasB) {
                a.out[e1 + k] = pa_.y;
                a.out[e1 + 1024 - k] = pq.y;
              }
            }
          } else {
            float2* o = reinterpret_cast<float2*>(a.out);
            if (!skip) {
              o[e0 + k] = make_float2(0.5f * xr.x, 0.5f * xi.x);
              o[e0 + 1024 - k] = make_float2(0.5f * yr.x, -0.5f * yi.x);
              if (hasB) {
                o[e1 + k] = make_float2(0.5f * xr.y, 0.5f * xi.y);
                o[e1 + 1024 - k] = make_float2(0.5f * yr.y, -0.5f * yi.y);
              }
            }
          }
        }
        if (lane == 0) {
          const float2 dc = padd(z0r, z0i), ny = psub(z0r, z0i);
          if constexpr (MODE == MODE_POWER) {
            const float2 p512 = pfma2(z16i, z16i, pmul(z16r, z16r));
            a.out[e0] = dc.x * dc.x; a.out[e0 + 1024] = ny.x * ny.x; a.out[e0 + 512] = p512.x;
            if (hasB) { a.out[e1] = dc.y * dc.y; a.out[e1 + 1024] = ny.y * ny.y; a.out[e1 + 512] = p512.y; }
          } else {
            float2* o = reinterpret_cast<float2*>(a.out);
            o[e0] = make_float2(dc.x, 0.f); o[e0 + 1024] = make_float2(ny.x, 0.f); o[e0 + 512] = make_float2(z16r.x, -z16i.x);
            if (hasB) { o[e1] = make_float2(dc.y, 0.f); o[e1 + 1024] = make_float2(ny.y, 0.f); o[e1 + 512] = make_float2(z16r.y, -z16i.y); }
          }
        }
        __syncwarp();   // partner rows consumed before the next tile reuses the buffer
      }
    }

    if constexpr (MODE == MODE_MEL) {
      __syncwarp();
      if (lane == 0) mbar_arrive(mbar + 2);
      if (!(a.diag & 2)) mbar_wait(mbar + 2, ph_pr);
      ph_pr ^= 1;
      // ---- banded mel: lanes = (frame, row); this warp walks its own list of group records over all 12 frames of the tile.
      // A record = up to 4 consecutive bins of one run of the band with their two taps; (L, H) accumulate as ONE packed FFMA2 per bin
      // (taps as the register pair, the power value as the broadcast scalar).  Records are fetched two ahead and power values one
      // ahead of their use, so that the walk is issue bound instead of a chain of shared-memory latencies.
      const int r = lane & 1;
      const int fl = lane >> 1;
      const bool lactive = fl < kV2W;
      const int f = lactive ? fl : 0;
      const uint32_t pbase = smem_u32(stage) + (uint32_t)f * (uint32_t)kV2Xb + (uint32_t)r * 4u;
      const bool can_store = lactive && (f0 + f < a.n_frames) && (rowA + r < a.rows);
      float* obase = a.out + ((size_t)(rowA + r) * (size_t)a.n_mels) * (size_t)a.n_frames + (size_t)(f0 + f);
      const int first_rec = reinterpret_cast<const int*>(s_groups)[warp];
      uint32_t rec = smem_u32(s_groups) + (uint32_t)first_rec * 16u;
      uint32_t pb_ = pbase;
      unsigned nfr4 = (unsigned)a.n_frames * 4u;
      unsigned char* ob = reinterpret_cast<unsigned char*>(obase);
      unsigned store_ok = can_store ? 1u : 0u;
      asm volatile("" : "+r"(rec), "+r"(pb_), "+r"(nfr4), "+l"(ob), "+r"(store_ok));   // keep the loop invariants in registers
      float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;        // (L, H) partial sums of the current run
      float carry = 0.f;
      // one step of the walk: consume record (H, WA, WB, P*) while fetching the record after the next one's header and the next
      // record's taps / power values; written as a macro so that two copies can ping-pong between two register sets (no moves)
#define AA_MEL_STEP(H, WA, WB, P0, P1, P2, P3, HN, HNN, WAN, WBN, Q0, Q1, Q2, Q3)                          \
      {                                                                                                     \
        rec += 48u;                                                                                         \
        WAN = lds_f4(rec + 16u); WBN = lds_f4(rec + 32u);                                                   \
        HNN = lds_i4(rec + 48u);                                                                            \
        const uint32_t qa = pb_ + (uint32_t)HN.x;                                                           \
        Q0 = lds_f32(qa); Q1 = lds_f32(qa + 8u); Q2 = lds_f32(qa + 16u); Q3 = lds_f32(qa + 24u);            \
        acc0 = pfma(make_float2(WA.x, WA.y), P0, acc0);                                                     \
        acc1 = pfma(make_float2(WA.z, WA.w), P1, acc1);                                                     \
        acc0 = pfma(make_float2(WB.x, WB.y), P2, acc0);                                                     \
        acc1 = pfma(make_float2(WB.z, WB.w), P3, acc1);                                                     \
        if (H.y & 1) {   /* end of a run: filter H.z = this run's L + the previous run's H */                \
          const float2 s_ = padd(acc0, acc1);                                                               \
          if (H.z >= 0 && store_ok) *reinterpret_cast<float*>(ob + (size_t)((unsigned)H.z * nfr4)) = s_.x + carry; \
          carry = s_.y;                                                                                     \
          acc0 = make_float2(0.f, 0.f);                                                                     \
          acc1 = acc0;                                                                                      \
        }                                                                                                   \
        if (H.y & 2) break;                                                                                 \
      }
      int4 hA = lds_i4(rec), hB = lds_i4(rec + 48u), hC;      // every list is followed by two dummy records
      float4 waA = lds_f4(rec + 16u), wbA = lds_f4(rec + 32u), waB, wbB;
      float pA0 = lds_f32(pb_ + (uint32_t)hA.x), pA1 = lds_f32(pb_ + (uint32_t)hA.x + 8u), pA2 = lds_f32(pb_ + (uint32_t)hA.x + 16u),
            pA3 = lds_f32(pb_ + (uint32_t)hA.x + 24u), pB0, pB1, pB2, pB3;
      if (a.diag & 1) hA.y = 2, hA.z = -1;
#pragma unroll 1
      while (true) {
        AA_MEL_STEP(hA, waA, wbA, pA0, pA1, pA2, pA3, hB, hC, waB, wbB, pB0, pB1, pB2, pB3)
        AA_MEL_STEP(hB, waB, wbB, pB0, pB1, pB2, pB3, hC, hA, waA, wbA, pA0, pA1, pA2, pA3)
        // now hC is the current record with its data in set A ... rotate names by one more step pair
        AA_MEL_STEP(hC, waA, wbA, pA0, pA1, pA2, pA3, hA, hB, waB, wbB, pB0, pB1, pB2, pB3)
        AA_MEL_STEP(hA, waB, wbB, pB0, pB1, pB2, pB3, hB, hC, waA, wbA, pA0, pA1, pA2, pA3)
        AA_MEL_STEP(hB, waA, wbA, pA0, pA1, pA2, pA3, hC, hA, waB, wbB, pB0, pB1, pB2, pB3)
        AA_MEL_STEP(hC, waB, wbB, pB0, pB1, pB2, pB3, hA, hB, waA, wbA, pA0, pA1, pA2, pA3)
      }
#undef AA_MEL_STEP
      __syncwarp();
      if (lane == 0) mbar_arrive(mbar + 3);
    }
  }
}
