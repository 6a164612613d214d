// stft2048_v3_kernel -- n_fft = 2048 STFT / power / mel front-end without ANY inter-warp synchronisation
// (included by stft.cu; same FFT arithmetic as stft2048_v2_kernel).
//
// Every warp owns whole frames of a row pair (two rows in the halves of packed fp32x2 registers) from the first load
// to the last store:
//   * samples come straight from global memory / L2 with coalesced 64-bit loads (a frame row is 8 KB of consecutive
//     floats; the 75 % overlap between consecutive frames is served by L1 / L2, the 12 warps of a CTA take 12
//     consecutive frames); the lines of the warp's NEXT frame are requested with prefetch.global.L2 while the
//     current one is transformed.  No shared sample buffers, no mbarriers, no TMA ring: a slow warp (chunk edges are
//     gathered with the reference's reflect / zero-pad index math) delays nobody;
//   * 1024-point complex FFT = 32 lanes x 32 registers, one transpose through the warp's own 8.4 KB buffer, even/odd
//     split against the partner lane through the same buffer (as in v2);
//   * complex / power: stored straight from registers, frequency-minor (torch.stft's own memory layout);
//   * mel: the frame's power line (float2 = (rowA, rowB) per bin, slot(k) = k + (k >> 5)) goes into the warp's buffer
//     and the SAME warp reduces it: lane g walks the 32 consecutive bins 32 g .. 32 g + 31 (conflict free: lanes are
//     33 float2 apart), two packed FFMAs per bin (tap of filter m_lo(k) and of m_lo(k) + 1 -- triangular banks are a
//     <= 2-adjacent-tap band), weights from a [step][lane] table.  A per-lane bit mask marks the last bin of every
//     run of equal m_lo: there the lane stores its (L, H) partial sums into the warp's run arrays and clears them.
//     The run a lane ends in the middle of is completed by adding its tail into the slot the next lane stored
//     (every 32-bin segment ends at least one run -- checked on the host, else the plan keeps the older kernels);
//     finally lane = filter: mel[m] = L[run m + 1] + H[run m] and the 128 values of a (row, frame) are written as one
//     contiguous 512-byte line -- [row][frame][mel], which is the memory layout of torchaudio's own result
//     (MelScale returns matmul(spec^T, fb)^T, a transposed view).
#pragma once

constexpr int kV3W = 12;                 // warps per CTA
constexpr int kV3Xb = 8456;              // per-warp buffer: 1057 float2
constexpr int kV3Runs = 136;             // slots of the per-warp run arrays (n_mels + 1 runs at most, last slot = permanent zero)
constexpr int kV3Tables = 32 * 256 + 512 + 256;   // tw1 [32][32] float2, Hann phases float4[32], W_2048^lane float2[32]
constexpr int kV3MelTab = 32 * 256 + 32 * 4 + 32 * 4 + 16;   // weights [32][32] float2, flush masks, first run per lane, w(bin 1024)

struct Stft3Args {
  const float* wav;          // [rows][n_in]
  float* out;
  int rows, n_in, n_pad, n_frames, hop, center_off;
  int n_items;               // row pairs x frames
  int n_freq, n_mels, wav_ok8, wav_ok16, prefetch;
  const float2* tw1;         // [32][32] W_1024^(k1 n2)
  const float* lane_consts;  // float4[32] Hann phases + float2[32] W_2048^lane
  const unsigned char* mel_tab;   // kV3MelTab bytes (see above), then uint32 [n_mels]: run slot of L | run slot of H << 16
};

__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f2(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }


// The warp-private banded mel walk over the power line in `xb_raw` (see the header): lane g = bins 32 g .. 32 g + 31.
__device__ __forceinline__ void v3_mel_walk(const Stft3Args& a, unsigned char* xb_raw, const unsigned char* s_mel, float4* s_runs,
                                            int lane, long long rowA, int frame, bool hasB) {
  // ---- the warp's own banded mel walk: lane g = bins 32 g .. 32 g + 31
  const uint32_t runs = smem_u32(s_runs);
  const uint32_t wbase = smem_u32(s_mel) + (uint32_t)lane * 8u;
  const uint32_t pbase = smem_u32(xb_raw) + (uint32_t)lane * (33u * 8u);
  const uint32_t mask = reinterpret_cast<const uint32_t*>(s_mel + 32 * 256)[lane];
  uint32_t mp = runs + reinterpret_cast<const uint32_t*>(s_mel + 32 * 256 + 128)[lane] * 16u;
  // One step = one bin: (L, H) += P * (w_lo, w_hi); a bin that closes a run stores (L, H) of both rows with one 128-bit store,
  // advances the run pointer and clears the sums (64-bit clears).
  unsigned long long aL = 0ull, aH = 0ull;   // packed (rowA, rowB) fp32 pairs
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    const float2 P = lds_f2(pbase + (uint32_t)t * 8u);
    const float2 w = lds_f2(wbase + (uint32_t)t * 256u);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(aL) : "l"(*reinterpret_cast<const unsigned long long*>(&P)), "l"(pack2(w.x, w.x)));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(aH) : "l"(*reinterpret_cast<const unsigned long long*>(&P)), "l"(pack2(w.y, w.y)));
    if (mask & (1u << t)) {
      asm volatile("st.shared.b64 [%0], %1;" ::"r"(mp), "l"(aL) : "memory");
      asm volatile("st.shared.b64 [%0+8], %1;" ::"r"(mp), "l"(aH) : "memory");
      mp += 16u;
      aL = 0ull;
      aH = 0ull;
    }
  }
  float2 aLf, aHf;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(aLf.x), "=f"(aLf.y) : "l"(aL));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(aHf.x), "=f"(aHf.y) : "l"(aH));
  const bool open_tail = (mask >> 31) == 0;   // my last bin did not close its run
  if (lane == 31) {   // bin 1024 closes the last run
    const float2 P = lds_f2(smem_u32(xb_raw) + 1056u * 8u);
    const float2 w = *reinterpret_cast<const float2*>(s_mel + 32 * 256 + 256);
    const float2 fL = open_tail ? pfma(P, w.x, aLf) : pmuls(P, w.x), fH = open_tail ? pfma(P, w.y, aHf) : pmuls(P, w.y);
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(mp), "f"(fL.x), "f"(fL.y), "f"(fH.x), "f"(fH.y) : "memory");
  }
  __syncwarp();
  if (lane < 31 && open_tail) {   // the unfinished run at the end of my segment was stored by the lane it ends in
    const float4 v = lds_f4(mp);
    const float2 nL = padd(make_float2(v.x, v.y), aLf), nH = padd(make_float2(v.z, v.w), aHf);
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(mp), "f"(nL.x), "f"(nL.y), "f"(nH.x), "f"(nH.y) : "memory");
  }
  __syncwarp();
  const uint32_t* s_filt = reinterpret_cast<const uint32_t*>(s_mel + kV3MelTab);
  float* oA = a.out + ((size_t)rowA * (size_t)a.n_frames + (size_t)frame) * (size_t)a.n_mels;
  float* oB = oA + (size_t)a.n_frames * (size_t)a.n_mels;
  for (int m = lane; m < a.n_mels; m += 32) {
    const uint32_t f = s_filt[m];
    const float2 v = padd(lds_f2(runs + (f & 0xffffu) * 16u), lds_f2(runs + (f >> 16) * 16u + 8u));
    oA[m] = v.x;
    if (hasB) oB[m] = v.y;
  }
  __syncwarp();   // run arrays and P line are consumed before the buffer is reused
}

template <int MODE>
__global__ void __launch_bounds__(kV3W * 32, 1) stft2048_v3_kernel(const __grid_constant__ Stft3Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hop = a.hop;
  unsigned char* xb_raw = smem + warp * kV3Xb;
  float2* XB = reinterpret_cast<float2*>(xb_raw);                         // transposes; later this frame's P line
  const uint32_t xb4 = (smem_u32(xb_raw) + 15u) & ~15u;                   // 16-byte aligned window for the float4 exchange
  unsigned char* tab = smem + kV3W * kV3Xb;
  float2* s_tw1 = reinterpret_cast<float2*>(tab);                          // [32][32] W_1024^(k1 lane)
  float4* s_lane = reinterpret_cast<float4*>(s_tw1 + 32 * 32);
  float2* s_tw2l = reinterpret_cast<float2*>(s_lane + 32);
  unsigned char* s_mel = reinterpret_cast<unsigned char*>(s_tw2l + 32);    // kV3MelTab bytes + filter slots (mel mode only)
  const int mel_bytes = (MODE == MODE_MEL) ? kV3MelTab + 4 * a.n_mels : 0;
  float4* s_runs = reinterpret_cast<float4*>(s_mel + ((mel_bytes + 15) & ~15)) + warp * kV3Runs;   // (L rowA, L rowB, H rowA, H rowB) per run

  for (int i = tid; i < 32 * 32; i += kV3W * 32) s_tw1[i] = __ldg(a.tw1 + i);
  if (tid < 48) reinterpret_cast<float4*>(s_lane)[tid] = __ldg(reinterpret_cast<const float4*>(a.lane_consts) + tid);
  if (MODE == MODE_MEL) {
    for (int i = tid; i < mel_bytes / 4; i += kV3W * 32)
      reinterpret_cast<uint32_t*>(s_mel)[i] = __ldg(reinterpret_cast<const uint32_t*>(a.mel_tab) + i);
    for (int i = lane; i < kV3Runs; i += 32) s_runs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = lane; i < kV3Xb / 8; i += 32) XB[i] = make_float2(0.f, 0.f);
  __syncthreads();

  const bool center = a.center_off != 0;
  const int plane = (32 - lane) & 31;
  const int pshift = (lane == 0) ? 16 : 15;
  const int stride_items = kV3W * (int)gridDim.x;

#pragma unroll 1
  for (int item = (int)blockIdx.x * kV3W + warp; item < a.n_items; item += stride_items) {
    const int pair = item / a.n_frames;
    const int frame = item - pair * a.n_frames;
    const long long rowA = 2LL * pair;
    const bool hasB = rowA + 1 < a.rows;
    const int sf = frame * hop - a.center_off;
    const float* __restrict__ pa = a.wav + (size_t)rowA * (size_t)a.n_in;
    const bool fast = hasB && a.wav_ok8 && sf >= 0 && sf + 2048 <= a.n_in;
    if (a.prefetch && lane == 0) {   // the part of this warp's NEXT frame nobody has touched yet -> L2, through the TMA unit (no LSU traffic)
      const int nitem = item + stride_items;
      if (nitem < a.n_items) {
        const int np = nitem / a.n_frames;
        const int nf = nitem - np * a.n_frames;
        const int ns = nf * hop - a.center_off + 2048 - hop;   // the last hop samples of the frame
        if (ns >= 0 && ns + hop <= a.n_in && a.wav_ok16 && 2 * np + 1 < a.rows) {
          const float* q = a.wav + (size_t)(2 * np) * (size_t)a.n_in + ns;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q), "r"(hop * 4) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q + a.n_in), "r"(hop * 4) : "memory");
        }
      }
    }
    float2 re[32], im[32];
    {
      const float4 ph = s_lane[lane];   // (cos phi0, sin phi0, cos phi1, sin phi1), phi = 2 pi (2 lane + {0,1}) / 2048
      if (fast) {
        const float2* FA = reinterpret_cast<const float2*>(pa + sf);
        const float2* FB = reinterpret_cast<const float2*>(pa + a.n_in + sf);
        float2 xa[32], xb[32];
#pragma unroll
        for (int sl = 0; sl < 32; ++sl) {   // in the order the first butterflies consume them (FFT slot order), rows interleaved
          xa[bitrev5(sl)] = __ldg(FA + 32 * bitrev5(sl) + lane);
          xb[bitrev5(sl)] = __ldg(FB + 32 * bitrev5(sl) + lane);
        }
#pragma unroll
        for (int sl = 0; sl < 32; ++sl) {
          const int n1 = bitrev5(sl);
          const float w0 = fmaf(0.5f * aa_consts::kSin32[n1], ph.y, fmaf(-0.5f * aa_consts::kCos32[n1], ph.x, 0.5f));
          const float w1 = fmaf(0.5f * aa_consts::kSin32[n1], ph.w, fmaf(-0.5f * aa_consts::kCos32[n1], ph.z, 0.5f));
          re[sl] = make_float2(xa[n1].x * w0, xb[n1].x * w0);
          im[sl] = make_float2(xa[n1].y * w1, xb[n1].y * w1);
        }
      } else {
        // chunk edge / odd last row / unaligned rows: gather the frame with the reference's reflect + zero-pad index math,
        // half a frame (1024 samples of both rows) at a time, through this warp's own buffer
        const float* __restrict__ pb = hasB ? pa + a.n_in : pa;
        const float bmask = hasB ? 1.f : 0.f;
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

    fft32_dit(re, im);
#pragma unroll
    for (int j = 1; j < 32; ++j) {
      const float2 tw = s_tw1[j * 32 + lane];
      const float2 r = re[j], i = im[j];
      re[j] = pfma(i, -tw.y, pmuls(r, tw.x));
      im[j] = pfma(i, tw.x, pmuls(r, tw.y));
    }
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
      float2 pk[16], pq[16];   // 4|X[k]|^2, 4|X[1024-k]|^2
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
        const float2 xr = pfma(o_i, -tw.y, pfma(o_r, tw.x, e_r));
        const float2 xi = pfma(o_r, tw.y, pfma(o_i, tw.x, e_i));
        const float2 yr = pfma(o_i, tw.y, pfma(o_r, -tw.x, e_r));
        const float2 yi = pfma(o_r, -tw.y, pfma(o_i, -tw.x, e_i));
        pk[i] = pfma2(xi, xi, pmul(xr, xr));
        pq[i] = pfma2(yi, yi, pmul(yr, yr));
      }
      if (lane == 0) {
        const float2 dc = padd(z0r, z0i), ny = psub(z0r, z0i);
        pk[0] = pmuls(pmul(dc, dc), 4.f);
        pq[0] = pmuls(pmul(ny, ny), 4.f);
      }
      __syncwarp();   // every partner row has been read: the buffer becomes the P line, slot(k) = k + (k >> 5)
      const uint32_t plo = smem_u32(xb_raw) + (uint32_t)lane * 8u;                                   // bin lane + 32 i -> slot lane + 33 i
      const uint32_t phi = smem_u32(xb_raw) + (uint32_t)(1055 - lane + (lane == 0 ? 1 : 0)) * 8u;    // bin 1024 - lane - 32 i -> this - 33 i
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        sts_f2(plo + (uint32_t)(33 * i) * 8u, pk[i]);
        sts_f2(phi - (uint32_t)(33 * i) * 8u, pq[i]);
      }
      if (lane == 0) sts_f2(smem_u32(xb_raw) + 528u * 8u, pmuls(pfma2(z16i, z16i, pmul(z16r, z16r)), 4.f));   // bin 512
      __syncwarp();
      v3_mel_walk(a, xb_raw, s_mel, s_runs, lane, rowA, frame, hasB);
    } else {
      // complex / power, frequency-minor output [row][frame][1025]: stored straight from registers
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
      __syncwarp();   // partner rows consumed before the next frame reuses the buffer
    }
  }
}
