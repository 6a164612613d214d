// stft_v3_kernel<NF, MODE> -- STFT / power / mel front-end for n_fft = 2048 and n_fft = 1024 (the reference's default
// geometry) without ANY inter-warp synchronisation (included by stft.cu; FFT arithmetic as in stft2048_v2_kernel).
//
// Every warp owns whole frames of a row pair (two rows in the halves of packed fp32x2 registers) from the first load
// to the last store.  One work item is 1024 complex points: ONE frame of n_fft = 2048 or TWO consecutive frames of
// n_fft = 1024 (FR = 2048 / NF frames per item):
//   * samples come straight from global memory / L2 with coalesced 64-bit loads (a frame row is NF consecutive
//     floats; the 75 % overlap between consecutive frames is served by L1 / L2, the 12 warps of a CTA take 12
//     consecutive items).  No shared sample buffers, no mbarriers, no TMA ring: a slow warp (chunk edges are
//     gathered with the reference's reflect / zero-pad index math) delays nobody;
//   * FFT = (32 / FR)-point FFTs in registers (FR of them, lane = n2) x a 32-point FFT after one transpose through the
//     warp's own 8.4 KB buffer (then lane = (frame, k1), slot = k2: Z[k1 + (32 / FR) k2]); the even/odd split pairs lane
//     (frame, k1) with lane (frame, -k1) through the same buffer;
//   * complex / power: stored straight from registers, frequency-minor (torch.stft's own memory layout);
//   * mel: the power lines (float2 = (rowA, rowB) per bin, slot(k) = k + (k >> 5) inside a frame's region) go into the
//     warp's buffer and the SAME warp reduces them: lane g walks 32 consecutive bins of one frame (conflict free: lanes
//     are 33 float2 apart), two packed FFMAs per bin (tap of filter m_lo(k) and of m_lo(k) + 1 -- triangular banks are a
//     <= 2-adjacent-tap band), weights from a [step][lane] table.  A per-lane bit mask marks the last bin of every
//     run of equal m_lo: there the lane stores its (L, H) partial sums into the warp's run arrays and clears them.
//     The run a lane ends in the middle of is completed by adding its tail into the slot the next lane stored
//     (every 32-bin segment ends at least one run -- checked on the host, else the plan keeps the older kernels);
//     finally lane = filter: mel[m] = L[run m + 1] + H[run m] and the n_mels values of a (row, frame) are written as one
//     contiguous line -- [row][frame][mel], which is the memory layout of torchaudio's own result
//     (MelScale returns matmul(spec^T, fb)^T, a transposed view).
// Measured alternatives that were NOT kept: deferring the walk behind the next frame's loads (rotated loop),
// hoisting the walk's loads / multiplying the sums by a keep flag instead of clearing them, the partner exchange by warp
// shuffles, and dedicated mel warps fed through mbarriers with setmaxnreg (16 warps per SM).
#pragma once

constexpr int kV3W = 12;                 // warps per CTA
constexpr int kV3Xb = 8464;              // per-warp buffer: 1058 float2 ([32][33] transposes; 1057 / 2 x 529 power-line slots)
constexpr int kV3Runs = 136;             // run slots per frame (n_mels + 1 runs at most, last slot = permanent zero)
constexpr int kV3Tables = 32 * 256 + 512 + 256;   // tw1 [32][32] float2, Hann phases float4[32], W_NF^k1 float2[32]
constexpr int kV3MelTab = 32 * 256 + 32 * 4 + 32 * 4 + 16;   // weights [32 steps][32 lanes] float2, close masks, first run per lane, w(last bin)

struct Stft3Args {
  const float* wav;          // [rows][n_in]
  float* out;
  int rows, n_in, n_pad, n_frames, hop, center_off;
  int n_items, items_per_pair;   // row pairs x ceil(frames / FR)
  int n_freq, n_mels, wav_ok8, wav_ok16, prefetch;
  const float2* tw1;         // NF = 2048: [32][32] W_1024^(k1 n2);  NF = 1024: [16][32] W_512^(k1 n2)
  const float* lane_consts;  // float4[32] Hann phases (cos phi0, cos phi1, sin phi0, sin phi1; phi_e = 2 pi (2 lane + e) / NF) + float2[32] W_NF^(k1 of the lane)
  const unsigned char* mel_tab;   // kV3MelTab bytes (see above), then uint32 [n_mels]: run slot of L | run slot of H << 16
};

__host__ __device__ constexpr int bitrev4(int n) { return ((n & 1) << 3) | ((n & 2) << 1) | ((n & 4) >> 1) | ((n & 8) >> 3); }

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

// The warp-private banded mel walk over the power line(s) in `xb_raw` (see the header).  Lane g = frame g / SEGS of the item,
// bins 32 (g % SEGS) .. + 31 of it; SEGS = 32 / FR segments per frame, the last one also takes the frame's last bin.
template <int NF>
__device__ __forceinline__ void v3_mel_walk(const Stft3Args& a, unsigned char* xb_raw, const unsigned char* s_mel, float4* s_runs,
                                            int lane, long long rowA, int frame0, bool hasB) {
  constexpr int FR = 2048 / NF, SEGS = 32 / FR, H = NF / 2;
  constexpr uint32_t PSTRIDE = FR == 1 ? 0u : 529u;           // float2 slots of one frame's power line (FR == 2)
  const int fr = lane / SEGS, seg = lane % SEGS;
  const uint32_t runs = smem_u32(s_runs);
  const uint32_t wbase = smem_u32(s_mel) + (uint32_t)lane * 8u;
  const uint32_t pline = smem_u32(xb_raw) + (uint32_t)fr * PSTRIDE * 8u;
  const uint32_t pbase = pline + (uint32_t)seg * (33u * 8u);
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
  if (seg == SEGS - 1) {   // the frame's last bin (n_fft / 2) closes the last run
    const float2 P = lds_f2(pline + (uint32_t)(H + (H >> 5)) * 8u);
    const float2 w = *reinterpret_cast<const float2*>(s_mel + 32 * 256 + 256);
    const float2 fL = open_tail ? pfma(P, w.x, aLf) : pmuls(P, w.x), fH = open_tail ? pfma(P, w.y, aHf) : pmuls(P, w.y);
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(mp), "f"(fL.x), "f"(fL.y), "f"(fH.x), "f"(fH.y) : "memory");
  }
  __syncwarp();
  if (seg < SEGS - 1 && open_tail) {   // the unfinished run at the end of my segment was stored by the lane it ends in
    const float4 v = lds_f4(mp);
    const float2 nL = padd(make_float2(v.x, v.y), aLf), nH = padd(make_float2(v.z, v.w), aHf);
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(mp), "f"(nL.x), "f"(nL.y), "f"(nH.x), "f"(nH.y) : "memory");
  }
  __syncwarp();
  const uint32_t* s_filt = reinterpret_cast<const uint32_t*>(s_mel + kV3MelTab);
#pragma unroll
  for (int f = 0; f < FR; ++f) {
    if (frame0 + f >= a.n_frames) break;
    float* oA = a.out + ((size_t)rowA * (size_t)a.n_frames + (size_t)(frame0 + f)) * (size_t)a.n_mels;
    float* oB = oA + (size_t)a.n_frames * (size_t)a.n_mels;
    const uint32_t rf = runs + (uint32_t)f * (kV3Runs * 16u);
    for (int m = lane; m < a.n_mels; m += 32) {
      const uint32_t fw = s_filt[m];
      const float2 v = padd(lds_f2(rf + (fw & 0xffffu) * 16u), lds_f2(rf + (fw >> 16) * 16u + 8u));
      oA[m] = v.x;
      if (hasB) oB[m] = v.y;
    }
  }
  __syncwarp();   // run arrays and P line are consumed before the buffer is reused
}

template <int NF, int MODE>
__global__ void __launch_bounds__(kV3W * 32, 1) stft_v3_kernel(const __grid_constant__ Stft3Args a) {
  static_assert(NF == 2048 || NF == 1024, "one item = 1024 complex points");
  constexpr int FR = 2048 / NF;        // frames per item
  constexpr int R1 = 32 / FR;          // points of the in-register first-stage FFTs = bins stride of the slots after stage 2
  constexpr int H = NF / 2;            // complex points per frame; bins 0 .. H
  constexpr uint32_t PSTRIDE = FR == 1 ? 0u : 529u;
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hop = a.hop;
  unsigned char* xb_raw = smem + warp * kV3Xb;
  float2* XB = reinterpret_cast<float2*>(xb_raw);                         // transposes; later this item's power line(s)
  const uint32_t xb4 = smem_u32(xb_raw);                                  // 16-byte aligned window for the float4 exchange
  unsigned char* tab = smem + kV3W * kV3Xb;
  float2* s_tw1 = reinterpret_cast<float2*>(tab);                          // [R1][32] W_H^(k1 lane)
  float4* s_lane = reinterpret_cast<float4*>(s_tw1 + 32 * 32);
  float2* s_tw2l = reinterpret_cast<float2*>(s_lane + 32);
  unsigned char* s_mel = reinterpret_cast<unsigned char*>(s_tw2l + 32);    // kV3MelTab bytes + filter slots (mel mode only)
  const int mel_bytes = (MODE == MODE_MEL) ? kV3MelTab + 4 * a.n_mels : 0;
  float4* s_runs = reinterpret_cast<float4*>(s_mel + ((mel_bytes + 15) & ~15)) + warp * (FR * kV3Runs);   // (L rowA, L rowB, H rowA, H rowB) per run

  for (int i = tid; i < R1 * 32; i += kV3W * 32) s_tw1[i] = __ldg(a.tw1 + i);
  if (tid < 48) reinterpret_cast<float4*>(s_lane)[tid] = __ldg(reinterpret_cast<const float4*>(a.lane_consts) + tid);
  if (MODE == MODE_MEL) {
    for (int i = tid; i < mel_bytes / 4; i += kV3W * 32)
      reinterpret_cast<uint32_t*>(s_mel)[i] = __ldg(reinterpret_cast<const uint32_t*>(a.mel_tab) + i);
    for (int i = lane; i < FR * kV3Runs; i += 32) s_runs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = lane; i < kV3Xb / 8; i += 32) XB[i] = make_float2(0.f, 0.f);
  __syncthreads();

  const bool center = a.center_off != 0;
  const int frL = lane / R1;                                   // this lane's frame of the item after the transpose
  const int k1 = lane % R1;                                    // ... and its k1: slot k2 holds Z[k1 + R1 k2]
  const int plane = frL * R1 + ((R1 - k1) % R1);               // partner lane of the even/odd split
  const int pshift = (k1 == 0) ? 16 : 15;
  const int stride_items = kV3W * (int)gridDim.x;

#pragma unroll 1
  for (int item = (int)blockIdx.x * kV3W + warp; item < a.n_items; item += stride_items) {
    const int pair = item / a.items_per_pair;
    const int frame0 = (item - pair * a.items_per_pair) * FR;
    const long long rowA = 2LL * pair;
    const bool hasB = rowA + 1 < a.rows;
    const int sf = frame0 * hop - a.center_off;                // first sample of the item's first frame
    const float* __restrict__ pa = a.wav + (size_t)rowA * (size_t)a.n_in;
    const bool fast = hasB && a.wav_ok8 && sf >= 0 && sf + (FR - 1) * hop + NF <= a.n_in && frame0 + FR <= a.n_frames;
    if (FR == 1 && a.prefetch && lane == 0) {   // the part of this warp's NEXT frame nobody has touched yet -> L2, through the TMA unit
      const int nitem = item + stride_items;
      if (nitem < a.n_items) {
        const int np = nitem / a.items_per_pair;
        const int nf = nitem - np * a.items_per_pair;
        const int ns = nf * hop - a.center_off + NF - hop;   // the last hop samples of the frame
        if (ns >= 0 && ns + hop <= a.n_in && a.wav_ok16 && 2 * np + 1 < a.rows) {
          const float* q = a.wav + (size_t)(2 * np) * (size_t)a.n_in + ns;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q), "r"(hop * 4) : "memory");
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q + a.n_in), "r"(hop * 4) : "memory");
        }
      }
    }
    float2 re[32], im[32];
    {
      // Hann window of sample 2 (32 n1 + lane) + e: 0.5 - 0.5 cos(2 pi n1 / R1 + phi_e), phi_e = 2 pi (2 lane + e) / NF
      const float4 ph = s_lane[lane];   // (cos phi0, cos phi1, sin phi0, sin phi1)
      const float2 phc = make_float2(ph.x, ph.y), phs = make_float2(ph.z, ph.w);
      if (fast) {
        float2 xa[32], xb[32];
#pragma unroll
        for (int s1 = 0; s1 < R1; ++s1)   // in the order the first butterflies consume them (FFT slot order), rows interleaved
#pragma unroll
          for (int f = 0; f < FR; ++f) {
            const int n1 = FR == 1 ? bitrev5(s1) : bitrev4(s1);
            const float2* FA = reinterpret_cast<const float2*>(pa + sf + f * hop);
            const float2* FB = reinterpret_cast<const float2*>(pa + a.n_in + sf + f * hop);
            xa[f * R1 + s1] = __ldg(FA + 32 * n1 + lane);
            xb[f * R1 + s1] = __ldg(FB + 32 * n1 + lane);
          }
#pragma unroll
        for (int s1 = 0; s1 < R1; ++s1) {
          const int n1 = FR == 1 ? bitrev5(s1) : bitrev4(s1);
          // both window values of the lane's sample pair in one packed FFMA chain (same fmaf arithmetic per component)
          const float2 w01 = pfma(phs, 0.5f * aa_consts::kSin32[FR * n1], pfma(phc, -0.5f * aa_consts::kCos32[FR * n1], make_float2(0.5f, 0.5f)));
          const float w0 = w01.x, w1 = w01.y;
#pragma unroll
          for (int f = 0; f < FR; ++f) {
            re[f * R1 + s1] = make_float2(xa[f * R1 + s1].x * w0, xb[f * R1 + s1].x * w0);
            im[f * R1 + s1] = make_float2(xa[f * R1 + s1].y * w1, xb[f * R1 + s1].y * w1);
          }
        }
      } else {
        // chunk edge / odd last row / unaligned rows / odd last frame: gather with the reference's reflect + zero-pad index
        // math, 1024 samples of both rows at a time (half a 2048-frame or one 1024-frame), through this warp's own buffer
        const float* __restrict__ pb = hasB ? pa + a.n_in : pa;
        const float bmask = hasB ? 1.f : 0.f;
        float* GA = reinterpret_cast<float*>(xb_raw);   // [1024] row A, then [1024] row B
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int s0 = FR == 1 ? sf + half * 1024 : sf + half * hop;
          __syncwarp();
#pragma unroll 8
          for (int j = lane; j < 1024; j += 32) {
            GA[j] = fetch_sample_nb(pa, s0 + j, a.n_in, a.n_pad, center);
            GA[1024 + j] = bmask * fetch_sample_nb(pb, s0 + j, a.n_in, a.n_pad, center);
          }
          __syncwarp();
          const float2* FA = reinterpret_cast<const float2*>(GA);
          const float2* FB = reinterpret_cast<const float2*>(GA + 1024);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int n1 = FR == 1 ? half * 16 + q : q;
            const int sl = FR == 1 ? bitrev5(n1) : half * 16 + bitrev4(q);
            const float2 xa = FA[32 * q + lane], xb = FB[32 * q + lane];
            const float2 w01 = pfma(phs, 0.5f * aa_consts::kSin32[FR * n1], pfma(phc, -0.5f * aa_consts::kCos32[FR * n1], make_float2(0.5f, 0.5f)));
            const float w0 = w01.x, w1 = w01.y;
            re[sl] = make_float2(xa.x * w0, xb.x * w0);
            im[sl] = make_float2(xa.y * w1, xb.y * w1);
          }
        }
        __syncwarp();
      }
    }

    if constexpr (FR == 1) {
      fft32_dit(re, im);
    } else {
      fft16_dit(*reinterpret_cast<float2(*)[16]>(&re[0]), *reinterpret_cast<float2(*)[16]>(&im[0]));
      fft16_dit(*reinterpret_cast<float2(*)[16]>(&re[16]), *reinterpret_cast<float2(*)[16]>(&im[16]));
    }
#pragma unroll
    for (int j = 1; j < R1; ++j) {
      const float2 tw = s_tw1[j * 32 + lane];
#pragma unroll
      for (int f = 0; f < FR; ++f) {
        const float2 r = re[f * R1 + j], i = im[f * R1 + j];
        re[f * R1 + j] = pfma(i, -tw.y, pmuls(r, tw.x));
        im[f * R1 + j] = pfma(i, tw.x, pmuls(r, tw.y));
      }
    }
#pragma unroll
    for (int s = 0; s < 32; ++s) XB[s * 33 + lane] = re[s];
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) re[bitrev5(n2)] = XB[lane * 33 + n2];
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 32; ++s) XB[s * 33 + lane] = im[s];
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) im[bitrev5(n2)] = XB[lane * 33 + n2];
    __syncwarp();
    fft32_dit(re, im);   // slot k2 of lane (frL, k1): Z_frL[k1 + R1 k2]
    // publish the upper half (k2 >= 16) for the partner lane
#pragma unroll
    for (int k2 = 16; k2 < 32; ++k2)
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(xb4 + (uint32_t)((k2 - 16) * 32 + lane) * 16u), "f"(re[k2].x),
                   "f"(re[k2].y), "f"(im[k2].x), "f"(im[k2].y)
                   : "memory");
    __syncwarp();

    // Even/odd split for the pair (k, H - k), k = k1 + R1 i (i < 16): own Z[k] in slot i, partner's Z[H - k] in lane
    // (frL, -k1) slot 31 - i (k1 = 0: slot 32 - i).  E2 = a + conj(b), O2 = (a - conj(b)) / i, T = W_NF^k O2:
    // 2 X[k] = E2 + T, 2 X[H - k] = conj(E2 - T); W_NF^k = W_NF^k1 W_64^i for both transform sizes.
    const float2 cl = s_tw2l[lane];
    const float2 z0r = re[0], z0i = im[0], z16r = re[16], z16i = im[16];
    const int frame = frame0 + frL;                            // the frame this lane's bins belong to
    const bool fvalid = FR == 1 || frame < a.n_frames;   // FR = 2: the second frame of the last item of an odd frame count
    if constexpr (MODE == MODE_MEL) {
      float2 pk[16], pq[16];   // 4|X[k]|^2, 4|X[H-k]|^2
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
      if (k1 == 0) {
        const float2 dc = padd(z0r, z0i), ny = psub(z0r, z0i);
        pk[0] = pmuls(pmul(dc, dc), 4.f);
        pq[0] = pmuls(pmul(ny, ny), 4.f);
      }
      __syncwarp();   // every partner row has been read: the buffer becomes the power line(s), slot(k) = k + (k >> 5)
      const uint32_t pline = smem_u32(xb_raw) + (uint32_t)frL * PSTRIDE * 8u;
      if constexpr (FR == 1) {
        const uint32_t plo = pline + (uint32_t)lane * 8u;                                   // bin lane + 32 i -> slot lane + 33 i
        const uint32_t phi = pline + (uint32_t)(1055 - lane + (lane == 0 ? 1 : 0)) * 8u;    // bin 1024 - lane - 32 i -> this - 33 i
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          sts_f2(plo + (uint32_t)(33 * i) * 8u, pk[i]);
          sts_f2(phi - (uint32_t)(33 * i) * 8u, pq[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          sts_f2(pline + (uint32_t)(k1 + 16 * i + (i >> 1)) * 8u, pk[i]);   // bin k1 + 16 i
          const int kk = H - k1 - 16 * i;                                     // bin H - k
          sts_f2(pline + (uint32_t)(kk + (kk >> 5)) * 8u, pq[i]);
        }
      }
      if (k1 == 0) sts_f2(pline + (uint32_t)(H / 2 + (H >> 6)) * 8u, pmuls(pfma2(z16i, z16i, pmul(z16r, z16r)), 4.f));   // bin H / 2
      __syncwarp();
      v3_mel_walk<NF>(a, xb_raw, s_mel, s_runs, lane, rowA, frame0, hasB);
    } else {
      // complex / power, frequency-minor output [row][frame][H + 1]: stored straight from registers
      const long long e0 = (rowA * (long long)a.n_frames + frame) * a.n_freq;
      const long long e1 = e0 + (long long)a.n_frames * a.n_freq;
      if (fvalid) {   // one (rarely divergent) branch around the whole epilogue keeps the stores inside it predicated, not branched
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
        const float2 yr = pfma(o_i, tw.y, pfma(o_r, -tw.x, e_r));   // 2 Re X[H-k]
        const float2 yi = pfma(o_r, -tw.y, pfma(o_i, -tw.x, e_i));  // -2 Im X[H-k]
        const int k = k1 + R1 * i;
        const bool skip = (k1 == 0 && i == 0);   // DC / Nyquist come from the k1 = 0 lanes below
        if constexpr (MODE == MODE_POWER) {
          const float2 pa_ = pmuls(pfma2(xi, xi, pmul(xr, xr)), 0.25f);
          const float2 pq = pmuls(pfma2(yi, yi, pmul(yr, yr)), 0.25f);
          if (!skip) {
            a.out[e0 + k] = pa_.x;
            a.out[e0 + H - k] = pq.x;
            if (hasB) {
              a.out[e1 + k] = pa_.y;
              a.out[e1 + H - k] = pq.y;
            }
          }
        } else {
          float2* o = reinterpret_cast<float2*>(a.out);
          if (!skip) {
            o[e0 + k] = make_float2(0.5f * xr.x, 0.5f * xi.x);
            o[e0 + H - k] = make_float2(0.5f * yr.x, -0.5f * yi.x);
            if (hasB) {
              o[e1 + k] = make_float2(0.5f * xr.y, 0.5f * xi.y);
              o[e1 + H - k] = make_float2(0.5f * yr.y, -0.5f * yi.y);
            }
          }
        }
      }
      if (k1 == 0) {
        const float2 dc = padd(z0r, z0i), ny = psub(z0r, z0i);
        if constexpr (MODE == MODE_POWER) {
          const float2 pmid = pfma2(z16i, z16i, pmul(z16r, z16r));
          a.out[e0] = dc.x * dc.x; a.out[e0 + H] = ny.x * ny.x; a.out[e0 + H / 2] = pmid.x;
          if (hasB) { a.out[e1] = dc.y * dc.y; a.out[e1 + H] = ny.y * ny.y; a.out[e1 + H / 2] = pmid.y; }
        } else {
          float2* o = reinterpret_cast<float2*>(a.out);
          o[e0] = make_float2(dc.x, 0.f); o[e0 + H] = make_float2(ny.x, 0.f); o[e0 + H / 2] = make_float2(z16r.x, -z16i.x);
          if (hasB) { o[e1] = make_float2(dc.y, 0.f); o[e1 + H] = make_float2(ny.y, 0.f); o[e1 + H / 2] = make_float2(z16r.y, -z16i.y); }
        }
      }
      }
      __syncwarp();   // partner rows consumed before the next item reuses the buffer
    }
  }
}
