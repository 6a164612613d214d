// STFT / power / mel front-end for sm_100a.
//
// Replaces torchaudio.transforms.{Spectrogram,MelSpectrogram} as used by the reference's
// SpectrogramAE / MagSpectrogramAE / MelSpectrogramAE.encode (audio_algebra/given_models.py:149-283)
// and GivenModelClass.zero_pad_po2 (:139-145).
//
// Fast path (n_fft = 2048, hop % 4 == 0, hop <= 1024) -- `stft2048_kernel`:
//   * one CTA = 4 warps = 4 consecutive frames of one ROW PAIR; the two rows ride in the two halves
//     of packed fp32x2 registers, so the FFT arithmetic is FADD2/FFMA2 (one issue slot per 2 flops);
//   * the tile's samples are brought into shared memory once per row by a TMA bulk copy
//     (cp.async.bulk + mbarrier; reflect padding and the zero_pad_po2 tail are handled in index
//     math on the few tiles at the chunk edges), so the 75 % frame overlap costs no extra HBM
//     traffic inside a tile;
//   * a real 2048-point FFT is one 1024-point complex FFT: lane n2 holds z[32*n1+n2] (32 packed
//     complex values), runs a 32-point in-register FFT, multiplies by W_1024^(n2*k1), transposes
//     through a warp-private shared-memory buffer, runs the second 32-point FFT (lane k1 then holds
//     Z[k1+32*k2]), and the even/odd split pairs lane k1 with lane 32-k1 through the same buffer;
//   * epilogues: |X|^2 -> sparse (<= 2 taps per bin) mel projection from shared memory, or
//     staged power / complex stores that write 16-byte runs along the frame axis.
// Generic path (any power-of-two n_fft in [64, 8192]) -- `stft_generic_kernel`: one CTA per
//   (row, frame), shared-memory radix-2 FFT; correct for every configuration, not tuned.
#include "aa_common.cuh"
#include "fft_gen.cuh"

#include <cmath>
#include <vector>

namespace {

constexpr int MODE_COMPLEX = 0, MODE_POWER = 1, MODE_MEL = 2;

__host__ __device__ constexpr int bitrev5(int n) {
  return ((n & 1) << 4) | ((n & 2) << 2) | (n & 4) | ((n & 8) >> 2) | ((n & 16) >> 4);
}

struct StftArgs {
  const float* wav;     // [rows][n_in]
  long long rows, n_in, n_pad;
  int n_frames, hop, center_off, tiles_per_pair;
  int n_freq, n_mels;
  const float2* window2;  // [n_fft/2] (w[2n], w[2n+1])
  const float2* tw1;      // fast: [32][32] W_1024^(k1*n2);  generic: [M/2] W_M^j
  const float2* tw2;      // [n_fft/4] W_{n_fft}^k
  const int* mel_start4;  // [n_mels] first bin of the filter, rounded down to a multiple of 4
  const int* mel_cnt4;    // [n_mels] number of float4 weight groups
  const int* mel_off4;    // [n_mels] offset (in float4) into mel_w4
  const float4* mel_w4;   // padded filter weights, pre-multiplied by 0.25 (P holds 4|X|^2)
  float* out;
  int wav_aligned16;
};

__device__ __forceinline__ float fetch_sample(const float* __restrict__ p, long long i, long long n_in,
                                              long long n_pad, bool center) {
  if (p == nullptr) return 0.0f;
  if (center) {  // torch.stft pad_mode="reflect" on the zero-padded signal of length n_pad
    if (i < 0) i = -i;
    if (i >= n_pad) i = 2 * (n_pad - 1) - i;
  }
  if (i < 0 || i >= n_in) return 0.0f;  // zero_pad_po2 tail (given_models.py:139-145)
  return __ldg(p + i);
}

// ------------------------------------------------------------------------------------------
// Fast kernel, n_fft = 2048.
// ------------------------------------------------------------------------------------------
constexpr int kWarps = 4;                       // frames per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kXbBytes = 32 * 33 * 8;           // warp-private exchange buffer (float2 [32][33])
constexpr int kPStride = 1028;                  // floats per (frame,row) line of the P staging area
constexpr int kStageBytes = kWarps * kXbBytes;  // 33792 >= 8 * 1028 * 4 = 32896

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "AA_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra AA_DONE;\n"
      "bra AA_WAIT;\n"
      "AA_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 3) stft2048_kernel(const StftArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hop = a.hop;
  const int span = 3 * hop + 2048;              // samples covered by the 4 frames of this tile
  float* SA = reinterpret_cast<float*>(smem);                        // row A samples [span]
  float* SB = SA + span;                                             // row B samples [span]
  unsigned char* stage = smem + (size_t)span * 8;                    // kStageBytes
  float2* XB = reinterpret_cast<float2*>(stage + warp * kXbBytes);   // this warp's exchange buffer
  uint64_t* mbar = reinterpret_cast<uint64_t*>(stage + kStageBytes);

  const long long pair = blockIdx.x / a.tiles_per_pair;
  const int tile = blockIdx.x % a.tiles_per_pair;
  const int f0 = tile * kWarps;
  const long long rowA = 2 * pair, rowB = rowA + 1;
  const bool hasB = rowB < a.rows;
  const float* __restrict__ pa = a.wav + rowA * a.n_in;
  const float* __restrict__ pb = hasB ? a.wav + rowB * a.n_in : nullptr;
  const long long s0 = (long long)f0 * hop - a.center_off;
  const bool center = a.center_off != 0;

  // ---- stage the tile's samples ---------------------------------------------------------------
  if (s0 >= 0 && s0 + span <= a.n_in && hasB && a.wav_aligned16 && (a.n_in & 3) == 0) {
    if (tid == 0) {
      mbar_init(mbar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(mbar, (uint32_t)span * 8u);
      bulk_g2s(SA, pa + s0, (uint32_t)span * 4u, mbar);
      bulk_g2s(SB, pb + s0, (uint32_t)span * 4u, mbar);
    }
    __syncthreads();          // barrier init visible to every waiter
    mbar_wait(mbar, 0);
  } else {                    // chunk edges / odd row count / unaligned rows: reflect + zero tail
    for (int j = tid; j < span; j += kThreads) {
      SA[j] = fetch_sample(pa, s0 + j, a.n_in, a.n_pad, center);
      SB[j] = fetch_sample(pb, s0 + j, a.n_in, a.n_pad, center);
    }
    __syncthreads();
  }

  const int frame = f0 + warp;
  const bool fvalid = frame < a.n_frames;

  float2 re[32], im[32];
  if (fvalid) {
    // ---- load + window: z[n] = x[2n] w[2n] + i x[2n+1] w[2n+1], n = 32*n1 + lane ------------
    const float2* FA = reinterpret_cast<const float2*>(SA + warp * hop);
    const float2* FB = reinterpret_cast<const float2*>(SB + warp * hop);
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
      const int n = 32 * n1 + lane;
      const float2 xa = FA[n], xb = FB[n];
      const float2 w = __ldg(a.window2 + n);
      re[bitrev5(n1)] = make_float2(xa.x * w.x, xb.x * w.x);
      im[bitrev5(n1)] = make_float2(xa.y * w.y, xb.y * w.y);
    }
    fft32_dit(re, im);  // slot k1: sum_n1 z[32 n1 + lane] W_32^(n1 k1)
    // ---- twiddle W_1024^(lane*k1) ------------------------------------------------------------
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) {
      const float2 t = __ldg(a.tw1 + k1 * 32 + lane);
      const float2 r = re[k1], i = im[k1];
      re[k1] = pfma(i, -t.y, pmuls(r, t.x));
      im[k1] = pfma(i, t.x, pmuls(r, t.y));
    }
    // ---- transpose through the warp-private buffer (re half, then im half) ----------------------
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
    fft32_dit(re, im);  // slot k2 of lane k1: Z[k1 + 32 k2]
    // ---- publish the upper half (k2 >= 16) for the partner lane ------------------------------------
    float4* XB4 = reinterpret_cast<float4*>(XB);
#pragma unroll
    for (int k2 = 16; k2 < 32; ++k2)
      XB4[(k2 - 16) * 32 + lane] = make_float4(re[k2].x, re[k2].y, im[k2].x, im[k2].y);
    __syncwarp();
  }

  // Even/odd split for the pair (k, 1024-k), k = lane + 32 i (i < 16): own Z[k] in slot i, partner's
  // Z[1024-k] in lane (32-lane)&31 slot 31-i (lane 0: slot 32-i).  With E2 = a + conj(b),
  // O2 = (a - conj(b))/i, T = W_2048^k O2:  2 X[k] = E2 + T,  2 X[1024-k] = conj(E2 - T).
  const int plane = (32 - lane) & 31;
  const int pshift = (lane == 0) ? 16 : 15;  // partner slot - 16 = pshift - i
  const float4* XB4r = reinterpret_cast<const float4*>(XB);

  if constexpr (MODE == MODE_COMPLEX) {
    // Four k-quarters, each staged as [8 (frame,row)][260] complex in the (now dead) sample area and
    // stored with lanes = (8 bins) x (4 frames) so that every (row, bin) gets one 32-byte run.
    float2* CS = reinterpret_cast<float2*>(smem);
    constexpr int kCStride = 260;
    const long long rowbase = rowA;
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      __syncthreads();  // sample area (q = 0) / previous quarter's staging is free
      if (fvalid) {
        const bool upper = q >= 2;               // quarters 2,3 emit X[1024-k]
        const int ibase = (q == 0 || q == 3) ? 0 : 8;
        const int kq0 = q * 256;                 // staged bins: [kq0, kq0+256) (q=3: +1024)
        float2* c0 = CS + (warp * 2 + 0) * kCStride;
        float2* c1 = CS + (warp * 2 + 1) * kCStride;
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
          // register arrays need compile-time indices: select by (uniform) ibase
          const float2 ar = (ibase == 0) ? re[ii] : re[ii + 8];
          const float2 ai = (ibase == 0) ? im[ii] : im[ii + 8];
          const int i = ibase + ii;
          int ps = pshift - i;
          ps = ps > 15 ? 15 : ps;
          const float4 b = XB4r[ps * 32 + plane];
          const float2 br = make_float2(b.x, b.y), bi = make_float2(b.z, b.w);
          const int k = lane + 32 * i;
          const float2 t = __ldg(a.tw2 + k);
          const float2 e_r = padd(ar, br), e_i = psub(ai, bi);
          const float2 o_r = padd(ai, bi), o_i = psub(br, ar);
          float2 xr, xi;
          if (!upper) {  // X[k] = (E2 + T)/2
            xr = pfma(o_i, -t.y, pfma(o_r, t.x, e_r));
            xi = pfma(o_r, t.y, pfma(o_i, t.x, e_i));
          } else {       // X[1024-k] = conj(E2 - T)/2
            xr = pfma(o_i, t.y, pfma(o_r, -t.x, e_r));
            xi = psub(pfma(o_r, t.y, pmuls(o_i, t.x)), e_i);
          }
          xr = pmuls(xr, 0.5f);
          xi = pmuls(xi, 0.5f);
          const int kk = (upper ? 1024 - k : k) - kq0;   // 0..256
          if (!(lane == 0 && i == 0)) {
            c0[kk] = make_float2(xr.x, xi.x);
            c1[kk] = make_float2(xr.y, xi.y);
          }
        }
        if (lane == 0) {
          if (q == 0) {            // DC
            c0[0] = make_float2(re[0].x + im[0].x, 0.f);
            c1[0] = make_float2(re[0].y + im[0].y, 0.f);
          } else if (q == 3) {     // Nyquist (bin 1024 -> index 256)
            c0[256] = make_float2(re[0].x - im[0].x, 0.f);
            c1[256] = make_float2(re[0].y - im[0].y, 0.f);
          } else if (q == 2) {     // bin 512 = conj(Z[512]) (lane 0, slot 16)
            c0[0] = make_float2(re[16].x, -im[16].x);
            c1[0] = make_float2(re[16].y, -im[16].y);
          }
        }
      }
      __syncthreads();
      // bins staged this quarter: q=0: [0,256) q=1: [256,512) q=2: [512,768]->idx 0..256 minus... see below
      // q=2 stages bins 1024-k for k in [256,512) => (512,768] plus bin 512 => idx 0..256
      // q=3 stages bins 1024-k for k in [0,256)   => (768,1024] plus 1024   => idx 1..256
      const int lo = (q == 3) ? 1 : 0;
      const int hi = (q >= 2) ? 257 : 256;
      const int fsub = tid & 3, ksub = (tid >> 2) & 7, grp = tid >> 5;  // 4 groups of (8 bins x 4 frames)
      const int fr = f0 + fsub;
      for (int r = 0; r < 2; ++r) {
        if (rowbase + r >= a.rows) break;
        float2* obase = reinterpret_cast<float2*>(a.out) + (rowbase + r) * (long long)a.n_freq * a.n_frames;
        for (int kb = lo + grp * 8 + ksub; kb < hi; kb += 32) {
          if (fr < a.n_frames) {
            const float2 v = CS[(fsub * 2 + r) * kCStride + kb];
            obase[(long long)(q * 256 + kb) * a.n_frames + fr] = v;
          }
        }
      }
    }
    return;
  } else {
    // ---- power of every bin, kept in registers until the staging area is free ----------------
    float2 pk[16], pq[16];   // 4|X[k]|^2 and 4|X[1024-k]|^2 for the two rows
    float2 p512 = make_float2(0.f, 0.f);
    if (fvalid) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        int ps = pshift - i;
        ps = ps > 15 ? 15 : ps;
        const float4 b = XB4r[ps * 32 + plane];
        const float2 br = make_float2(b.x, b.y), bi = make_float2(b.z, b.w);
        const float2 ar = re[i], ai = im[i];
        const float2 t = __ldg(a.tw2 + lane + 32 * i);
        const float2 e_r = padd(ar, br), e_i = psub(ai, bi);
        const float2 o_r = padd(ai, bi), o_i = psub(br, ar);
        const float2 xr = pfma(o_i, -t.y, pfma(o_r, t.x, e_r));   // Re(E2 + T)
        const float2 xi = pfma(o_r, t.y, pfma(o_i, t.x, e_i));    // Im(E2 + T)
        const float2 yr = pfma(o_i, t.y, pfma(o_r, -t.x, e_r));   // Re(E2 - T)
        const float2 yi = pfma(o_r, -t.y, pfma(o_i, -t.x, e_i));  // Im(E2 - T)
        pk[i] = pfma2(xi, xi, pmul(xr, xr));
        pq[i] = pfma2(yi, yi, pmul(yr, yr));
      }
      if (lane == 0) {  // DC, Nyquist, bin 512
        const float2 dc = padd(re[0], im[0]), ny = psub(re[0], im[0]);
        pk[0] = pmuls(pmul(dc, dc), 4.f);
        pq[0] = pmuls(pmul(ny, ny), 4.f);
        p512 = pmuls(pfma2(im[16], im[16], pmul(re[16], re[16])), 4.f);
      }
    }
    __syncthreads();  // every warp is done with its exchange buffer: the staging area becomes P
    float* P = reinterpret_cast<float*>(stage);   // [8 (frame,row)][kPStride]
    if (fvalid) {
      float* p0 = P + (warp * 2 + 0) * kPStride;
      float* p1 = P + (warp * 2 + 1) * kPStride;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = lane + 32 * i;
        p0[k] = pk[i].x;
        p1[k] = pk[i].y;
        p0[1024 - k] = pq[i].x;
        p1[1024 - k] = pq[i].y;
      }
      if (lane == 0) {
        p0[512] = p512.x;
        p1[512] = p512.y;
        p0[1025] = p0[1026] = p0[1027] = 0.f;
        p1[1025] = p1[1026] = p1[1027] = 0.f;
      }
    }
    __syncthreads();

    if constexpr (MODE == MODE_MEL) {
      // lanes = 8 (frame,row) lines x 4 adjacent mel bins per warp; 16 bins per round
      const int line = tid & 7;                 // frame = line>>1, row = line&1
      const int fr = f0 + (line >> 1);
      const long long row = rowA + (line & 1);
      const bool ok = fr < a.n_frames && row < a.rows;
      const float4* P4 = reinterpret_cast<const float4*>(P + line * kPStride);
      for (int m = tid >> 3; m < a.n_mels; m += kThreads / 8) {
        const int start4 = __ldg(a.mel_start4 + m) >> 2, cnt = __ldg(a.mel_cnt4 + m);
        const float4* w4 = a.mel_w4 + __ldg(a.mel_off4 + m);
        float acc0 = 0.f, acc1 = 0.f;
        int j = 0;
        for (; j + 1 < cnt; j += 2) {
          const float4 p = P4[start4 + j], w = __ldg(w4 + j);
          const float4 p2 = P4[start4 + j + 1], w2 = __ldg(w4 + j + 1);
          acc0 = fmaf(p.x, w.x, acc0); acc1 = fmaf(p.y, w.y, acc1);
          acc0 = fmaf(p.z, w.z, acc0); acc1 = fmaf(p.w, w.w, acc1);
          acc0 = fmaf(p2.x, w2.x, acc0); acc1 = fmaf(p2.y, w2.y, acc1);
          acc0 = fmaf(p2.z, w2.z, acc0); acc1 = fmaf(p2.w, w2.w, acc1);
        }
        if (j < cnt) {
          const float4 p = P4[start4 + j], w = __ldg(w4 + j);
          acc0 = fmaf(p.x, w.x, acc0); acc1 = fmaf(p.y, w.y, acc1);
          acc0 = fmaf(p.z, w.z, acc0); acc1 = fmaf(p.w, w.w, acc1);
        }
        if (ok) a.out[(row * a.n_mels + m) * (long long)a.n_frames + fr] = acc0 + acc1;
      }
    } else {  // MODE_POWER: lanes = (8 bins) x (4 frames) -> 16-byte runs along the frame axis
      const int fsub = tid & 3, ksub = (tid >> 2) & 7, grp = tid >> 5;
      const int fr = f0 + fsub;
      if (fr < a.n_frames) {
        for (int r = 0; r < 2; ++r) {
          if (rowA + r >= a.rows) break;
          float* obase = a.out + (rowA + r) * (long long)a.n_freq * a.n_frames + fr;
          const float* pl = P + (fsub * 2 + r) * kPStride;
          for (int k = grp * 8 + ksub; k < 1025; k += 32) obase[(long long)k * a.n_frames] = 0.25f * pl[k];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Generic kernel: one CTA per (row, frame); M = n_fft/2 complex points in shared memory.
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128) stft_generic_kernel(const StftArgs a, int log2m) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int M = 1 << log2m, n_fft = 2 * M;
  float2* z = reinterpret_cast<float2*>(smem);               // [M]
  float* P = reinterpret_cast<float*>(smem + (size_t)M * 8);  // [M + 8]
  const int tid = threadIdx.x;
  const long long row = blockIdx.x / a.n_frames;
  const int frame = blockIdx.x % a.n_frames;
  const float* __restrict__ p = a.wav + row * a.n_in;
  const long long s0 = (long long)frame * a.hop - a.center_off;
  const bool center = a.center_off != 0;
  for (int n = tid; n < M; n += blockDim.x) {
    const float2 w = __ldg(a.window2 + n);
    const float x0 = fetch_sample(p, s0 + 2 * n, a.n_in, a.n_pad, center);
    const float x1 = fetch_sample(p, s0 + 2 * n + 1, a.n_in, a.n_pad, center);
    z[__brev((unsigned)n) >> (32 - log2m)] = make_float2(x0 * w.x, x1 * w.y);
  }
  __syncthreads();
  for (int s = 1; s <= log2m; ++s) {
    const int half = 1 << (s - 1);
    for (int b = tid; b < M / 2; b += blockDim.x) {
      const int j = b & (half - 1);
      const int i0 = ((b >> (s - 1)) << s) + j, i1 = i0 + half;
      const float2 w = __ldg(a.tw1 + ((size_t)j << (log2m - s)));
      const float2 u = z[i0], v = z[i1];
      const float tr = fmaf(v.x, w.x, -v.y * w.y), ti = fmaf(v.x, w.y, v.y * w.x);
      z[i0] = make_float2(u.x + tr, u.y + ti);
      z[i1] = make_float2(u.x - tr, u.y - ti);
    }
    __syncthreads();
  }
  // split: k in [0, M/2]; emits bins k and M-k (and DC/Nyquist for k = 0)
  const long long orow = row * (long long)a.n_freq;
  auto emit = [&](int k, float xr, float xi) {
    if (MODE == MODE_COMPLEX) {
      reinterpret_cast<float2*>(a.out)[(orow + k) * a.n_frames + frame] = make_float2(xr, xi);
    } else if (MODE == MODE_POWER) {
      a.out[(orow + k) * a.n_frames + frame] = xr * xr + xi * xi;
    } else {
      P[k] = 4.0f * (xr * xr + xi * xi);
    }
  };
  for (int k = tid; k <= M / 2; k += blockDim.x) {
    if (k == 0) {
      emit(0, z[0].x + z[0].y, 0.f);
      emit(M, z[0].x - z[0].y, 0.f);
    } else {
      const float2 u = z[k], v = z[M - k];
      const float2 t = __ldg(a.tw2 + k);
      const float er = u.x + v.x, ei = u.y - v.y, orr = u.y + v.y, oi = v.x - u.x;
      const float tr = fmaf(orr, t.x, -oi * t.y), ti = fmaf(orr, t.y, oi * t.x);
      emit(k, 0.5f * (er + tr), 0.5f * (ei + ti));
      if (k != M - k) emit(M - k, 0.5f * (er - tr), 0.5f * (ti - ei));
    }
  }
  if (MODE == MODE_MEL) {
    if (tid < 7) P[M + 1 + tid] = 0.f;
    __syncthreads();
    const float4* P4 = reinterpret_cast<const float4*>(P);
    for (int m = tid; m < a.n_mels; m += blockDim.x) {
      const int start4 = __ldg(a.mel_start4 + m) >> 2, cnt = __ldg(a.mel_cnt4 + m);
      const float4* w4 = a.mel_w4 + __ldg(a.mel_off4 + m);
      float acc = 0.f;
      for (int j = 0; j < cnt; ++j) {
        const float4 pv = P4[start4 + j], w = __ldg(w4 + j);
        acc = fmaf(pv.x, w.x, acc); acc = fmaf(pv.y, w.y, acc);
        acc = fmaf(pv.z, w.z, acc); acc = fmaf(pv.w, w.w, acc);
      }
      a.out[(row * a.n_mels + m) * (long long)a.n_frames + frame] = acc;
    }
  }
  (void)n_fft;
}

// MagDPhaseSpectrogramAE epilogue (given_models.py:214-231)
__global__ void magdphase_kernel(const float2* __restrict__ spec, long long cf, int n_frames, long long half_elems,
                                 float* __restrict__ out) {
  const float two_pi = 2.0f * 3.141592653589f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cf * n_frames;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % n_frames);
    const float2 v = spec[i];
    const float th = atan2f(v.y, v.x);
    float d = th;
    if (t > 0) {
      const float2 u = spec[i - 1];
      d = th - atan2f(u.y, u.x);
      if (d < 0.f) d += two_pi;
    }
    out[i] = hypotf(v.x, v.y);
    out[half_elems + i] = d;
  }
}

double hz_to_mel_htk(double f) { return 2595.0 * std::log10(1.0 + f / 700.0); }

}  // namespace

struct AaStftPlan {
  int n_fft = 0, hop = 0, center = 1, n_mels = 0, n_freq = 0, log2m = 0, device = 0;
  bool fast = false;
  float2* d_window2 = nullptr;
  float2* d_tw1 = nullptr;
  float2* d_tw2 = nullptr;
  int* d_mel_meta = nullptr;   // start4 | cnt4 | off4
  float4* d_mel_w4 = nullptr;
  // resources of aa_stft_mel_f32_host (created lazily)
  cudaStream_t hstream[2] = {nullptr, nullptr};
  float* hbuf_in[2] = {nullptr, nullptr};
  float* hbuf_out[2] = {nullptr, nullptr};
  long long hbuf_rows = 0, hbuf_nin = 0, hbuf_out_per_row = 0;
};

extern "C" {
#pragma GCC visibility push(default)

int aa_stft_plan_create(AaStftPlan** plan_out, int n_fft, int hop, int center, const float* window_host,
                        int n_mels, float sample_rate, float f_min, float f_max, const float* fb_host) {
  AA_REQUIRE(plan_out != nullptr, "plan is NULL");
  AA_REQUIRE(n_fft >= 64 && n_fft <= 8192 && (n_fft & (n_fft - 1)) == 0,
             "n_fft=%d: only powers of two in [64, 8192] are supported", n_fft);
  AA_REQUIRE(hop >= 1, "hop=%d must be >= 1", hop);
  AA_REQUIRE(n_mels >= 0 && n_mels <= 4096, "n_mels=%d out of range", n_mels);
  int rc = aa_check_device();
  if (rc != AA_OK) return rc;
  AaStftPlan* p = new AaStftPlan();
  p->n_fft = n_fft; p->hop = hop; p->center = center ? 1 : 0; p->n_mels = n_mels; p->n_freq = n_fft / 2 + 1;
  const int M = n_fft / 2;
  while ((1 << p->log2m) < M) p->log2m++;
  p->fast = (n_fft == 2048) && (hop % 4 == 0) && (hop <= 1024);
  AA_CUDA(cudaGetDevice(&p->device));
  const double PI = 3.14159265358979323846;
  // window as (w[2n], w[2n+1]) pairs
  std::vector<float> win(n_fft);
  for (int i = 0; i < n_fft; ++i)
    win[i] = window_host ? window_host[i] : (float)(0.5 - 0.5 * std::cos(2.0 * PI * i / n_fft));
  AA_CUDA(cudaMalloc(&p->d_window2, sizeof(float) * n_fft));
  AA_CUDA(cudaMemcpy(p->d_window2, win.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
  // twiddles
  std::vector<float2> tw1, tw2(std::max(n_fft / 4, 1) + 1);
  if (p->fast) {
    tw1.resize(32 * 32);
    for (int k1 = 0; k1 < 32; ++k1)
      for (int n2 = 0; n2 < 32; ++n2) {
        const double th = -2.0 * PI * (double)(k1 * n2) / 1024.0;
        tw1[k1 * 32 + n2] = make_float2((float)std::cos(th), (float)std::sin(th));
      }
  } else {
    tw1.resize(std::max(M / 2, 1));
    for (int j = 0; j < (int)tw1.size(); ++j) {
      const double th = -2.0 * PI * (double)j / (double)M;
      tw1[j] = make_float2((float)std::cos(th), (float)std::sin(th));
    }
  }
  for (int k = 0; k < (int)tw2.size(); ++k) {
    const double th = -2.0 * PI * (double)k / (double)n_fft;
    tw2[k] = make_float2((float)std::cos(th), (float)std::sin(th));
  }
  AA_CUDA(cudaMalloc(&p->d_tw1, sizeof(float2) * tw1.size()));
  AA_CUDA(cudaMemcpy(p->d_tw1, tw1.data(), sizeof(float2) * tw1.size(), cudaMemcpyHostToDevice));
  AA_CUDA(cudaMalloc(&p->d_tw2, sizeof(float2) * tw2.size()));
  AA_CUDA(cudaMemcpy(p->d_tw2, tw2.data(), sizeof(float2) * tw2.size(), cudaMemcpyHostToDevice));
  // mel filterbank -> per-filter (start4, cnt4, off4) + float4-aligned weights (x 0.25)
  if (n_mels > 0) {
    const int F = p->n_freq;
    std::vector<float> fb((size_t)F * n_mels);
    if (fb_host) {
      std::copy(fb_host, fb_host + (size_t)F * n_mels, fb.begin());
    } else {  // torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk")
      AA_REQUIRE(sample_rate > 0 && f_max > f_min, "bad mel parameters sr=%f f_min=%f f_max=%f", sample_rate, f_min, f_max);
      const double m_min = hz_to_mel_htk(f_min), m_max = hz_to_mel_htk(f_max);
      std::vector<double> f_pts(n_mels + 2);
      for (int i = 0; i < n_mels + 2; ++i) {
        const double m = m_min + (m_max - m_min) * (double)i / (double)(n_mels + 1);
        f_pts[i] = 700.0 * (std::pow(10.0, m / 2595.0) - 1.0);
      }
      const double nyq = (double)((long long)sample_rate / 2);
      for (int k = 0; k < F; ++k) {
        const double f = nyq * (double)k / (double)(F - 1);
        for (int m = 0; m < n_mels; ++m) {
          const double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
          const double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
          fb[(size_t)k * n_mels + m] = (float)std::max(0.0, std::min(down, up));
        }
      }
    }
    std::vector<int> meta(3 * (size_t)n_mels);
    std::vector<float> w;
    for (int m = 0; m < n_mels; ++m) {
      int lo = -1, hi = -1;
      for (int k = 0; k < F; ++k)
        if (fb[(size_t)k * n_mels + m] != 0.0f) { if (lo < 0) lo = k; hi = k; }
      if (lo < 0) { meta[m] = 0; meta[n_mels + m] = 0; meta[2 * n_mels + m] = 0; continue; }
      const int s4 = lo & ~3, cnt = (hi - s4) / 4 + 1;
      meta[m] = s4; meta[n_mels + m] = cnt; meta[2 * n_mels + m] = (int)(w.size() / 4);
      for (int k = s4; k < s4 + 4 * cnt; ++k) w.push_back(k < F ? 0.25f * fb[(size_t)k * n_mels + m] : 0.0f);
    }
    if (w.empty()) w.resize(4, 0.f);
    AA_CUDA(cudaMalloc(&p->d_mel_meta, sizeof(int) * meta.size()));
    AA_CUDA(cudaMemcpy(p->d_mel_meta, meta.data(), sizeof(int) * meta.size(), cudaMemcpyHostToDevice));
    AA_CUDA(cudaMalloc(&p->d_mel_w4, sizeof(float) * w.size()));
    AA_CUDA(cudaMemcpy(p->d_mel_w4, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
  }
  if (p->fast) {
    const int smem = (3 * hop + 2048) * 8 + kStageBytes + 16;
    AA_CUDA(cudaFuncSetAttribute(stft2048_kernel<MODE_COMPLEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AA_CUDA(cudaFuncSetAttribute(stft2048_kernel<MODE_POWER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AA_CUDA(cudaFuncSetAttribute(stft2048_kernel<MODE_MEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  } else if (M * 12 + 64 > 48 * 1024) {
    const int smem = M * 12 + 64;
    AA_CUDA(cudaFuncSetAttribute(stft_generic_kernel<MODE_COMPLEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AA_CUDA(cudaFuncSetAttribute(stft_generic_kernel<MODE_POWER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AA_CUDA(cudaFuncSetAttribute(stft_generic_kernel<MODE_MEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  *plan_out = p;
  return AA_OK;
}

int aa_stft_plan_destroy(AaStftPlan* p) {
  if (!p) return AA_OK;
  cudaFree(p->d_window2); cudaFree(p->d_tw1); cudaFree(p->d_tw2); cudaFree(p->d_mel_meta); cudaFree(p->d_mel_w4);
  for (int i = 0; i < 2; ++i) {
    if (p->hbuf_in[i]) cudaFree(p->hbuf_in[i]);
    if (p->hbuf_out[i]) cudaFree(p->hbuf_out[i]);
    if (p->hstream[i]) cudaStreamDestroy(p->hstream[i]);
  }
  delete p;
  return AA_OK;
}

int aa_stft_out_shape(const AaStftPlan* p, int64_t n_in, int zero_pad, int64_t* n_pad_out, int64_t* n_frames_out) {
  AA_REQUIRE(p != nullptr, "plan is NULL");
  AA_REQUIRE(n_in >= 1, "n_in=%lld must be >= 1", (long long)n_in);
  int64_t n_pad = n_in;
  if (zero_pad) { n_pad = 1; while (n_pad < n_in) n_pad <<= 1; }
  int64_t frames;
  if (p->center) {
    AA_REQUIRE(n_pad > p->n_fft / 2, "reflect padding needs more than n_fft/2=%d samples, got %lld", p->n_fft / 2, (long long)n_pad);
    frames = 1 + n_pad / p->hop;
  } else {
    AA_REQUIRE(n_pad >= p->n_fft, "center=False needs at least n_fft=%d samples, got %lld", p->n_fft, (long long)n_pad);
    frames = 1 + (n_pad - p->n_fft) / p->hop;
  }
  if (n_pad_out) *n_pad_out = n_pad;
  if (n_frames_out) *n_frames_out = frames;
  return AA_OK;
}

static int stft_launch(const AaStftPlan* p, int mode, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                       float* out, cudaStream_t st) {
  AA_REQUIRE(p != nullptr, "plan is NULL");
  AA_REQUIRE(rows >= 0, "rows=%lld", (long long)rows);
  if (rows == 0) return AA_OK;
  AA_REQUIRE(wav != nullptr && out != nullptr, "NULL tensor pointer");
  AA_REQUIRE(mode != MODE_MEL || p->n_mels > 0, "plan was created without a mel stage");
  int64_t n_pad = 0, n_frames = 0;
  int rc = aa_stft_out_shape(p, n_in, zero_pad, &n_pad, &n_frames);
  if (rc != AA_OK) return rc;
  AA_REQUIRE(n_frames < (1LL << 30), "too many frames");
  StftArgs a;
  a.wav = wav; a.rows = rows; a.n_in = n_in; a.n_pad = n_pad; a.n_frames = (int)n_frames; a.hop = p->hop;
  a.center_off = p->center ? p->n_fft / 2 : 0;
  a.n_freq = p->n_freq; a.n_mels = p->n_mels;
  a.window2 = p->d_window2; a.tw1 = p->d_tw1; a.tw2 = p->d_tw2;
  a.mel_start4 = p->d_mel_meta; a.mel_cnt4 = p->d_mel_meta ? p->d_mel_meta + p->n_mels : nullptr;
  a.mel_off4 = p->d_mel_meta ? p->d_mel_meta + 2 * p->n_mels : nullptr;
  a.mel_w4 = p->d_mel_w4; a.out = out;
  a.wav_aligned16 = ((reinterpret_cast<uintptr_t>(wav) & 15) == 0) ? 1 : 0;
  if (p->fast) {
    a.tiles_per_pair = (int)((n_frames + kWarps - 1) / kWarps);
    const int64_t pairs = (rows + 1) / 2, grid = pairs * a.tiles_per_pair;
    AA_REQUIRE(grid < (1LL << 31), "grid too large (%lld tiles)", (long long)grid);
    const int smem = (3 * p->hop + 2048) * 8 + kStageBytes + 16;
    if (mode == MODE_COMPLEX) stft2048_kernel<MODE_COMPLEX><<<(unsigned)grid, kThreads, smem, st>>>(a);
    else if (mode == MODE_POWER) stft2048_kernel<MODE_POWER><<<(unsigned)grid, kThreads, smem, st>>>(a);
    else stft2048_kernel<MODE_MEL><<<(unsigned)grid, kThreads, smem, st>>>(a);
  } else {
    a.tiles_per_pair = 0;
    const int64_t grid = rows * n_frames;
    AA_REQUIRE(grid < (1LL << 31), "grid too large (%lld frames)", (long long)grid);
    const int M = p->n_fft / 2, smem = M * 12 + 64;
    if (mode == MODE_COMPLEX) stft_generic_kernel<MODE_COMPLEX><<<(unsigned)grid, 128, smem, st>>>(a, p->log2m);
    else if (mode == MODE_POWER) stft_generic_kernel<MODE_POWER><<<(unsigned)grid, 128, smem, st>>>(a, p->log2m);
    else stft_generic_kernel<MODE_MEL><<<(unsigned)grid, 128, smem, st>>>(a, p->log2m);
  }
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_stft_complex_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                        float* out, void* stream) {
  return stft_launch(plan, MODE_COMPLEX, wav, rows, n_in, zero_pad, out, (cudaStream_t)stream);
}
int aa_stft_power_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                      float* out, void* stream) {
  return stft_launch(plan, MODE_POWER, wav, rows, n_in, zero_pad, out, (cudaStream_t)stream);
}
int aa_stft_mel_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                    float* out, void* stream) {
  return stft_launch(plan, MODE_MEL, wav, rows, n_in, zero_pad, out, (cudaStream_t)stream);
}

int aa_magdphase_f32(const float* spec, int64_t c, int64_t n_freq, int64_t n_frames, float* out, void* stream) {
  AA_REQUIRE(spec && out, "NULL tensor pointer");
  AA_REQUIRE(c >= 1 && n_freq >= 1 && n_frames >= 1 && n_frames < (1LL << 31), "bad shape");
  const long long cf = c * n_freq, total = cf * n_frames;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)aa::num_sms() * 8);
  magdphase_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(spec), cf, (int)n_frames, total, out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_stft_mel_f32_host(const AaStftPlan* plan_c, const float* wav_host, int64_t rows, int64_t n_in, int zero_pad,
                         float* out_host, int64_t rows_per_chunk) {
  AaStftPlan* p = const_cast<AaStftPlan*>(plan_c);
  AA_REQUIRE(p != nullptr && wav_host && out_host, "NULL argument");
  AA_REQUIRE(p->n_mels > 0, "plan was created without a mel stage");
  if (rows == 0) return AA_OK;
  int64_t n_pad = 0, n_frames = 0;
  int rc = aa_stft_out_shape(p, n_in, zero_pad, &n_pad, &n_frames);
  if (rc != AA_OK) return rc;
  if (rows_per_chunk <= 0) rows_per_chunk = 64;
  rows_per_chunk = (rows_per_chunk + 1) & ~1LL;  // keep row pairs inside a chunk
  const long long out_per_row = (long long)p->n_mels * n_frames;
  if (p->hbuf_rows < rows_per_chunk || p->hbuf_nin != n_in || p->hbuf_out_per_row != out_per_row) {
    for (int i = 0; i < 2; ++i) {
      if (p->hbuf_in[i]) cudaFree(p->hbuf_in[i]);
      if (p->hbuf_out[i]) cudaFree(p->hbuf_out[i]);
      p->hbuf_in[i] = p->hbuf_out[i] = nullptr;
      if (!p->hstream[i]) AA_CUDA(cudaStreamCreateWithFlags(&p->hstream[i], cudaStreamNonBlocking));
      AA_CUDA(cudaMalloc(&p->hbuf_in[i], sizeof(float) * rows_per_chunk * n_in));
      AA_CUDA(cudaMalloc(&p->hbuf_out[i], sizeof(float) * rows_per_chunk * out_per_row));
    }
    p->hbuf_rows = rows_per_chunk; p->hbuf_nin = n_in; p->hbuf_out_per_row = out_per_row;
  }
  int slot = 0;
  for (int64_t r0 = 0; r0 < rows; r0 += rows_per_chunk, slot ^= 1) {
    const int64_t nr = std::min<int64_t>(rows_per_chunk, rows - r0);
    cudaStream_t st = p->hstream[slot];
    AA_CUDA(cudaMemcpyAsync(p->hbuf_in[slot], wav_host + r0 * n_in, sizeof(float) * nr * n_in, cudaMemcpyHostToDevice, st));
    rc = stft_launch(p, MODE_MEL, p->hbuf_in[slot], nr, n_in, zero_pad, p->hbuf_out[slot], st);
    if (rc != AA_OK) return rc;
    AA_CUDA(cudaMemcpyAsync(out_host + r0 * out_per_row, p->hbuf_out[slot], sizeof(float) * nr * out_per_row,
                            cudaMemcpyDeviceToHost, st));
  }
  AA_CUDA(cudaStreamSynchronize(p->hstream[0]));
  AA_CUDA(cudaStreamSynchronize(p->hstream[1]));
  return AA_OK;
}

#pragma GCC visibility pop
}  // extern "C"
