// STFT / power / mel front-end for sm_100a.
//
// Replaces torchaudio.transforms.{Spectrogram,MelSpectrogram} as used by the reference's
// SpectrogramAE / MagSpectrogramAE / MelSpectrogramAE.encode (audio_algebra/given_models.py:149-283)
// and GivenModelClass.zero_pad_po2 (:139-145).
//
// Fast path (n_fft = 2048, hop % 4 == 0, hop <= 1024) -- `stft2048_kernel`:
//   * persistent CTAs (2 per SM) of 6 warps walk over tiles of 6 consecutive frames of one ROW PAIR;
//     the two rows ride in the two halves of packed fp32x2 registers, so the FFT arithmetic is
//     FADD2/FFMA2 (one issue slot per 2 flops); twiddle / window tables live in shared memory;
//   * the tile's samples are brought into shared memory once per row by a TMA bulk copy
//     (cp.async.bulk + mbarrier; reflect padding and the zero_pad_po2 tail are handled in index
//     math on the few tiles at the chunk edges), so the 75 % frame overlap costs no extra HBM
//     traffic inside a tile; the next tile's copy is issued as soon as the current tile's samples
//     are in registers and overlaps with the FFT;
//   * a real 2048-point FFT is one 1024-point complex FFT: lane n2 holds z[32*n1+n2] (32 packed
//     complex values), runs a 32-point in-register FFT, multiplies by W_1024^(n2*k1), transposes
//     through a warp-private shared-memory buffer, runs the second 32-point FFT (lane k1 then holds
//     Z[k1+32*k2]), and the even/odd split pairs lane k1 with lane 32-k1 through the same buffer;
//   * epilogues: |X|^2 -> sparse (<= 2 taps per bin) mel projection from shared memory (quarter-warp =
//     one filter x the tile's 6 frames, packed (rowA,rowB) FFMA2, conflict-free skewed P lines), or
//     staged power / complex stores that write 24/48-byte runs along the frame axis.
// Warp path (n_fft = 64 .. 1024, any hop / window: the reference defaults n_fft = 1024, hop = 256) -- `stft_warp_kernel`:
//   one warp per frame of a row pair, R-point FFT in registers x 32-point FFT across lanes by shuffles.
// Generic path (any other power-of-two n_fft up to 8192, or n_fft = 2048 with a non-Hann window / odd hop) --
//   `stft_generic_kernel`: one CTA per (row, frame), shared-memory radix-2 FFT; correct for every configuration, not tuned.
#include "aa_common.cuh"
#include "fft_gen.cuh"
#include "stft_consts.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
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
  int out_tf;           // complex / power output layout: 0 = [row][freq][frame], 1 = [row][frame][freq] (torch.stft's own memory layout)
  int wav_aligned16;
  long long n_tiles;      // fast path: row pairs x tiles_per_pair
  const int4* mel_steps;  // fast path: [kWarps][mel_steps_per_warp + 1][4] step headers, then the step weights
  int mel_steps_per_warp, mel_hdr_bytes, mel_w_bytes;
  const float* lane_consts;  // fast path: float4[32] Hann phases + float2[32] W_2048^lane
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
// Fast kernel, n_fft = 2048: persistent CTAs of 6 warps; one warp = one frame of one row pair.
// ------------------------------------------------------------------------------------------
constexpr int kWarps = 6;                       // frames per tile
constexpr int kThreads = kWarps * 32;
constexpr int kXbBytes = 32 * 33 * 8 + 16;      // warp-private exchange buffer (float2 [32][33]) + 16 B skew:
                                                // consecutive frames' P lines start 4 banks apart
constexpr int kPLine = kXbBytes / 8;            // float2 stride between the P lines of consecutive frames (1058)
constexpr int kStageBytes = kWarps * kXbBytes;  // 50688
constexpr int kTableBytes = 8192 + 512 + 256;   // tw1 [32][32] float2, per-lane Hann phases float4[32], W_2048^lane float2[32]
constexpr int kCStride = 260;                   // complex-mode staging: [12 lines][260] float2

static_assert(kPLine >= 1028, "a P line (1025 bins + pad) must fit in one exchange buffer");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ int4 lds_i4(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
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

struct TileGeom {
  int rowA;         // first row of the pair
  int s0;           // first sample (signal coordinates, may be negative / past the end)
  int f0;
  int lo, hi;       // [lo, hi): part of the tile's span that lies inside the row -> TMA (if `tma`)
  bool hasB, tma;
};

// 32-bit on purpose (the host checks n_tiles, rows and n_pad fit).
__device__ __forceinline__ TileGeom tile_geom(const StftArgs& a, int t, int span) {
  TileGeom g;
  const int pair = t / a.tiles_per_pair;
  g.f0 = (t - pair * a.tiles_per_pair) * kWarps;
  g.rowA = 2 * pair;
  g.hasB = g.rowA + 1 < (int)a.rows;
  g.s0 = g.f0 * a.hop - a.center_off;
  g.lo = g.s0 < 0 ? -g.s0 : 0;
  g.hi = min(span, (int)a.n_in - g.s0);
  // bulk copies need 16-byte aligned source / destination / size: rows of n_in % 4 == 0 floats from an
  // aligned base (hop % 4 == 0 and the centre offset 1024 keep s0 a multiple of 4)
  g.tma = g.hasB && a.wav_aligned16 && ((int)a.n_in & 3) == 0 && g.hi - g.lo >= 4 && g.lo < span;
  return g;
}

__device__ __forceinline__ void issue_tma(float* SA, float* SB, const float* src, int lo, int cnt, long long n_in,
                                          uint64_t* mbar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(mbar, (uint32_t)cnt * 8u);
  bulk_g2s(SA + lo, src, (uint32_t)cnt * 4u, mbar);
  bulk_g2s(SB + lo, src + n_in, (uint32_t)cnt * 4u, mbar);
}

// Samples of the span outside [lo, hi) (chunk borders: reflect padding, zero_pad_po2 tail,
// given_models.py:139-145) -- or the whole span when the tile cannot use TMA (odd last row, unaligned rows).
__device__ __forceinline__ void load_tile_manual(const StftArgs& a, const TileGeom& g, int span, float* SA, float* SB,
                                                 int tid) {
  const float* __restrict__ pa = a.wav + (size_t)g.rowA * (size_t)a.n_in;
  const int n_in = (int)a.n_in, n_pad = (int)a.n_pad;
  const bool center = a.center_off != 0;
  const int skip_lo = g.tma ? g.lo : 0, skip_hi = g.tma ? g.hi : 0;   // [skip_lo, skip_hi) is covered by TMA
  const int n_out = span - (skip_hi - skip_lo);
  for (int jj = tid; jj < n_out; jj += kThreads) {
    const int j = jj < skip_lo ? jj : jj + (skip_hi - skip_lo);
    int i = g.s0 + j;
    if (center) {
      if (i < 0) i = -i;
      if (i >= n_pad) i = 2 * (n_pad - 1) - i;
    }
    const bool ok = i >= 0 && i < n_in;
    SA[j] = ok ? __ldg(pa + i) : 0.f;
    SB[j] = (ok && g.hasB) ? __ldg(pa + n_in + i) : 0.f;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) stft2048_kernel(const StftArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hop = a.hop;
  const int span = (kWarps - 1) * hop + 2048;   // samples covered by the frames of one tile
  float* SA = reinterpret_cast<float*>(smem);                        // row A samples [span]
  float* SB = SA + span;                                             // row B samples [span]
  unsigned char* stage = smem + (size_t)span * 8;                    // kStageBytes
  float2* XB = reinterpret_cast<float2*>(stage + warp * kXbBytes);   // this warp's exchange buffer
  float2* s_tw1 = reinterpret_cast<float2*>(stage + kStageBytes);    // [32][32] W_1024^(k1 n2)
  float4* s_lane = reinterpret_cast<float4*>(s_tw1 + 1024);          // [32] (cos,sin) of 2pi(2 lane + {0,1})/2048
  float2* s_tw2l = reinterpret_cast<float2*>(s_lane + 32);           // [32] W_2048^lane
  int4* s_melh = reinterpret_cast<int4*>(s_tw2l + 32);               // mel step headers [kWarps][steps + 1][4 quarter-warps]
  float4* s_melw = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(s_melh) + a.mel_hdr_bytes);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_melw) + a.mel_w_bytes);
  constexpr bool kPrefetch = (MODE != MODE_COMPLEX);  // complex mode reuses the sample area as staging

  for (int i = tid; i < 1024; i += kThreads) s_tw1[i] = __ldg(a.tw1 + i);
  if (tid < 48) reinterpret_cast<float4*>(s_lane)[tid] = __ldg(reinterpret_cast<const float4*>(a.lane_consts) + tid);
  for (int i = tid; i < (a.mel_hdr_bytes + a.mel_w_bytes) / 16; i += kThreads)
    s_melh[i] = __ldg(a.mel_steps + i);   // headers and weights are one contiguous device block
  if (tid == 0) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t phase = 0;
  const int n_tiles = (int)a.n_tiles;
  if ((int)blockIdx.x >= n_tiles) return;
  if (tid == 0) {
    const TileGeom g0 = tile_geom(a, blockIdx.x, span);
    if (g0.tma)
      issue_tma(SA, SB, a.wav + (size_t)g0.rowA * (size_t)a.n_in + (g0.s0 + g0.lo), g0.lo, g0.hi - g0.lo, a.n_in, mbar);
  }

#pragma unroll 1
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    // ---- this tile's samples: the in-row part was prefetched by TMA; the reflected / zero-padded rest of
    // an edge tile (or a whole tile that cannot use TMA) is filled in here --------------------------------
    const TileGeom g = tile_geom(a, t, span);
    if (!g.tma || g.lo > 0 || g.hi < span) {
      load_tile_manual(a, g, span, SA, SB, tid);
      __syncthreads();
    }
    if (g.tma) {
      mbar_wait(mbar, phase);
      phase ^= 1;
    }
    if (kPrefetch && tid == 0) {   // decide the next prefetch now, while registers are free; park it in smem
      const int tn = t + (int)gridDim.x;
      unsigned long long src = 0ull, rng = 0ull;
      if (tn < n_tiles) {
        const TileGeom gn = tile_geom(a, tn, span);
        if (gn.tma) {
          src = reinterpret_cast<unsigned long long>(a.wav + (size_t)gn.rowA * (size_t)a.n_in + (gn.s0 + gn.lo));
          rng = (unsigned long long)(unsigned)gn.lo | ((unsigned long long)(unsigned)(gn.hi - gn.lo) << 32);
        }
      }
      mbar[1] = src;
      mbar[2] = rng;
    }
    const int f0 = g.f0;
    const long long rowA = g.rowA;   // 64-bit for output addressing
    const int frame = f0 + warp;
    const bool fvalid = frame < a.n_frames;

    float2 re[32], im[32];
    if (fvalid) {
      // ---- load + window: z[n] = x[2n] w[2n] + i x[2n+1] w[2n+1], n = 32*n1 + lane ------------
      // periodic Hann on the fly: w[j] = 0.5 - 0.5 cos(2 pi n1/32 + phi), phi = 2 pi (2 lane + {0,1})/2048
      const float2* FA = reinterpret_cast<const float2*>(SA + warp * hop);
      const float2* FB = reinterpret_cast<const float2*>(SB + warp * hop);
      const float4 ph = s_lane[lane];   // (cos phi0, sin phi0, cos phi1, sin phi1)
#pragma unroll
      for (int n1 = 0; n1 < 32; ++n1) {
        const int n = 32 * n1 + lane;
        const float2 xa = FA[n], xb = FB[n];
        const float w0 = fmaf(0.5f * aa_consts::kSin32[n1], ph.y, fmaf(-0.5f * aa_consts::kCos32[n1], ph.x, 0.5f));
        const float w1 = fmaf(0.5f * aa_consts::kSin32[n1], ph.w, fmaf(-0.5f * aa_consts::kCos32[n1], ph.z, 0.5f));
        re[bitrev5(n1)] = make_float2(xa.x * w0, xb.x * w0);
        im[bitrev5(n1)] = make_float2(xa.y * w1, xb.y * w1);
      }
    }
    __syncthreads();  // samples consumed by every warp; previous tile's epilogue has left the staging area
    if (kPrefetch && tid == 0) {   // next interior tile: TMA overlaps with the FFT below
      const float* src = reinterpret_cast<const float*>(mbar[1]);
      const unsigned long long rng = mbar[2];
      if (src != nullptr) issue_tma(SA, SB, src, (int)(rng & 0xffffffffu), (int)(rng >> 32), a.n_in, mbar);
    }

    if (fvalid) {
      fft32_dit(re, im);  // slot k1: sum_n1 z[32 n1 + lane] W_32^(n1 k1)
      // ---- twiddle W_1024^(lane*k1) ------------------------------------------------------------
#pragma unroll
      for (int k1 = 1; k1 < 32; ++k1) {
        const float2 tw = s_tw1[k1 * 32 + lane];
        const float2 r = re[k1], i = im[k1];
        re[k1] = pfma(i, -tw.y, pmuls(r, tw.x));
        im[k1] = pfma(i, tw.x, pmuls(r, tw.y));
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
      // Four k-quarters, each staged as [12 (frame,row)][260] complex in the (now dead) sample area and
      // stored with lanes = (32 bins) x (6 frames) so that every (row, bin) gets one 48-byte run.
      float2* CS = reinterpret_cast<float2*>(smem);
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        if (q > 0) __syncthreads();  // previous quarter's staging has been stored
        if (fvalid) {
          const bool upper = q >= 2;               // quarters 2,3 emit X[1024-k]
          const int ibase = (q == 0 || q == 3) ? 0 : 8;
          const int kq0 = q * 256;
          float2* c0 = CS + (warp * 2 + 0) * kCStride;
          float2* c1 = CS + (warp * 2 + 1) * kCStride;
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {
            const float2 ar = (ibase == 0) ? re[ii] : re[ii + 8];
            const float2 ai = (ibase == 0) ? im[ii] : im[ii + 8];
            const int i = ibase + ii;
            int ps = pshift - i;
            ps = ps > 15 ? 15 : ps;
            const float4 b = XB4r[ps * 32 + plane];
            const float2 br = make_float2(b.x, b.y), bi = make_float2(b.z, b.w);
            const int k = lane + 32 * i;
            const float2 cl = s_tw2l[lane];
            const float c64 = (ibase == 0) ? aa_consts::kCos64[ii] : aa_consts::kCos64[ii + 8];
            const float s64 = (ibase == 0) ? aa_consts::kSin64[ii] : aa_consts::kSin64[ii + 8];
            const float2 tw = make_float2(fmaf(cl.y, s64, cl.x * c64), fmaf(cl.y, c64, -cl.x * s64));
            const float2 e_r = padd(ar, br), e_i = psub(ai, bi);
            const float2 o_r = padd(ai, bi), o_i = psub(br, ar);
            float2 xr, xi;
            if (!upper) {  // X[k] = (E2 + T)/2
              xr = pfma(o_i, -tw.y, pfma(o_r, tw.x, e_r));
              xi = pfma(o_r, tw.y, pfma(o_i, tw.x, e_i));
            } else {       // X[1024-k] = conj(E2 - T)/2
              xr = pfma(o_i, tw.y, pfma(o_r, -tw.x, e_r));
              xi = psub(pfma(o_r, tw.y, pmuls(o_i, tw.x)), e_i);
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
        // q=0: bins [0,256)  q=1: [256,512)  q=2: [512,768] (idx 0..256)  q=3: (768,1024] (idx 1..256)
        const int lo = (q == 3) ? 1 : 0;
        const int hi = (q >= 2) ? 257 : 256;
        if (a.out_tf) {
          // frequency-minor output: warp = frame of the tile, lanes run along the bins -> 256-byte runs per store
          const int fr = f0 + warp;
          if (fr < a.n_frames) {
            for (int r = 0; r < 2; ++r) {
              if (rowA + r >= a.rows) break;
              float2* obase = reinterpret_cast<float2*>(a.out) + ((rowA + r) * (long long)a.n_frames + fr) * a.n_freq + q * 256;
              const float2* cl = CS + (warp * 2 + r) * kCStride;
              for (int kb = lo + lane; kb < hi; kb += 32) obase[kb] = cl[kb];
            }
          }
          continue;
        }
        const int fsub = tid % kWarps, ksub = tid / kWarps;   // 32 bins x 6 frames per pass
        const int fr = f0 + fsub;
        if (fr < a.n_frames) {
          for (int r = 0; r < 2; ++r) {
            if (rowA + r >= a.rows) break;
            float2* obase = reinterpret_cast<float2*>(a.out) + (rowA + r) * (long long)a.n_freq * a.n_frames + fr;
            const float2* cl = CS + (fsub * 2 + r) * kCStride;
            for (int kb = lo + ksub; kb < hi; kb += 32) obase[(long long)(q * 256 + kb) * a.n_frames] = cl[kb];
          }
        }
      }
      __syncthreads();  // staging (= sample area) free again
      if (t + (int)gridDim.x < n_tiles && tid == 0) {
        const TileGeom gn = tile_geom(a, t + gridDim.x, span);
        if (gn.tma)
          issue_tma(SA, SB, a.wav + (size_t)gn.rowA * (size_t)a.n_in + (gn.s0 + gn.lo), gn.lo, gn.hi - gn.lo, a.n_in, mbar);
      }
    } else {
      // ---- power of every bin.  The P line of this frame (float2 = (rowA,rowB) per bin, 4|X|^2) is
      // written into this warp's OWN exchange buffer: 4|X[1024-k]|^2 right away (those addresses are
      // already consumed), 4|X[k]|^2 after the last partner read. ---------------------------------
      float2* pl = reinterpret_cast<float2*>(stage + warp * kXbBytes);
      if (fvalid) {
        float2 pk[16];
        const float2 z0r = re[0], z0i = im[0], z16r = re[16], z16i = im[16];
        const float2 cl = s_tw2l[lane];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          int ps = pshift - i;
          ps = ps > 15 ? 15 : ps;
          const float4 b = XB4r[ps * 32 + plane];
          __syncwarp();   // all lanes have read slot row `ps` before anyone overwrites it below
          const float2 br = make_float2(b.x, b.y), bi = make_float2(b.z, b.w);
          const float2 ar = re[i], ai = im[i];
          // W_2048^(lane + 32 i) = W_2048^lane * W_64^i
          const float2 tw = make_float2(fmaf(cl.y, aa_consts::kSin64[i], cl.x * aa_consts::kCos64[i]),
                                        fmaf(cl.y, aa_consts::kCos64[i], -cl.x * aa_consts::kSin64[i]));
          const float2 e_r = padd(ar, br), e_i = psub(ai, bi);
          const float2 o_r = padd(ai, bi), o_i = psub(br, ar);
          const float2 xr = pfma(o_i, -tw.y, pfma(o_r, tw.x, e_r));   // Re(E2 + T)
          const float2 xi = pfma(o_r, tw.y, pfma(o_i, tw.x, e_i));    // Im(E2 + T)
          const float2 yr = pfma(o_i, tw.y, pfma(o_r, -tw.x, e_r));   // Re(E2 - T)
          const float2 yi = pfma(o_r, -tw.y, pfma(o_i, -tw.x, e_i));  // Im(E2 - T)
          pk[i] = pfma2(xi, xi, pmul(xr, xr));
          const float2 pq = pfma2(yi, yi, pmul(yr, yr));
          // bins 1024-k >= 513: byte offsets >= 4104 - 256 i, above every partner row still to be read
          if (!(lane == 0 && i == 0)) pl[1024 - (lane + 32 * i)] = pq;
          // partner rows still unread after this iteration cover bytes [0, 512 (16-i)) (lane 0 runs one
          // row behind): pk[j] (bytes [256 j, 256 j + 256)) may go out once j >= 32 - 2 i
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j <= i && j >= 32 - 2 * i && (j == i || j < 34 - 2 * i)) pl[lane + 32 * j] = pk[j];
        }
        __syncwarp();
        pl[lane] = pk[0];
        pl[lane + 32] = pk[1];
        __syncwarp();
        if (lane == 0) {  // DC, Nyquist, bin 512, zero tail of the last float4 group
          const float2 dc = padd(z0r, z0i), ny = psub(z0r, z0i);
          pl[0] = pmuls(pmul(dc, dc), 4.f);
          pl[1024] = pmuls(pmul(ny, ny), 4.f);
          pl[512] = pmuls(pfma2(z16i, z16i, pmul(z16r, z16r)), 4.f);
          pl[1025] = pl[1026] = pl[1027] = make_float2(0.f, 0.f);
        }
      }
      __syncthreads();
      const float2* P = reinterpret_cast<const float2*>(stage);   // frame f at P + f * kPLine

      if constexpr (MODE == MODE_MEL) {
        // Each warp walks its own list of steps; a step = 4 adjacent mel filters (one per quarter-warp),
        // lanes of a quarter-warp = the 6 frames of the tile (2 lanes idle).  A quarter-warp reads 16 B from
        // 6 different P lines (skewed by 16 B) -> conflict-free; step headers and weights (one float4 per 4
        // bins, zero-padded to the longest filter of the step) live in shared memory; accumulators are
        // packed (rowA,rowB) FFMA2.
        const int q = lane >> 3, fl = lane & 7;
        const bool factive = fl < kWarps;
        // 32-bit output indexing (the host routes outputs of >= 2^31 elements to the generic kernel)
        unsigned plane_sz = (unsigned)a.n_mels * (unsigned)a.n_frames;
        unsigned n_frames_u = (unsigned)a.n_frames;
        float* outA = a.out + (size_t)rowA * plane_sz + (f0 + fl);
        // bit 0: this lane stores row A, bit 1: row B exists
        unsigned flags = ((factive && (f0 + fl < a.n_frames)) ? 1u : 0u) | ((rowA + 1 < a.rows) ? 2u : 0u);
        // shared-window addresses of this lane's P line, the step weights and this quarter-warp's header list:
        // per-(warp, step, quarter) header {P offset (bytes), n_iter / 2, weight offset (bytes), filter or -1};
        // the list of a warp ends with an n_iter = 0 entry
        unsigned pf = smem_u32(P + (factive ? fl : 0) * kPLine);
        unsigned wbase = smem_u32(s_melw);
        unsigned hp = smem_u32(s_melh + warp * (a.mel_steps_per_warp + 1) * 4 + q);
        // keep the loop invariants in registers (otherwise they are rematerialised in every step)
        asm volatile("" : "+r"(plane_sz), "+r"(n_frames_u), "+l"(outA), "+r"(flags), "+r"(pf), "+r"(wbase), "+r"(hp));
        int4 h = lds_i4(hp);
#pragma unroll 1
        while (h.y != 0) {
          hp += 64;
          const int4 hn = lds_i4(hp);   // next step's header: its latency hides behind this step's loop
          unsigned wq = wbase + (unsigned)h.z;
          unsigned pp = pf + (unsigned)h.x;
          // four independent accumulation chains (one per bin of the float4 group); loads run one
          // iteration ahead of the FFMA2s (the loop is load-latency bound)
          float2 acc = make_float2(0.f, 0.f), acc1 = acc, acc2 = acc, acc3 = acc;
          float4 w = lds_f4(wq), p0 = lds_f4(pp), p1 = lds_f4(pp + 16);
          int it = h.y;   // pairs of iterations (the host pads every step to an even count)
#pragma unroll 1
          do {
            const float4 wb = lds_f4(wq + 64), p0b = lds_f4(pp + 32), p1b = lds_f4(pp + 48);
            acc = pfma(make_float2(p0.x, p0.y), w.x, acc);
            acc1 = pfma(make_float2(p0.z, p0.w), w.y, acc1);
            acc2 = pfma(make_float2(p1.x, p1.y), w.z, acc2);
            acc3 = pfma(make_float2(p1.z, p1.w), w.w, acc3);
            wq += 128;
            pp += 64;
            w = lds_f4(wq); p0 = lds_f4(pp); p1 = lds_f4(pp + 16);   // over-read past the last pair stays in smem
            acc = pfma(make_float2(p0b.x, p0b.y), wb.x, acc);
            acc1 = pfma(make_float2(p0b.z, p0b.w), wb.y, acc1);
            acc2 = pfma(make_float2(p1b.x, p1b.y), wb.z, acc2);
            acc3 = pfma(make_float2(p1b.z, p1b.w), wb.w, acc3);
          } while (--it);
          acc = padd(padd(acc, acc1), padd(acc2, acc3));
          if ((flags & 1u) && h.w >= 0) {
            float* o = outA + (unsigned)h.w * n_frames_u;
            asm volatile("st.global.f32 [%0], %1;" ::"l"(o), "f"(acc.x) : "memory");
            if (flags & 2u) asm volatile("st.global.f32 [%0], %1;" ::"l"(o + plane_sz), "f"(acc.y) : "memory");
          }
          h = hn;
        }
      } else if (a.out_tf) {
        // MODE_POWER, frequency-minor output: warp = frame of the tile, lanes run along the bins.  The lane -> bin map is shifted so
        // that every warp store starts on a 32-byte sector boundary (rows are 1025 floats long: the row start moves by 4 bytes
        // per row), i.e. each instruction writes whole sectors except at the two ends of the row.
        const int fr = f0 + warp;
        if (fr < a.n_frames) {
          const float2* pl = P + warp * kPLine;
          for (int r = 0; r < 2; ++r) {
            if (rowA + r >= a.rows) break;
            const long long e0 = ((rowA + r) * (long long)a.n_frames + fr) * a.n_freq;   // element offset of bin 0
            float* o = a.out + e0;
            const int shift = (int)((8 - (e0 & 7)) & 7);                                  // bins before the first sector boundary
            for (int k = lane + shift - 32; k < 1025; k += 32)
              if (k >= 0) o[k] = 0.25f * (r == 0 ? pl[k].x : pl[k].y);
          }
        }
      } else {  // MODE_POWER: lanes = (32 bins) x (6 frames) -> 24-byte runs along the frame axis
        const int fsub = tid % kWarps, ksub = tid / kWarps;
        const int fr = f0 + fsub;
        if (fr < a.n_frames) {
          const bool hasB = rowA + 1 < a.rows;
          float* o0 = a.out + rowA * (long long)a.n_freq * a.n_frames + fr;
          float* o1 = o0 + (long long)a.n_freq * a.n_frames;
          const float2* pl = P + fsub * kPLine;
          for (int k = ksub; k < 1025; k += 32) {
            const float2 v = pl[k];
            o0[(long long)k * a.n_frames] = 0.25f * v.x;
            if (hasB) o1[(long long)k * a.n_frames] = 0.25f * v.y;
          }
        }
      }
    }
  }
}

#include "stft2048_v2.cuh"
#include "stft2048_v3.cuh"

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
    if (MODE != MODE_MEL && a.out_tf) {
      const long long e = (row * a.n_frames + frame) * (long long)a.n_freq + k;
      if (MODE == MODE_COMPLEX) reinterpret_cast<float2*>(a.out)[e] = make_float2(xr, xi);
      else a.out[e] = xr * xr + xi * xi;
    } else if (MODE == MODE_COMPLEX) {
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

// ------------------------------------------------------------------------------------------
// Warp kernel, n_fft = 64 R with R in {1, 2, 4, 8, 16} (n_fft 64 .. 1024: covers the reference defaults n_fft = 1024,
// hop = 256, given_models.py:151-160,259-264): one warp = one frame of one ROW PAIR (rows packed in fp32x2 registers).
// M = 32 R complex points, n = 32 r + lane, k = k1 + R k2:
//   Z[k1 + R k2] = sum_l W_32^(l k2) [ W_M^(l k1) sum_r z[32 r + l] W_R^(r k1) ]
// = an R-point FFT over the lane's registers, a twiddle, and a 32-point FFT ACROSS LANES done with shuffles (radix-2 DIF:
// lane p ends up with k2 = bitrev5(p)).  The real-FFT split needs Z[M - k]: register R - k1 of lane p ^ 31 (k1 >= 1), or
// register 0 of the lane holding k2' = (32 - k2) & 31 (k1 = 0).  No shared memory except the mel power line.
// ------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void fft_regs(float2 (&re)[R], float2 (&im)[R]) {
  if constexpr (R == 2) {
    const float2 tr = re[1], ti = im[1];
    re[1] = psub(re[0], tr); im[1] = psub(im[0], ti); re[0] = padd(re[0], tr); im[0] = padd(im[0], ti);
  } else if constexpr (R == 4) {   // slots hold x0, x2, x1, x3 (bit-reversed input), natural output
    float2 ar = padd(re[0], re[1]), ai = padd(im[0], im[1]), br = psub(re[0], re[1]), bi = psub(im[0], im[1]);
    float2 cr = padd(re[2], re[3]), ci = padd(im[2], im[3]), dr = psub(re[2], re[3]), di = psub(im[2], im[3]);
    re[0] = padd(ar, cr); im[0] = padd(ai, ci); re[2] = psub(ar, cr); im[2] = psub(ai, ci);
    re[1] = padd(br, di); im[1] = psub(bi, dr);   // b + (-i) d
    re[3] = psub(br, di); im[3] = padd(bi, dr);   // b - (-i) d
  } else if constexpr (R == 8) {
    fft8_dit(re, im);
  } else if constexpr (R == 16) {
    fft16_dit(re, im);
  }
}

// Output staging of the warp kernel: a block = 8 warps = 8 CONSECUTIVE frames of one row pair.  Every warp leaves its frame's
// bins in shared memory (index k + k / R: lanes hold k = k1 + R bitrev5(lane), the skew makes the writes conflict free), then the
// block stores them with lanes = (4 bins) x (8 frames): 32-byte runs along the frame axis instead of 4-byte scattered stores
// (the output layout [.., F, T] is frame-minor; see the tile kernel's power epilogue for the same reason).
template <int R>
struct WkCfg {
  static constexpr int M = 32 * R;
  static constexpr int ROWLEN = ((M + M / (R > 1 ? R : 1) + 8 + 15) / 16) * 16 + 4;   // float2 units; = 4 mod 16: the store pass reads conflict free
  static constexpr int kThreadsW = 256;
};
template <int R>
__device__ __forceinline__ int wk_idx(int k) { return (R > 1) ? k + k / R : 2 * k; }

template <int R, int MODE>
__global__ void __launch_bounds__(256, 2) stft_warp_kernel(const StftArgs a, const float2* __restrict__ tw_lane,
                                                           const float2* __restrict__ tw_split, long long n_groups_total, int mel_w4_count) {
  constexpr int M = 32 * R, LOGR = (R == 1) ? 0 : (R == 2) ? 1 : (R == 4) ? 2 : (R == 8) ? 3 : 4;
  constexpr int ROWLEN = (R > 1) ? WkCfg<R>::ROWLEN : ((2 * M + 8 + 15) / 16) * 16 + 4;
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // staging: power / mel: SA[8 frames][ROWLEN] float2 (rowA, rowB); complex: SA = real parts, SB = imaginary parts
  float2* SA = reinterpret_cast<float2*>(smem);
  float2* SB = SA + 8 * ROWLEN;
  float2* P = SA + warp * ROWLEN;                                         // this warp's frame line
  float2* PI = SB + warp * ROWLEN;
  // mel: results [8 frames][n_mels] float2, then the filter tables staged once per (persistent) block
  float2* SMEL = SA + 8 * ROWLEN;
  int* s_meta = reinterpret_cast<int*>(SMEL + 8 * ((a.n_mels + 3) & ~3));
  float4* s_w4 = reinterpret_cast<float4*>(s_meta + ((3 * a.n_mels + 3) & ~3));
  if constexpr (MODE == MODE_MEL) {
    for (int i = threadIdx.x; i < 3 * a.n_mels; i += 256) s_meta[i] = __ldg(a.mel_start4 + i);   // start | cnt | off are contiguous
    for (int i = threadIdx.x; i < mel_w4_count; i += 256) s_w4[i] = __ldg(a.mel_w4 + i);
    __syncthreads();
  }
  // per-lane twiddles of the 5 lane-FFT stages (distance s = 16, 8, 4, 2, 1): upper lanes (lane & s) multiply by W_{2s}^(lane & (s-1))
  float2 stw[5];
  float ssg[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const int sdist = 16 >> q;
    const bool upper = (lane & sdist) != 0;
    float sn, cs;
    sincospif(-(float)(lane & (sdist - 1)) / (float)sdist, &sn, &cs);
    stw[q] = upper ? make_float2(cs, sn) : make_float2(1.f, 0.f);
    ssg[q] = upper ? -1.f : 1.f;
  }
  const int k2 = (int)(__brev((unsigned)lane) >> 27);
  const int src0 = (int)(__brev((unsigned)((32 - k2) & 31)) >> 27);   // lane that holds k2' = (32 - k2) & 31
  const bool center = a.center_off != 0;
  const long long n_freq = a.n_freq;
  const bool vec_ok = a.wav_aligned16 && (a.n_in & 1) == 0 && (a.hop & 1) == 0;
  const int groups_per_pair = (a.n_frames + 7) / 8;

  for (long long grp = blockIdx.x; grp < n_groups_total; grp += gridDim.x) {
    const long long pair = grp / groups_per_pair;
    const int f0 = (int)(grp - pair * groups_per_pair) * 8;
    const int frame = f0 + warp;
    const long long rowA = 2 * pair;
    const bool hasB = rowA + 1 < a.rows;
    if (frame < a.n_frames) {
    const float* __restrict__ pa = a.wav + rowA * a.n_in;
    const float* __restrict__ pb = hasB ? pa + a.n_in : nullptr;
    const long long s0 = (long long)frame * a.hop - a.center_off;
    float2 re[R], im[R];
    if (vec_ok && hasB && s0 >= 0 && s0 + 2 * M <= a.n_in) {   // interior frame: 8-byte loads of (x[2n], x[2n+1])
      const float2* FA = reinterpret_cast<const float2*>(pa + s0);
      const float2* FB = reinterpret_cast<const float2*>(pb + s0);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int n = 32 * r + lane;
        const float2 w = __ldg(a.window2 + n);
        const float2 xa = __ldg(FA + n), xb = __ldg(FB + n);
        const int slot = (LOGR == 0) ? 0 : (int)(__brev((unsigned)r) >> (32 - (LOGR == 0 ? 1 : LOGR)));
        re[slot] = make_float2(xa.x * w.x, xb.x * w.x);
        im[slot] = make_float2(xa.y * w.y, xb.y * w.y);
      }
    } else {                                                    // chunk borders: reflect padding, zero_pad_po2 tail, odd last row
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int n = 32 * r + lane;
        const float2 w = __ldg(a.window2 + n);
        const float xa0 = fetch_sample(pa, s0 + 2 * n, a.n_in, a.n_pad, center), xa1 = fetch_sample(pa, s0 + 2 * n + 1, a.n_in, a.n_pad, center);
        const float xb0 = fetch_sample(pb, s0 + 2 * n, a.n_in, a.n_pad, center), xb1 = fetch_sample(pb, s0 + 2 * n + 1, a.n_in, a.n_pad, center);
        const int slot = (LOGR == 0) ? 0 : (int)(__brev((unsigned)r) >> (32 - (LOGR == 0 ? 1 : LOGR)));
        re[slot] = make_float2(xa0 * w.x, xb0 * w.x);
        im[slot] = make_float2(xa1 * w.y, xb1 * w.y);
      }
    }
    fft_regs<R>(re, im);                       // slot k1: sum_r z[32 r + lane] W_R^(r k1)
#pragma unroll
    for (int k1 = 1; k1 < R; ++k1) {           // twiddle W_M^(lane k1)
      const float2 tw = __ldg(tw_lane + k1 * 32 + lane);
      const float2 r_ = re[k1], i_ = im[k1];
      re[k1] = pfma(i_, -tw.y, pmuls(r_, tw.x));
      im[k1] = pfma(i_, tw.x, pmuls(r_, tw.y));
    }
    // ---- 32-point FFT across lanes, radix-2 DIF: lower lane: a + b; upper lane: (a_lower - a_upper) W ----
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int sdist = 16 >> q;
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) {
        const float2 orr = make_float2(__shfl_xor_sync(0xffffffffu, re[k1].x, sdist), __shfl_xor_sync(0xffffffffu, re[k1].y, sdist));
        const float2 oi = make_float2(__shfl_xor_sync(0xffffffffu, im[k1].x, sdist), __shfl_xor_sync(0xffffffffu, im[k1].y, sdist));
        const float2 dr = pfma(re[k1], ssg[q], orr), di = pfma(im[k1], ssg[q], oi);   // lower: mine + other; upper: other - mine
        if (q < 4) {
          re[k1] = pfma(di, -stw[q].y, pmuls(dr, stw[q].x));
          im[k1] = pfma(di, stw[q].x, pmuls(dr, stw[q].y));
        } else {
          re[k1] = dr; im[k1] = di;
        }
      }
    }
    // lane holds Z[k1 + R k2], k2 = bitrev5(lane).  Real-FFT split: 2 X[k] = (Z[k] + conj Z[M-k]) - i W_{2M}^k (Z[k] - conj Z[M-k])
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) {
      const int kp = (k1 == 0) ? 0 : R - k1;                 // partner register
      float2 br, bi;
      if (k1 == 0) {
        br = make_float2(__shfl_sync(0xffffffffu, re[0].x, src0), __shfl_sync(0xffffffffu, re[0].y, src0));
        bi = make_float2(__shfl_sync(0xffffffffu, im[0].x, src0), __shfl_sync(0xffffffffu, im[0].y, src0));
      } else {
        br = make_float2(__shfl_xor_sync(0xffffffffu, re[kp].x, 31), __shfl_xor_sync(0xffffffffu, re[kp].y, 31));
        bi = make_float2(__shfl_xor_sync(0xffffffffu, im[kp].x, 31), __shfl_xor_sync(0xffffffffu, im[kp].y, 31));
      }
      const int k = k1 + R * k2;
      const float2 tw = __ldg(tw_split + k);                 // W_{2M}^k
      const float2 ar = re[k1], ai = im[k1];
      const float2 e_r = padd(ar, br), e_i = psub(ai, bi);   // E2 = a + conj(b)
      const float2 o_r = padd(ai, bi), o_i = psub(br, ar);   // O2 = (a - conj(b)) / i
      float2 xr = pmuls(pfma(o_i, -tw.y, pfma(o_r, tw.x, e_r)), 0.5f);
      float2 xi = pmuls(pfma(o_r, tw.y, pfma(o_i, tw.x, e_i)), 0.5f);
      if (k == 0) { xr = padd(ar, ai); xi = make_float2(0.f, 0.f); }   // DC
      if constexpr (MODE == MODE_COMPLEX) {
        P[wk_idx<R>(k)] = xr;
        PI[wk_idx<R>(k)] = xi;
      } else {
        P[wk_idx<R>(k)] = pfma2(xi, xi, pmul(xr, xr));
      }
    }
    if (lane == 0) {   // Nyquist bin M = Re Z[0] - Im Z[0] (lane 0 holds k2 = 0)
      const float2 ny = psub(re[0], im[0]);
      if constexpr (MODE == MODE_COMPLEX) {
        P[wk_idx<R>(M)] = ny;
        PI[wk_idx<R>(M)] = make_float2(0.f, 0.f);
      } else {
        P[wk_idx<R>(M)] = pmul(ny, ny);
      }
    }
    if constexpr (MODE == MODE_MEL) {
      __syncwarp();
      for (int m = lane; m < a.n_mels; m += 32) {
        const int start = s_meta[m], cnt = s_meta[a.n_mels + m];
        const float4* w4 = s_w4 + s_meta[2 * a.n_mels + m];      // weights x 0.25 (the fast kernel's P holds 4|X|^2)
        float2 acc = make_float2(0.f, 0.f);
        for (int j = 0; j < cnt; ++j) {
          const float4 w = w4[j];
          const int kb = start + 4 * j;
          if (kb <= M) acc = pfma(P[wk_idx<R>(kb)], w.x, acc);
          if (kb + 1 <= M) acc = pfma(P[wk_idx<R>(kb + 1)], w.y, acc);
          if (kb + 2 <= M) acc = pfma(P[wk_idx<R>(kb + 2)], w.z, acc);
          if (kb + 3 <= M) acc = pfma(P[wk_idx<R>(kb + 3)], w.w, acc);
        }
        SMEL[warp * ((a.n_mels + 3) & ~3) + m] = pmuls(acc, 4.0f);
      }
    }
    }   // frame < n_frames
    __syncthreads();
    // ---- block store: lanes = (4 bins / filters) x (8 frames): 32-byte runs along the frame axis ----
    {
      const int f = threadIdx.x & 7, kb = threadIdx.x >> 3;
      const bool fok = f0 + f < a.n_frames;
      if constexpr (MODE == MODE_MEL) {
        const int mstride = (a.n_mels + 3) & ~3;
        float* o = a.out + (rowA * a.n_mels) * a.n_frames + f0 + f;
        for (int m = kb; m < a.n_mels && fok; m += 32) {
          const float2 v = SMEL[f * mstride + m];
          o[(long long)m * a.n_frames] = v.x;
          if (hasB) o[((long long)a.n_mels + m) * a.n_frames] = v.y;
        }
      } else if (a.out_tf) {
        // frequency-minor output: warp = frame of the block, lanes run along the bins
        const int fw = threadIdx.x >> 5, ln = threadIdx.x & 31;
        if (f0 + fw < a.n_frames) {
          for (int r = 0; r < (hasB ? 2 : 1); ++r) {
            const long long e0 = ((rowA + r) * (long long)a.n_frames + f0 + fw) * n_freq;
            if constexpr (MODE == MODE_POWER) {
              float* o = a.out + e0;
              const int shift = (int)((8 - (e0 & 7)) & 7);
              for (int k = ln + shift - 32; k <= M; k += 32)
                if (k >= 0) { const float2 v = SA[fw * ROWLEN + wk_idx<R>(k)]; o[k] = r == 0 ? v.x : v.y; }
            } else {
              float2* o = reinterpret_cast<float2*>(a.out) + e0;
              for (int k = ln; k <= M; k += 32) {
                const float2 vr = SA[fw * ROWLEN + wk_idx<R>(k)], vi = SB[fw * ROWLEN + wk_idx<R>(k)];
                o[k] = r == 0 ? make_float2(vr.x, vi.x) : make_float2(vr.y, vi.y);
              }
            }
          }
        }
      } else if constexpr (MODE == MODE_POWER) {
        float* o = a.out + (rowA * n_freq) * a.n_frames + f0 + f;
        for (int k = kb; k <= M && fok; k += 32) {
          const float2 v = SA[f * ROWLEN + wk_idx<R>(k)];
          o[(long long)k * a.n_frames] = v.x;
          if (hasB) o[(n_freq + k) * a.n_frames] = v.y;
        }
      } else {
        float2* o = reinterpret_cast<float2*>(a.out) + (rowA * n_freq) * a.n_frames + f0 + f;
        for (int k = kb; k <= M && fok; k += 32) {
          const float2 vr = SA[f * ROWLEN + wk_idx<R>(k)], vi = SB[f * ROWLEN + wk_idx<R>(k)];
          o[(long long)k * a.n_frames] = make_float2(vr.x, vi.x);
          if (hasB) o[(n_freq + k) * a.n_frames] = make_float2(vr.y, vi.y);
        }
      }
    }
    __syncthreads();
  }
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
  int mel_w4_count = 0;
  int4* d_mel_steps = nullptr;   // fast path mel step lists followed by the step weights
  float* d_lane_consts = nullptr;
  int mel_steps_per_warp = 0, mel_hdr_bytes = 0, mel_w_bytes = 0;
  int fast_smem = 0, fast_grid_per_sm = 2;
  float2* d_wk = nullptr;        // warp kernel (n_fft <= 1024): [R][32] W_M^(lane k1), then [M] W_2M^k
  int wk_r = 0;
  // stft2048_v2_kernel (decoupled warps; banded mel): per KW in {6, 12}
  bool v2_mel_ok = false;        // the filterbank is a <= 2-adjacent-tap band with non-decreasing filter index
  int4* d_v2_groups = nullptr;   // warp start indices + 48-byte group records of the banded mel walk
  int v2_groups_len = 0;
  // stft2048_v3_kernel (no inter-warp synchronisation; warp-private mel walk, [row][frame][mel] output)
  bool v3 = false;                     // n_fft = 2048 / 1024, periodic Hann window, even hop: stft_v3_kernel can run the plan
  bool v3_mel_ok = false;
  float2* d_v3_tw1 = nullptr;          // [32 / FR][32] W_(n_fft/2)^(k1 n2)
  float* d_v3_lane = nullptr;          // float4[32] Hann phases + float2[32] W_n_fft^(k1 of the lane)
  unsigned char* d_v3_tab = nullptr;   // kV3MelTab bytes + uint32 [n_mels] run slots (see stft2048_v3.cuh)
  // resources of aa_stft_mel_f32_host (created lazily)
  cudaStream_t hstream[2] = {nullptr, nullptr};
  float* hbuf_in[2] = {nullptr, nullptr};
  float* hbuf_out[2] = {nullptr, nullptr};
  long long hbuf_rows = 0, hbuf_nin = 0, hbuf_out_per_row = 0;
};

// Banded form of the mel filterbank for stft2048_v2_kernel: every bin k feeds at most the two ADJACENT filters m_lo(k) and
// m_lo(k) + 1 with m_lo non-decreasing in k (true for torchaudio's triangular HTK / Slaney banks; verified here on the actual
// matrix, so a custom bank that is not of this form simply keeps the v1 kernel).  The runs of equal m_lo become "steps"
// {first bin, bins, filter to store}: out[m] = sum_{run m} w_lo P + sum_{run m-1} w_hi P.  Steps are dealt to the KW warps as
// contiguous ranges of about equal cost; a warp first re-walks the run before its range to get that run's w_hi sum.
static int v2_build_mel(AaStftPlan* p, const std::vector<float>& fb, int F, int n_mels) {
  p->v2_mel_ok = false;
  if (F != 1025 || n_mels < 1) return AA_OK;
  std::vector<int> mlo(F);
  std::vector<float2> w(1028, make_float2(0.f, 0.f));
  int prev = -1;
  for (int k = 0; k < F; ++k) {
    int nz[3], nn = 0;
    for (int m = 0; m < n_mels && nn < 3; ++m)
      if (fb[(size_t)k * n_mels + m] != 0.0f) nz[nn++] = m;
    int m;
    if (nn > 2) return AA_OK;
    if (nn == 2) {
      if (nz[1] != nz[0] + 1) return AA_OK;
      m = nz[0];
    } else if (nn == 1) {
      m = (prev == nz[0] - 1) ? prev : nz[0];
    } else {
      m = prev;
    }
    if (m < prev) return AA_OK;
    mlo[k] = m;
    prev = m;
    w[k].x = (m >= 0) ? 0.25f * fb[(size_t)k * n_mels + m] : 0.f;
    w[k].y = (m + 1 < n_mels) ? 0.25f * fb[(size_t)k * n_mels + m + 1] : 0.f;
  }
  // runs: s = m + 1 for m = -1 .. n_mels - 1 (possibly empty); groups of <= 4 bins per run
  const int ns = n_mels + 1;
  std::vector<int> k0(ns, 0), len(ns, 0), ng(ns, 1);
  for (int k = 0; k < F; ++k) {
    const int s_ = mlo[k] + 1;
    if (len[s_] == 0) k0[s_] = k;
    len[s_]++;
  }
  for (int s_ = 0; s_ < ns; ++s_) ng[s_] = std::max(1, (len[s_] + 3) / 4);
  // contiguous partition of the runs over the 12 warps, minimising the largest cost (2 per group, 1 per run end, plus the groups of
  // the run BEFORE the range, which the warp re-walks for its w_hi sum): binary search on the bound + greedy fill
  auto overlap = [&](int a0) { return (a0 > 0 && len[a0 - 1] > 0) ? 2 * ng[a0 - 1] : 0; };
  auto parts_needed = [&](long long bound, std::vector<int>* firsts) {
    int parts = 0, s_ = 0;
    if (firsts) firsts->clear();
    while (s_ < ns) {
      long long c = overlap(s_);
      if (firsts) firsts->push_back(s_);
      int taken = 0;
      while (s_ < ns && (taken == 0 || c + 2 * ng[s_] + 1 <= bound)) { c += 2 * ng[s_] + 1; ++s_; ++taken; }
      ++parts;
    }
    return parts;
  };
  long long lo_b = 1, hi_b = 0;
  for (int s_ = 0; s_ < ns; ++s_) hi_b += 2 * ng[s_] + 1;
  hi_b += 64;
  while (lo_b < hi_b) {
    const long long mid = (lo_b + hi_b) / 2;
    if (parts_needed(mid, nullptr) <= kV2W) hi_b = mid; else lo_b = mid + 1;
  }
  std::vector<int> firsts;
  parts_needed(lo_b, &firsts);
  while ((int)firsts.size() < kV2W) firsts.push_back(ns);   // idle warps get an empty range
  firsts.push_back(ns);
  std::vector<int4> tab(4, make_int4(0, 0, 0, 0));
  auto f2i = [](float v) { int i; std::memcpy(&i, &v, 4); return i; };
  auto push_run = [&](int s_, int m_store, bool last_of_warp) {
    const int n = ng[s_];
    for (int gi = 0; gi < n; ++gi) {
      const int kb = k0[s_] + 4 * gi;
      float wv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int j = 0; j < 4; ++j)
        if (4 * gi + j < len[s_]) { wv[2 * j] = w[kb + j].x; wv[2 * j + 1] = w[kb + j].y; }
      const bool last_g = gi == n - 1;
      tab.push_back(make_int4(len[s_] > 0 ? kb * 8 : 0, (last_g ? 1 : 0) | ((last_g && last_of_warp) ? 2 : 0), m_store, 0));
      tab.push_back(make_int4(f2i(wv[0]), f2i(wv[1]), f2i(wv[2]), f2i(wv[3])));
      tab.push_back(make_int4(f2i(wv[4]), f2i(wv[5]), f2i(wv[6]), f2i(wv[7])));
    }
  };
  auto push_dummy = [&](int flags) {
    tab.push_back(make_int4(0, flags, -1, 0));
    tab.push_back(make_int4(0, 0, 0, 0));
    tab.push_back(make_int4(0, 0, 0, 0));
  };
  for (int wq = 0; wq < kV2W; ++wq) {
    reinterpret_cast<int*>(tab.data())[wq] = (int)tab.size();
    const int a0 = firsts[wq], a1 = firsts[wq + 1];
    if (a0 >= a1) { push_dummy(3); continue; }
    if (a0 > 0 && len[a0 - 1] > 0) push_run(a0 - 1, -1, false);               // the run before the range: only its w_hi sum is used
    for (int s_ = a0; s_ < a1; ++s_) push_run(s_, s_ - 1, s_ == a1 - 1);      // s_ = 0 (bins below the first filter) stores nothing
  }
  push_dummy(3);
  push_dummy(3);   // the walk prefetches two records past the end of a list
  p->v2_groups_len = (int)tab.size();
  AA_CUDA(cudaMalloc(&p->d_v2_groups, sizeof(int4) * tab.size()));
  AA_CUDA(cudaMemcpy(p->d_v2_groups, tab.data(), sizeof(int4) * tab.size(), cudaMemcpyHostToDevice));
  p->v2_mel_ok = true;
  return AA_OK;
}

// Tables of the warp-private mel walk of stft2048_v3_kernel (see the header of stft2048_v3.cuh).  The bank must be a
// <= 2-adjacent-tap band with non-decreasing lower filter index (torchaudio's triangular HTK / Slaney banks are); runs of equal
// m_lo are numbered in bin order (only the non-empty ones: a filter whose run is empty points at a permanent zero slot).
static int v3_build_mel(AaStftPlan* p, const std::vector<float>& fb, int F, int n_mels) {
  p->v3_mel_ok = false;
  if (!p->v3 || (F != 1025 && F != 513) || n_mels < 1 || n_mels + 2 > kV3Runs) return AA_OK;
  const int segs = (F - 1) / 32;       // 32-bin segments per frame: 32 (n_fft = 2048) or 16 (n_fft = 1024: lanes 16..31 = second frame)
  std::vector<int> run(F);
  std::vector<float2> w(F, make_float2(0.f, 0.f));
  int prev = -1;
  for (int k = 0; k < F; ++k) {
    int nz[3], nn = 0;
    for (int m = 0; m < n_mels && nn < 3; ++m)
      if (fb[(size_t)k * n_mels + m] != 0.0f) nz[nn++] = m;
    int m;
    if (nn > 2) return AA_OK;
    if (nn == 2) {
      if (nz[1] != nz[0] + 1) return AA_OK;
      m = nz[0];
    } else if (nn == 1) {
      m = (prev == nz[0] - 1) ? prev : nz[0];
    } else {
      m = prev;
    }
    if (m < prev) return AA_OK;
    run[k] = m + 1;
    prev = m;
    w[k].x = (m >= 0) ? 0.25f * fb[(size_t)k * n_mels + m] : 0.f;          // P holds 4 |X|^2
    w[k].y = (m + 1 < n_mels) ? 0.25f * fb[(size_t)k * n_mels + m + 1] : 0.f;
  }
  std::vector<int> slot(F), slot_of_run(n_mels + 1, kV3Runs - 1);
  int c = 0;
  for (int k = 0; k < F; ++k) {
    if (k > 0 && run[k] != run[k - 1]) ++c;
    slot[k] = c;
    slot_of_run[run[k]] = c;
  }
  if (c + 1 > kV3Runs - 1) return AA_OK;
  std::vector<unsigned char> tab(kV3MelTab + 4 * (size_t)n_mels, 0);
  float2* tw = reinterpret_cast<float2*>(tab.data());
  uint32_t* masks = reinterpret_cast<uint32_t*>(tab.data() + 32 * 256);
  uint32_t* first = reinterpret_cast<uint32_t*>(tab.data() + 32 * 256 + 128);
  for (int g = 0; g < 32; ++g) {
    const int fr = g / segs, sg = g % segs;
    uint32_t mk = 0;
    for (int t = 0; t < 32; ++t) {
      const int k = 32 * sg + t;
      tw[t * 32 + g] = w[k];
      if (run[k + 1] != run[k]) mk |= 1u << t;
    }
    masks[g] = mk;
    first[g] = (uint32_t)(fr * kV3Runs + slot[32 * sg]);
    if (sg >= 1 && sg <= segs - 2 && mk == 0) return AA_OK;   // a segment without a run end: tail targets would collide
  }
  *reinterpret_cast<float2*>(tab.data() + 32 * 256 + 256) = w[F - 1];
  uint32_t* filt = reinterpret_cast<uint32_t*>(tab.data() + kV3MelTab);
  for (int m = 0; m < n_mels; ++m) filt[m] = (uint32_t)slot_of_run[m + 1] | ((uint32_t)slot_of_run[m] << 16);
  AA_CUDA(cudaMalloc(&p->d_v3_tab, tab.size()));
  AA_CUDA(cudaMemcpy(p->d_v3_tab, tab.data(), tab.size(), cudaMemcpyHostToDevice));
  p->v3_mel_ok = true;
  return AA_OK;
}

// twiddle / window tables of stft_v3_kernel for n_fft = 2048 / 1024
static int v3_build_tables(AaStftPlan* p) {
  const double PI = 3.14159265358979323846;
  const int nf = p->n_fft, fr = 2048 / nf, r1 = 32 / fr, h = nf / 2;
  std::vector<float2> tw((size_t)r1 * 32);
  for (int k1 = 0; k1 < r1; ++k1)
    for (int n2 = 0; n2 < 32; ++n2) {
      const double th = -2.0 * PI * (double)(k1 * n2) / (double)h;
      tw[(size_t)k1 * 32 + n2] = make_float2((float)std::cos(th), (float)std::sin(th));
    }
  std::vector<float> lc(32 * 4 + 32 * 2);
  for (int l = 0; l < 32; ++l) {
    for (int e = 0; e < 2; ++e) {
      const double phi = 2.0 * PI * (double)(2 * l + e) / (double)nf;
      lc[4 * l + e] = (float)std::cos(phi);         // (cos phi0, cos phi1, sin phi0, sin phi1): packed pairs for the window FFMA2s
      lc[4 * l + 2 + e] = (float)std::sin(phi);
    }
    const int k1 = l % r1;
    lc[128 + 2 * l] = (float)std::cos(2.0 * PI * k1 / (double)nf);
    lc[128 + 2 * l + 1] = (float)(-std::sin(2.0 * PI * k1 / (double)nf));
  }
  AA_CUDA(cudaMalloc(&p->d_v3_tw1, sizeof(float2) * tw.size()));
  AA_CUDA(cudaMemcpy(p->d_v3_tw1, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
  AA_CUDA(cudaMalloc(&p->d_v3_lane, sizeof(float) * lc.size()));
  AA_CUDA(cudaMemcpy(p->d_v3_lane, lc.data(), sizeof(float) * lc.size(), cudaMemcpyHostToDevice));
  return AA_OK;
}

static int v3_smem_bytes(int mode, int n_mels, int n_fft) {
  const int mel_bytes = mode == MODE_MEL ? ((kV3MelTab + 4 * n_mels + 15) & ~15) + kV3W * (2048 / n_fft) * kV3Runs * 16 : 0;
  return kV3W * kV3Xb + kV3Tables + mel_bytes + 16;
}

static int v2_smem_bytes(int nbuf, int hop, int groups_len) {
  const int span = (kV2W - 1) * hop + 2048;
  return nbuf * span * 8 + kV2W * kV2Xb + kV2Tables + groups_len * 16 + 4 * 8 + 16;
}

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
  AA_CUDA(cudaGetDevice(&p->device));
  const double PI = 3.14159265358979323846;
  // window as (w[2n], w[2n+1]) pairs
  std::vector<float> win(n_fft);
  for (int i = 0; i < n_fft; ++i)
    win[i] = window_host ? window_host[i] : (float)(0.5 - 0.5 * std::cos(2.0 * PI * i / n_fft));
  bool is_hann = true;   // the fast kernel synthesises the periodic Hann window in registers
  for (int i = 0; i < n_fft; ++i)
    if (std::fabs(win[i] - (float)(0.5 - 0.5 * std::cos(2.0 * PI * i / n_fft))) > 2e-6f) { is_hann = false; break; }
  p->fast = (n_fft == 2048) && (hop % 4 == 0) && (hop <= 1024) && is_hann;
  p->v3 = (n_fft == 2048 || n_fft == 1024) && (hop % 2 == 0) && is_hann;
  if (p->v3) {
    rc = v3_build_tables(p);
    if (rc != AA_OK) return rc;
  }
  AA_CUDA(cudaMalloc(&p->d_window2, sizeof(float) * n_fft));
  AA_CUDA(cudaMemcpy(p->d_window2, win.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
  // twiddles
  std::vector<float2> tw1, tw2(std::max(n_fft / 4, 1) + 1);
  if (p->fast) {
    std::vector<float> lc(32 * 4 + 32 * 2);
    for (int l = 0; l < 32; ++l) {
      for (int e = 0; e < 2; ++e) {
        const double phi = 2.0 * PI * (double)(2 * l + e) / 2048.0;
        lc[4 * l + 2 * e] = (float)std::cos(phi);
        lc[4 * l + 2 * e + 1] = (float)std::sin(phi);
      }
      lc[128 + 2 * l] = (float)std::cos(2.0 * PI * l / 2048.0);
      lc[128 + 2 * l + 1] = (float)(-std::sin(2.0 * PI * l / 2048.0));
    }
    AA_CUDA(cudaMalloc(&p->d_lane_consts, sizeof(float) * lc.size()));
    AA_CUDA(cudaMemcpy(p->d_lane_consts, lc.data(), sizeof(float) * lc.size(), cudaMemcpyHostToDevice));
    tw1.resize(33 * 32);   // row 32 = W_1024^(32 lane) = W_32^lane (stft2048_v2_kernel derives rows 17..31 from rows 15..1 with it)
    for (int k1 = 0; k1 < 33; ++k1)
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
  if (!p->fast && n_fft <= 1024 && getenv("AA_STFT_CTA_KERNEL") == nullptr) {
    const int R = n_fft / 64;
    std::vector<float2> wk((size_t)R * 32 + M);
    for (int k1 = 0; k1 < R; ++k1)
      for (int l = 0; l < 32; ++l) {
        const double th = -2.0 * PI * (double)(k1 * l) / (double)M;
        wk[(size_t)k1 * 32 + l] = make_float2((float)std::cos(th), (float)std::sin(th));
      }
    for (int k = 0; k < M; ++k) {
      const double th = -PI * (double)k / (double)M;
      wk[(size_t)R * 32 + k] = make_float2((float)std::cos(th), (float)std::sin(th));
    }
    AA_CUDA(cudaMalloc(&p->d_wk, sizeof(float2) * wk.size()));
    AA_CUDA(cudaMemcpy(p->d_wk, wk.data(), sizeof(float2) * wk.size(), cudaMemcpyHostToDevice));
    p->wk_r = R;
  }
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
    p->mel_w4_count = (int)(w.size() / 4);
    AA_CUDA(cudaMalloc(&p->d_mel_w4, sizeof(float) * w.size()));
    AA_CUDA(cudaMemcpy(p->d_mel_w4, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
    rc = v3_build_mel(p, fb, F, n_mels);
    if (rc != AA_OK) return rc;
    if (p->fast) {
      // Step lists of the fused mel epilogue.  A step = 4 adjacent filters; its weight block is
      // [n_iter][4 filters] float4 (4 bins each), n_iter = the longest filter of the step, shorter filters
      // zero-padded; a filter whose padded range would leave the 1028-bin P line is shifted down.  Steps are
      // dealt to the 6 warps longest-first onto the least loaded warp.
      struct Step { int4 h0, h1; };
      std::vector<Step> steps;
      std::vector<float> sw;
      for (int m0 = 0; m0 < n_mels; m0 += 4) {
        const int nv = std::min(4, n_mels - m0);
        int n_iter = 1;
        for (int qq = 0; qq < nv; ++qq) n_iter = std::max(n_iter, meta[n_mels + m0 + qq]);
        n_iter = (n_iter + 1) & ~1;   // the device loop is unrolled by two
        AA_REQUIRE(4 * n_iter <= 1028, "mel filter too wide for the fused epilogue");
        int bins[4] = {0, 0, 0, 0};
        const int wofs = (int)(sw.size() / 4);
        sw.resize(sw.size() + (size_t)n_iter * 16, 0.f);
        for (int qq = 0; qq < nv; ++qq) {
          const int m = m0 + qq, cnt = meta[n_mels + m];
          int b0 = cnt > 0 ? meta[m] : 0;
          const int shift = std::max(0, b0 + 4 * n_iter - 1028) / 4;   // groups to shift down
          b0 -= 4 * shift;
          bins[qq] = b0;
          for (int c = 0; c < cnt; ++c)
            for (int j = 0; j < 4; ++j)
              sw[((size_t)wofs + (size_t)(c + shift) * 4 + qq) * 4 + j] = w[((size_t)meta[2 * n_mels + m] + c) * 4 + j];
        }
        steps.push_back({make_int4(bins[0], bins[1], bins[2], bins[3]), make_int4(n_iter, wofs, m0, nv)});
      }
      std::vector<int> order(steps.size()), load(kWarps, 0);
      for (size_t i = 0; i < steps.size(); ++i) order[i] = (int)i;
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return steps[x].h1.x > steps[y].h1.x; });
      std::vector<std::vector<int>> per_warp(kWarps);
      for (int i : order) {
        int best = 0;
        for (int wq = 1; wq < kWarps; ++wq) if (load[wq] < load[best]) best = wq;
        per_warp[best].push_back(i);
        load[best] += steps[i].h1.x + 2;
      }
      size_t spw = 1;
      for (auto& l : per_warp) spw = std::max(spw, l.size());
      // headers [kWarps][spw + 1][4 quarters] = {P offset in bytes (8 per bin), n_iter / 2, weight offset in bytes, filter or -1}
      std::vector<int4> tab((size_t)kWarps * (spw + 1) * 4, make_int4(0, 0, 0, -1));
      for (int wq = 0; wq < kWarps; ++wq)
        for (size_t i = 0; i < per_warp[wq].size(); ++i) {
          const Step& s = steps[per_warp[wq][i]];
          const int b[4] = {s.h0.x, s.h0.y, s.h0.z, s.h0.w};
          for (int qq = 0; qq < 4; ++qq)
            tab[((size_t)wq * (spw + 1) + i) * 4 + qq] =
                make_int4((b[qq] >> 1) * 16, s.h1.x / 2, (s.h1.y + qq) * 16, qq < s.h1.w ? s.h1.z + qq : -1);
        }
      sw.resize(sw.size() + 32, 0.f);   // the device loop prefetches past the last step
      p->mel_steps_per_warp = (int)spw;
      p->mel_hdr_bytes = (int)(sizeof(int4) * tab.size());
      p->mel_w_bytes = (int)(sizeof(float) * sw.size());
      AA_CUDA(cudaMalloc(&p->d_mel_steps, p->mel_hdr_bytes + p->mel_w_bytes));
      AA_CUDA(cudaMemcpy(p->d_mel_steps, tab.data(), p->mel_hdr_bytes, cudaMemcpyHostToDevice));
      AA_CUDA(cudaMemcpy(reinterpret_cast<unsigned char*>(p->d_mel_steps) + p->mel_hdr_bytes, sw.data(), p->mel_w_bytes,
                         cudaMemcpyHostToDevice));
      rc = v2_build_mel(p, fb, F, n_mels);
      if (rc != AA_OK) return rc;
    }
  }
  if (p->fast) {
    const int smem = ((kWarps - 1) * hop + 2048) * 8 + kStageBytes + kTableBytes + p->mel_hdr_bytes + p->mel_w_bytes + 32;
    p->fast_smem = smem;
    AA_CUDA(cudaFuncSetAttribute(stft2048_kernel<MODE_COMPLEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AA_CUDA(cudaFuncSetAttribute(stft2048_kernel<MODE_POWER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    AA_CUDA(cudaFuncSetAttribute(stft2048_kernel<MODE_MEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nb = 0;
    AA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, stft2048_kernel<MODE_MEL>, kThreads, smem));
    p->fast_grid_per_sm = std::max(1, nb);
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
  cudaFree(p->d_window2); cudaFree(p->d_tw1); cudaFree(p->d_tw2); cudaFree(p->d_mel_meta); cudaFree(p->d_mel_w4); cudaFree(p->d_mel_steps); cudaFree(p->d_lane_consts); cudaFree(p->d_wk);
  cudaFree(p->d_v2_groups);
  cudaFree(p->d_v3_tab); cudaFree(p->d_v3_tw1); cudaFree(p->d_v3_lane);
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
                       float* out, cudaStream_t st, int out_tf = 0) {
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
  a.mel_w4 = p->d_mel_w4; a.out = out; a.out_tf = out_tf;
  a.wav_aligned16 = ((reinterpret_cast<uintptr_t>(wav) & 15) == 0) ? 1 : 0;
  const bool big_out = rows * (int64_t)std::max(p->n_freq, p->n_mels) * n_frames >= (1LL << 31);
  const int v3_mode = getenv("AA_STFT_V3") ? atoi(getenv("AA_STFT_V3")) : 1;      // 0: older kernels only
  AA_REQUIRE(!(mode == MODE_MEL && out_tf) || (p->v3 && p->v3_mel_ok && !big_out),
             "the [row][frame][mel] layout needs n_fft = 2048 / 1024 with a Hann window, an even hop and a banded filterbank (aa_stft_mel_tf_supported)");
  if (p->v3 && !big_out && out_tf && (v3_mode || mode == MODE_MEL) && rows < (1LL << 30) && n_pad < (1LL << 30)) {
    const int fr = 2048 / p->n_fft;
    Stft3Args b;
    b.wav = wav; b.out = out; b.rows = (int)rows; b.n_in = (int)n_in; b.n_pad = (int)n_pad; b.n_frames = (int)n_frames;
    b.hop = p->hop; b.center_off = a.center_off;
    b.items_per_pair = (int)((n_frames + fr - 1) / fr);
    const int64_t ni = ((rows + 1) / 2) * b.items_per_pair;
    AA_REQUIRE(ni < (1LL << 31) - 4096 * kV3W, "problem too large for the fast STFT path");
    b.n_items = (int)ni; b.n_freq = p->n_freq; b.n_mels = p->n_mels;
    b.wav_ok8 = ((reinterpret_cast<uintptr_t>(wav) & 7) == 0 && (n_in & 1) == 0 && (p->hop & 1) == 0) ? 1 : 0;
    b.wav_ok16 = ((reinterpret_cast<uintptr_t>(wav) & 15) == 0 && (n_in & 3) == 0 && (p->hop & 3) == 0) ? 1 : 0;
    static const int v3_prefetch = getenv("AA_STFT_PREFETCH") ? atoi(getenv("AA_STFT_PREFETCH")) : 0;
    b.prefetch = v3_prefetch;
    b.tw1 = p->d_v3_tw1; b.lane_consts = p->d_v3_lane; b.mel_tab = p->d_v3_tab;
    const int smem = v3_smem_bytes(mode, p->n_mels, p->n_fft);
    const int64_t tiles = (ni + kV3W - 1) / kV3W;
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)aa::num_sms());
#define AA_V3(NFFT, MD)                                                         \
  do {                                                                          \
    AA_CUDA(aa::ensure_dyn_smem(stft_v3_kernel<NFFT, MD>, smem));               \
    stft_v3_kernel<NFFT, MD><<<grid, kV3W * 32, smem, st>>>(b);                 \
  } while (0)
    if (p->n_fft == 2048) {
      if (mode == MODE_COMPLEX) AA_V3(2048, MODE_COMPLEX); else if (mode == MODE_POWER) AA_V3(2048, MODE_POWER); else AA_V3(2048, MODE_MEL);
    } else {
      if (mode == MODE_COMPLEX) AA_V3(1024, MODE_COMPLEX); else if (mode == MODE_POWER) AA_V3(1024, MODE_POWER); else AA_V3(1024, MODE_MEL);
    }
#undef AA_V3
    AA_LAUNCH_CHECK();
    return AA_OK;
  }
  const int v2_mode = getenv("AA_STFT_V2") ? atoi(getenv("AA_STFT_V2")) : 1;       // 0: v1 kernel only
  static const int v2_nbuf = getenv("AA_STFT_NBUF") ? atoi(getenv("AA_STFT_NBUF")) : 1;     // sample ring depth wanted (1 or 2)
  // mel: the v2 kernel's banded walk is correct but still slower end to end than the v1 tile kernel (393 vs 375 us on the headline
  // workload: every warp walks right after the P-line rendezvous, so the walk's latency is not hidden) -- opt in with AA_STFT_V2_MEL=1
  const bool v2_mel = getenv("AA_STFT_V2_MEL") != nullptr && atoi(getenv("AA_STFT_V2_MEL")) != 0;
  if (p->fast && !big_out && v2_mode && rows < (1LL << 30) && n_pad < (1LL << 30) &&
      ((mode == MODE_MEL && p->v2_mel_ok && v2_mel) || (mode != MODE_MEL && out_tf))) {
    const int groups_len = mode == MODE_MEL ? p->v2_groups_len : 1;
    int dev = 0, max_smem = 0;
    AA_CUDA(cudaGetDevice(&dev));
    AA_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int nbuf = (v2_nbuf >= 2 && v2_smem_bytes(2, p->hop, groups_len) <= max_smem) ? 2 : 1;
    const int smem = v2_smem_bytes(nbuf, p->hop, groups_len);
    AA_REQUIRE(smem <= max_smem, "hop=%d needs %d bytes of shared memory (limit %d)", p->hop, smem, max_smem);
    Stft2Args b;
    b.wav = wav; b.out = out; b.rows = (int)rows; b.n_in = (int)n_in; b.n_pad = (int)n_pad; b.n_frames = (int)n_frames;
    b.hop = p->hop; b.center_off = a.center_off; b.tiles_per_pair = (int)((n_frames + kV2W - 1) / kV2W);
    const int64_t nt = ((rows + 1) / 2) * b.tiles_per_pair;
    AA_REQUIRE(nt < (1LL << 31), "problem too large for the fast STFT path");
    b.n_tiles = (int)nt; b.n_freq = p->n_freq; b.n_mels = p->n_mels; b.wav_aligned16 = a.wav_aligned16; b.nbuf = nbuf;
    static const int v2_diag = getenv("AA_STFT_DIAG") ? atoi(getenv("AA_STFT_DIAG")) : 0;
    b.diag = v2_diag;
    b.tw1 = p->d_tw1; b.lane_consts = p->d_lane_consts;
    b.mel_groups = mode == MODE_MEL ? p->d_v2_groups : reinterpret_cast<const int4*>(p->d_tw1);
    b.mel_groups_len = groups_len;
    const unsigned grid = (unsigned)std::min<int64_t>(nt, (int64_t)aa::num_sms());
#define AA_V2(MD)                                                               \
  do {                                                                          \
    AA_CUDA(aa::ensure_dyn_smem(stft2048_v2_kernel<MD>, smem));                 \
    stft2048_v2_kernel<MD><<<grid, kV2W * 32, smem, st>>>(b);                   \
  } while (0)
    if (mode == MODE_COMPLEX) AA_V2(MODE_COMPLEX); else if (mode == MODE_POWER) AA_V2(MODE_POWER); else AA_V2(MODE_MEL);
#undef AA_V2
    AA_LAUNCH_CHECK();
    return AA_OK;
  }
  if (p->fast && !big_out) {
    a.tiles_per_pair = (int)((n_frames + kWarps - 1) / kWarps);
    const int64_t pairs = (rows + 1) / 2;
    a.n_tiles = pairs * a.tiles_per_pair;
    a.mel_steps = p->d_mel_steps; a.mel_steps_per_warp = p->mel_steps_per_warp;
    a.mel_hdr_bytes = p->mel_hdr_bytes; a.mel_w_bytes = p->mel_w_bytes; a.lane_consts = p->d_lane_consts;
    AA_REQUIRE(a.n_tiles < (1LL << 31) && rows < (1LL << 30) && n_pad < (1LL << 30), "problem too large for the fast STFT path");
    const int64_t grid = std::min<int64_t>(a.n_tiles, (int64_t)p->fast_grid_per_sm * aa::num_sms());
    const int smem = p->fast_smem;
    if (mode == MODE_COMPLEX) stft2048_kernel<MODE_COMPLEX><<<(unsigned)grid, kThreads, smem, st>>>(a);
    else if (mode == MODE_POWER) stft2048_kernel<MODE_POWER><<<(unsigned)grid, kThreads, smem, st>>>(a);
    else stft2048_kernel<MODE_MEL><<<(unsigned)grid, kThreads, smem, st>>>(a);
  } else {
    a.tiles_per_pair = 0; a.n_tiles = 0; a.mel_steps = nullptr; a.mel_steps_per_warp = 0;
    a.mel_hdr_bytes = a.mel_w_bytes = 0; a.lane_consts = nullptr;
    if (p->wk_r > 0) {
      const int R = p->wk_r;
      const long long n_groups = ((rows + 1) / 2) * ((n_frames + 7) / 8);
      const unsigned wgrid = (unsigned)std::min<long long>(n_groups, (long long)aa::num_sms() * 8);
      const int M_ = 32 * R;
      const int rowlen = (R > 1) ? ((M_ + M_ / R + 8 + 15) / 16) * 16 + 4 : ((2 * M_ + 8 + 15) / 16) * 16 + 4;
      const int mstride = (p->n_mels + 3) & ~3;
      const int wsmem = (mode == MODE_COMPLEX) ? 2 * 8 * rowlen * 8
                        : (mode == MODE_POWER) ? 8 * rowlen * 8
                                               : 8 * rowlen * 8 + 8 * mstride * 8 + ((3 * p->n_mels + 3) & ~3) * 4 + p->mel_w4_count * 16;
      const int w4c = p->mel_w4_count;
      const long long n_items = n_groups;
      const float2* twl = p->d_wk;
      const float2* tws = p->d_wk + (size_t)R * 32;
      AA_REQUIRE(wsmem <= 200 * 1024, "filterbank too large for the warp STFT kernel (%d bytes of shared memory)", wsmem);
#define AA_WK(RR)                                                                                                        \
  do {                                                                                                                   \
    if (mode == MODE_COMPLEX) {                                                                                          \
      AA_CUDA(cudaFuncSetAttribute(stft_warp_kernel<RR, MODE_COMPLEX>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsmem)); \
      stft_warp_kernel<RR, MODE_COMPLEX><<<wgrid, 256, wsmem, st>>>(a, twl, tws, n_items, w4c);                          \
    } else if (mode == MODE_POWER) {                                                                                     \
      AA_CUDA(cudaFuncSetAttribute(stft_warp_kernel<RR, MODE_POWER>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsmem));   \
      stft_warp_kernel<RR, MODE_POWER><<<wgrid, 256, wsmem, st>>>(a, twl, tws, n_items, w4c);                            \
    } else {                                                                                                             \
      AA_CUDA(cudaFuncSetAttribute(stft_warp_kernel<RR, MODE_MEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsmem));     \
      stft_warp_kernel<RR, MODE_MEL><<<wgrid, 256, wsmem, st>>>(a, twl, tws, n_items, w4c);                              \
    }                                                                                                                    \
  } while (0)
      if (R == 1) AA_WK(1); else if (R == 2) AA_WK(2); else if (R == 4) AA_WK(4); else if (R == 8) AA_WK(8); else AA_WK(16);
#undef AA_WK
      AA_LAUNCH_CHECK();
      return AA_OK;
    }
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
int aa_stft_complex_tf_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                           float* out, void* stream) {
  return stft_launch(plan, MODE_COMPLEX, wav, rows, n_in, zero_pad, out, (cudaStream_t)stream, 1);
}
int aa_stft_power_tf_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                         float* out, void* stream) {
  return stft_launch(plan, MODE_POWER, wav, rows, n_in, zero_pad, out, (cudaStream_t)stream, 1);
}
int aa_stft_mel_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                    float* out, void* stream) {
  return stft_launch(plan, MODE_MEL, wav, rows, n_in, zero_pad, out, (cudaStream_t)stream);
}

int aa_stft_mel_tf_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                       float* out, void* stream) {
  return stft_launch(plan, MODE_MEL, wav, rows, n_in, zero_pad, out, (cudaStream_t)stream, 1);
}
int aa_stft_mel_tf_supported(const AaStftPlan* plan) { return (plan && plan->v3 && plan->v3_mel_ok) ? 1 : 0; }

int aa_magdphase_f32(const float* spec, int64_t c, int64_t n_freq, int64_t n_frames, float* out, void* stream) {
  AA_REQUIRE(spec && out, "NULL tensor pointer");
  AA_REQUIRE(c >= 1 && n_freq >= 1 && n_frames >= 1 && n_frames < (1LL << 31), "bad shape");
  const long long cf = c * n_freq, total = cf * n_frames;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)aa::num_sms() * 8);
  magdphase_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(spec), cf, (int)n_frames, total, out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

static int mel_host_impl(const AaStftPlan* plan_c, const float* wav_host, int64_t rows, int64_t n_in, int zero_pad,
                         float* out_host, int64_t rows_per_chunk, int out_tf) {
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
    rc = stft_launch(p, MODE_MEL, p->hbuf_in[slot], nr, n_in, zero_pad, p->hbuf_out[slot], st, out_tf);
    if (rc != AA_OK) return rc;
    AA_CUDA(cudaMemcpyAsync(out_host + r0 * out_per_row, p->hbuf_out[slot], sizeof(float) * nr * out_per_row,
                            cudaMemcpyDeviceToHost, st));
  }
  AA_CUDA(cudaStreamSynchronize(p->hstream[0]));
  AA_CUDA(cudaStreamSynchronize(p->hstream[1]));
  return AA_OK;
}

int aa_stft_mel_f32_host(const AaStftPlan* plan, const float* wav_host, int64_t rows, int64_t n_in, int zero_pad,
                         float* out_host, int64_t rows_per_chunk) {
  return mel_host_impl(plan, wav_host, rows, n_in, zero_pad, out_host, rows_per_chunk, 0);
}
int aa_stft_mel_tf_f32_host(const AaStftPlan* plan, const float* wav_host, int64_t rows, int64_t n_in, int zero_pad,
                            float* out_host, int64_t rows_per_chunk) {
  return mel_host_impl(plan, wav_host, rows, n_in, zero_pad, out_host, rows_per_chunk, 1);
}

#pragma GCC visibility pop
}  // extern "C"
