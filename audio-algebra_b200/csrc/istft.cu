// Decoder side of the STFT given models (given_models.py:159,168 InverseSpectrogram, :181,189 GriffinLim,
// :268-269,278-280 InverseMelScale + GriffinLim) -- round-trip demos, not the hot path: simple, deterministic kernels.
//
//   aa_istft_f32          torch.istft(center, Hann or given window, onesided, unnormalised, length=None):
//                         frame kernel  = Hermitian extension + radix-2 inverse FFT in shared memory (one CTA per (row, frame)),
//                                         times window / n_fft -> scratch [rows][frames][n_fft];
//                         gather kernel = overlap-add of the <= ceil(n_fft / hop) frames that cover a sample in frame order
//                                         (no float atomics), divided by the window-square envelope, centre trim.
//   aa_griffinlim_update_c64   one phase update of torchaudio.functional.griffinlim (momentum form), fused with the product
//                              magnitude * phase that feeds the next inverse transform.
//   aa_inverse_mel_f32    relu(P mel) with P = pinv(fb^T) (minimum-norm least squares, what InverseMelScale's lstsq returns for
//                         the full-rank underdetermined system), P computed once on the host in float64.
#include "aa_common.cuh"

#include <algorithm>

namespace {

constexpr int kIstftThreads = 256;

__device__ __forceinline__ int bitrev_n(int v, int bits) { return (int)(__brev((unsigned)v) >> (32 - bits)); }

__global__ void __launch_bounds__(kIstftThreads) istft_frame_kernel(const float2* __restrict__ spec, long long row_stride, long long stride_f,
                                                                    long long stride_t, int n_fft, int logn, long long n_frames,
                                                                    const float* __restrict__ window, float* __restrict__ scratch) {
  extern __shared__ float2 sx[];   // n_fft complex points
  const long long t = blockIdx.x, row = blockIdx.y;
  const int half = n_fft >> 1;
  const float2* src = spec + row * row_stride + t * stride_t;
  for (int k = threadIdx.x; k < n_fft; k += blockDim.x) {
    const int kk = k <= half ? k : n_fft - k;
    float2 v = src[(long long)kk * stride_f];
    if (k > half) v.y = -v.y;
    if (kk == 0 || kk == half) v.y = 0.f;   // irfft ignores the imaginary parts of DC and Nyquist
    sx[bitrev_n(k, logn)] = v;
  }
  __syncthreads();
  for (int len = 2; len <= n_fft; len <<= 1) {
    const int h = len >> 1;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
      const int grp = i / h, j = i - grp * h;
      const int a = grp * len + j, b = a + h;
      float sn, cs;
      sincospif(2.0f * (float)j / (float)len, &sn, &cs);   // e^{+2 pi i j / len}: inverse transform
      const float2 xb = sx[b], xa = sx[a];
      const float2 tb = make_float2(xb.x * cs - xb.y * sn, xb.x * sn + xb.y * cs);
      sx[b] = make_float2(xa.x - tb.x, xa.y - tb.y);
      sx[a] = make_float2(xa.x + tb.x, xa.y + tb.y);
    }
    __syncthreads();
  }
  float* dst = scratch + (row * n_frames + t) * (long long)n_fft;
  const float inv_n = 1.0f / (float)n_fft;
  for (int m = threadIdx.x; m < n_fft; m += blockDim.x) {
    const float w = window ? window[m] : 0.5f - 0.5f * cospif(2.0f * (float)m / (float)n_fft);
    dst[m] = sx[m].x * inv_n * w;
  }
}

__global__ void istft_ola_kernel(const float* __restrict__ scratch, const float* __restrict__ window, int n_fft, int hop, long long n_frames,
                                 int start, long long out_len, float* __restrict__ out) {
  const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long row = blockIdx.y;
  if (n >= out_len) return;
  const long long j = n + start;
  long long t0 = (j - n_fft + hop) / hop;   // ceil((j - n_fft + 1) / hop) for j - n_fft + 1 > 0
  if (j - n_fft + 1 <= 0) t0 = 0;
  long long t1 = j / hop;
  if (t1 > n_frames - 1) t1 = n_frames - 1;
  float acc = 0.f, env = 0.f;
  for (long long t = t0; t <= t1; ++t) {
    const int m = (int)(j - t * hop);
    const float w = window ? window[m] : 0.5f - 0.5f * cospif(2.0f * (float)m / (float)n_fft);
    acc += scratch[(row * n_frames + t) * (long long)n_fft + m];
    env = fmaf(w, w, env);
  }
  out[row * out_len + n] = env > 1e-11f ? acc / env : 0.f;
}

// angles = rebuilt - m * tprev; angles /= |angles| + 1e-16; tprev <- rebuilt; prod = mag * angles   (functional.griffinlim)
__global__ void gl_update_kernel(const float2* __restrict__ rebuilt, float2* __restrict__ tprev, const float* __restrict__ mag,
                                 float2* __restrict__ prod, long long n, float momentum, int first) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float2 r = rebuilt[i];
    float2 a = r;
    if (!first && momentum != 0.f) {
      const float2 p = tprev[i];
      a.x -= momentum * p.x;
      a.y -= momentum * p.y;
    }
    const float inv = 1.0f / (sqrtf(a.x * a.x + a.y * a.y) + 1e-16f);
    tprev[i] = r;
    const float mg = mag[i];
    prod[i] = make_float2(mg * a.x * inv, mg * a.y * inv);
  }
}

// MagDPhaseSpectrogramAE.decode (given_models.py:233-254): theta integrated along time with the reference's wrap
// (theta >= 2 pi -> theta - 2 pi, pi = 3.141592653589), spec = mag (cos theta + i sin theta).  One thread per (channel, bin).
__global__ void magdphase_decode_kernel(const float* __restrict__ reps, int c, int f, long long t, int init_mode, const float* __restrict__ theta0,
                                        float2* __restrict__ spec) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)c * f) return;
  const float two_pi = (float)(2.0 * 3.141592653589);
  const float* mag = reps + idx * t;
  const float* dth = reps + ((long long)c * f + idx) * t;
  float2* out = spec + idx * t;
  float th = init_mode == 0 ? dth[0] : (init_mode == 1 ? theta0[idx] : 0.f);   // 'true' | 'rand' | 'zero'
  for (long long k = 0; k < t; ++k) {
    if (k > 0) {
      th = th + dth[k];
      th = th < two_pi ? th : th - two_pi;
    }
    float sn, cs;
    sincosf(th, &sn, &cs);
    const float m = mag[k];
    out[k] = make_float2(m * cs, m * sn);
  }
}

// out[r][f][t] = relu(sum_m P[f][m] mel[r][m][t]); tile 32 f x 32 t, mel strides in elements
constexpr int IM_T = 32;
__global__ void __launch_bounds__(IM_T * 8) inverse_mel_kernel(const float* __restrict__ P, const float* __restrict__ mel, long long mel_row_stride,
                                                               long long mel_stride_m, long long mel_stride_t, int n_mels, int n_freq,
                                                               long long n_frames, float* __restrict__ out) {
  __shared__ float sP[IM_T][33], sM[32][IM_T + 1];
  const long long r = blockIdx.z;
  const int f0 = blockIdx.y * IM_T;
  const long long t0 = (long long)blockIdx.x * IM_T;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int m0 = 0; m0 < n_mels; m0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ff = ty + 8 * i;
      sP[ff][tx] = (f0 + ff < n_freq && m0 + tx < n_mels) ? P[(long long)(f0 + ff) * n_mels + m0 + tx] : 0.f;
      const int mm = ty + 8 * i;
      sM[mm][tx] = (m0 + mm < n_mels && t0 + tx < n_frames) ? mel[r * mel_row_stride + (long long)(m0 + mm) * mel_stride_m + (t0 + tx) * mel_stride_t] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int m = 0; m < 32; ++m) {
      const float mv = sM[m][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(sP[ty + 8 * i][m], mv, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ff = f0 + ty + 8 * i;
    if (ff < n_freq && t0 + tx < n_frames) out[(r * n_freq + ff) * n_frames + t0 + tx] = fmaxf(acc[i], 0.f);
  }
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int64_t aa_istft_workspace_floats(int64_t rows, int n_fft, int64_t n_frames) { return rows * n_frames * (int64_t)n_fft; }

int aa_istft_f32(const void* spec, int64_t rows, int n_fft, int hop, int center, int64_t n_frames, int64_t row_stride,
                 int64_t stride_f, int64_t stride_t, const float* window, float* out, int64_t out_len, float* workspace,
                 void* stream) {
  AA_REQUIRE(spec && out && workspace, "NULL argument");
  AA_REQUIRE(n_fft >= 8 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0, "n_fft=%d: the inverse STFT supports powers of two in [8, 4096]", n_fft);
  AA_REQUIRE(hop >= 1 && hop <= n_fft && n_frames >= 1 && rows >= 1 && rows < 65536, "bad shape rows=%lld frames=%lld hop=%d", (long long)rows,
             (long long)n_frames, hop);
  const int64_t full = (int64_t)n_fft + (int64_t)hop * (n_frames - 1);
  const int start = center ? n_fft / 2 : 0;
  AA_REQUIRE(out_len >= 1 && out_len + start <= full, "out_len=%lld does not fit the %lld overlap-added samples", (long long)out_len, (long long)full);
  int logn = 0;
  while ((1 << logn) < n_fft) ++logn;
  cudaStream_t st = (cudaStream_t)stream;
  AA_REQUIRE(n_frames < (1LL << 31), "too many frames");
  istft_frame_kernel<<<dim3((unsigned)n_frames, (unsigned)rows), kIstftThreads, (size_t)n_fft * sizeof(float2), st>>>(
      reinterpret_cast<const float2*>(spec), row_stride, stride_f, stride_t, n_fft, logn, n_frames, window, workspace);
  AA_LAUNCH_CHECK();
  istft_ola_kernel<<<dim3((unsigned)((out_len + 255) / 256), (unsigned)rows), 256, 0, st>>>(workspace, window, n_fft, hop, n_frames, start, out_len, out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_griffinlim_update_c64(const void* rebuilt, void* tprev, const float* mag, void* prod, int64_t n, float momentum, int first,
                             void* stream) {
  AA_REQUIRE(rebuilt && tprev && mag && prod && n >= 1, "bad argument");
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)aa::num_sms() * 8);
  gl_update_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(rebuilt), reinterpret_cast<float2*>(tprev), mag,
                                                            reinterpret_cast<float2*>(prod), n, momentum, first);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_magdphase_decode_f32(const float* reps, int64_t c, int64_t f, int64_t t, int init_mode, const float* theta0, void* spec, void* stream) {
  AA_REQUIRE(reps && spec && c >= 1 && f >= 1 && t >= 1, "bad argument");
  AA_REQUIRE(init_mode == 0 || init_mode == 2 || (init_mode == 1 && theta0), "init_mode 1 ('rand') needs theta0 [c][f]");
  const long long n = c * f;
  magdphase_decode_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(reps, (int)c, (int)f, t, init_mode, theta0,
                                                                                       reinterpret_cast<float2*>(spec));
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_inverse_mel_f32(const float* pinv, const float* mel, int64_t rows, int n_mels, int n_freq, int64_t n_frames, int64_t mel_row_stride,
                       int64_t mel_stride_m, int64_t mel_stride_t, float* out, void* stream) {
  AA_REQUIRE(pinv && mel && out, "NULL argument");
  AA_REQUIRE(rows >= 1 && rows < 65536 && n_mels >= 1 && n_freq >= 1 && n_frames >= 1, "bad shape");
  inverse_mel_kernel<<<dim3((unsigned)((n_frames + IM_T - 1) / IM_T), (unsigned)((n_freq + IM_T - 1) / IM_T), (unsigned)rows), IM_T * 8, 0,
                       (cudaStream_t)stream>>>(pinv, mel, mel_row_stride, mel_stride_m, mel_stride_t, n_mels, n_freq, n_frames, out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

#pragma GCC visibility pop
}  // extern "C"
