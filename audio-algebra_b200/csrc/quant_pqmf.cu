// The two dormant branches of DiffusionDVAE.encode_it (aa_mixer.py:178-179, 189-192):
//   * PQMF analysis front-end (pqmf_bands > 1): the reference calls the third-party `diffusion.pqmf.PQMF(2, 70, bands)`;
//     its forward is a strided convolution of every audio channel with a cosine-modulated filterbank hk [bands][taps]
//     (stride = bands, padding = taps/2, last output dropped) followed by the "reverse half" sign flip of the even samples
//     of the odd bands.  The filterbank itself is designed on the host (audio-algebra_b200/DiffusionDVAE.py).
//   * Memcodes quantiser (num_quantizers > 0, third-party `nwt_pytorch.Memcodes` / `dvae.residual_memcodes`): per head,
//     the code whose key has the largest dot product with the (scaled) query slice is looked up and its value emitted
//     (eval-mode path: argmax + one-hot; the training path samples with gumbel-softmax and is not on encode_it's route).
// Both are restated from memory of un-vendored upstream code: PARITY UNPINNED, same status as the
// SoundStreamXL encoder.  These branches are off at the reference defaults (pqmf_bands = 1, num_quantizers = 0).
#include "aa_common.cuh"

#include <algorithm>
#include <cstdint>

namespace {

constexpr int kT = 128;   // outputs (time steps) per block

// out[row][k][t] = sign(k, t) * sum_j hk[k][j] * xpad[row][t * m + j],  xpad = x zero-padded by taps/2 on both sides
__global__ void __launch_bounds__(kT) pqmf_analysis_kernel(const float* __restrict__ x, const float* __restrict__ hk, int m, int taps, long long n,
                                                           long long t_out, float* __restrict__ out) {
  extern __shared__ float sh[];
  const int span = kT * m + taps;
  float* xs = sh;                  // [span] input window of this block
  float* hs = sh + span;           // [m][taps]
  const long long row = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * kT;
  const long long s0 = t0 * m - taps / 2;
  const float* xr = x + row * n;
  for (int i = threadIdx.x; i < span; i += kT) {
    const long long s = s0 + i;
    xs[i] = (s >= 0 && s < n) ? xr[s] : 0.f;
  }
  for (int i = threadIdx.x; i < m * taps; i += kT) hs[i] = hk[i];
  __syncthreads();
  const long long t = t0 + threadIdx.x;
  if (t >= t_out) return;
  const float* xw = xs + threadIdx.x * m;
  for (int k = 0; k < m; ++k) {
    const float* h = hs + k * taps;
    float a0 = 0.f, a1 = 0.f;
    int j = 0;
    for (; j + 1 < taps; j += 2) { a0 = fmaf(h[j], xw[j], a0); a1 = fmaf(h[j + 1], xw[j + 1], a1); }
    if (j < taps) a0 = fmaf(h[j], xw[j], a0);
    float v = a0 + a1;
    if ((k & 1) && !(t & 1)) v = -v;   // reverse_half: mask[..., 1::2, ::2] = -1
    out[(row * m + k) * t_out + t] = v;
  }
}

// k_out[h][j][o] = sum_i wk[h*d + o][i] * codes[h][j][i]   (grouped 1x1 Conv1d over the codes, groups = heads; same for v)
__global__ void memcodes_kv_kernel(const float* __restrict__ codes, const float* __restrict__ wk, const float* __restrict__ wv, int heads, int n_codes,
                                   int d, float* __restrict__ k_out, float* __restrict__ v_out) {
  const long long total = (long long)heads * n_codes * d;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(idx % d);
    const long long hj = idx / d;
    const int h = (int)(hj / n_codes);
    const float* c = codes + hj * d;
    const float* rk = wk + ((size_t)h * d + o) * d;
    const float* rv = wv + ((size_t)h * d + o) * d;
    float ak = 0.f, av = 0.f;
    for (int i = 0; i < d; ++i) { ak = fmaf(rk[i], c[i], ak); av = fmaf(rv[i], c[i], av); }
    k_out[idx] = ak;
    v_out[idx] = av;
  }
}

// One thread per (batch, head, position): q = x[b][h*d .. h*d+d)[n] * scale; j* = argmax_j <q, k[h][j]> (first maximum);
// quantized = v[h][j*].  x / outputs are channel-major [B][heads*d][N] (the reference rearranges to [B][N][D] and back).
// Optional fused residual-VQ bookkeeping: resid_out = x - quantized; acc (in/out) += quantized, tanh applied to acc when final_tanh.
template <int D>
__global__ void __launch_bounds__(128) memcodes_quantize_kernel(const float* __restrict__ x, const float* __restrict__ kk, const float* __restrict__ vv,
                                                                int heads, int n_codes, long long n_pos, float scale, float* __restrict__ q_out,
                                                                float* __restrict__ resid_out, float* __restrict__ acc, int final_tanh,
                                                                long long* __restrict__ idx_out) {
  extern __shared__ float ks[];   // [n_codes][D] keys of this head
  const int h = blockIdx.y;
  const long long b = blockIdx.z;
  const float* kh = kk + (size_t)h * n_codes * D;
  for (int i = threadIdx.x; i < n_codes * D; i += blockDim.x) ks[i] = kh[i];
  __syncthreads();
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_pos) return;
  const size_t base = ((size_t)b * heads * D + (size_t)h * D) * n_pos + n;
  float q[D];
#pragma unroll
  for (int i = 0; i < D; ++i) q[i] = x[base + (size_t)i * n_pos] * scale;
  float best = -INFINITY;
  int bj = 0;
  for (int j = 0; j < n_codes; ++j) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) a = fmaf(q[i], ks[j * D + i], a);
    if (a > best) { best = a; bj = j; }
  }
  const float* vh = vv + ((size_t)h * n_codes + bj) * D;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const size_t o = base + (size_t)i * n_pos;
    const float qv = vh[i];
    if (q_out) q_out[o] = qv;
    if (resid_out) resid_out[o] = x[o] - qv;
    if (acc) {
      const float s = acc[o] + qv;
      acc[o] = final_tanh ? tanhf(s) : s;
    }
  }
  if (idx_out) idx_out[((size_t)b * heads + h) * n_pos + n] = bj;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int aa_pqmf_analysis_f32(const float* x, int64_t rows, int64_t n, const float* hk, int bands, int taps, float* out, void* stream) {
  AA_REQUIRE(x && hk && out, "NULL tensor");
  AA_REQUIRE(bands >= 2 && bands <= 64 && taps >= 2 && taps <= 4096, "bands=%d taps=%d out of range", bands, taps);
  AA_REQUIRE(rows >= 0 && rows <= 65535 && n >= 1, "rows=%lld must be <= 65535", (long long)rows);
  if (rows == 0) return AA_OK;
  // conv1d(stride = bands, padding = taps/2) gives (n + 2*(taps/2) - taps) / bands + 1 outputs; the reference drops the last one
  const long long t_out = (n + 2 * (taps / 2) - taps) / bands;
  AA_REQUIRE(t_out >= 1, "input too short for the filterbank");
  const int smem = (int)(sizeof(float) * ((size_t)kT * bands + taps + (size_t)bands * taps));
  AA_REQUIRE(smem <= 200 * 1024, "filterbank too large for shared memory (%d bytes)", smem);
  AA_CUDA(aa::ensure_dyn_smem(pqmf_analysis_kernel, std::max(smem, 48 * 1024)));
  dim3 grid((unsigned)((t_out + kT - 1) / kT), (unsigned)rows);
  pqmf_analysis_kernel<<<grid, kT, smem, (cudaStream_t)stream>>>(x, hk, bands, taps, n, t_out, out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_memcodes_kv_f32(const float* codes, const float* wk, const float* wv, int heads, int n_codes, int d, float* k_out, float* v_out, void* stream) {
  AA_REQUIRE(codes && wk && wv && k_out && v_out, "NULL tensor");
  AA_REQUIRE(heads >= 1 && n_codes >= 1 && d >= 1, "bad shape");
  const long long total = (long long)heads * n_codes * d;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)aa::num_sms() * 8);
  memcodes_kv_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(codes, wk, wv, heads, n_codes, d, k_out, v_out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_memcodes_quantize_f32(const float* x, int64_t batch, int heads, int d, int64_t n_pos, const float* k, const float* v, int n_codes, float scale,
                             float* q_out, float* resid_out, float* acc, int final_tanh, int64_t* idx_out, void* stream) {
  AA_REQUIRE(x && k && v, "NULL tensor");
  AA_REQUIRE(batch >= 0 && batch <= 65535 && heads >= 1 && heads <= 65535 && n_pos >= 1 && n_codes >= 1, "bad shape");
  AA_REQUIRE(d == 1 || d == 2 || d == 4 || d == 8 || d == 16 || d == 32 || d == 64, "dim / heads = %d: supported head widths are powers of two up to 64", d);
  if (batch == 0) return AA_OK;
  const int smem = (int)(sizeof(float) * (size_t)n_codes * d);
  AA_REQUIRE(smem <= 200 * 1024, "codebook of one head (%d bytes) does not fit shared memory", smem);
  dim3 grid((unsigned)((n_pos + 127) / 128), (unsigned)heads, (unsigned)batch);
  cudaStream_t st = (cudaStream_t)stream;
  static_assert(sizeof(long long) == sizeof(int64_t), "index type");
  long long* io = reinterpret_cast<long long*>(idx_out);
#define AA_MQ(DD)                                                                                              \
  do {                                                                                                         \
    AA_CUDA(aa::ensure_dyn_smem(memcodes_quantize_kernel<DD>, std::max(smem, 48 * 1024)));                     \
    memcodes_quantize_kernel<DD><<<grid, 128, smem, st>>>(x, k, v, heads, n_codes, n_pos, scale, q_out, resid_out, acc, final_tanh, io); \
  } while (0)
  switch (d) {
    case 1: AA_MQ(1); break;
    case 2: AA_MQ(2); break;
    case 4: AA_MQ(4); break;
    case 8: AA_MQ(8); break;
    case 16: AA_MQ(16); break;
    case 32: AA_MQ(32); break;
    default: AA_MQ(64); break;
  }
#undef AA_MQ
  AA_LAUNCH_CHECK();
  return AA_OK;
}

#pragma GCC visibility pop
}  // extern "C"
