// Remaining latent-space operations of Destructo.ipynb cells 22 / 49, the use_cos / debug branches of
// MagDPhaseSpectrogramAE.encode (given_models.py:214-231) and a standalone EmbedBlock (aa_mixer.py:205-221):
// single-pass kernels where the notebook runs one full-tensor ATen kernel per arithmetic op (the "reverb" cell runs
// T of them).  None of this is bandwidth critical at Destructo sizes ([8, 64, 512] latents); the point is that every
// row of the path has a device implementation behind the C ABI.
#include "aa_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdint>

namespace {

constexpr int kThreads = 256;

inline int grid_for(long long n) {
  return (int)std::max<long long>(1, std::min<long long>((n + kThreads - 1) / kThreads, (long long)aa::num_sms() * 8));
}

// z * vec[t]   ("wavy": vec = cos(linspace(0, 4*6.28, T)) computed by the caller exactly as the notebook does)
__global__ void mul_time_kernel(const float* __restrict__ z, const float* __restrict__ vec, float* __restrict__ out, long long n, int t) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = z[i] * vec[(int)(i % t)];
}
// z + flip(z, -1)   ("flippy")
__global__ void add_flip_time_kernel(const float* __restrict__ z, float* __restrict__ out, long long n, int t) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ti = (int)(i % t);
    out[i] = z[i] + z[i - ti + (t - 1 - ti)];
  }
}
// z[:, c0:c1, :] = 0   ("kill_half": c0 = 33, c1 = C - 1)
__global__ void zero_channels_kernel(const float* __restrict__ z, float* __restrict__ out, long long n, int c, int t, int c0, int c1) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)((i / t) % c);
    out[i] = (ci >= c0 && ci < c1) ? 0.f : z[i];
  }
}
// a * x + b * z * (2 u - 1)   ("call_and_response": x = z, a = -1, b = rand_fac; "hurt_drums": x = embeddings, a = 1 - rand_fac)
// evaluated in the notebook's order of operations: (a*x) + ((b*z) * (2*u - 1))
__global__ void randmix_kernel(const float* __restrict__ x, const float* __restrict__ z, const float* __restrict__ u, float a, float b,
                               float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float r = __fadd_rn(__fmul_rn(2.0f, u[i]), -1.0f);
    out[i] = __fadd_rn(__fmul_rn(a, x[i]), __fmul_rn(__fmul_rn(b, z[i]), r));
  }
}

// "reverb_time": for i in range(T): z = z + coef[i] * shift_right(z, i + 1), every step on the UPDATED z.
// One block per (b, c) row; the row ping-pongs between two shared-memory lines, one __syncthreads per step.
// Step i only changes samples t >= i + 1, with the same unfused multiply-then-add as the ATen ops.
__global__ void reverb_kernel(const float* __restrict__ z, const float* __restrict__ coef, float* __restrict__ out, int t) {
  extern __shared__ float sh[];
  float* cur = sh;
  float* nxt = sh + t;
  const float* row = z + (size_t)blockIdx.x * t;
  for (int i = threadIdx.x; i < t; i += blockDim.x) cur[i] = row[i];
  __syncthreads();
  for (int s = 0; s < t; ++s) {
    const float c = coef[s];
    const int d = s + 1;
    for (int i = threadIdx.x; i < t; i += blockDim.x) nxt[i] = (i >= d) ? __fadd_rn(cur[i], __fmul_rn(c, cur[i - d])) : cur[i];
    __syncthreads();
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  float* orow = out + (size_t)blockIdx.x * t;
  for (int i = threadIdx.x; i < t; i += blockDim.x) orow[i] = cur[i];
}

// Destructo.ipynb cell 49, time_avg = False branch with unequal lengths: diff [Bw][C][Td] is left-padded with zeros
// (F.pad(diff, (length_difference, 0, ...)) -- the notebook's comment says "end", the code pads the front) or truncated to
// the embedding length Tz, averaged over Bw and added to every batch entry.
__global__ void effect_transfer_ex_kernel(const float* __restrict__ emb, long long b, int c, int tz, const float* __restrict__ wet,
                                          const float* __restrict__ dry, long long bw, int td, float* __restrict__ out) {
  const long long ct = (long long)c * tz;
  const int shift = tz > td ? tz - td : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ct; i += (long long)gridDim.x * blockDim.x) {
    const int ti = (int)(i % tz);
    const long long ci = i / tz;
    const int tsrc = ti - shift;
    float d = 0.f;
    if (tsrc >= 0 && tsrc < td) {
      for (long long j = 0; j < bw; ++j) d += wet[(j * c + ci) * td + tsrc] - dry[(j * c + ci) * td + tsrc];
    }
    d /= (float)bw;
    for (long long j = 0; j < b; ++j) out[j * ct + i] = emb[j * ct + i] + d;
  }
}
// mean over the last axis of (wet - dry): one warp per row
__global__ void row_mean_diff_kernel(const float* __restrict__ wet, const float* __restrict__ dry, long long rows, int t, float* __restrict__ out) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc = 0.f;
  for (int i = lane; i < t; i += 32) acc += wet[row * t + i] - dry[row * t + i];
  acc = aa::warp_sum(acc);
  if (lane == 0) out[row] = acc / (float)t;
}
// out[b][c][t] = z[b][c][t] + d[c or 0][t or 0]   (torch broadcasting of a 2-D tensor against [B, C, T])
__global__ void add_bcast2_kernel(const float* __restrict__ z, long long n, int c, int t, const float* __restrict__ d, int d0, int d1,
                                  float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ti = (int)(i % t), ci = (int)((i / t) % c);
    out[i] = z[i] + d[(size_t)(d0 > 1 ? ci : 0) * d1 + (d1 > 1 ? ti : 0)];
  }
}

// MagDPhaseSpectrogramAE.encode (given_models.py:214-231), all branches: spec [c*F][T] complex64 ->
// out[0 .. cf*T) = |spec|, out[cf*T ..) = dtheta.  wrap_theta = the `debug` branch (theta < 0 -> theta + 2 pi BEFORE the
// difference); use_cos = acos of the clipped normalised dot product of consecutive frames; column 0 keeps theta.
__global__ void magdphase_ex_kernel(const float2* __restrict__ spec, long long cf, int n_frames, long long half_elems, int use_cos,
                                    int wrap_theta, float* __restrict__ out) {
  const float two_pi = 2.0f * 3.141592653589f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cf * n_frames; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % n_frames);
    const float2 v = spec[i];
    const float mag = hypotf(v.x, v.y);
    float th = atan2f(v.y, v.x);
    if (wrap_theta && th < 0.f) th += two_pi;
    float d = th;
    if (t > 0) {
      const float2 u = spec[i - 1];
      if (use_cos) {
        const float num = __fadd_rn(__fmul_rn(v.x, u.x), __fmul_rn(v.y, u.y));
        const float den = __fmul_rn(mag, hypotf(u.x, u.y));
        float a = (den == 0.f) ? 1.f : num / den;
        a = fminf(fmaxf(a, -1.f), 1.f);
        d = acosf(a);
      } else {
        float thp = atan2f(u.y, u.x);
        if (wrap_theta && thp < 0.f) thp += two_pi;
        d = th - thp;
        if (d < 0.f) d += two_pi;
      }
    }
    out[i] = mag;
    out[half_elems + i] = d;
  }
}

// ---- standalone EmbedBlock (aa_mixer.py:205-221): y = Linear(x); y = GELU_erf(y) if act; [BatchNorm outside]; y = x + y if resid ----
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float v) {
  const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
  return cdf + v * 0.39894228040143267794f * expf(-0.5f * v * v);
}
// one thread per (token, out feature); pre = Linear(x) saved for the backward when `pre_out` != NULL
__global__ void embed_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, long long n_tok,
                                 int din, int dout, int act, int resid, float* __restrict__ y, float* __restrict__ pre_out) {
  const long long total = n_tok * dout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long tok = i / dout;
    const int o = (int)(i % dout);
    const float* xr = x + tok * din;
    const float* wr = w + (size_t)o * din;
    float acc = bias ? bias[o] : 0.f;
    for (int k = 0; k < din; ++k) acc = fmaf(xr[k], wr[k], acc);
    if (pre_out) pre_out[i] = acc;
    float v = act ? gelu_erf(acc) : acc;
    if (resid) v += xr[o];
    y[i] = v;
  }
}
// gpre = gy * act'(pre);  gx[tok][k] = sum_o gpre[tok][o] w[o][k] (+ gy[tok][k] if resid)
__global__ void embed_bwd_pre_kernel(const float* __restrict__ gy, const float* __restrict__ pre, long long total, int act,
                                     float* __restrict__ gpre) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    gpre[i] = act ? gy[i] * gelu_erf_grad(pre[i]) : gy[i];
}
__global__ void embed_bwd_x_kernel(const float* __restrict__ gpre, const float* __restrict__ gy, const float* __restrict__ w, long long n_tok,
                                   int din, int dout, int resid, float* __restrict__ gx) {
  const long long total = n_tok * din;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long tok = i / din;
    const int k = (int)(i % din);
    float acc = resid ? gy[tok * dout + k] : 0.f;
    for (int o = 0; o < dout; ++o) acc = fmaf(gpre[tok * dout + o], w[(size_t)o * din + k], acc);
    gx[i] = acc;
  }
}
// gw[o][k] = sum_tok gpre[tok][o] x[tok][k];  gb[o] = sum_tok gpre[tok][o]: one block per (o, k) pair (k = din: the bias),
// fixed-order tree reduction (deterministic)
__global__ void embed_bwd_w_kernel(const float* __restrict__ gpre, const float* __restrict__ x, long long n_tok, int din, int dout,
                                   float* __restrict__ gw, float* __restrict__ gb) {
  __shared__ float sh[kThreads];
  const int o = blockIdx.x / (din + 1), k = blockIdx.x % (din + 1);
  float acc = 0.f;
  for (long long tok = threadIdx.x; tok < n_tok; tok += blockDim.x)
    acc = fmaf(gpre[tok * dout + o], k < din ? x[tok * din + k] : 1.0f, acc);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (k < din) gw[(size_t)o * din + k] = sh[0];
    else if (gb) gb[o] = sh[0];
  }
}

// ---- BatchNorm1d over [N, C] (the only input rank EmbedBlock(use_bn=True) accepts in the reference) ----
// one block per feature: batch mean / biased variance (training) or running statistics (eval); save[0][c] = mean, save[1][c] = rstd
__global__ void bn_fwd_kernel(const float* __restrict__ x, long long n, int c, const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* __restrict__ run_mean, float* __restrict__ run_var, int training, float momentum, float eps,
                              float* __restrict__ y, float* __restrict__ save) {
  __shared__ float sh[kThreads];
  __shared__ float s_mean, s_rstd;
  const int f = blockIdx.x;
  if (training) {
    float acc = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += x[i * c + f];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kThreads / 2; s > 0; s >>= 1) { if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s]; __syncthreads(); }
    const float mean = sh[0] / (float)n;
    __syncthreads();
    float q = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) { const float d = x[i * c + f] - mean; q = fmaf(d, d, q); }
    sh[threadIdx.x] = q;
    __syncthreads();
    for (int s = kThreads / 2; s > 0; s >>= 1) { if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) {
      const float var = sh[0] / (float)n;
      s_mean = mean;
      s_rstd = rsqrtf(var + eps);
      if (run_mean) run_mean[f] = (1.f - momentum) * run_mean[f] + momentum * mean;
      if (run_var) run_var[f] = (1.f - momentum) * run_var[f] + momentum * (n > 1 ? sh[0] / (float)(n - 1) : var);
    }
  } else if (threadIdx.x == 0) {
    s_mean = run_mean[f];
    s_rstd = rsqrtf(run_var[f] + eps);
  }
  __syncthreads();
  const float mean = s_mean, rstd = s_rstd, g = gamma ? gamma[f] : 1.f, b = beta ? beta[f] : 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) y[i * c + f] = (x[i * c + f] - mean) * rstd * g + b;
  if (threadIdx.x == 0 && save) { save[f] = mean; save[c + f] = rstd; }
}
// training-mode backward: gx = g rstd (gy - mean(gy) - xhat mean(gy xhat)); ggamma = sum gy xhat; gbeta = sum gy
// eval-mode backward (training = 0): gx = gy g rstd
__global__ void bn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, long long n, int c, const float* __restrict__ gamma,
                              const float* __restrict__ save, int training, float* __restrict__ gx, float* __restrict__ ggamma,
                              float* __restrict__ gbeta) {
  __shared__ float sh[kThreads], sh2[kThreads];
  const int f = blockIdx.x;
  const float mean = save[f], rstd = save[c + f], g = gamma ? gamma[f] : 1.f;
  float a = 0.f, b = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float xh = (x[i * c + f] - mean) * rstd, d = gy[i * c + f];
    a += d;
    b = fmaf(d, xh, b);
  }
  sh[threadIdx.x] = a; sh2[threadIdx.x] = b;
  __syncthreads();
  for (int s = kThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { sh[threadIdx.x] += sh[threadIdx.x + s]; sh2[threadIdx.x] += sh2[threadIdx.x + s]; }
    __syncthreads();
  }
  const float sum_gy = sh[0], sum_gyxh = sh2[0];
  if (threadIdx.x == 0) {
    if (ggamma) ggamma[f] = sum_gyxh;
    if (gbeta) gbeta[f] = sum_gy;
  }
  if (gx) {
    const float m1 = sum_gy / (float)n, m2 = sum_gyxh / (float)n;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const float xh = (x[i * c + f] - mean) * rstd, d = gy[i * c + f];
      gx[i * c + f] = training ? g * rstd * (d - m1 - xh * m2) : d * g * rstd;
    }
  }
}

// ---- GroupNorm (+ SiLU) over channel-major [B][C][L]: the C/G channels of a group are one contiguous run of (C/G)*L floats ----
// one block per (batch, group): mean, then centred second moment (two passes over data that sits in L2), then normalise,
// per-channel affine and x * sigmoid(x)
__global__ void __launch_bounds__(kThreads) groupnorm_act_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, int c, long long l, int groups, float eps,
                                                                 int silu, float* __restrict__ out) {
  __shared__ float sh[kThreads];
  __shared__ float s_mean, s_rstd;
  const int g = blockIdx.x % groups;
  const long long b = blockIdx.x / groups;
  const int cpg = c / groups;
  const long long n = (long long)cpg * l;
  const float* xg = x + (b * c + (long long)g * cpg) * l;
  float* og = out + (b * c + (long long)g * cpg) * l;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += kThreads) acc += xg[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s_ = kThreads / 2; s_ > 0; s_ >>= 1) { if ((int)threadIdx.x < s_) sh[threadIdx.x] += sh[threadIdx.x + s_]; __syncthreads(); }
  if (threadIdx.x == 0) s_mean = sh[0] / (float)n;
  __syncthreads();
  const float mean = s_mean;
  float q = 0.f;
  for (long long i = threadIdx.x; i < n; i += kThreads) { const float d = xg[i] - mean; q = fmaf(d, d, q); }
  __syncthreads();
  sh[threadIdx.x] = q;
  __syncthreads();
  for (int s_ = kThreads / 2; s_ > 0; s_ >>= 1) { if ((int)threadIdx.x < s_) sh[threadIdx.x] += sh[threadIdx.x + s_]; __syncthreads(); }
  if (threadIdx.x == 0) s_rstd = rsqrtf(sh[0] / (float)n + eps);
  __syncthreads();
  const float rstd = s_rstd;
  for (long long i = threadIdx.x; i < n; i += kThreads) {
    const int ch = g * cpg + (int)(i / l);
    float v = (xg[i] - mean) * rstd * (gamma ? gamma[ch] : 1.f) + (beta ? beta[ch] : 0.f);
    if (silu) v = v / (1.0f + expf(-v));
    og[i] = v;
  }
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int aa_latent_mul_time_f32(const float* z, const float* vec, float* out, int64_t b, int64_t c, int64_t t, void* stream) {
  AA_REQUIRE(z && vec && out, "NULL tensor");
  AA_REQUIRE(b >= 0 && c >= 1 && t >= 1 && t < (1LL << 31), "bad shape");
  const long long n = b * c * t;
  if (n == 0) return AA_OK;
  mul_time_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(z, vec, out, n, (int)t);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_latent_add_flip_time_f32(const float* z, float* out, int64_t b, int64_t c, int64_t t, void* stream) {
  AA_REQUIRE(z && out && z != out, "NULL tensor or in-place call");
  AA_REQUIRE(b >= 0 && c >= 1 && t >= 1 && t < (1LL << 31), "bad shape");
  const long long n = b * c * t;
  if (n == 0) return AA_OK;
  add_flip_time_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(z, out, n, (int)t);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_latent_zero_channels_f32(const float* z, float* out, int64_t b, int64_t c, int64_t t, int64_t c0, int64_t c1, void* stream) {
  AA_REQUIRE(z && out, "NULL tensor");
  AA_REQUIRE(b >= 0 && c >= 1 && c < (1LL << 31) && t >= 1 && t < (1LL << 31), "bad shape");
  const long long n = b * c * t;
  if (n == 0) return AA_OK;
  zero_channels_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(z, out, n, (int)c, (int)t, (int)std::max<int64_t>(c0, 0),
                                                                            (int)std::min<int64_t>(c1, c));
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_latent_randmix_f32(const float* x, const float* z, const float* u, float a, float b, float* out, int64_t n, void* stream) {
  AA_REQUIRE(x && z && u && out, "NULL tensor");
  if (n <= 0) return AA_OK;
  randmix_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(x, z, u, a, b, out, n);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_latent_reverb_f32(const float* z, const float* coef, float* out, int64_t rows, int64_t t, void* stream) {
  AA_REQUIRE(z && coef && out, "NULL tensor");
  AA_REQUIRE(rows >= 0 && rows < (1LL << 31) && t >= 1 && t <= 16384, "rows=%lld t=%lld: t must be in [1, 16384]", (long long)rows, (long long)t);
  if (rows == 0) return AA_OK;
  const int smem = (int)(2 * t * sizeof(float));
  AA_CUDA(aa::ensure_dyn_smem(reverb_kernel, std::max(smem, 48 * 1024)));
  reverb_kernel<<<(unsigned)rows, (unsigned)std::min<int64_t>(1024, ((t + 31) / 32) * 32), smem, (cudaStream_t)stream>>>(z, coef, out, (int)t);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_effect_transfer_ex_f32(const float* emb, int64_t b, int64_t c, int64_t t_emb, const float* wet, const float* dry, int64_t bw,
                              int64_t t_diff, float* out, void* stream) {
  AA_REQUIRE(emb && wet && dry && out, "NULL tensor");
  AA_REQUIRE(bw >= 1 && b >= 0 && c >= 1 && c < (1LL << 31) && t_emb >= 1 && t_diff >= 1 && t_emb < (1LL << 31) && t_diff < (1LL << 31), "bad shape");
  if (b == 0) return AA_OK;
  effect_transfer_ex_kernel<<<grid_for(c * t_emb), kThreads, 0, (cudaStream_t)stream>>>(emb, b, (int)c, (int)t_emb, wet, dry, bw, (int)t_diff, out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_latent_row_mean_diff_f32(const float* wet, const float* dry, int64_t rows, int64_t t, float* out, void* stream) {
  AA_REQUIRE(wet && dry && out, "NULL tensor");
  AA_REQUIRE(rows >= 0 && t >= 1 && t < (1LL << 31), "bad shape");
  if (rows == 0) return AA_OK;
  row_mean_diff_kernel<<<(unsigned)((rows * 32 + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(wet, dry, rows, (int)t, out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_latent_add_bcast2_f32(const float* z, int64_t b, int64_t c, int64_t t, const float* d, int64_t d0, int64_t d1, float* out, void* stream) {
  AA_REQUIRE(z && d && out, "NULL tensor");
  AA_REQUIRE(b >= 0 && c >= 1 && t >= 1 && c < (1LL << 31) && t < (1LL << 31), "bad shape");
  AA_REQUIRE((d0 == 1 || d0 == c) && (d1 == 1 || d1 == t), "a [%lld, %lld] tensor does not broadcast against [.., %lld, %lld]", (long long)d0,
             (long long)d1, (long long)c, (long long)t);
  const long long n = b * c * t;
  if (n == 0) return AA_OK;
  add_bcast2_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(z, n, (int)c, (int)t, d, (int)d0, (int)d1, out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_magdphase_ex_f32(const float* spec, int64_t c, int64_t n_freq, int64_t n_frames, int use_cos, int wrap_theta, float* out, void* stream) {
  AA_REQUIRE(spec && out, "NULL tensor pointer");
  AA_REQUIRE(c >= 1 && n_freq >= 1 && n_frames >= 1 && n_frames < (1LL << 31), "bad shape");
  const long long cf = c * n_freq, total = cf * n_frames;
  magdphase_ex_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(spec), cf, (int)n_frames, total, use_cos,
                                                                              wrap_theta, out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_embed_block_fwd_f32(const float* x, const float* w, const float* bias, int64_t n_tok, int din, int dout, int act, int resid, float* y,
                           float* pre_out, void* stream) {
  AA_REQUIRE(x && w && y, "NULL tensor");
  AA_REQUIRE(n_tok >= 0 && din >= 1 && dout >= 1, "bad shape");
  AA_REQUIRE(!resid || din == dout, "a residual block needs in_dims == out_dims");
  if (n_tok == 0) return AA_OK;
  embed_fwd_kernel<<<grid_for(n_tok * dout), kThreads, 0, (cudaStream_t)stream>>>(x, w, bias, n_tok, din, dout, act, resid, y, pre_out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

/* scratch: n_tok * dout floats */
int aa_embed_block_bwd_f32(const float* x, const float* w, const float* pre, const float* gy, int64_t n_tok, int din, int dout, int act, int resid,
                           float* gx, float* gw, float* gb, float* scratch, void* stream) {
  AA_REQUIRE(x && w && pre && gy && scratch, "NULL tensor");
  AA_REQUIRE(n_tok >= 0 && din >= 1 && dout >= 1, "bad shape");
  if (n_tok == 0) return AA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  embed_bwd_pre_kernel<<<grid_for(n_tok * dout), kThreads, 0, st>>>(gy, pre, n_tok * dout, act, scratch);
  AA_LAUNCH_CHECK();
  if (gx) {
    embed_bwd_x_kernel<<<grid_for(n_tok * din), kThreads, 0, st>>>(scratch, gy, w, n_tok, din, dout, resid, gx);
    AA_LAUNCH_CHECK();
  }
  if (gw || gb) {
    AA_REQUIRE(gw != nullptr, "gw is NULL");
    embed_bwd_w_kernel<<<(unsigned)(dout * (din + 1)), kThreads, 0, st>>>(scratch, x, n_tok, din, dout, gw, gb);
    AA_LAUNCH_CHECK();
  }
  return AA_OK;
}

int aa_batchnorm_fwd_f32(const float* x, int64_t n, int c, const float* gamma, const float* beta, float* run_mean, float* run_var, int training,
                         float momentum, float eps, float* y, float* save, void* stream) {
  AA_REQUIRE(x && y && save, "NULL tensor");
  AA_REQUIRE(n >= 1 && c >= 1, "bad shape");
  AA_REQUIRE(training || (run_mean && run_var), "eval mode needs running statistics");
  bn_fwd_kernel<<<(unsigned)c, kThreads, 0, (cudaStream_t)stream>>>(x, n, c, gamma, beta, run_mean, run_var, training, momentum, eps, y, save);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_batchnorm_bwd_f32(const float* x, const float* gy, int64_t n, int c, const float* gamma, const float* save, int training, float* gx,
                         float* ggamma, float* gbeta, void* stream) {
  AA_REQUIRE(x && gy && save, "NULL tensor");
  AA_REQUIRE(n >= 1 && c >= 1, "bad shape");
  bn_bwd_kernel<<<(unsigned)c, kThreads, 0, (cudaStream_t)stream>>>(x, gy, n, c, gamma, save, training, gx, ggamma, gbeta);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_groupnorm_act_f32(const float* x, int64_t batch, int c, int64_t l, const float* gamma, const float* beta, int groups, float eps, int silu,
                         float* out, void* stream) {
  AA_REQUIRE(x && out, "NULL tensor");
  AA_REQUIRE(batch >= 0 && c >= 1 && l >= 1 && groups >= 1 && c % groups == 0, "channels %d not divisible by %d groups", c, groups);
  AA_REQUIRE(batch * groups < (1LL << 31), "too many (batch, group) pairs");
  if (batch == 0) return AA_OK;
  groupnorm_act_kernel<<<(unsigned)(batch * groups), kThreads, 0, (cudaStream_t)stream>>>(x, gamma, beta, c, l, groups, eps, silu, out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

#pragma GCC visibility pop
}  // extern "C"
