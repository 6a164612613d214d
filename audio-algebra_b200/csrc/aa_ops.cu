// Latent algebra, losses, PCA scatter accumulation and Adam for the audio-algebra hot path (sm_100a).
//
// Replaces, with fused / deterministic kernels:
//   aa_mixer.py:307 (zsum), train_aa_effects.py:70-71 (effect guesses), Destructo.ipynb cells 22,48-49
//   aa_mixer.py:344 mseloss, :351-353 vicreg_var_loss (+ L2 hinge train_aa_effects.py:42-46),
//   aa_mixer.py:355-364 off_diagonal / vicreg_cov_loss  -> Gram identity, no DxD covariance,
//   calc_effects_pca.py:81-89 running covariance numerator, train_aa_mixer_accel.py:481,533 Adam.
// All reductions are two-stage with a fixed summation order (no float atomics): same bits every run.
#include "aa_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstdint>
#include <cmath>

namespace {

constexpr int kRedThreads = 256;
constexpr int kMaxParts = 2048;   // partial sums per reduction

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = aa::warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) r = aa::warp_sum(r);
  __syncthreads();
  return r;   // valid in warp 0
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  v = aa::warp_max(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : -INFINITY;
  if (w == 0) r = aa::warp_max(r);
  __syncthreads();
  return r;
}

// out[0] = scale * sum(parts[0..n))   (one block, fixed order)
__global__ void finalize_sum_kernel(const float* __restrict__ parts, int n, float scale, float* __restrict__ out,
                                    const float* __restrict__ sub, int n_sub, float sub_scale) {
  __shared__ float sh[32];
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += parts[i];
  float s = block_sum(v, sh);
  float u = 0.f;
  if (sub != nullptr) {
    float w = 0.f;
    for (int i = threadIdx.x; i < n_sub; i += blockDim.x) w += sub[i];
    u = block_sum(w, sh);
  }
  if (threadIdx.x == 0) out[0] = scale * s - sub_scale * u;
}

// ------------------------------------------------------------------------------------------ algebra
struct LinArgs {
  const float* z[8];
  float c[8];
  int n_terms;
};
__global__ void lincomb_kernel(LinArgs a, float* __restrict__ out, long long n) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < a.n_terms) {
        const float4 v = reinterpret_cast<const float4*>(a.z[j])[i];
        acc.x = fmaf(a.c[j], v.x, acc.x); acc.y = fmaf(a.c[j], v.y, acc.y);
        acc.z = fmaf(a.c[j], v.z, acc.z); acc.w = fmaf(a.c[j], v.w, acc.w);
      }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < a.n_terms; ++j) acc = fmaf(a.c[j], a.z[j][i], acc);
    out[i] = acc;
  }
}
__global__ void lincomb_scalar_kernel(LinArgs a, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < a.n_terms; ++j) acc = fmaf(a.c[j], a.z[j][i], acc);
    out[i] = acc;
  }
}

// parts[2*blk] = max(z), parts[2*blk+1] = max|z| over the block's slice
__global__ void minmax_part_kernel(const float* __restrict__ z, long long n, float* __restrict__ parts) {
  __shared__ float sh[32];
  float m = -INFINITY, am = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = z[i];
    m = fmaxf(m, v);
    am = fmaxf(am, fabsf(v));
  }
  m = block_max(m, sh);
  am = block_max(am, sh);
  if (threadIdx.x == 0) { parts[2 * blockIdx.x] = m; parts[2 * blockIdx.x + 1] = am; }
}
__global__ void minmax_final_kernel(const float* __restrict__ parts, int n, float* __restrict__ out2) {
  __shared__ float sh[32];
  float m = -INFINITY, am = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { m = fmaxf(m, parts[2 * i]); am = fmaxf(am, parts[2 * i + 1]); }
  m = block_max(m, sh);
  am = block_max(am, sh);
  if (threadIdx.x == 0) { out2[0] = m; out2[1] = am; }
}
__global__ void unary_kernel(int op, const float* __restrict__ z, float* __restrict__ out, long long n, float param,
                             const float* __restrict__ mm) {
  const float zmax = mm[0], zabsmax = mm[1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = z[i];
    float r;
    if (op == AA_UNARY_SIGN_FOLD) {
      const float sg = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
      r = zmax * (sg - v);
    } else if (op == AA_UNARY_ABSMAX_MINUS) {
      r = zabsmax - v;
    } else {
      r = zmax * tanhf(param * v);
    }
    out[i] = r;
  }
}
__global__ void flip_kernel(int op, const float* __restrict__ z, float* __restrict__ out, long long b, int c, int t) {
  const long long n = b * c * (long long)t;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ti = (int)(i % t);
    const long long bc = i / t;
    const int ci = (int)(bc % c);
    const long long bi = bc / c;
    const long long src = (op == AA_UNARY_FLIP_CHANNELS) ? ((bi * c + (c - 1 - ci)) * t + ti) : ((bi * c + ci) * t + (t - 1 - ti));
    out[i] = z[src];
  }
}
// out[b][i] = emb[b][i] + mean_b'(wet[b'][i] - dry[b'][i])
__global__ void effect_transfer_kernel(const float* __restrict__ emb, long long b, const float* __restrict__ wet,
                                       const float* __restrict__ dry, long long bw, long long ct, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ct; i += (long long)gridDim.x * blockDim.x) {
    float d = 0.f;
    for (long long j = 0; j < bw; ++j) d += wet[j * ct + i] - dry[j * ct + i];
    d /= (float)bw;
    for (long long j = 0; j < b; ++j) out[j * ct + i] = emb[j * ct + i] + d;
  }
}

// ------------------------------------------------------------------------------------------ mse
__global__ void mse_part_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ parts) {
  __shared__ float sh[32];
  float acc = 0.f;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    acc = fmaf(d0, d0, acc); acc = fmaf(d1, d1, acc); acc = fmaf(d2, d2, acc); acc = fmaf(d3, d3, acc);
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    acc = fmaf(d, d, acc);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) parts[blockIdx.x] = acc;
}
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, const float* __restrict__ gloss,
                               float gscale, float* __restrict__ ga, float* __restrict__ gb, int accumulate) {
  const float k = 2.0f * gscale * (gloss ? gloss[0] : 1.0f) / (float)n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = k * (a[i] - b[i]);
    if (ga) ga[i] = accumulate ? ga[i] + g : g;
    if (gb) gb[i] = accumulate ? gb[i] - g : -g;
  }
}

// ------------------------------------------------------------------------------------------ variance loss
// A block owns 32 feature columns (lane = column, coalesced 128-byte rows); its 8 warps split the batch into 8 contiguous slices.  Every slice accumulates sum and sum of squares of (v - shift) with the column's first sample
// as shift (4 independent chains: the loop is bandwidth bound, not latency bound like a per-sample Welford update); the slices
// are merged in a fixed order, so the result is deterministic.  mean = shift + S/n, M2 = Q - S^2/n.
__global__ void __launch_bounds__(kRedThreads) var_stats_kernel(const float* __restrict__ z, int b, long long d, float gamma, float eps,
                                                                int hinge_l2, float* __restrict__ stats, float* __restrict__ parts) {
  __shared__ float sh[32];
  __shared__ float sS[8][32], sQ[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i0 = (int)(((long long)b * w) / 8), i1 = (int)(((long long)b * (w + 1)) / 8);
  float hsum = 0.f;
  {
    const long long col = blockIdx.x * 32LL + lane;
    const bool ok = col < d;
    const float* p = z + (ok ? col : 0);
    const float shift = ok ? __ldg(p) : 0.f;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
    int i = i0;
    if (ok) {
#pragma unroll 2
      for (; i + 3 < i1; i += 4) {
        const float v0 = __ldg(p + (long long)i * d) - shift, v1 = __ldg(p + (long long)(i + 1) * d) - shift;
        const float v2 = __ldg(p + (long long)(i + 2) * d) - shift, v3 = __ldg(p + (long long)(i + 3) * d) - shift;
        s0 += v0; s1 += v1; s2 += v2; s3 += v3;
        q0 = fmaf(v0, v0, q0); q1 = fmaf(v1, v1, q1); q2 = fmaf(v2, v2, q2); q3 = fmaf(v3, v3, q3);
      }
      for (; i < i1; ++i) {
        const float v0 = __ldg(p + (long long)i * d) - shift;
        s0 += v0; q0 = fmaf(v0, v0, q0);
      }
    }
    sS[w][lane] = (s0 + s1) + (s2 + s3);
    sQ[w][lane] = (q0 + q1) + (q2 + q3);
    __syncthreads();
    if (w == 0 && ok) {
      float S = 0.f, Q = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { S += sS[k][lane]; Q += sQ[k][lane]; }
      const float mean_s = S / (float)b;
      const float m2 = fmaxf(Q - S * mean_s, 0.f);
      const float var = (b > 1) ? m2 / (float)(b - 1) : 0.f;
      if (stats) { stats[col] = shift + mean_s; stats[d + col] = var; }
      const float r = fmaxf(gamma - sqrtf(var + eps), 0.f);
      hsum += hinge_l2 ? r * r : r;
    }
    __syncthreads();
  }
  const float h = block_sum(hsum, sh);
  if (parts && threadIdx.x == 0) parts[blockIdx.x] = h;
}
__global__ void var_bwd_kernel(const float* __restrict__ z, const float* __restrict__ stats, int b, long long d, float gamma,
                               float eps, int hinge_l2, const float* __restrict__ gloss, float gscale,
                               float* __restrict__ gz, int accumulate) {
  const float g0 = gscale * (gloss ? gloss[0] : 1.0f) / (float)d;
  const long long n = (long long)b * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long col = i % d;
    const float mean = stats[col], var = stats[d + col];
    const float sd = sqrtf(var + eps);
    const float r = gamma - sd;
    float g = 0.f;
    if (r > 0.f) {
      const float dl = hinge_l2 ? -2.0f * r : -1.0f;                 // dL/dstd (before the 1/D)
      g = g0 * dl * (z[i] - mean) / ((float)(b - 1) * sd);           // dstd/dz = (z-mean)/((B-1) std)
    }
    gz[i] = accumulate ? gz[i] + g : g;
  }
}

// ------------------------------------------------------------------------------------------ covariance loss
// G = Xc Xc^T  (Xc = z - column mean), [B][B], split over D with fixed-order partial tiles.
constexpr int GT = 64, GK = 16;   // output tile, K step; 256 threads, 4x4 micro tile
__global__ void __launch_bounds__(256) gram_part_kernel(const float* __restrict__ z, const float* __restrict__ mean, int b,
                                                        long long d, long long d_per_split, float* __restrict__ parts) {
  __shared__ float As[GK][GT + 4], Bs[GK][GT + 4];
  const int ti = blockIdx.x, tj = blockIdx.y, split = blockIdx.z;
  if (tj > ti) return;   // lower triangle only; mirrored by the reducer
  const long long k0 = split * d_per_split, k1 = min(d, k0 + d_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;   // loader: row 0..63, 4 consecutive k
  for (long long k = k0; k < k1; k += GK) {
    {
      const int ra = ti * GT + lr, rb = tj * GT + lr;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long kk = k + lk + e;
        const bool okk = kk < k1;
        const float mu = okk ? mean[kk] : 0.f;
        As[lk + e][lr] = (okk && ra < b) ? z[(long long)ra * d + kk] - mu : 0.f;
        Bs[lk + e][lr] = (okk && rb < b) ? z[(long long)rb * d + kk] - mu : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { av[e] = As[kk][ty * 4 + e]; bv[e] = Bs[kk][tx * 4 + e]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* p = parts + (long long)split * b * b;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = ti * GT + ty * 4 + i, c = tj * GT + tx * 4 + j;
      if (r < b && c < b) p[(long long)r * b + c] = acc[i][j];
    }
}
// gram[r][c] = sum_s parts[s][max][min]; parts2[blk] = sum of gram^2 over the block's elements
__global__ void gram_reduce_kernel(const float* __restrict__ parts, int n_split, int b, float* __restrict__ gram,
                                   float* __restrict__ sq_parts) {
  __shared__ float sh[32];
  float acc2 = 0.f;
  const long long n = (long long)b * b;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / b), c = (int)(i % b);
    const long long src = (r >= c) ? (long long)r * b + c : (long long)c * b + r;
    float g = 0.f;
    for (int s = 0; s < n_split; ++s) g += parts[(long long)s * n + src];
    gram[i] = g;
    acc2 = fmaf(g, g, acc2);
  }
  acc2 = block_sum(acc2, sh);
  if (threadIdx.x == 0) sq_parts[blockIdx.x] = acc2;
}
// parts[blk] = sum_d ((B-1) var_d)^2 over the block's columns
__global__ void diag_sq_kernel(const float* __restrict__ stats, int b, long long d, float* __restrict__ parts) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < d; i += (long long)gridDim.x * blockDim.x) {
    const float s = (float)(b - 1) * stats[d + i];
    acc = fmaf(s, s, acc);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) parts[blockIdx.x] = acc;
}
// grad[r][col] (+)= k * (sum_j G[r][j] Xc[j][col] - s_col Xc[r][col]);  tile: 64 rows x 64 cols, K = B
__global__ void __launch_bounds__(256) cov_bwd_kernel(const float* __restrict__ z, const float* __restrict__ stats,
                                                      const float* __restrict__ gram, int b, long long d,
                                                      const float* __restrict__ gloss, float gscale, float* __restrict__ gz,
                                                      int accumulate) {
  __shared__ float Gs[GK][GT + 4], Xs[GK][GT + 4];
  const int tr = blockIdx.y;            // row tile
  const long long c0 = blockIdx.x * (long long)GT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k = 0; k < b; k += GK) {
    {  // G tile: rows tr*64.., k..k+15  -> Gs[kk][row];  X tile: rows k..k+15, cols c0..  -> Xs[kk][col]
      const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;
      const int r = tr * GT + lr;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kk = k + lk + e;
        Gs[lk + e][lr] = (r < b && kk < b) ? gram[(long long)r * b + kk] : 0.f;
      }
      const int xr = threadIdx.x >> 4, xc = (threadIdx.x & 15) * 4;
      const int kr = k + xr;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long col = c0 + xc + e;
        Xs[xr][xc + e] = (kr < b && col < d) ? z[(long long)kr * d + col] - stats[col] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float gv[4], xv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { gv[e] = Gs[kk][ty * 4 + e]; xv[e] = Xs[kk][tx * 4 + e]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], xv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float kf = 4.0f * gscale * (gloss ? gloss[0] : 1.0f) / ((float)(b - 1) * (float)(b - 1) * (float)d);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = tr * GT + ty * 4 + i;
      const long long col = c0 + tx * 4 + j;
      if (r < b && col < d) {
        const float xc = z[(long long)r * d + col] - stats[col];
        const float s = (float)(b - 1) * stats[d + col];
        const float g = kf * (acc[i][j] - s * xc);
        const long long o = (long long)r * d + col;
        gz[o] = accumulate ? gz[o] + g : g;
      }
    }
}

// ------------------------------------------------------------------------------------------ PCA scatter
// sums[c] over all (b, t) of y[b][c][t]  (one block per (c, slice); fixed-order second stage on host-free path)
__global__ void chan_sum_part_kernel(const float* __restrict__ y, long long b, int c, long long t, float* __restrict__ parts,
                                     int slices) {
  // block (ch, slice): the rows y[bi][ch][0..t) of every slices-th batch element, read as contiguous runs (no div / mod per element)
  __shared__ float sh[32];
  const int ch = blockIdx.x, sl = blockIdx.y;
  float acc0 = 0.f, acc1 = 0.f;
  for (long long bi = sl; bi < b; bi += slices) {
    const float* row = y + (bi * c + ch) * t;
    if ((t & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
      const float4* r4 = reinterpret_cast<const float4*>(row);
      for (long long i = threadIdx.x; i < t / 4; i += blockDim.x) {
        const float4 v = __ldg(r4 + i);
        acc0 += v.x + v.y;
        acc1 += v.z + v.w;
      }
    } else {
      for (long long i = threadIdx.x; i < t; i += blockDim.x) acc0 += __ldg(row + i);
    }
  }
  const float acc = block_sum(acc0 + acc1, sh);
  if (threadIdx.x == 0) parts[ch * slices + sl] = acc;
}
__global__ void chan_mean_kernel(const float* __restrict__ parts, int c, int slices, double n, float* __restrict__ mean) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch < c) {
    double s = 0.0;
    for (int i = 0; i < slices; ++i) s += (double)parts[ch * slices + i];
    mean[ch] = (float)(s / n);
  }
}
// each block: a run of (b, t-chunk) tiles; 64x64 channel tile of the scatter; writes its partial
constexpr int PT = 32;   // points per smem tile
constexpr int PLD = 64 + 4;   // leading dim of the staged [point][channel] tiles: rows stay 16-byte aligned for float4 loads
__global__ void __launch_bounds__(256) scatter_part_kernel(const float* __restrict__ y, long long b, int c, long long t,
                                                           const float* __restrict__ mean, float* __restrict__ parts) {
  __shared__ __align__(16) float Ys[PT][PLD];
  __shared__ __align__(16) float Yt[PT][PLD];
  const int ci = blockIdx.y, cj = blockIdx.z;   // 64-channel tiles
  const bool same = (ci == cj);                 // diagonal tile (always, for C <= 64): one staged copy serves both operands
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  const long long tiles_t = (t + PT - 1) / PT, n_tiles = b * tiles_t;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long bi = tile / tiles_t, t0 = (tile % tiles_t) * PT;
    for (int e = threadIdx.x; e < 64 * PT; e += 256) {
      const int ch = e / PT, p = e % PT;
      const long long tt = t0 + p;
      const int ca = ci * 64 + ch, cb = cj * 64 + ch;
      Ys[p][ch] = (tt < t && ca < c) ? __ldg(y + (bi * c + ca) * t + tt) - mean[ca] : 0.f;
      if (!same) Yt[p][ch] = (tt < t && cb < c) ? __ldg(y + (bi * c + cb) * t + tt) - mean[cb] : 0.f;
    }
    __syncthreads();
    const float (*Yb)[PLD] = same ? Ys : Yt;
#pragma unroll 8
    for (int p = 0; p < PT; ++p) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Ys[p][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Yb[p][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* p = parts + (long long)blockIdx.x * c * c;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = ci * 64 + ty * 4 + i, cc = cj * 64 + tx * 4 + j;
      if (r < c && cc < c) p[(long long)r * c + cc] = acc[i][j];
    }
}
// cov_num[i] += sum_p parts[p][i] in a fixed order: a block owns 32 outputs (lane), its 8 warps split the partials into 8
// contiguous slices (coalesced 128-byte rows, 4 independent chains), the slices are merged in warp order.
__global__ void __launch_bounds__(256) scatter_reduce_kernel(const float* __restrict__ parts, int n_parts, int c,
                                                             float* __restrict__ cov_num, double* __restrict__ count, double n_points) {
  __shared__ float sS[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane, cc = c * c;
  const int p0 = (int)(((long long)n_parts * w) / 8), p1 = (int)(((long long)n_parts * (w + 1)) / 8);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < cc) {
    int p = p0;
    for (; p + 3 < p1; p += 4) {
      s0 += __ldg(parts + (long long)p * cc + i);
      s1 += __ldg(parts + (long long)(p + 1) * cc + i);
      s2 += __ldg(parts + (long long)(p + 2) * cc + i);
      s3 += __ldg(parts + (long long)(p + 3) * cc + i);
    }
    for (; p < p1; ++p) s0 += __ldg(parts + (long long)p * cc + i);
  }
  sS[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && i < cc) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sS[k][lane];
    cov_num[i] += s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && count) count[0] += n_points;
}

// ------------------------------------------------------------------------------------------ Adam
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;       // torch.optim.Adam: sqrt(v)/sqrt(1-b2^t) + eps
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}

inline int grid_for(long long n, int per_block = kRedThreads) {
  const long long g = (n + per_block - 1) / per_block;
  return (int)std::max<long long>(1, std::min<long long>(g, (long long)aa::num_sms() * 8));
}
inline int parts_for(long long n) {
  return (int)std::max<long long>(1, std::min<long long>((n + 4 * kRedThreads - 1) / (4 * kRedThreads), kMaxParts));
}

}  // namespace

namespace aa {   // csrc/cov_tc.cu
bool cov_tc_eligible(const float* z, const float* stats, int64_t b, int64_t d);
int cov_tc_splits(int64_t b, int64_t d);
int gram_bb_tc(const float* z, const float* mean, int64_t b, int64_t d, float* parts, int splits, cudaStream_t stream);
int cov_bwd_tc(const float* z, const float* stats, const float* gram, int64_t b, int64_t d, const float* gloss, float gscale, float* gz,
               int accumulate, cudaStream_t stream);
}
namespace aa {   // csrc/gram_tc.cu
int64_t gram_tc_workspace_floats();
int gram_tc(const float* y, int64_t b, int64_t t, float* cov_num, double* count, float* workspace, cudaStream_t stream);
}

extern "C" {
#pragma GCC visibility push(default)

int64_t aa_reduce_workspace_floats(void) { return 4 * kMaxParts + 64; }

int aa_latent_lincomb_f32(int n_terms, const float* const* zs_host, const float* coeffs_host, float* out, int64_t n,
                          void* stream) {
  AA_REQUIRE(n_terms >= 1 && n_terms <= 8, "n_terms=%d must be in [1, 8]", n_terms);
  AA_REQUIRE(zs_host && coeffs_host && out, "NULL argument");
  if (n == 0) return AA_OK;
  LinArgs a;
  bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int j = 0; j < 8; ++j) {
    a.z[j] = j < n_terms ? zs_host[j] : nullptr;
    a.c[j] = j < n_terms ? coeffs_host[j] : 0.f;
    if (j < n_terms) {
      AA_REQUIRE(zs_host[j] != nullptr, "zs[%d] is NULL", j);
      aligned = aligned && (reinterpret_cast<uintptr_t>(zs_host[j]) & 15) == 0;
    }
  }
  a.n_terms = n_terms;
  if (aligned) lincomb_kernel<<<grid_for(n / 4 + 1), kRedThreads, 0, (cudaStream_t)stream>>>(a, out, n);
  else lincomb_scalar_kernel<<<grid_for(n), kRedThreads, 0, (cudaStream_t)stream>>>(a, out, n);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_latent_unary_f32(int op, const float* z, float* out, int64_t b, int64_t c, int64_t t, float param, float* workspace,
                        void* stream) {
  AA_REQUIRE(z && out, "NULL tensor");
  AA_REQUIRE(b >= 0 && c >= 1 && t >= 1, "bad shape");
  const long long n = b * c * t;
  if (n == 0) return AA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (op == AA_UNARY_FLIP_CHANNELS || op == AA_UNARY_FLIP_TIME) {
    AA_REQUIRE(z != out, "flips cannot run in place");
    AA_REQUIRE(c < (1LL << 31) && t < (1LL << 31), "dims too large");
    flip_kernel<<<grid_for(n), kRedThreads, 0, st>>>(op, z, out, b, (int)c, (int)t);
    AA_LAUNCH_CHECK();
    return AA_OK;
  }
  AA_REQUIRE(op == AA_UNARY_SIGN_FOLD || op == AA_UNARY_ABSMAX_MINUS || op == AA_UNARY_TANH_DRIVE, "unknown op %d", op);
  AA_REQUIRE(workspace != nullptr, "workspace is NULL (need aa_reduce_workspace_floats() floats)");
  const int parts = parts_for(n);
  minmax_part_kernel<<<parts, kRedThreads, 0, st>>>(z, n, workspace + 2);
  AA_LAUNCH_CHECK();
  minmax_final_kernel<<<1, kRedThreads, 0, st>>>(workspace + 2, parts, workspace);
  AA_LAUNCH_CHECK();
  unary_kernel<<<grid_for(n), kRedThreads, 0, st>>>(op, z, out, n, param, workspace);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_effect_transfer_f32(const float* emb, int64_t b, const float* wet, const float* dry, int64_t bw, int64_t ct, float* out,
                           void* stream) {
  AA_REQUIRE(emb && wet && dry && out, "NULL tensor");
  AA_REQUIRE(bw >= 1 && b >= 0 && ct >= 1, "bad shape");
  if (b == 0) return AA_OK;
  effect_transfer_kernel<<<grid_for(ct), kRedThreads, 0, (cudaStream_t)stream>>>(emb, b, wet, dry, bw, ct, out);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_mse_fwd_f32(const float* a, const float* b, int64_t n, float* loss, float* workspace, void* stream) {
  AA_REQUIRE(a && b && loss && workspace, "NULL argument");
  AA_REQUIRE(n >= 1, "n must be >= 1");
  AA_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0, "tensors must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int parts = parts_for(n);
  mse_part_kernel<<<parts, kRedThreads, 0, st>>>(a, b, n, workspace);
  AA_LAUNCH_CHECK();
  finalize_sum_kernel<<<1, kRedThreads, 0, st>>>(workspace, parts, 1.0f / (float)n, loss, nullptr, 0, 0.f);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_mse_bwd_f32(const float* a, const float* b, int64_t n, const float* gloss, float gscale, float* grad_a, float* grad_b,
                   int accumulate, void* stream) {
  AA_REQUIRE(a && b, "NULL argument");
  AA_REQUIRE(n >= 1, "n must be >= 1");
  if (!grad_a && !grad_b) return AA_OK;
  mse_bwd_kernel<<<grid_for(n), kRedThreads, 0, (cudaStream_t)stream>>>(a, b, n, gloss, gscale, grad_a, grad_b, accumulate);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_vicreg_var_fwd_f32(const float* z, int64_t b, int64_t d, float gamma, float eps, int hinge_l2, float* loss, float* stats,
                          float* workspace, void* stream) {
  AA_REQUIRE(z && workspace, "NULL argument");
  AA_REQUIRE(b >= 2 && d >= 1 && b < (1LL << 31), "need batch >= 2 (unbiased variance), got b=%lld d=%lld", (long long)b, (long long)d);
  cudaStream_t st = (cudaStream_t)stream;
  const long long blocks = (d + 31) / 32;
  AA_REQUIRE(blocks <= (1LL << 30), "d too large");
  // partial sums: one per block; more than kMaxParts blocks are folded by the finalizer anyway (any count works)
  float* parts = workspace;
  const bool fits = blocks <= aa_reduce_workspace_floats();
  AA_REQUIRE(fits || loss == nullptr, "d=%lld too large for the reduction workspace", (long long)d);
  var_stats_kernel<<<(unsigned)blocks, kRedThreads, 0, st>>>(z, (int)b, d, gamma, eps, hinge_l2, stats, loss ? parts : nullptr);
  AA_LAUNCH_CHECK();
  if (loss) {
    finalize_sum_kernel<<<1, kRedThreads, 0, st>>>(parts, (int)blocks, 1.0f / (float)d, loss, nullptr, 0, 0.f);
    AA_LAUNCH_CHECK();
  }
  return AA_OK;
}

int aa_vicreg_var_bwd_f32(const float* z, const float* stats, int64_t b, int64_t d, float gamma, float eps, int hinge_l2,
                          const float* gloss, float gscale, float* grad_z, int accumulate, void* stream) {
  AA_REQUIRE(z && stats && grad_z, "NULL argument");
  AA_REQUIRE(b >= 2 && d >= 1, "bad shape");
  var_bwd_kernel<<<grid_for(b * d), kRedThreads, 0, (cudaStream_t)stream>>>(z, stats, (int)b, d, gamma, eps, hinge_l2, gloss,
                                                                          gscale, grad_z, accumulate);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

// workspace layout for cov fwd: [0, 4*kMaxParts+64) reduction scratch, then n_split * b * b partial Gram tiles
static int cov_splits(int64_t b, int64_t d) {
  const int64_t tiles = ((b + GT - 1) / GT) * ((b + GT - 1) / GT + 1) / 2;
  int64_t s = std::max<int64_t>(1, (2LL * aa::num_sms() + tiles - 1) / tiles);
  s = std::min<int64_t>(s, std::max<int64_t>(1, d / 256));
  return (int)std::min<int64_t>(s, 64);
}

int64_t aa_cov_loss_workspace_floats(int64_t b, int64_t d) {
  return aa_reduce_workspace_floats() + (int64_t)std::max(cov_splits(b, d), aa::cov_tc_splits(b, d)) * b * b;
}

int aa_vicreg_cov_fwd_f32(const float* z, int64_t b, int64_t d, const float* stats, float* stats_out, float* gram, float* loss,
                          float* workspace, void* stream) {
  AA_REQUIRE(z && gram && loss && workspace, "NULL argument");
  AA_REQUIRE(b >= 2 && d >= 1 && b < 65536, "bad shape b=%lld d=%lld", (long long)b, (long long)d);
  cudaStream_t st = (cudaStream_t)stream;
  if (stats == nullptr) {
    AA_REQUIRE(stats_out != nullptr, "either stats or stats_out must be given");
    int rc = aa_vicreg_var_fwd_f32(z, b, d, 1.0f, 1e-4f, 0, nullptr, stats_out, workspace, stream);
    if (rc != AA_OK) return rc;
    stats = stats_out;
  }
  float* red = workspace;
  float* gparts = workspace + aa_reduce_workspace_floats();
  int splits;
  if (aa::cov_tc_eligible(z, stats, b, d)) {   // B x B x D Gram on tcgen05 (3-term TF32 split, csrc/cov_tc.cu)
    splits = aa::cov_tc_splits(b, d);
    int rc = aa::gram_bb_tc(z, stats, b, d, gparts, splits, st);
    if (rc != AA_OK) return rc;
  } else {
    splits = cov_splits(b, d);
    long long dps = (d + splits - 1) / splits;
    dps = (dps + GK - 1) / GK * GK;
    const int nt = (int)((b + GT - 1) / GT);
    gram_part_kernel<<<dim3(nt, nt, splits), 256, 0, st>>>(z, stats, (int)b, d, dps, gparts);
    AA_LAUNCH_CHECK();
  }
  const int p1 = parts_for(b * b);
  gram_reduce_kernel<<<p1, kRedThreads, 0, st>>>(gparts, splits, (int)b, gram, red);
  AA_LAUNCH_CHECK();
  const int p2 = parts_for(d);
  diag_sq_kernel<<<p2, kRedThreads, 0, st>>>(stats, (int)b, d, red + kMaxParts);
  AA_LAUNCH_CHECK();
  const float k = 1.0f / ((float)(b - 1) * (float)(b - 1) * (float)d);
  finalize_sum_kernel<<<1, kRedThreads, 0, st>>>(red, p1, k, loss, red + kMaxParts, p2, k);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_vicreg_cov_bwd_f32(const float* z, const float* stats, const float* gram, int64_t b, int64_t d, const float* gloss,
                          float gscale, float* grad_z, int accumulate, void* stream) {
  AA_REQUIRE(z && stats && gram && grad_z, "NULL argument");
  AA_REQUIRE(b >= 2 && d >= 1 && b < 65536, "bad shape");
  if (aa::cov_tc_eligible(z, stats, b, d) && b % 4 == 0 && ((uintptr_t)gram & 15) == 0 && ((uintptr_t)grad_z & 15) == 0)
    return aa::cov_bwd_tc(z, stats, gram, b, d, gloss, gscale, grad_z, accumulate, (cudaStream_t)stream);
  const long long ct = (d + GT - 1) / GT;
  AA_REQUIRE(ct < (1LL << 31), "d too large");
  cov_bwd_kernel<<<dim3((unsigned)ct, (unsigned)((b + GT - 1) / GT)), 256, 0, (cudaStream_t)stream>>>(
      z, stats, gram, (int)b, d, gloss, gscale, grad_z, accumulate);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

static int scatter_blocks() { return 2 * aa::num_sms(); }

int64_t aa_cov_workspace_floats(int64_t c) {
  const int64_t slices = 64;
  return std::max<int64_t>((int64_t)scatter_blocks() * c * c + c * slices + c + 64, aa::gram_tc_workspace_floats());
}

int aa_cov_accumulate_f32(const float* y, int64_t b, int64_t c, int64_t t, float* cov_num, double* count, float* workspace,
                          void* stream) {
  AA_REQUIRE(y && cov_num && workspace, "NULL argument");
  AA_REQUIRE(b >= 1 && c >= 1 && t >= 1 && c <= 4096, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  // 64 channels (every given model's latent width): one pass over the latents, rank-n update on tcgen05 (csrc/gram_tc.cu)
  if (c == 64 && (t & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && getenv("AA_PCA_CUDA_CORES") == nullptr)
    return aa::gram_tc(y, b, t, cov_num, count, workspace, st);
  const int slices = 64;
  float* parts = workspace;                                   // [blocks][c][c]
  float* sums = workspace + (int64_t)scatter_blocks() * c * c;  // [c][slices]
  float* mean = sums + c * slices;
  chan_sum_part_kernel<<<dim3((unsigned)c, slices), kRedThreads, 0, st>>>(y, b, (int)c, t, sums, slices);
  AA_LAUNCH_CHECK();
  chan_mean_kernel<<<(unsigned)((c + 127) / 128), 128, 0, st>>>(sums, (int)c, slices, (double)b * (double)t, mean);
  AA_LAUNCH_CHECK();
  const int ctiles = (int)((c + 63) / 64);
  const long long n_tiles = b * ((t + PT - 1) / PT);
  const int blocks = (int)std::min<long long>(n_tiles, std::max(1, scatter_blocks() / (ctiles * ctiles)));
  scatter_part_kernel<<<dim3(blocks, ctiles, ctiles), 256, 0, st>>>(y, b, (int)c, t, mean, parts);
  AA_LAUNCH_CHECK();
  scatter_reduce_kernel<<<(unsigned)((c * c + 31) / 32), 256, 0, st>>>(parts, blocks, (int)c, cov_num, count,
                                                                        (double)b * (double)t);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_adam_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, int64_t step, void* stream) {
  AA_REQUIRE(params && grads && exp_avg && exp_avg_sq, "NULL argument");
  AA_REQUIRE(step >= 1, "step counts from 1");
  if (n == 0) return AA_OK;
  const double bc1 = 1.0 - std::pow((double)beta1, (double)step);
  const double bc2 = 1.0 - std::pow((double)beta2, (double)step);
  adam_kernel<<<grid_for(n), kRedThreads, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                   (float)bc1, (float)std::sqrt(bc2));
  AA_LAUNCH_CHECK();
  return AA_OK;
}

#pragma GCC visibility pop
}  // extern "C"
