// AudioAlgebra projector (h: y -> z and its inverse): one fused kernel per half.
//
// Reference: EmbedBlock (aa_mixer.py:205-221) and AudioAlgebra.encode/decode (aa_mixer.py:252-260):
//   h0 = x^T;  h_{l+1} = [h_l +] act_l(W_l h_l + b_l)  (inner residual iff in==out, GELU(erf) on the first
//   three blocks, none on the last);  out = h_4^T [+ x]   (outer residual iff resid)
// The reference runs 4 cuBLAS GEMMs, ~13 elementwise kernels and 2 transpose copies per half; here the
// channel-major [B][C][T] tensor is read once, all four layers are applied to a tile of tokens held in
// shared memory (weights resident in shared memory for the whole persistent CTA) and the result is
// written once.  fp32 FMA version (exact-parity path); dims, hidden <= 64 (zero padded to 64).
//
// Backward recomputes the forward activations of the tile in shared memory (nothing is saved by the
// forward), back-propagates, and accumulates the weight / bias gradients per CTA; a second kernel sums
// the per-CTA partials in a fixed order (deterministic).
#include "aa_common.cuh"
#include "packed_f32x2.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdlib>

namespace {

constexpr int PD = 64;        // padded feature dim
constexpr int TT = 32;        // tokens per tile
constexpr int TSD = TT + 4;   // row stride of the activation tiles: 16-byte aligned rows, and 16 consecutive feature rows hit distinct bank groups
constexpr int PTHREADS = 256; // 16 x 16 threads, 4 (features) x 2 (tokens) micro tile
constexpr int WLD = PD + 4;   // padded leading dim of weight tiles in smem (W[o][i]): rows stay 16-byte aligned for float4 loads

struct ProjArgs {
  const float* w[4];
  const float* b[4];
  int in_dim[4], out_dim[4];
  int resid_inner[4];   // layer l adds its input (resid && in == out)
  int resid_outer;
  int dims;             // channels of x / out
  const float* x;
  float* out;
  long long batch;
  int t;
  long long n_tiles;    // batch * ceil(t / TT)
};

__device__ __forceinline__ float gelu_erf(float u) { return 0.5f * u * (1.0f + erff(u * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float u) {
  const float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * u * u);
  return cdf + u * pdf;
}

__device__ __forceinline__ void load_weights(const ProjArgs& a, float* sW, float* sB) {
  for (int l = 0; l < 4; ++l) {
    for (int e = threadIdx.x; e < PD * PD; e += blockDim.x) {
      const int o = e / PD, i = e % PD;
      sW[l * PD * WLD + o * WLD + i] = (o < a.out_dim[l] && i < a.in_dim[l]) ? a.w[l][o * a.in_dim[l] + i] : 0.f;
    }
    for (int e = threadIdx.x; e < PD; e += blockDim.x) sB[l * PD + e] = (e < a.out_dim[l]) ? a.b[l][e] : 0.f;
  }
}

// H_out[o][t] = [H_in[o][t] +] act(sum_i W[o][i] H_in[i][t] + b[o]); optionally also stores the pre-activation.
// Thread (ty, tx): features o = ty*4..+3, tokens t = tx*2..+1.
template <bool ACT, bool STORE_U>
__device__ __forceinline__ void layer_fwd(const float* __restrict__ sW, const float* __restrict__ sB, const float* __restrict__ Hin,
                                          float* __restrict__ Hout, float* __restrict__ U, bool resid) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float2 acc[4];   // the two tokens of the micro tile ride in one packed fp32x2 accumulator (FFMA2: one issue slot per 2 FMAs)
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float2(sB[ty * 4 + i], sB[ty * 4 + i]);
  // four input features per step: one float4 of weights per output feature (broadcast within the half-warp that shares ty)
  // and one float2 of activations per input feature -> 8 shared loads per 32 FMAs
#pragma unroll 4
  for (int k = 0; k < PD; k += 4) {
    float2 h[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) h[kk] = *reinterpret_cast<const float2*>(Hin + (k + kk) * TSD + tx * 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 w = *reinterpret_cast<const float4*>(sW + (ty * 4 + i) * WLD + k);
      acc[i] = pfma(h[0], w.x, acc[i]);
      acc[i] = pfma(h[1], w.y, acc[i]);
      acc[i] = pfma(h[2], w.z, acc[i]);
      acc[i] = pfma(h[3], w.w, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = ty * 4 + i;
    float2 r;
    if (STORE_U) *reinterpret_cast<float2*>(U + o * TSD + tx * 2) = acc[i];
    r.x = ACT ? gelu_erf(acc[i].x) : acc[i].x;
    r.y = ACT ? gelu_erf(acc[i].y) : acc[i].y;
    if (resid) {
      const float2 h = *reinterpret_cast<const float2*>(Hin + o * TSD + tx * 2);
      r.x += h.x;
      r.y += h.y;
    }
    *reinterpret_cast<float2*>(Hout + o * TSD + tx * 2) = r;
  }
}

__device__ __forceinline__ void load_tile(const ProjArgs& a, const float* __restrict__ src, long long tile, float* H, int& bi,
                                          int& t0) {
  const int tiles_t = (a.t + TT - 1) / TT;
  bi = (int)(tile / tiles_t);
  t0 = (int)(tile % tiles_t) * TT;
  for (int e = threadIdx.x; e < PD * TT; e += blockDim.x) {
    const int c = e / TT, tt = e % TT;
    H[c * TSD + tt] = (c < a.dims && t0 + tt < a.t) ? src[((long long)bi * a.dims + c) * a.t + t0 + tt] : 0.f;
  }
}

__global__ void __launch_bounds__(PTHREADS) proj_fwd_kernel(const ProjArgs a) {
  extern __shared__ float sm[];
  float* sW = sm;                      // [4][PD][WLD]
  float* sB = sW + 4 * PD * WLD;       // [4][PD]
  float* H0 = sB + 4 * PD;             // [PD][TSD] tile of x
  float* Ha = H0 + PD * TSD;
  float* Hb = Ha + PD * TSD;
  load_weights(a, sW, sB);
  for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    int bi, t0;
    __syncthreads();
    load_tile(a, a.x, tile, H0, bi, t0);
    __syncthreads();
    layer_fwd<true, false>(sW, sB, H0, Ha, nullptr, a.resid_inner[0]);
    __syncthreads();
    layer_fwd<true, false>(sW + PD * WLD, sB + PD, Ha, Hb, nullptr, a.resid_inner[1]);
    __syncthreads();
    layer_fwd<true, false>(sW + 2 * PD * WLD, sB + 2 * PD, Hb, Ha, nullptr, a.resid_inner[2]);
    __syncthreads();
    layer_fwd<false, false>(sW + 3 * PD * WLD, sB + 3 * PD, Ha, Hb, nullptr, a.resid_inner[3]);
    __syncthreads();
    for (int e = threadIdx.x; e < PD * TT; e += blockDim.x) {
      const int c = e / TT, tt = e % TT;
      if (c < a.dims && t0 + tt < a.t)
        a.out[((long long)bi * a.dims + c) * a.t + t0 + tt] = Hb[c * TSD + tt] + (a.resid_outer ? H0[c * TSD + tt] : 0.f);
    }
  }
}

struct ProjBwdArgs {
  ProjArgs f;
  const float* gout;   // dL/dout [B][dims][T]
  float* gx;           // dL/dx (may be NULL)
  float* partials;     // [grid][4][PD*PD + PD]
  int accumulate_gx;
};

// dHin[i][t] = sum_o W[o][i] dU[o][t]  (+ dHout[i][t] if resid); thread (ty,tx): i = ty*4.., t = tx*2..
__device__ __forceinline__ void layer_bwd_data(const float* __restrict__ sW, const float* __restrict__ dU,
                                               const float* __restrict__ dHout, float* __restrict__ dHin, bool resid) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float2 acc[4] = {};
#pragma unroll 8
  for (int o = 0; o < PD; ++o) {
    const float2 g = *reinterpret_cast<const float2*>(dU + o * TSD + tx * 2);
    const float4 w = *reinterpret_cast<const float4*>(sW + o * WLD + ty * 4);   // W[o][i .. i+3]
    acc[0] = pfma(g, w.x, acc[0]);
    acc[1] = pfma(g, w.y, acc[1]);
    acc[2] = pfma(g, w.z, acc[2]);
    acc[3] = pfma(g, w.w, acc[3]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = ty * 4 + i;
    float2 r = acc[i];
    if (resid) {
      const float2 h = *reinterpret_cast<const float2*>(dHout + c * TSD + tx * 2);
      r.x += h.x;
      r.y += h.y;
    }
    *reinterpret_cast<float2*>(dHin + c * TSD + tx * 2) = r;
  }
}

// gW[o][i] += sum_t dU[o][t] Hin[i][t]; thread (ty,tx): o = ty*4 + a, i = tx + 16 b   (registers, persistent over tiles).
// The 16 lanes of a half-warp read 16 consecutive feature rows (stride TSD = 36 floats -> distinct 16-byte bank groups).
__device__ __forceinline__ void layer_bwd_weight(const float* __restrict__ dU, const float* __restrict__ Hin, float (&gw)[4][4],
                                                 float (&gb)[4]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll 2
  for (int t = 0; t < TT; t += 4) {      // four tokens per step: 8 float4 loads per 64 FMAs
    float4 du[4], h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      du[e] = *reinterpret_cast<const float4*>(dU + (ty * 4 + e) * TSD + t);
      h[e] = *reinterpret_cast<const float4*>(Hin + (tx + 16 * e) * TSD + t);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        gw[i][j] = fmaf(du[i].x, h[j].x, gw[i][j]);
        gw[i][j] = fmaf(du[i].y, h[j].y, gw[i][j]);
        gw[i][j] = fmaf(du[i].z, h[j].z, gw[i][j]);
        gw[i][j] = fmaf(du[i].w, h[j].w, gw[i][j]);
      }
      if (tx == 0) gb[i] += (du[i].x + du[i].y) + (du[i].z + du[i].w);
    }
  }
}

__global__ void __launch_bounds__(PTHREADS) proj_bwd_kernel(const ProjBwdArgs ba) {
  const ProjArgs& a = ba.f;
  extern __shared__ float sm[];
  float* sW = sm;                      // [4][PD][WLD]
  float* sB = sW + 4 * PD * WLD;       // [4][PD]
  constexpr int TILE = PD * TSD;
  float* H = sB + 4 * PD;              // H[0..3]: inputs of the four layers, [PD][TSD] each
  float* U = H + 4 * TILE;             // U[0..2]: pre-activations of the GELU layers
  float* G = U + 3 * TILE;             // dL/dout tile (kept for the outer residual)
  float* GA = G + TILE;                // gradient ping
  float* GB = GA + TILE;               // gradient pong
  float* DU = GB + TILE;               // dL/dU of the current layer
  load_weights(a, sW, sB);
  float gw[4][4][4] = {};
  float gb[4][4] = {};
  for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    int bi, t0;
    __syncthreads();
    load_tile(a, a.x, tile, H, bi, t0);
    load_tile(a, ba.gout, tile, G, bi, t0);
    __syncthreads();
    // forward recompute: inputs of every layer and the pre-activations of the GELU layers
    layer_fwd<true, true>(sW, sB, H, H + TILE, U, a.resid_inner[0]);
    __syncthreads();
    layer_fwd<true, true>(sW + PD * WLD, sB + PD, H + TILE, H + 2 * TILE, U + TILE, a.resid_inner[1]);
    __syncthreads();
    layer_fwd<true, true>(sW + 2 * PD * WLD, sB + 2 * PD, H + 2 * TILE, H + 3 * TILE, U + 2 * TILE, a.resid_inner[2]);
    __syncthreads();
    // layer 3 (no activation): dU3 = dH4 = gout
    layer_bwd_weight(G, H + 3 * TILE, gw[3], gb[3]);
    layer_bwd_data(sW + 3 * PD * WLD, G, G, GA, a.resid_inner[3]);   // GA = dH3
    __syncthreads();
    float* gin = GA;
    float* gnext = GB;
#pragma unroll
    for (int l = 2; l >= 0; --l) {
      for (int e = threadIdx.x; e < PD * TT; e += blockDim.x) {
        const int idx = (e / TT) * TSD + (e % TT);
        DU[idx] = gin[idx] * gelu_erf_grad(U[l * TILE + idx]);
      }
      __syncthreads();
      layer_bwd_weight(DU, H + l * TILE, gw[l], gb[l]);
      layer_bwd_data(sW + l * PD * WLD, DU, gin, gnext, a.resid_inner[l]);   // gnext = dH_l
      __syncthreads();
      float* tmp = gin; gin = gnext; gnext = tmp;
    }
    // gin = dH0; dL/dx = dH0 (+ gout through the outer residual)
    if (ba.gx) {
      for (int e = threadIdx.x; e < PD * TT; e += blockDim.x) {
        const int c = e / TT, tt = e % TT;
        if (c < a.dims && t0 + tt < a.t) {
          const long long o = ((long long)bi * a.dims + c) * a.t + t0 + tt;
          const float g = gin[c * TSD + tt] + (a.resid_outer ? G[c * TSD + tt] : 0.f);
          ba.gx[o] = ba.accumulate_gx ? ba.gx[o] + g : g;
        }
      }
    }
  }
  // per-CTA partial gradients
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float* p = ba.partials + (long long)blockIdx.x * 4 * (PD * PD + PD);
  for (int l = 0; l < 4; ++l) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) p[l * (PD * PD + PD) + (ty * 4 + i) * PD + tx + 16 * j] = gw[l][i][j];
      if (tx == 0) p[l * (PD * PD + PD) + PD * PD + ty * 4 + i] = gb[l][i];
    }
  }
}

struct ProjReduceArgs {
  float* gw[4];
  float* gb[4];
  int in_dim[4], out_dim[4];
  const float* partials;
  int n_parts;
  float scale;
  int accumulate;
};
__global__ void proj_grad_reduce_kernel(const ProjReduceArgs r) {
  const int per = PD * PD + PD;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * per) return;
  const int l = idx / per, e = idx % per;
  float s = 0.f;
  for (int p = 0; p < r.n_parts; ++p) s += r.partials[(long long)p * 4 * per + idx];
  s *= r.scale;
  if (e < PD * PD) {
    const int o = e / PD, i = e % PD;
    if (o < r.out_dim[l] && i < r.in_dim[l] && r.gw[l]) {
      float* dst = r.gw[l] + o * r.in_dim[l] + i;
      *dst = r.accumulate ? *dst + s : s;
    }
  } else {
    const int o = e - PD * PD;
    if (o < r.out_dim[l] && r.gb[l]) {
      float* dst = r.gb[l] + o;
      *dst = r.accumulate ? *dst + s : s;
    }
  }
}

int fill_args(ProjArgs& a, const float* const* w, const float* const* b, int dims, int hidden, int resid, const float* x,
              float* out, int64_t batch, int64_t t) {
  AA_REQUIRE(w && b && x, "NULL argument");
  AA_REQUIRE(dims >= 1 && dims <= PD && hidden >= 1 && hidden <= PD, "dims=%d hidden=%d: the fused projector supports <= %d", dims,
             hidden, PD);
  AA_REQUIRE(batch >= 0 && t >= 1 && t < (1LL << 31), "bad shape");
  const int ind[4] = {dims, hidden, hidden, hidden}, outd[4] = {hidden, hidden, hidden, dims};
  for (int l = 0; l < 4; ++l) {
    AA_REQUIRE(w[l] && b[l], "weight/bias %d is NULL", l);
    a.w[l] = w[l]; a.b[l] = b[l];
    a.in_dim[l] = ind[l]; a.out_dim[l] = outd[l];
    a.resid_inner[l] = (resid && ind[l] == outd[l]) ? 1 : 0;
  }
  a.resid_outer = resid ? 1 : 0;
  a.dims = dims; a.x = x; a.out = out; a.batch = batch; a.t = (int)t;
  a.n_tiles = batch * ((t + TT - 1) / TT);
  return AA_OK;
}

constexpr int kFwdSmem = (4 * PD * WLD + 4 * PD + 3 * PD * TSD) * 4;
constexpr int kBwdSmem = (4 * PD * WLD + 4 * PD + (4 + 3 + 4) * PD * TSD) * 4;

}  // namespace

namespace aa {
int proj_fwd_tc(const float* const* w, const float* const* b, const float* x, int64_t batch, int64_t t, float* out, cudaStream_t stream);
// csrc/proj_bwd_tc.cu
int64_t proj_bwd_tc_workspace_floats();
bool proj_bwd_tc_enabled();
int proj_bwd_tc(const float* const* w, const float* const* b, const float* x, const float* gout, int64_t batch, int64_t t, float* gx,
                int accumulate_gx, float* const* gw, float* const* gb, int accumulate_gw, float gscale, float* workspace,
                cudaStream_t stream);
}

extern "C" {
#pragma GCC visibility push(default)

int aa_projector_half_fwd_f32(const float* const* w_host, const float* const* b_host, int dims, int hidden, int resid,
                              const float* x, int64_t batch, int64_t t, float* out, void* stream) {
  ProjArgs a;
  int rc = fill_args(a, w_host, b_host, dims, hidden, resid, x, out, batch, t);
  if (rc != AA_OK) return rc;
  AA_REQUIRE(out != nullptr, "out is NULL");
  if (a.n_tiles == 0) return AA_OK;
  // the standard 64 -> 64 residual projector runs on the tensor core (tcgen05, 3-term TF32 split: fp32-accurate, proj_tc.cu);
  // other shapes (toy dims, no residual, very short sequences) and AA_PROJ_FP32=1 (tests) use the CUDA-core kernel below
  static const bool force_fp32 = getenv("AA_PROJ_FP32") != nullptr;
  if (!force_fp32 && dims == PD && hidden == PD && resid && t >= 64 && (reinterpret_cast<uintptr_t>(w_host[0]) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(w_host[1]) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_host[2]) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(w_host[3]) & 15) == 0)
    return aa::proj_fwd_tc(w_host, b_host, x, batch, t, out, (cudaStream_t)stream);
  AA_CUDA(aa::ensure_dyn_smem(proj_fwd_kernel, kFwdSmem));   // per (kernel, device)
  const int grid = (int)std::min<long long>(a.n_tiles, 2LL * aa::num_sms());
  proj_fwd_kernel<<<grid, PTHREADS, kFwdSmem, (cudaStream_t)stream>>>(a);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int64_t aa_projector_bwd_workspace_floats(void) {
  return std::max<int64_t>((int64_t)aa::num_sms() * 4 * (PD * PD + PD), aa::proj_bwd_tc_workspace_floats());
}

int aa_projector_half_bwd_f32(const float* const* w_host, const float* const* b_host, int dims, int hidden, int resid,
                              const float* x, const float* gout, int64_t batch, int64_t t, float* gx, int accumulate_gx,
                              float* const* gw_host, float* const* gb_host, int accumulate_gw, float gscale, float* workspace,
                              void* stream) {
  ProjBwdArgs ba;
  int rc = fill_args(ba.f, w_host, b_host, dims, hidden, resid, x, nullptr, batch, t);
  if (rc != AA_OK) return rc;
  AA_REQUIRE(gout && workspace && gw_host && gb_host, "NULL argument");
  if (ba.f.n_tiles == 0) return AA_OK;
  // the standard 64 -> 64 residual projector runs on the tensor core (tcgen05, bf16x3 split: fp32-accurate, proj_bwd_tc.cu);
  // other shapes and AA_PROJ_FP32=1 / AA_PROJ_BWD_TC=0 use the CUDA-core kernel below
  {
    static const bool force_fp32 = getenv("AA_PROJ_FP32") != nullptr;
    bool ok = !force_fp32 && aa::proj_bwd_tc_enabled() && dims == PD && hidden == PD && resid && t >= 64;
    for (int l = 0; l < 4 && ok; ++l) ok = (reinterpret_cast<uintptr_t>(w_host[l]) & 15) == 0 && gw_host[l] != nullptr && gb_host[l] != nullptr;
    if (ok)
      return aa::proj_bwd_tc(w_host, b_host, x, gout, batch, t, gx, accumulate_gx, gw_host, gb_host, accumulate_gw, gscale, workspace,
                             (cudaStream_t)stream);
  }
  AA_CUDA(aa::ensure_dyn_smem(proj_bwd_kernel, kBwdSmem));   // per (kernel, device)
  ba.gout = gout; ba.gx = gx; ba.partials = workspace; ba.accumulate_gx = accumulate_gx;
  const int grid = (int)std::min<long long>(ba.f.n_tiles, (long long)aa::num_sms());
  proj_bwd_kernel<<<grid, PTHREADS, kBwdSmem, (cudaStream_t)stream>>>(ba);
  AA_LAUNCH_CHECK();
  ProjReduceArgs r;
  for (int l = 0; l < 4; ++l) {
    r.gw[l] = gw_host[l]; r.gb[l] = gb_host[l];
    r.in_dim[l] = ba.f.in_dim[l]; r.out_dim[l] = ba.f.out_dim[l];
  }
  r.partials = workspace; r.n_parts = grid; r.scale = gscale; r.accumulate = accumulate_gw;
  proj_grad_reduce_kernel<<<(4 * (PD * PD + PD) + 255) / 256, 256, 0, (cudaStream_t)stream>>>(r);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

#pragma GCC visibility pop
}  // extern "C"
