// Given-model conv encoder ("f: audio -> y"): the restated SoundStreamXLEncoder of the reference's
// DiffusionDVAE (aa_mixer.py:118-131 constructor arguments; architecture per SURVEY.md Appendix A --
// third-party source, PARITY UNPINNED upstream):
//   Conv1d(in->cap,k7,p3) ELU  [ResUnit(d=1) ELU ResUnit(3) ELU ResUnit(9) ELU Conv1d(k=2s,stride s,pad ceil(s/2)) ELU] x5
//   Conv1d(c_last->latent,k3,p1)  [tanh for encode_it / DVAEWrapper.encode, none for DiffusionDVAE.encode]
//   ResUnit(x) = x + Conv1d(k1)(ELU(Conv1d(k7, dilation d, padding 3d)(x)))
// The layer list is data (a table of {cin, cout, k, stride, dilation, pad, residual role}) so a corrected
// upstream architecture is a table change.  Every conv is one kernel with a fused epilogue:
//   bias (+ residual) -> ELU (-> tanh on the last layer); the first layer also fuses the fader-scaled sum
//   of stems (aa_mixer.py:303,309: fadedstem = s*f; mix += fadedstem) into its load.
// This file: layer table, weight storage, sequencing, and the fp32 CUDA-core conv kernel (exact-parity
// path).  The bf16 tensor-core path (tcgen05 implicit GEMM) lives in conv_tc.cu.
#include "aa_common.cuh"
#include "encoder.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace {

constexpr int CT = 64;    // output-channel tile
constexpr int LT = 64;    // output-position tile
constexpr int CK = 8;     // input channels per smem chunk
constexpr int XW = 272;   // max input strip width: 63*stride + (k-1)*dil + 1 <= 63*4 + 7 + 1 = 260; 63 + 6*9 + 1 = 118

struct ConvArgs {
  const float* x[4];      // up to 4 stems (layer 0) or one activation tensor
  float fader[4];
  int n_in;
  const float* w;         // [cout][cin][k]
  const float* bias;      // [cout]
  const float* res;       // residual [B][cout][lout] or NULL
  float* out;             // [B][cout][lout]
  int cin, cout, lin, lout, k, stride, dil, pad;
  int elu, tanh_out;
};

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }

__global__ void __launch_bounds__(256) conv1d_f32_kernel(const ConvArgs a) {
  __shared__ float Xs[CK][XW];
  __shared__ __align__(16) float Ws[CK][8][CT + 4];   // [ci][k][co]: rows 16-byte aligned (float4 loads of 4 output channels)
  const int b = blockIdx.z;
  const int co0 = blockIdx.y * CT, l0 = blockIdx.x * LT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // tx: positions, ty: channels
  const int span = (LT - 1) * a.stride + (a.k - 1) * a.dil + 1;
  const int in0 = l0 * a.stride - a.pad;
  float acc[4][4] = {};
  for (int c0 = 0; c0 < a.cin; c0 += CK) {
    for (int e = threadIdx.x; e < CK * span; e += 256) {
      const int ci = e / span, j = e % span;
      const int c = c0 + ci, pos = in0 + j;
      float v = 0.f;
      if (c < a.cin && pos >= 0 && pos < a.lin) {
        const long long off = ((long long)b * a.cin + c) * a.lin + pos;
        v = a.fader[0] * a.x[0][off];
        for (int s = 1; s < a.n_in; ++s) v = fmaf(a.fader[s], a.x[s][off], v);
      }
      Xs[ci][j] = v;
    }
    for (int e = threadIdx.x; e < CK * a.k * CT; e += 256) {
      const int co = e % CT, r = e / CT, kk = r % a.k, ci = r / a.k;
      const int c = c0 + ci, o = co0 + co;
      Ws[ci][kk][co] = (c < a.cin && o < a.cout) ? a.w[((long long)o * a.cin + c) * a.k + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < CK; ++ci) {
      for (int kk = 0; kk < a.k; ++kk) {
        const float4 w4 = *reinterpret_cast<const float4*>(&Ws[ci][kk][ty * 4]);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
        float xv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) xv[e] = Xs[ci][(tx * 4 + e) * a.stride + kk * a.dil];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = co0 + ty * 4 + i;
    if (o >= a.cout) continue;
    const float bv = a.bias[o];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = l0 + tx * 4 + j;
      if (l >= a.lout) continue;
      const long long off = ((long long)b * a.cout + o) * a.lout + l;
      float v = acc[i][j] + bv;
      if (a.res) v += a.res[off];
      if (a.elu) v = elu1(v);
      if (a.tanh_out) v = tanhf(v);
      a.out[off] = v;
    }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
int aa::build_layer_table(const AaEncoderCfg& cfg, std::vector<aa::ConvLayer>& L) {
  L.clear();
  AA_REQUIRE(cfg.n_blocks >= 1 && cfg.n_blocks <= 8, "n_blocks=%d", cfg.n_blocks);
  AA_REQUIRE(cfg.in_channels >= 1 && cfg.capacity >= 1 && cfg.latent_dim >= 1, "bad encoder config");
  auto push = [&](int cin, int cout, int k, int stride, int dil, int pad, int elu, int role) {
    aa::ConvLayer l{};
    l.cin = cin; l.cout = cout; l.k = k; l.stride = stride; l.dil = dil; l.pad = pad; l.elu = elu; l.role = role;
    L.push_back(l);
  };
  int c = cfg.capacity;
  push(cfg.in_channels, c, 7, 1, 1, 3, 1, aa::ROLE_PLAIN);
  for (int i = 0; i < cfg.n_blocks; ++i) {
    const int cin = c, cout = cfg.c_mults[i] * cfg.capacity, s = cfg.strides[i];
    AA_REQUIRE(s >= 1 && s <= 4, "stride %d not supported (1..4)", s);
    const int dils[3] = {1, 3, 9};
    for (int d : dils) {
      push(cin, cin, 7, 1, d, 3 * d, 1, aa::ROLE_RES_FIRST);   // h = ELU(conv_k7(x))
      push(cin, cin, 1, 1, 1, 0, 1, aa::ROLE_RES_SECOND);      // y = ELU(conv_k1(h) + x)
    }
    push(cin, cout, 2 * s, s, 1, (s + 1) / 2, 1, aa::ROLE_PLAIN);
    c = cout;
  }
  push(c, cfg.latent_dim, 3, 1, 1, 1, 0, aa::ROLE_PLAIN);
  return AA_OK;
}

struct AaEncoder {
  AaEncoderCfg cfg;
  std::vector<aa::ConvLayer> layers;
  std::vector<float*> w, b;          // device fp32 copies, [cout][cin][k] / [cout]
  aa::TcState* tc = nullptr;         // bf16 tensor-core path state (conv_tc.cu), created lazily
  aa::TfState* tf = nullptr;         // 3xTF32 tensor-core path state (conv_tf32.cuh), created lazily
  int total_stride = 1;
  bool tf_ok = false;                // layer table supported by the 3xTF32 path
};

extern "C" {
#pragma GCC visibility push(default)

int aa_encoder_create(const AaEncoderCfg* cfg, AaEncoder** out) {
  AA_REQUIRE(cfg && out, "NULL argument");
  int rc = aa_check_device();
  if (rc != AA_OK) return rc;
  AaEncoder* e = new AaEncoder();
  e->cfg = *cfg;
  rc = aa::build_layer_table(*cfg, e->layers);
  if (rc != AA_OK) { delete e; return rc; }
  e->w.assign(e->layers.size(), nullptr);
  e->b.assign(e->layers.size(), nullptr);
  for (int i = 0; i < cfg->n_blocks; ++i) e->total_stride *= cfg->strides[i];
  e->tf_ok = aa::tf_eligible(e->layers);
  for (size_t i = 0; i < e->layers.size(); ++i) {
    const auto& l = e->layers[i];
    AA_CUDA(cudaMalloc(&e->w[i], sizeof(float) * (size_t)l.cout * l.cin * l.k));
    AA_CUDA(cudaMalloc(&e->b[i], sizeof(float) * l.cout));
    AA_CUDA(cudaMemset(e->w[i], 0, sizeof(float) * (size_t)l.cout * l.cin * l.k));
    AA_CUDA(cudaMemset(e->b[i], 0, sizeof(float) * l.cout));
  }
  *out = e;
  return AA_OK;
}

int aa_encoder_destroy(AaEncoder* e) {
  if (!e) return AA_OK;
  for (auto p : e->w) cudaFree(p);
  for (auto p : e->b) cudaFree(p);
  aa::tc_destroy(e->tc);
  aa::tf_destroy(e->tf);
  delete e;
  return AA_OK;
}

int aa_encoder_num_layers(const AaEncoder* e) { return e ? (int)e->layers.size() : 0; }

int aa_encoder_layer_shape(const AaEncoder* e, int layer, int* cout, int* cin, int* k) {
  AA_REQUIRE(e && layer >= 0 && layer < (int)e->layers.size(), "bad layer index %d", layer);
  if (cout) *cout = e->layers[layer].cout;
  if (cin) *cin = e->layers[layer].cin;
  if (k) *k = e->layers[layer].k;
  return AA_OK;
}

int aa_encoder_set_weights(AaEncoder* e, int layer, const float* w_dev, const float* b_dev, void* stream) {
  AA_REQUIRE(e && layer >= 0 && layer < (int)e->layers.size(), "bad layer index %d", layer);
  AA_REQUIRE(w_dev && b_dev, "NULL weights");
  const auto& l = e->layers[layer];
  AA_CUDA(cudaMemcpyAsync(e->w[layer], w_dev, sizeof(float) * (size_t)l.cout * l.cin * l.k, cudaMemcpyDeviceToDevice,
                          (cudaStream_t)stream));
  AA_CUDA(cudaMemcpyAsync(e->b[layer], b_dev, sizeof(float) * l.cout, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  if (e->tc) aa::tc_invalidate_weights(e->tc);
  if (e->tf) aa::tf_invalidate_weights(e->tf);
  return AA_OK;
}

static int64_t max_act_elems(const AaEncoder* e, int64_t batch, int64_t n) {
  int64_t l = n, mx = 0;
  for (const auto& ly : e->layers) {
    const int64_t lout = (l + 2 * ly.pad - (int64_t)ly.dil * (ly.k - 1) - 1) / ly.stride + 1;
    mx = std::max(mx, batch * ly.cout * lout);
    l = lout;
  }
  return mx;
}

// AA_DTYPE_F32 asks for fp32-grade results, not for a particular pipe: layer tables the 3xTF32 kernels support run there (same
// measured accuracy against float64 as the CUDA-core kernel, ~10x its speed); AA_DTYPE_F32_CUDA_CORES or the environment
// variable AA_ENC_FP32_CUDA_CORES=1 selects the CUDA-core kernel.
static bool fp32_on_tensor_cores(const AaEncoder* e) {
  return e->tf_ok && getenv("AA_ENC_FP32_CUDA_CORES") == nullptr;
}

// Generic fp32 Conv1d on channel-major tensors (the same CUDA-core kernel the fp32 encoder path uses), exported for the given
// models whose layers are not in the SoundStreamXL table (StackedDiffAE's Encoder1d: 1x1 / k3 / strided k5 convs).
// out = act(conv(x) + bias (+ res)); act: 0 none, 1 ELU, 2 tanh.  x [B][cin][lin], w [cout][cin][k], res / out [B][cout][lout].
int aa_conv1d_f32(const float* x, int64_t batch, int cin, int64_t lin, const float* w, const float* bias, int cout, int k, int stride,
                  int dil, int pad, const float* res, int act, float* out, void* stream) {
  AA_REQUIRE(x && w && bias && out, "NULL tensor");
  AA_REQUIRE(batch >= 0 && batch <= 65535 && cin >= 1 && cout >= 1 && lin >= 1 && lin < (1LL << 31), "bad shape");
  AA_REQUIRE(k >= 1 && k <= 8 && stride >= 1 && dil >= 1 && pad >= 0, "bad conv geometry");
  AA_REQUIRE((LT - 1) * stride + (k - 1) * dil + 1 <= XW, "stride %d / kernel %d / dilation %d: input strip too wide for the kernel", stride, k, dil);
  AA_REQUIRE(act >= 0 && act <= 2, "act=%d", act);
  if (batch == 0) return AA_OK;
  const int64_t lout = (lin + 2 * pad - (int64_t)dil * (k - 1) - 1) / stride + 1;
  AA_REQUIRE(lout >= 1, "input too short");
  ConvArgs a{};
  a.n_in = 1; a.x[0] = x; a.fader[0] = 1.0f;
  a.w = w; a.bias = bias; a.res = res; a.out = out;
  a.cin = cin; a.cout = cout; a.lin = (int)lin; a.lout = (int)lout; a.k = k; a.stride = stride; a.dil = dil; a.pad = pad;
  a.elu = act == 1; a.tanh_out = act == 2;
  dim3 grid((unsigned)((lout + LT - 1) / LT), (unsigned)((cout + CT - 1) / CT), (unsigned)batch);
  conv1d_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

int aa_encoder_out_length(const AaEncoder* e, int64_t n, int64_t* t_out) {
  AA_REQUIRE(e && t_out, "NULL argument");
  int64_t l = n;
  for (const auto& ly : e->layers) l = (l + 2 * ly.pad - (int64_t)ly.dil * (ly.k - 1) - 1) / ly.stride + 1;
  *t_out = l;
  return AA_OK;
}

// The tensor-core paths walk a large batch in sub-batches through the WHOLE layer stack (not layer by layer over the full batch), so
// that the activation workspace stops growing with the batch: 2^25 samples per channel per sub-batch for bf16 (256 chunks of 2^17:
// 8 GB of workspace instead of 160 GB at B = 4096), 2^23 for the fp32 planes.  Speed is unchanged (B = 512: 49.0 ms either way;
// sub-batches of 64 chunks and less lose 1-5 %, `tools/time_subbatch.py`) -- the 9 % that B = 64 gains over B = 512 in
// profiles/encode_sweep_r01.json is burst clocks, not L2 residency.  AA_ENC_SUB_SAMPLES overrides the target.
static int64_t sub_batch(int64_t batch, int64_t n, bool bf16) {
  static const int64_t forced = getenv("AA_ENC_SUB_SAMPLES") ? std::max<int64_t>(1, atoll(getenv("AA_ENC_SUB_SAMPLES"))) : 0;
  const int64_t target = forced ? forced : (bf16 ? (256LL << 17) : (64LL << 17));
  return std::min(batch, std::max<int64_t>(1, target / std::max<int64_t>(n, 1)));
}

int64_t aa_encoder_workspace_bytes(const AaEncoder* e, int64_t batch, int64_t n, int dtype) {
  if (!e) return 0;
  const int64_t elems = max_act_elems(e, batch, n);
  const int64_t sb = sub_batch(batch, n, dtype == AA_DTYPE_BF16);
  if (dtype == AA_DTYPE_BF16) return aa::tc_workspace_bytes(e->layers, sb, n);
  if (dtype == AA_DTYPE_TF32X3) return aa::tf_workspace_bytes(e->layers, sb, n);
  const int64_t cuda_core_bytes = 3 * elems * (int64_t)sizeof(float) + 256;
  if (dtype == AA_DTYPE_F32 && fp32_on_tensor_cores(e)) return std::max(cuda_core_bytes, aa::tf_workspace_bytes(e->layers, sb, n));
  return cuda_core_bytes;
}

int aa_encoder_forward(AaEncoder* e, const float* const* stems_host, const float* faders_host, int n_stems, int64_t batch,
                       int64_t n, int apply_tanh, int dtype, float* y, void* workspace, void* stream) {
  AA_REQUIRE(e && stems_host && y && workspace, "NULL argument");
  AA_REQUIRE(n_stems >= 1 && n_stems <= 4, "n_stems=%d must be in [1,4]", n_stems);
  AA_REQUIRE(batch >= 0 && n >= 1, "bad shape");
  if (batch == 0) return AA_OK;
  for (int s = 0; s < n_stems; ++s) AA_REQUIRE(stems_host[s] != nullptr, "stem %d is NULL", s);
  const bool bf16 = dtype == AA_DTYPE_BF16;
  if (bf16 || dtype == AA_DTYPE_TF32X3 || (dtype == AA_DTYPE_F32 && fp32_on_tensor_cores(e))) {
    if (bf16 && !e->tc) {
      int rc = aa::tc_create(&e->tc, e->layers);
      if (rc != AA_OK) return rc;
    }
    if (!bf16 && !e->tf) {
      int rc = aa::tf_create(&e->tf, e->layers);
      if (rc != AA_OK) return rc;
    }
    int64_t t_out = 0;
    int rc = aa_encoder_out_length(e, n, &t_out);
    if (rc != AA_OK) return rc;
    const int64_t sb = sub_batch(batch, n, bf16);
    for (int64_t b0 = 0; b0 < batch; b0 += sb) {
      const int64_t nb = std::min(sb, batch - b0);
      const float* stems_sub[4];
      for (int s = 0; s < n_stems; ++s) stems_sub[s] = stems_host[s] + b0 * e->cfg.in_channels * n;
      float* y_sub = y + b0 * e->cfg.latent_dim * t_out;
      rc = bf16 ? aa::tc_forward(e->tc, e->layers, e->w, e->b, stems_sub, faders_host, n_stems, nb, n, apply_tanh, y_sub, workspace,
                                 (cudaStream_t)stream)
                : aa::tf_forward(e->tf, e->layers, e->w, e->b, stems_sub, faders_host, n_stems, nb, n, apply_tanh, y_sub, workspace,
                                 (cudaStream_t)stream);
      if (rc != AA_OK) return rc;
    }
    return AA_OK;
  }
  AA_REQUIRE(dtype == AA_DTYPE_F32 || dtype == AA_DTYPE_F32_CUDA_CORES, "unknown dtype %d", dtype);
  AA_REQUIRE(batch <= 65535, "fp32 path: batch <= 65535 per call");
  const int64_t elems = max_act_elems(e, batch, n);
  float* buf[3] = {reinterpret_cast<float*>(workspace), reinterpret_cast<float*>(workspace) + elems,
                   reinterpret_cast<float*>(workspace) + 2 * elems};
  int cur = -1;            // buffer holding the current activation (-1: the input stems)
  int res_buf = -1;        // buffer holding the residual-unit input
  int64_t l = n;
  for (size_t i = 0; i < e->layers.size(); ++i) {
    const auto& ly = e->layers[i];
    const int64_t lout = (l + 2 * ly.pad - (int64_t)ly.dil * (ly.k - 1) - 1) / ly.stride + 1;
    AA_REQUIRE(lout >= 1 && l < (1LL << 31), "input too short / too long for layer %zu", i);
    ConvArgs a{};
    if (cur < 0) {
      a.n_in = n_stems;
      for (int s = 0; s < n_stems; ++s) { a.x[s] = stems_host[s]; a.fader[s] = faders_host ? faders_host[s] : 1.0f; }
    } else {
      a.n_in = 1; a.x[0] = buf[cur]; a.fader[0] = 1.0f;
    }
    const bool last = (i + 1 == e->layers.size());
    int dst = 0;
    while (dst == cur || dst == res_buf) ++dst;
    if (ly.role == aa::ROLE_RES_FIRST) res_buf = cur;          // keep x of the residual unit
    a.w = e->w[i]; a.bias = e->b[i];
    a.res = (ly.role == aa::ROLE_RES_SECOND) ? buf[res_buf] : nullptr;
    a.out = last ? y : buf[dst];
    a.cin = ly.cin; a.cout = ly.cout; a.lin = (int)l; a.lout = (int)lout; a.k = ly.k; a.stride = ly.stride; a.dil = ly.dil;
    a.pad = ly.pad; a.elu = ly.elu; a.tanh_out = (last && apply_tanh) ? 1 : 0;
    dim3 grid((unsigned)((lout + LT - 1) / LT), (unsigned)((ly.cout + CT - 1) / CT), (unsigned)batch);
    conv1d_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    AA_LAUNCH_CHECK();
    if (ly.role == aa::ROLE_RES_SECOND) res_buf = -1;
    cur = dst;
    l = lout;
  }
  return AA_OK;
}

#pragma GCC visibility pop
}  // extern "C"
