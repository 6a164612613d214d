/*
 * aa_b200.h -- C ABI of libaa_b200.so: the B200-native (sm_100a) implementation of the
 * audio-algebra hot path (batched STFT/mel front-end, given-model conv encoder, AudioAlgebra
 * projector, latent algebra and mixer/effects losses).
 *
 * The reference (drscotthawley/audio-algebra) is pure Python and has no FFI: its "plugin
 * interface" for this path is the duck-typed Python protocol in audio_algebra/given_models.py
 * (GivenModelClass.encode) and audio_algebra/aa_mixer.py (AudioAlgebra, do_mixing, losses).
 * Every entry point below cites the reference lines whose arithmetic it replaces; the Python
 * package `audio-algebra_b200/` binds them with ctypes and keeps the
 * reference's class / function names.
 *
 * Conventions
 *  - plain pointers and sizes only; all tensor pointers are DEVICE pointers unless the name ends
 *    in `_host`; tensors are dense, row-major, the layouts given per function;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every call is
 *    asynchronous and stream-ordered unless stated otherwise;
 *  - return value: 0 = AA_OK, negative = AA_ERR_*; aa_last_error() returns a thread-local message;
 *  - handles (AaStftPlan, AaEncoder) belong to the CUDA device that was current at creation and
 *    are not internally locked; distinct handles / streams may be used concurrently.
 */
#ifndef AA_B200_H
#define AA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AA_OK 0
#define AA_ERR_INVALID_ARG (-1)  /* bad shape / null pointer / unsupported parameter combination */
#define AA_ERR_CUDA (-2)         /* a CUDA runtime call failed; message holds cudaGetErrorString  */
#define AA_ERR_UNSUPPORTED (-3)  /* valid request that this build has no kernel for               */
#define AA_ERR_ARCH (-4)         /* device is not sm_100 (the library ships sm_100a code only)    */

#define AA_DTYPE_F32 0
#define AA_DTYPE_BF16 1
#define AA_DTYPE_TF32X3 2 /* encoder only: fp32 operands split into two TF32 parts, three tcgen05 MMAs per K step, fp32 accumulate */
#define AA_DTYPE_F32_CUDA_CORES 3 /* encoder only: the CUDA-core fp32 kernel, whatever the layer shapes */

int aa_version(void);
const char* aa_last_error(void);
/* 0 if the current device can run this library (compute capability 10.x), else AA_ERR_ARCH. */
int aa_check_device(void);
/* number of kernels this library has launched from the calling process (all threads) */
int64_t aa_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * STFT / power / mel front-end.
 * Replaces torchaudio.transforms.{Spectrogram,MelSpectrogram} as called from
 * given_models.py:158 (SpectrogramAE, power=None), :180 (MagSpectrogramAE, power=2) and
 * :267 (MelSpectrogramAE), and GivenModelClass.zero_pad_po2 (given_models.py:139-145), which is
 * folded into the load (n_in -> n_pad = next power of two, never materialised).
 * Semantics: torch.stft(center, pad_mode="reflect", win_length=n_fft, onesided, normalized=False);
 * mel = fb^T |X|^2 with fb = melscale_fbanks(n_fft/2+1, f_min, f_max, n_mels, sample_rate,
 * norm=None, mel_scale="htk").
 * ------------------------------------------------------------------------------------------ */
typedef struct AaStftPlan AaStftPlan;

/* window_host: n_fft floats or NULL (periodic Hann).  fb_host: [n_fft/2+1][n_mels] floats or NULL
 * (HTK triangular filterbank built internally).  n_mels = 0 for a plan without a mel stage.
 * n_fft must be a power of two in [64, 8192]; hop >= 1. */
int aa_stft_plan_create(AaStftPlan** plan, int n_fft, int hop, int center, const float* window_host,
                        int n_mels, float sample_rate, float f_min, float f_max, const float* fb_host);
int aa_stft_plan_destroy(AaStftPlan* plan);
/* n_pad = zero_pad ? next_pow2(n_in) : n_in;  n_frames = center ? 1 + n_pad/hop : 1 + (n_pad-n_fft)/hop */
int aa_stft_out_shape(const AaStftPlan* plan, int64_t n_in, int zero_pad, int64_t* n_pad, int64_t* n_frames);

/* wav [rows][n_in] f32 -> out [rows][n_fft/2+1][n_frames] complex64 (interleaved re,im) */
int aa_stft_complex_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                        float* out, void* stream);
/* -> out [rows][n_fft/2+1][n_frames] f32 = |X|^2 */
int aa_stft_power_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                      float* out, void* stream);
/* The same two, written frequency-minor: out [rows][n_frames][n_fft/2+1].  This is the memory layout of the reference's own
 * tensors -- torch.stft / torchaudio Spectrogram return [.., freq, time] as a transposed VIEW of a [.., time, freq] buffer
 * (given_models.py:161,182 call them unchanged) -- and the one the kernels write at full sector width (rows of n_fft/2+1
 * contiguous values instead of 24..32-byte runs along a 257-frame axis).  The Python wrappers call these and return
 * out.transpose(-1, -2): same shape, values and strides as the reference. */
int aa_stft_complex_tf_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                           float* out, void* stream);
int aa_stft_power_tf_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                         float* out, void* stream);
/* -> out [rows][n_mels][n_frames] f32 */
int aa_stft_mel_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                    float* out, void* stream);
/* The mel spectrogram written mel-minor: out [rows][n_frames][n_mels].  Again the reference's own memory layout: torchaudio's
 * MelScale returns matmul(spec^T, fb)^T, i.e. MelSpectrogramAE.encode (given_models.py:267,277) hands back a [.., n_mels, time]
 * VIEW of a [.., time, n_mels] buffer (strides (.., 1, n_mels)).  One warp produces all n_mels values of a (row, frame) and
 * writes them as one contiguous line.  Needs n_fft = 2048 with the Hann window and a triangular (<= 2 adjacent taps per bin)
 * filterbank: aa_stft_mel_tf_supported() returns 1 when the plan can run it, else callers use aa_stft_mel_f32. */
int aa_stft_mel_tf_f32(const AaStftPlan* plan, const float* wav, int64_t rows, int64_t n_in, int zero_pad,
                       float* out, void* stream);
int aa_stft_mel_tf_supported(const AaStftPlan* plan);
/* MagDPhaseSpectrogramAE.encode epilogue (given_models.py:214-231, use_cos=False): spec [c][F][T]
 * complex64 -> out [2c][F][T] f32: magnitudes then phase differences along T (negative differences
 * wrapped by +2*3.141592653589, frame 0 keeps the phase). */
int aa_magdphase_f32(const float* spec, int64_t c, int64_t n_freq, int64_t n_frames, float* out, void* stream);
/* End-to-end helper for the bulk encode loop (xae_dataset.ipynb cell 50): host wav -> host mel, chunked
 * over rows, H2D / kernel / D2H overlapped on internal streams; synchronous on return.  Host buffers
 * should be pinned for full PCIe bandwidth (pageable memory works, slower). */
int aa_stft_mel_f32_host(const AaStftPlan* plan, const float* wav_host, int64_t rows, int64_t n_in, int zero_pad,
                         float* out_host, int64_t rows_per_chunk);
/* same, out_host [rows][n_frames][n_mels] (the layout of aa_stft_mel_tf_f32) */
int aa_stft_mel_tf_f32_host(const AaStftPlan* plan, const float* wav_host, int64_t rows, int64_t n_in, int zero_pad,
                            float* out_host, int64_t rows_per_chunk);

/* ------------------------------------------------------------------------------------------
 * Latent algebra (aa_mixer.py:307 zsum; train_aa_effects.py:70-71 guesses; Destructo.ipynb
 * cells 22, 48-49).  All tensors f32, n elements, contiguous.
 * ------------------------------------------------------------------------------------------ */
/* out[i] = sum_j coeffs_host[j] * zs[j][i], 1 <= n_terms <= 8; `zs_host` is a host array of device
 * pointers.  out may alias any input. */
int aa_latent_lincomb_f32(int n_terms, const float* const* zs_host, const float* coeffs_host, float* out,
                          int64_t n, void* stream);
#define AA_UNARY_SIGN_FOLD 0    /* max(z) * (sign(z) - z)        Destructo.ipynb cell 22 */
#define AA_UNARY_ABSMAX_MINUS 1 /* max|z| - z                                            */
#define AA_UNARY_TANH_DRIVE 2   /* max(z) * tanh(param * z)                              */
#define AA_UNARY_FLIP_CHANNELS 3 /* z.flip(dims=[1]) on [B][C][T]; needs c,t               */
#define AA_UNARY_FLIP_TIME 4    /* z.flip(dims=[2])                                      */
/* workspace: >= 2 floats of device scratch (global max reductions). */
int aa_latent_unary_f32(int op, const float* z, float* out, int64_t b, int64_t c, int64_t t, float param,
                        float* workspace, void* stream);
/* Destructo.ipynb cell 48-49: out[b] = emb[b] + mean_b'(wet[b'] - dry[b']); emb [B][C*T], wet/dry [Bw][C*T] */
int aa_effect_transfer_f32(const float* emb, int64_t b, const float* wet, const float* dry, int64_t bw,
                           int64_t ct, float* out, void* stream);

/* --- remaining Destructo.ipynb cell 22 operations (csrc/latent_ext.cu); z, out are [B][C][T] f32 --------------------------- */
/* "wavy": out = z * vec[t]; the caller passes vec = cos(linspace(0, 4*6.28, T)) computed as the notebook computes it */
int aa_latent_mul_time_f32(const float* z, const float* vec, float* out, int64_t b, int64_t c, int64_t t, void* stream);
/* "flippy": out = z + flip(z, -1) (not in place) */
int aa_latent_add_flip_time_f32(const float* z, float* out, int64_t b, int64_t c, int64_t t, void* stream);
/* "kill_half": out = z with channels [c0, c1) zeroed (the notebook: z[:, 33:-1, :] = 0) */
int aa_latent_zero_channels_f32(const float* z, float* out, int64_t b, int64_t c, int64_t t, int64_t c0, int64_t c1, void* stream);
/* "call_and_response" / "hurt_drums": out = a*x + (b*z) * (2*u - 1) with u = torch.rand_like(z) supplied by the caller
 * (x = z, a = -1, b = rand_fac  /  x = embeddings, a = 1 - rand_fac, b = rand_fac) */
int aa_latent_randmix_f32(const float* x, const float* z, const float* u, float a, float b, float* out, int64_t n, void* stream);
/* "reverb_time": for i in range(T): z = z + coef[i] * shift_right(z, i+1) on the updated z; coef[i] = (float)exp(-i/reverb_time);
 * z, out [rows][T] */
int aa_latent_reverb_f32(const float* z, const float* coef, float* out, int64_t rows, int64_t t, void* stream);
/* Destructo.ipynb cell 49 with unequal lengths: diff = wet - dry [Bw][C][t_diff] is left-padded with zeros (F.pad(diff,
 * (length_difference, 0, ...))) or truncated to t_emb, averaged over Bw, and added to every emb[b] ([B][C][t_emb]). */
int aa_effect_transfer_ex_f32(const float* emb, int64_t b, int64_t c, int64_t t_emb, const float* wet, const float* dry,
                              int64_t bw, int64_t t_diff, float* out, void* stream);
/* cell 49 time_avg=True: out[row] = mean_t(wet[row][t] - dry[row][t]); then torch broadcasting of that 2-D tensor [d0][d1]
 * against [B][C][T] (d0 in {1, C}, d1 in {1, T}; anything else is the broadcasting error the notebook would raise) */
int aa_latent_row_mean_diff_f32(const float* wet, const float* dry, int64_t rows, int64_t t, float* out, void* stream);
int aa_latent_add_bcast2_f32(const float* z, int64_t b, int64_t c, int64_t t, const float* d, int64_t d0, int64_t d1, float* out,
                             void* stream);
/* MagDPhaseSpectrogramAE.encode with its use_cos (acos of the clipped normalised dot product of consecutive frames,
 * given_models.py:218-225) and debug (theta < 0 -> theta + 2 pi before differencing, :215) branches */
int aa_magdphase_ex_f32(const float* spec, int64_t c, int64_t n_freq, int64_t n_frames, int use_cos, int wrap_theta, float* out,
                        void* stream);
/* Standalone EmbedBlock (aa_mixer.py:205-221) on [n_tok][din] rows: y = Linear(x) (+ exact-erf GELU if act) (+ x if resid);
 * pre_out (optional) keeps Linear(x) for the backward.  BatchNorm1d(out_dims) of the use_bn variant is the separate pair below
 * (the reference applies it to [N, C] inputs -- the only rank its BatchNorm1d accepts there). */
int aa_embed_block_fwd_f32(const float* x, const float* w, const float* bias, int64_t n_tok, int din, int dout, int act, int resid,
                           float* y, float* pre_out, void* stream);
/* scratch: n_tok * dout floats; gx / gw / gb may be NULL (gb needs gw) */
int aa_embed_block_bwd_f32(const float* x, const float* w, const float* pre, const float* gy, int64_t n_tok, int din, int dout,
                           int act, int resid, float* gx, float* gw, float* gb, float* scratch, void* stream);
/* x, y [n][c]; save [2][c] (mean, rstd); training updates run_mean / run_var with `momentum` (unbiased variance) */
int aa_batchnorm_fwd_f32(const float* x, int64_t n, int c, const float* gamma, const float* beta, float* run_mean, float* run_var,
                         int training, float momentum, float eps, float* y, float* save, void* stream);
int aa_batchnorm_bwd_f32(const float* x, const float* gy, int64_t n, int c, const float* gamma, const float* save, int training,
                         float* gx, float* ggamma, float* gbeta, void* stream);

/* --- dormant branches of DiffusionDVAE.encode_it (aa_mixer.py:178-179, 189-192; csrc/quant_pqmf.cu) ------------------------ */
/* PQMF analysis (third-party diffusion.pqmf.PQMF(2, 70, bands).forward, classic form): x [rows][n] -> out [rows][bands][t_out],
 * t_out = (n + 2*(taps/2) - taps) / bands; out[row][k][t] = sum_j hk[k][j] xpad[row][t*bands + j] with xpad zero-padded by taps/2,
 * then the even samples of the odd bands are negated ("reverse_half").  hk [bands][taps] is designed on the host. */
int aa_pqmf_analysis_f32(const float* x, int64_t rows, int64_t n, const float* hk, int bands, int taps, float* out, void* stream);
/* Memcodes (third-party nwt_pytorch.Memcodes, eval path): keys / values of the codes, k[h][j][:] = Wk[h] codes[h][j][:] (grouped
 * 1x1 Conv1d, wk [heads*d][d]) ... */
int aa_memcodes_kv_f32(const float* codes, const float* wk, const float* wv, int heads, int n_codes, int d, float* k_out,
                       float* v_out, void* stream);
/* ... and the lookup: x [B][heads*d][N] (channel-major, what encoder_ema returns); per (b, head, n): j* = argmax_j <scale * x_slice,
 * k[h][j]>, quantized = v[h][j*].  Any of q_out (quantized), resid_out (x - quantized), acc (+= quantized, tanh afterwards when
 * final_tanh), idx_out ([B][heads][N] int64) may be NULL: one call per layer of dvae.residual_memcodes.ResidualMemcodes. */
int aa_memcodes_quantize_f32(const float* x, int64_t batch, int heads, int d, int64_t n_pos, const float* k, const float* v,
                             int n_codes, float scale, float* q_out, float* resid_out, float* acc, int final_tanh,
                             int64_t* idx_out, void* stream);

/* --- generic layers for given models outside the SoundStreamXL layer table (StackedDiffAEWrapper, given_models.py:361-385) --------- */
/* out = act(conv1d(x) + bias (+ res)), channel-major fp32: x [B][cin][lin], w [cout][cin][k], res / out [B][cout][lout];
 * act: 0 none, 1 ELU, 2 tanh; k <= 8 */
int aa_conv1d_f32(const float* x, int64_t batch, int cin, int64_t lin, const float* w, const float* bias, int cout, int k, int stride,
                  int dil, int pad, const float* res, int act, float* out, void* stream);
/* nn.GroupNorm(groups, c) (+ SiLU when silu != 0) on [B][c][l]; gamma / beta [c] or NULL */
int aa_groupnorm_act_f32(const float* x, int64_t batch, int c, int64_t l, const float* gamma, const float* beta, int groups, float eps,
                         int silu, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Losses (aa_mixer.py:344-364; L2-hinge variant train_aa_effects.py:42-46).
 * z is [B][D] f32 (D = C*T flattened features), the batch dimension is the statistics dimension.
 * Each *_fwd writes ONE f32 scalar to `loss` (device) and is deterministic (no float atomics).
 * ------------------------------------------------------------------------------------------ */
/* mean((a-b)^2); workspace >= aa_reduce_workspace_floats() floats */
int64_t aa_reduce_workspace_floats(void);
int aa_mse_fwd_f32(const float* a, const float* b, int64_t n, float* loss, float* workspace, void* stream);
/* grad_a[i] (+)= gscale * 2 (a-b)/n ; grad_b likewise with the opposite sign (either may be NULL) */
int aa_mse_bwd_f32(const float* a, const float* b, int64_t n, const float* gloss, float gscale,
                   float* grad_a, float* grad_b, int accumulate, void* stream);
/* vicreg_var_loss: mean_d(relu(gamma - sqrt(var_unbiased_b(z[:,d]) + eps))) (hinge_l2: squared relu).
 * Also emits per-feature mean and var (needed by bwd and by the cov loss): stats [2][D]. */
int aa_vicreg_var_fwd_f32(const float* z, int64_t b, int64_t d, float gamma, float eps, int hinge_l2,
                          float* loss, float* stats, float* workspace, void* stream);
int aa_vicreg_var_bwd_f32(const float* z, const float* stats, int64_t b, int64_t d, float gamma, float eps,
                          int hinge_l2, const float* gloss, float gscale, float* grad_z, int accumulate,
                          void* stream);
/* vicreg_cov_loss via the Gram identity: with X = z - mean_b(z) ([B][D]), G = X X^T ([B][B]),
 *   sum_{i!=j} cov_ij^2 = (||G||_F^2 - sum_d (sum_b X_bd^2)^2) / (B-1)^2,  loss = that / D,
 * instead of materialising the DxD covariance (aa_mixer.py:360-364).  gram: [B][B] f32 scratch that
 * also feeds the backward; stats as produced by aa_vicreg_var_fwd_f32 (or NULL: computed internally
 * into `stats_out`, [2][D]). */
/* workspace for aa_vicreg_cov_fwd_f32: >= aa_cov_loss_workspace_floats(b, d) floats */
int64_t aa_cov_loss_workspace_floats(int64_t b, int64_t d);
int aa_vicreg_cov_fwd_f32(const float* z, int64_t b, int64_t d, const float* stats, float* stats_out,
                          float* gram, float* loss, float* workspace, void* stream);
/* grad_z (+)= gscale * dL/dz,  dL/dX = 4/((B-1)^2 D) (G X - X diag(s)), s_d = sum_b X_bd^2, projected
 * onto zero-mean columns. */
int aa_vicreg_cov_bwd_f32(const float* z, const float* stats, const float* gram, int64_t b, int64_t d,
                          const float* gloss, float gscale, float* grad_z, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------
 * Given-model conv encoder: SoundStreamXLEncoder as constructed by DiffusionDVAE.__init__
 * (aa_mixer.py:118-131) and called by DiffusionDVAE.encode (:165-168, no tanh) / encode_it (:175-195,
 * tanh) / DVAEWrapper.encode (given_models.py:313-338).  The architecture is restated from the
 * third-party `audio-diffusion` package (parity unpinned upstream).
 * Weights: nn.Conv1d layout [cout][cin][k] f32 + bias [cout], layer order = the order
 * aa_encoder_layer_shape enumerates (conv_in; per block: 3 x (res conv k7, res conv k1), down conv; conv_out).
 * ------------------------------------------------------------------------------------------ */
typedef struct AaEncoder AaEncoder;
typedef struct AaEncoderCfg {
  int in_channels;  /* 2 * pqmf_bands                       */
  int capacity;     /* 32                                   */
  int latent_dim;   /* 64                                   */
  int n_blocks;     /* 5                                    */
  int c_mults[8];   /* {2,4,8,16,32}                        */
  int strides[8];   /* {4,4,2,2,2}                          */
} AaEncoderCfg;
int aa_encoder_create(const AaEncoderCfg* cfg, AaEncoder** enc);
int aa_encoder_destroy(AaEncoder* enc);
int aa_encoder_num_layers(const AaEncoder* enc);
int aa_encoder_layer_shape(const AaEncoder* enc, int layer, int* cout, int* cin, int* k);
/* copies device tensors into the handle (and invalidates packed bf16 copies) */
int aa_encoder_set_weights(AaEncoder* enc, int layer, const float* w_dev, const float* b_dev, void* stream);
int aa_encoder_out_length(const AaEncoder* enc, int64_t n, int64_t* t_out);
int64_t aa_encoder_workspace_bytes(const AaEncoder* enc, int64_t batch, int64_t n, int dtype);
/* y [B][latent][T'] f32 = encoder(sum_j faders_host[j] * stems[j]) (tanh applied iff apply_tanh).
 * stems_host: host array of n_stems (1..4) device pointers to [B][in_channels][n] f32 tensors; the
 * fader-scaled sum (aa_mixer.py:303,309) is fused into the first layer's load.  dtype selects the
 * arithmetic: AA_DTYPE_F32 (fp32-grade results: the 3xTF32 tensor-core kernels when the layer shapes allow -- capacity 32,
 * channel counts multiples of 32, even strides -- else the CUDA-core fp32 kernel), AA_DTYPE_TF32X3 (tcgen05 kind::tf32 with
 * hi/lo operand split, error if the shapes do not allow), AA_DTYPE_F32_CUDA_CORES (always the CUDA-core kernel) or
 * AA_DTYPE_BF16 (tcgen05 kind::f16, fp32 accumulate). */
int aa_encoder_forward(AaEncoder* enc, const float* const* stems_host, const float* faders_host, int n_stems,
                       int64_t batch, int64_t n, int apply_tanh, int dtype, float* y, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Projector (aa_mixer.py:205-267): one half (encode or decode) of AudioAlgebra =
 *   4 EmbedBlocks (Linear + exact-erf GELU on the first three, inner residual iff in==out) applied to
 *   x^T, transposed back, plus the outer residual -- fused in one kernel on the channel-major tensor.
 * x, out: [B][dims][T] f32.  w_host / b_host: host arrays of 4 device pointers, nn.Linear layout
 * ([out][in], [out]) with shapes dims->hidden, hidden->hidden, hidden->hidden, hidden->dims; dims, hidden <= 64.
 * ------------------------------------------------------------------------------------------ */
int aa_projector_half_fwd_f32(const float* const* w_host, const float* const* b_host, int dims, int hidden, int resid,
                              const float* x, int64_t batch, int64_t t, float* out, void* stream);
int64_t aa_projector_bwd_workspace_floats(void);
/* Backward of the same half: recomputes the activations from x.  gx (+)= dL/dx (may be NULL);
 * gw_host[l] / gb_host[l] (+)= gscale * dL/dW_l, dL/db_l (deterministic two-stage reduction). */
int aa_projector_half_bwd_f32(const float* const* w_host, const float* const* b_host, int dims, int hidden, int resid,
                              const float* x, const float* gout, int64_t batch, int64_t t, float* gx, int accumulate_gx,
                              float* const* gw_host, float* const* gb_host, int accumulate_gw, float gscale,
                              float* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Decoder side of the STFT given models (round-trip demos).
 * aa_istft_f32 replaces T.InverseSpectrogram / torch.istft (given_models.py:159,168): spec = complex64
 * (interleaved re, im) addressed as spec[row*row_stride + f*stride_f + t*stride_t] (strides in complex
 * elements: both torch.stft's [T][F]-memory view and a contiguous [F][T] tensor work); window = device
 * pointer to n_fft floats or NULL (periodic Hann); out [rows][out_len], out_len <= hop*(n_frames-1) for
 * center != 0 (torch's length=None); workspace >= aa_istft_workspace_floats floats.  Deterministic
 * (gather overlap-add, no float atomics).
 * aa_griffinlim_update_c64: one iteration's phase update of torchaudio.functional.griffinlim
 * (given_models.py:181,189,269): angles = rebuilt - momentum*tprev (skipped when first != 0),
 * angles /= |angles| + 1e-16, tprev <- rebuilt, prod = mag * angles; n complex elements, momentum already
 * divided by (1 + momentum).
 * aa_inverse_mel_f32 replaces T.InverseMelScale (given_models.py:268,279): out [rows][n_freq][n_frames] =
 * relu(pinv [n_freq][n_mels] x mel), mel addressed by (row, mel, frame) strides in elements.
 * ------------------------------------------------------------------------------------------ */
int64_t aa_istft_workspace_floats(int64_t rows, int n_fft, int64_t n_frames);
int aa_istft_f32(const void* spec, int64_t rows, int n_fft, int hop, int center, int64_t n_frames, int64_t row_stride,
                 int64_t stride_f, int64_t stride_t, const float* window, float* out, int64_t out_len, float* workspace,
                 void* stream);
int aa_griffinlim_update_c64(const void* rebuilt, void* tprev, const float* mag, void* prod, int64_t n, float momentum, int first,
                             void* stream);
/* MagDPhaseSpectrogramAE.decode (given_models.py:233-254): reps [2c][f][t] (magnitudes, then phase differences) -> complex64
 * spec [c][f][t] = mag * exp(i theta), theta integrated along time with the reference's 2 pi wrap; init_mode 0 = 'true'
 * (theta_0 = dtheta_0), 1 = 'rand' (theta_0 from theta0 [c][f]), 2 = 'zero'.  Feed spec to aa_istft_f32. */
int aa_magdphase_decode_f32(const float* reps, int64_t c, int64_t f, int64_t t, int init_mode, const float* theta0, void* spec, void* stream);
int aa_inverse_mel_f32(const float* pinv, const float* mel, int64_t rows, int n_mels, int n_freq, int64_t n_frames, int64_t mel_row_stride,
                       int64_t mel_stride_m, int64_t mel_stride_t, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused loss entry points of the two training steps: one call evaluates every loss term
 * and one call writes every gradient; the per-term kernels above run back to back on `stream` and
 * accumulate straight into the gradient buffers.  Replaces train_aa_mixer_accel.py:504-517 (mixer) and
 * train_aa_effects.py:66-82 (effects; L2-hinge variance loss).  All tensors [B][C][T] f32, d = C*T.
 * losses[5] (device) = {loss, mix_loss, var_loss, cov_loss, aa_recon_loss} = the reference's log_dict.
 * saved: >= aa_*_loss_saved_floats(b, d) floats, written by fwd and read by bwd (feature statistics,
 * Gram matrices, the two effect guesses); workspace: >= aa_fused_loss_workspace_floats(b, d) floats.
 * bwd: gloss = device scalar dL/dloss or NULL (= 1); every gradient output is overwritten; y / ymix / ys
 * carry no gradient (the given model is frozen).
 * ------------------------------------------------------------------------------------------ */
int64_t aa_mixer_loss_saved_floats(int64_t b, int64_t d);
int64_t aa_effects_loss_saved_floats(int64_t b, int64_t d);
int64_t aa_fused_loss_workspace_floats(int64_t b, int64_t d);
int aa_mixer_loss_fwd_f32(const float* zsum, const float* zmix, const float* y, const float* y_recon, const float* ymix,
                          const float* ymix_recon, int64_t b, int64_t c, int64_t t, int hinge_l2, float gamma, float eps,
                          float* losses, float* saved, float* workspace, void* stream);
int aa_mixer_loss_bwd_f32(const float* zsum, const float* zmix, const float* y, const float* y_recon, const float* ymix,
                          const float* ymix_recon, int64_t b, int64_t c, int64_t t, int hinge_l2, float gamma, float eps,
                          const float* saved, const float* gloss, float* g_zsum, float* g_zmix, float* g_y_recon,
                          float* g_ymix_recon, void* stream);
/* zs = {za1, zb1, za2, zb2}; ys / yrecons / g_zs / g_yrecons likewise: host arrays of 4 device pointers */
int aa_effects_loss_fwd_f32(const float* const* zs, const float* const* ys, const float* const* yrecons, int64_t b, int64_t c,
                            int64_t t, float gamma, float eps, float* losses, float* saved, float* workspace, void* stream);
int aa_effects_loss_bwd_f32(const float* const* zs, const float* const* ys, const float* const* yrecons, int64_t b, int64_t c,
                            int64_t t, float gamma, float eps, const float* saved, const float* gloss, float* const* g_zs,
                            float* const* g_yrecons, void* stream);

/* ------------------------------------------------------------------------------------------
 * PCA accumulation (calc_effects_pca.py:81-89): y [B][C][T] f32 -> cov_num [C][C] += C-by-C
 * batch-centred scatter of the B*T points; count += B*T.  (eigh of the 64x64 result stays on host.)
 * workspace: >= C*C*grid + 2*C floats, see aa_cov_workspace_floats.
 * ------------------------------------------------------------------------------------------ */
int64_t aa_cov_workspace_floats(int64_t c);
int aa_cov_accumulate_f32(const float* y, int64_t b, int64_t c, int64_t t, float* cov_num, double* count,
                          float* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser (train_aa_mixer_accel.py:481-482,533): torch.optim.Adam semantics on one flat f32
 * buffer; step counts from 1.
 * ------------------------------------------------------------------------------------------ */
int aa_adam_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                     float lr, float beta1, float beta2, float eps, int64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AA_B200_H */
