// Fused loss entry points of the two training steps: ONE C call evaluates all loss terms of
//   train_aa_mixer_accel.py:504-517   loss = mse(zsum, zmix) + (var(zsum) + var(zmix)) / 2 + (cov(zsum) + cov(zmix)) / 2
//                                            + mse(y, y_recon) + mse(ymix, ymix_recon)
//   train_aa_effects.py:66-82         loss = (mse(zb2 - zb1 + za1, za2) + mse(za2 - za1 + zb1, zb2)) / 2 + mean_4 var_l2(z) + mean_4 cov(z)
//                                            + sum_4 mse(yrecon_i, y_i)
// and ONE C call writes every gradient; the per-term kernels (aa_ops.cu, cov_tc.cu) are launched back to back on the caller's
// stream and accumulate straight into the gradient buffers (no autograd graph of ten Functions, no ATen adds between them).
// losses[5] = {loss, mix_loss, var_loss, cov_loss, aa_recon_loss} (device floats, the reference's log_dict entries).
#include "aa_common.cuh"

#include <algorithm>

namespace {

// out = {sum, t[0] * k[0] .. } style combine of up to 16 partial scalars: term j = sum_i w[j][i] * t[i]
struct CombineArgs {
  const float* t;
  int n;
  float w[4][16];
  float* losses;
};
__global__ void loss_combine_kernel(const CombineArgs a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float term[4];
  for (int j = 0; j < 4; ++j) {
    float s = 0.f;
    for (int i = 0; i < a.n; ++i)
      if (a.w[j][i] != 0.f) s += a.w[j][i] * a.t[i];
    term[j] = s;
  }
  a.losses[1] = term[0]; a.losses[2] = term[1]; a.losses[3] = term[2]; a.losses[4] = term[3];
  a.losses[0] = ((term[0] + term[1]) + term[2]) + term[3];   // mix + var + cov + recon, the reference's order
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int64_t aa_mixer_loss_saved_floats(int64_t b, int64_t d) { return 4 * d + 2 * b * b; }
int64_t aa_effects_loss_saved_floats(int64_t b, int64_t d) { return 8 * d + 4 * b * b + 2 * b * d; }
int64_t aa_fused_loss_workspace_floats(int64_t b, int64_t d) { return aa_cov_loss_workspace_floats(b, d) + 64; }

#define AA_TRY(call)            \
  do {                          \
    int _rc = (call);           \
    if (_rc != AA_OK) return _rc; \
  } while (0)

int aa_mixer_loss_fwd_f32(const float* zsum, const float* zmix, const float* y, const float* y_recon, const float* ymix,
                          const float* ymix_recon, int64_t b, int64_t c, int64_t t, int hinge_l2, float gamma, float eps,
                          float* losses, float* saved, float* workspace, void* stream) {
  AA_REQUIRE(zsum && zmix && y && y_recon && ymix && ymix_recon && losses && saved && workspace, "NULL argument");
  const int64_t d = c * t, n = b * d;
  float* tmp = workspace;            // 7 partial scalars
  float* ws = workspace + 64;
  float* st_s = saved; float* st_m = saved + 2 * d; float* gr_s = saved + 4 * d; float* gr_m = gr_s + b * b;
  AA_TRY(aa_mse_fwd_f32(zsum, zmix, n, tmp + 0, ws, stream));
  AA_TRY(aa_vicreg_var_fwd_f32(zsum, b, d, gamma, eps, hinge_l2, tmp + 1, st_s, ws, stream));
  AA_TRY(aa_vicreg_var_fwd_f32(zmix, b, d, gamma, eps, hinge_l2, tmp + 2, st_m, ws, stream));
  AA_TRY(aa_vicreg_cov_fwd_f32(zsum, b, d, st_s, nullptr, gr_s, tmp + 3, ws, stream));
  AA_TRY(aa_vicreg_cov_fwd_f32(zmix, b, d, st_m, nullptr, gr_m, tmp + 4, ws, stream));
  AA_TRY(aa_mse_fwd_f32(y, y_recon, n, tmp + 5, ws, stream));
  AA_TRY(aa_mse_fwd_f32(ymix, ymix_recon, n, tmp + 6, ws, stream));
  CombineArgs ca{};
  ca.t = tmp; ca.n = 7; ca.losses = losses;
  ca.w[0][0] = 1.f; ca.w[1][1] = 0.5f; ca.w[1][2] = 0.5f; ca.w[2][3] = 0.5f; ca.w[2][4] = 0.5f; ca.w[3][5] = 1.f; ca.w[3][6] = 1.f;
  loss_combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ca);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

// g_* = gloss * d loss / d *  (gloss: device scalar or NULL = 1); every output is overwritten.  y and ymix carry no gradient
// (the given model is frozen: train_aa_mixer_accel.py:512-515 encodes under no_grad / the archive is detached).
int aa_mixer_loss_bwd_f32(const float* zsum, const float* zmix, const float* y, const float* y_recon, const float* ymix,
                          const float* ymix_recon, int64_t b, int64_t c, int64_t t, int hinge_l2, float gamma, float eps,
                          const float* saved, const float* gloss, float* g_zsum, float* g_zmix, float* g_y_recon,
                          float* g_ymix_recon, void* stream) {
  AA_REQUIRE(zsum && zmix && y && y_recon && ymix && ymix_recon && saved && g_zsum && g_zmix && g_y_recon && g_ymix_recon,
             "NULL argument");
  const int64_t d = c * t, n = b * d;
  const float* st_s = saved; const float* st_m = saved + 2 * d; const float* gr_s = saved + 4 * d; const float* gr_m = gr_s + b * b;
  AA_TRY(aa_mse_bwd_f32(zsum, zmix, n, gloss, 1.0f, g_zsum, g_zmix, 0, stream));
  AA_TRY(aa_vicreg_var_bwd_f32(zsum, st_s, b, d, gamma, eps, hinge_l2, gloss, 0.5f, g_zsum, 1, stream));
  AA_TRY(aa_vicreg_var_bwd_f32(zmix, st_m, b, d, gamma, eps, hinge_l2, gloss, 0.5f, g_zmix, 1, stream));
  AA_TRY(aa_vicreg_cov_bwd_f32(zsum, st_s, gr_s, b, d, gloss, 0.5f, g_zsum, 1, stream));
  AA_TRY(aa_vicreg_cov_bwd_f32(zmix, st_m, gr_m, b, d, gloss, 0.5f, g_zmix, 1, stream));
  AA_TRY(aa_mse_bwd_f32(y, y_recon, n, gloss, 1.0f, nullptr, g_y_recon, 0, stream));
  AA_TRY(aa_mse_bwd_f32(ymix, ymix_recon, n, gloss, 1.0f, nullptr, g_ymix_recon, 0, stream));
  return AA_OK;
}

// zs = {za1, zb1, za2, zb2}, ys / yrecons likewise (host arrays of 4 device pointers)
int aa_effects_loss_fwd_f32(const float* const* zs, const float* const* ys, const float* const* yrecons, int64_t b, int64_t c,
                            int64_t t, float gamma, float eps, float* losses, float* saved, float* workspace, void* stream) {
  AA_REQUIRE(zs && ys && yrecons && losses && saved && workspace, "NULL argument");
  const int64_t d = c * t, n = b * d;
  float* tmp = workspace;            // 14 partial scalars
  float* ws = workspace + 64;
  float* guess1 = saved + 8 * d + 4 * b * b;   // za2_guess = zb2 - zb1 + za1
  float* guess2 = guess1 + n;                  // zb2_guess = za2 - za1 + zb1
  const float cf[3] = {1.f, -1.f, 1.f};
  const float* g1[3] = {zs[3], zs[1], zs[0]};
  const float* g2[3] = {zs[2], zs[0], zs[1]};
  AA_TRY(aa_latent_lincomb_f32(3, g1, cf, guess1, n, stream));
  AA_TRY(aa_latent_lincomb_f32(3, g2, cf, guess2, n, stream));
  AA_TRY(aa_mse_fwd_f32(guess1, zs[2], n, tmp + 0, ws, stream));
  AA_TRY(aa_mse_fwd_f32(guess2, zs[3], n, tmp + 1, ws, stream));
  for (int i = 0; i < 4; ++i) {
    float* st = saved + 2 * d * i;
    float* gr = saved + 8 * d + (int64_t)i * b * b;
    AA_TRY(aa_vicreg_var_fwd_f32(zs[i], b, d, gamma, eps, 1, tmp + 2 + i, st, ws, stream));
    AA_TRY(aa_vicreg_cov_fwd_f32(zs[i], b, d, st, nullptr, gr, tmp + 6 + i, ws, stream));
    AA_TRY(aa_mse_fwd_f32(yrecons[i], ys[i], n, tmp + 10 + i, ws, stream));
  }
  CombineArgs ca{};
  ca.t = tmp; ca.n = 14; ca.losses = losses;
  ca.w[0][0] = 0.5f; ca.w[0][1] = 0.5f;
  for (int i = 0; i < 4; ++i) { ca.w[1][2 + i] = 0.25f; ca.w[2][6 + i] = 0.25f; ca.w[3][10 + i] = 1.f; }
  loss_combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(ca);
  AA_LAUNCH_CHECK();
  return AA_OK;
}

// g_zs[4], g_yrecons[4]: host arrays of device pointers, every output overwritten; ys carry no gradient (frozen given model)
int aa_effects_loss_bwd_f32(const float* const* zs, const float* const* ys, const float* const* yrecons, int64_t b, int64_t c,
                            int64_t t, float gamma, float eps, const float* saved, const float* gloss, float* const* g_zs,
                            float* const* g_yrecons, void* stream) {
  AA_REQUIRE(zs && ys && yrecons && saved && g_zs && g_yrecons, "NULL argument");
  const int64_t d = c * t, n = b * d;
  const float* guess1 = saved + 8 * d + 4 * b * b;
  const float* guess2 = guess1 + n;
  // mix term: e1 = guess1 - za2 enters za1 (+), zb1 (-), zb2 (+) through the guess and za2 (-) as the target; e2 = guess2 - zb2
  // enters za2 (+), za1 (-), zb1 (+) and zb2 (-).  mse_bwd writes +g into grad_a and -g into grad_b.
  //   pass 1: g(e1) -> g_zb2 (as grad_a, overwrite), g_za2 (as grad_b, overwrite)
  //   pass 2: g(e2) -> g_za2 (+=), g_zb2 (+=, negative sign through grad_b)
  //   then g_za1 = g(e1) - g(e2) = g_zb2_total restricted ... computed by two more passes below (same arithmetic, no temporaries)
  AA_TRY(aa_mse_bwd_f32(guess1, zs[2], n, gloss, 0.5f, g_zs[3], g_zs[2], 0, stream));   // g_zb2 = +g1, g_za2 = -g1
  AA_TRY(aa_mse_bwd_f32(guess2, zs[3], n, gloss, 0.5f, g_zs[2], g_zs[3], 1, stream));   // g_za2 += g2, g_zb2 -= g2
  AA_TRY(aa_mse_bwd_f32(guess1, zs[2], n, gloss, 0.5f, g_zs[0], g_zs[1], 0, stream));   // g_za1 = +g1, g_zb1 = -g1
  AA_TRY(aa_mse_bwd_f32(guess2, zs[3], n, gloss, 0.5f, g_zs[1], g_zs[0], 1, stream));   // g_zb1 += g2, g_za1 -= g2
  for (int i = 0; i < 4; ++i) {
    const float* st = saved + 2 * d * i;
    const float* gr = saved + 8 * d + (int64_t)i * b * b;
    AA_TRY(aa_vicreg_var_bwd_f32(zs[i], st, b, d, gamma, eps, 1, gloss, 0.25f, g_zs[i], 1, stream));
    AA_TRY(aa_vicreg_cov_bwd_f32(zs[i], st, gr, b, d, gloss, 0.25f, g_zs[i], 1, stream));
    AA_TRY(aa_mse_bwd_f32(yrecons[i], ys[i], n, gloss, 1.0f, g_yrecons[i], nullptr, 0, stream));
  }
  return AA_OK;
}

#pragma GCC visibility pop
}  // extern "C"
