"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the REAL reference code.

Run in the build container (where /root/reference is mounted):
    python -m oracle.make_golden
It imports the reference's own given_models.py / aa_mixer.py / aa_effects.py behind stub modules
(oracle/refload.py), feeds them seeded inputs on CPU in fp32, and stores inputs + outputs.
The committed fixtures are what pins oracle/aa_oracle.py (tests/test_oracle_golden.py) and, on the
GPU box, the CUDA kernels (tests/test_gpu_*.py) to the reference.  /root/reference is never read
at test time on the GPU box.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.refload import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def synth(shape, seed):
    "tonal + noise test signal (SURVEY.md section 8d), fp32 in [-1,1]"
    g = torch.Generator().manual_seed(seed)
    *lead, n = shape
    t = torch.arange(n, dtype=torch.float64) / 48000.0
    f = 55.0 + (8000.0 - 55.0) * torch.rand(*lead, 1, generator=g, dtype=torch.float64)
    ph = 2 * np.pi * torch.rand(*lead, 1, generator=g, dtype=torch.float64)
    x = 0.5 * torch.sin(2 * np.pi * f * t + ph) + 0.1 * torch.randn(*lead, n, generator=g, dtype=torch.float64)
    return x.clamp(-1, 1).float()


class ToyGivenModel(torch.nn.Module):
    "deterministic stand-in for f: [B,2,N] -> [B,64,N/64] (a strided, fixed random linear map + tanh)"

    def __init__(self, seed=7):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.register_buffer("w", torch.randn(64, 2, 64, generator=g) / 8.0)

    def encode(self, x):
        return torch.tanh(torch.nn.functional.conv1d(x, self.w, stride=64))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    gm, mx, fx = ref["given_models"], ref["aa_mixer"], ref["aa_effects"]
    torch.set_grad_enabled(False)

    # ---------------- STFT wrappers ----------------
    d = {}
    x_small = synth((2, 2, 8192), 11)
    d["x_small"] = x_small.numpy()
    for tag, (n_fft, hop) in {"2048_512": (2048, 512), "1024_256": (1024, 256)}.items():
        d[f"complex_{tag}"] = gm.SpectrogramAE(n_fft=n_fft, hop_length=hop).encode(x_small).numpy()
        d[f"power_{tag}"] = gm.MagSpectrogramAE(n_fft=n_fft, hop_length=hop).encode(x_small).numpy()
        d[f"mel_{tag}"] = gm.MelSpectrogramAE(sample_rate=48000, n_fft=n_fft, hop_length=hop).encode(x_small).numpy()
    # non power-of-two length: exercises zero_pad_po2 (given-models.ipynb cell 14 KAT shape [2,513,257])
    x_np2 = synth((2, 55728), 12)
    s = gm.SpectrogramAE().encode(x_np2)
    assert tuple(s.shape) == (2, 513, 257) and s.dtype == torch.complex64
    d["x_np2"] = x_np2.numpy()
    d["complex_np2_strided"] = s[:, ::4, ::4].contiguous().numpy()
    d["power_np2_strided"] = gm.MagSpectrogramAE().encode(x_np2)[:, ::4, ::4].contiguous().numpy()
    d["mel_np2"] = gm.MelSpectrogramAE().encode(x_np2).numpy()
    # odd length that is not a multiple of 4 (unaligned rows) and mono
    x_odd = synth((3, 1, 5001), 13)
    d["x_odd"] = x_odd.numpy()
    d["mel_odd_2048_512"] = gm.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512).encode(x_odd).numpy()
    d["power_odd_1024_256"] = gm.MagSpectrogramAE().encode(x_odd)[:, :, ::4, :].contiguous().numpy()
    # headline-shaped chunk: one [1,2,131072] mel (config 2 geometry)
    x_big = synth((1, 2, 131072), 14)
    d["x_big_seed"] = np.array([14])
    d["mel_big_2048_512"] = gm.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512).encode(x_big).numpy()
    # mag + dphase (unbatched semantics)
    x_mdp = synth((2, 4096), 15)
    d["x_mdp"] = x_mdp.numpy()
    d["magdphase"] = gm.MagDPhaseSpectrogramAE().encode(x_mdp).numpy()
    d["magdphase_use_cos"] = gm.MagDPhaseSpectrogramAE(use_cos=True).encode(x_mdp).numpy()
    d["magdphase_debug"] = gm.MagDPhaseSpectrogramAE(debug=True).encode(x_mdp).numpy()
    # mel filterbank itself
    import torchaudio
    d["fb_2048_48k_128"] = torchaudio.functional.melscale_fbanks(1025, 0.0, 24000.0, 128, 48000).numpy()
    d["fb_1024_48k_128"] = torchaudio.functional.melscale_fbanks(513, 0.0, 24000.0, 128, 48000).numpy()
    np.savez_compressed(os.path.join(OUT, "stft.npz"), **d)

    # ---------------- projector ----------------
    d = {}
    torch.manual_seed(2)  # train_aa_mixer_accel.py:51,479
    aa = mx.AudioAlgebra(dims=64, hidden_dims=64)
    for k, v in aa.state_dict().items():
        d["sd." + k] = v.numpy()
    g = torch.Generator().manual_seed(21)
    y = torch.randn(3, 64, 32, generator=g)
    z, yr = aa(y)
    d["y"], d["z"], d["y_recon"] = y.numpy(), z.numpy(), yr.numpy()
    d["z_encode"] = aa.encode(y).numpy()
    d["y_decode_of_y"] = aa.decode(y).numpy()
    # toy-size projector (dims=2, hidden=16; aa-mixer-toy.ipynb) for the generic-dims path
    torch.manual_seed(2)
    aat = mx.AudioAlgebra(dims=2, hidden_dims=16)
    for k, v in aat.state_dict().items():
        d["toy_sd." + k] = v.numpy()
    yt = torch.randn(5, 2, 7, generator=g)
    zt, yrt = aat(yt)
    d["toy_y"], d["toy_z"], d["toy_y_recon"] = yt.numpy(), zt.numpy(), yrt.numpy()
    # gradients of a scalar through the projector (for the backward kernels)
    torch.set_grad_enabled(True)
    aa.zero_grad()
    yg = y.clone().requires_grad_(True)
    zg, yrg = aa(yg)
    gz = torch.randn(3, 64, 32, generator=g)
    gyr = torch.randn(3, 64, 32, generator=g)
    ((zg * gz).sum() + (yrg * gyr).sum()).backward()
    d["gz"], d["gyr"], d["grad_y"] = gz.numpy(), gyr.numpy(), yg.grad.numpy()
    for k, p in aa.named_parameters():
        d["grad." + k] = p.grad.numpy()
    torch.set_grad_enabled(False)
    # standalone EmbedBlock (aa_mixer.py:205-221): 3-D rows without BN; 2-D rows with BatchNorm1d in training then eval mode
    torch.manual_seed(5)
    blk = mx.EmbedBlock(64, 64)
    xb = torch.randn(3, 7, 64)
    d["blk_w"], d["blk_b"], d["blk_x"], d["blk_y"] = blk.lin.weight.numpy(), blk.lin.bias.numpy(), xb.numpy(), blk(xb).numpy()
    blk2 = mx.EmbedBlock(16, 24, act=None)           # no residual (in != out), no activation
    xb2 = torch.randn(5, 16)
    d["blk2_w"], d["blk2_b"], d["blk2_x"], d["blk2_y"] = blk2.lin.weight.numpy(), blk2.lin.bias.numpy(), xb2.numpy(), blk2(xb2).numpy()
    blk3 = mx.EmbedBlock(8, 8, use_bn=True)
    blk3.bn.weight.data = torch.rand(8) + 0.5
    blk3.bn.bias.data = torch.randn(8) * 0.1
    xb3 = torch.randn(10, 8)
    d["blk3_w"], d["blk3_b"], d["blk3_bn_w"], d["blk3_bn_b"], d["blk3_x"] = (blk3.lin.weight.numpy(), blk3.lin.bias.numpy(),
                                                                               blk3.bn.weight.numpy().copy(), blk3.bn.bias.numpy().copy(), xb3.numpy())
    blk3.train()
    d["blk3_y_train"] = blk3(xb3).numpy()
    d["blk3_run_mean"], d["blk3_run_var"] = blk3.bn.running_mean.numpy().copy(), blk3.bn.running_var.numpy().copy()
    blk3.eval()
    d["blk3_y_eval"] = blk3(xb3).numpy()
    np.savez_compressed(os.path.join(OUT, "projector.npz"), **d)

    # ---------------- losses ----------------
    d = {}
    g = torch.Generator().manual_seed(31)
    za = torch.randn(8, 64, 16, generator=g)
    zb = za + 0.3 * torch.randn(8, 64, 16, generator=g)
    d["za"], d["zb"] = za.numpy(), zb.numpy()
    d["mse"] = mx.mseloss(za, zb).numpy()
    d["var_a"] = mx.vicreg_var_loss(za).numpy()
    d["var_b_small"] = mx.vicreg_var_loss(0.2 * zb).numpy()  # hinge active
    d["cov_a"] = mx.vicreg_cov_loss(za.clone()).numpy()
    d["cov_b"] = mx.vicreg_cov_loss(zb.clone()).numpy()
    std = torch.sqrt((0.2 * zb).var(dim=0) + 1e-4)
    d["var_b_small_l2"] = torch.mean(torch.relu(1 - std) ** 2).numpy()  # train_aa_effects.py:42-44
    # gradients
    torch.set_grad_enabled(True)
    for name, fn, inp in [("var", mx.vicreg_var_loss, 0.2 * zb), ("cov", mx.vicreg_cov_loss, za)]:
        t = inp.clone().requires_grad_(True)
        fn(t).backward()
        d[f"grad_{name}"] = t.grad.numpy()
    a = za.clone().requires_grad_(True)
    mx.mseloss(a, zb).backward()
    d["grad_mse_a"] = a.grad.numpy()
    torch.set_grad_enabled(False)
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **d)

    # ---------------- do_mixing (both flavours), faders ----------------
    d = {}
    toy = ToyGivenModel()
    d["toy_w"] = toy.w.numpy()
    g = torch.Generator().manual_seed(41)
    stems = [0.5 * torch.randn(4, 2, 2048, generator=g) for _ in range(2)]
    faders = torch.tensor([1.4630, -0.5718])  # aa-mixer-toy.ipynb cell 39
    zsum, zmix, arch = mx.do_mixing(stems, faders, toy, aa, "cpu")
    d["stem0"], d["stem1"], d["faders"] = stems[0].numpy(), stems[1].numpy(), faders.numpy()
    d["zsum"], d["zmix"] = zsum.numpy(), zmix.numpy()
    d["ymix"], d["ymix_recon"], d["mix"], d["ysum"] = (arch["ymix"].numpy(), arch["ymix_recon"].numpy(),
                                                      arch["mix"].numpy(), arch["ysum"].numpy())
    for i in range(2):
        d[f"zs{i}"], d[f"ys{i}"], d[f"yrecons{i}"] = arch["zs"][i].numpy(), arch["ys"][i].numpy(), arch["yrecons"][i].numpy()
    # mixer loss terms exactly as train_aa_mixer_accel.py:504-517
    y = toy.encode(stems[0])
    z, yrecon = aa(y)
    d["L_mix"] = mx.mseloss(zsum, zmix).numpy()
    d["L_var"] = ((mx.vicreg_var_loss(zsum) + mx.vicreg_var_loss(zmix)) / 2).numpy()
    d["L_cov"] = ((mx.vicreg_cov_loss(zsum.clone()) + mx.vicreg_cov_loss(zmix.clone())) / 2).numpy()
    d["L_rec"] = (mx.mseloss(y, yrecon) + mx.mseloss(arch["ymix"], arch["ymix_recon"])).numpy()
    # effects flavour
    batch = {k: 0.5 * torch.randn(4, 2, 2048, generator=g) for k in ("a1", "b1", "a2", "b2")}
    arch2 = fx.do_mixing(batch, toy, aa, "cpu")
    for k in ("a1", "b1", "a2", "b2"):
        d["fx_" + k] = batch[k].numpy()
    for i in range(4):
        d[f"fx_ys{i}"], d[f"fx_zs{i}"], d[f"fx_yrecons{i}"] = (arch2["ys"][i].numpy(), arch2["zs"][i].numpy(),
                                                                arch2["yrecons"][i].numpy())
    # faders from the reference's RNG recipe
    random.seed(0)
    torch.manual_seed(0)
    st, fd, _ = mx.get_stems_faders(stems[0], iter([stems[1]]), [stems[1]], maxstems=2)
    d["faders_seed0"] = fd.numpy()
    torch.manual_seed(0)
    d["faders_seed0_u"] = torch.stack([torch.rand(2), torch.rand(2)]).numpy()
    np.savez_compressed(os.path.join(OUT, "mixing.npz"), **d)

    # ---------------- PCA accumulation (calc_effects_pca.py:76-94) ----------------
    d = {}
    g = torch.Generator().manual_seed(51)
    from einops import rearrange
    cov_num, npoints = None, 0
    for bi in range(2):
        ys = torch.tanh(torch.randn(4, 64, 32, generator=g) + 0.1 * bi)
        d[f"ys{bi}"] = ys.numpy()
        yy = rearrange(ys, 'b d n -> d (b n)')
        npoints += yy.shape[1]
        c = torch.cov(yy) * (yy.shape[1] - 1)
        cov_num = c if cov_num is None else cov_num + c
    cov = cov_num / (npoints - 1)
    lam = torch.sort(torch.linalg.eigh(cov)[0], descending=True)[0]
    d["cov_numerator"], d["npoints"], d["lambdas"] = cov_num.numpy(), np.array([npoints]), lam.numpy()
    np.savez_compressed(os.path.join(OUT, "pca.npz"), **d)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
