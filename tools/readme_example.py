"""The README quick-start snippet, runnable (dev check)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, audio_algebra_b200 as aab
x = torch.rand(32, 2, 131072, device="cuda") - 0.5
mel = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512).encode(x)
spec = aab.SpectrogramAE(n_fft=2048, hop_length=512).encode(x)
print(mel.shape, spec.shape, spec.dtype, spec.stride()[-2:])
dvae = aab.DVAEWrapper(debug=False).cuda()
fast = aab.DVAEWrapper(debug=False, compute_dtype="bf16").cuda()
reps = dvae.encode(x[:8])
print(reps.shape)
aa = aab.AudioAlgebra(dims=64, hidden_dims=64).cuda()
stems, faders = [x[:8], x[8:16]], [1.4630, -0.5718]
zsum, zmix, archive = aab.do_mixing(stems, faders, fast.model, aa, "cuda")
loss = aab.mseloss(zsum, zmix) + aab.vicreg_var_loss(zsum) + aab.vicreg_cov_loss(zsum)
print(float(loss), sorted(archive.keys()))
