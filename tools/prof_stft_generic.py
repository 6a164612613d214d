"""Launches the n_fft=1024 / hop=256 STFT (warp kernel) a few times at batch 256 x [2,131072] (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab
mode = sys.argv[1] if len(sys.argv) > 1 else "mel"
x = torch.rand(int(os.environ.get("B", 256)), 2, 131072, device="cuda") - 0.5
cls = {"mel": aab.MelSpectrogramAE, "power": aab.MagSpectrogramAE, "complex": aab.SpectrogramAE}[mode]
m = cls(n_fft=1024, hop_length=256, **(dict(sample_rate=48000) if mode == "mel" else {}))
for _ in range(3):
    out = m.encode(x)
torch.cuda.synchronize()
print("ok", out.shape)
