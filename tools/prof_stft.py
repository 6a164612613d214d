"""Launches the mel (default) / power / complex STFT kernel a few times at BASELINE config 2 (for ncu).
env: NFFT (2048), HOP (512), B (256)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab
mode = sys.argv[1] if len(sys.argv) > 1 else "mel"
B = int(os.environ.get("B", 256))
x = torch.rand(B, 2, 131072, device="cuda") - 0.5
cls = {"mel": aab.MelSpectrogramAE, "power": aab.MagSpectrogramAE, "complex": aab.SpectrogramAE}[mode]
m = cls(n_fft=int(os.environ.get("NFFT", 2048)), hop_length=int(os.environ.get("HOP", 512)), **(dict(sample_rate=48000) if mode == "mel" else {}))
for _ in range(4):
    out = m.encode(x)
torch.cuda.synchronize()
print("ok", out.shape)
