"""Device-side timing of the STFT kernels at the reference's default n_fft=1024 / hop=256 (generic path), dev tool."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab
B, n = int(os.environ.get("B", 256)), 131072
x = torch.rand(B, 2, n, device="cuda") - 0.5
res = {}
for name, cls in [("mel", aab.MelSpectrogramAE), ("power", aab.MagSpectrogramAE), ("complex", aab.SpectrogramAE)]:
    kw = dict(sample_rate=48000) if name == "mel" else {}
    m = cls(n_fft=int(os.environ.get("NFFT", 1024)), hop_length=int(os.environ.get("HOP", 256)), **kw)
    for _ in range(2):
        out = m.encode(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = m.encode(x); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    res[name] = dict(us_med=ts[len(ts) // 2], audio_s_per_s=B * n / 48000 / (ts[len(ts) // 2] * 1e-6))
    del out
print(json.dumps(res, indent=1))
