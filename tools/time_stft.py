"""Quick device-side timing of the STFT kernels at BASELINE config 2 (not the bench; a dev tool)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab

B, n = int(os.environ.get("B", 256)), 131072
x = torch.rand(B, 2, n, device="cuda") - 0.5
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for name, cls in [("mel", aab.MelSpectrogramAE), ("power", aab.MagSpectrogramAE), ("complex", aab.SpectrogramAE)]:
    if name not in os.environ.get("MODES", "mel,power,complex").split(","):
        continue
    kw = dict(sample_rate=48000) if name == "mel" else {}
    m = cls(n_fft=int(os.environ.get("NFFT", 2048)), hop_length=int(os.environ.get("HOP", 512)), center=os.environ.get("CENTER", "1") == "1", **kw)
    ekw = dict(freq_major=True) if (name == "mel" and os.environ.get("FREQ_MAJOR")) else {}
    for _ in range(3):
        out = m.encode(x, **ekw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = m.encode(x, **ekw); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    nbytes = x.numel() * 4 + out.numel() * out.element_size()
    res[name] = dict(us_min=ts[0], us_med=ts[len(ts) // 2], GBps_med=nbytes / ts[len(ts) // 2] / 1e3,
                     audio_s_per_s=B * n / 48000 / (ts[len(ts) // 2] * 1e-6))
    del out
print(json.dumps(res, indent=1))
