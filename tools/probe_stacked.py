"""dev probe: does the SoundStreamXL encoder handle accept the StackedDiffAE first-stage geometry (capacity 64, strides [2]*5, latent 32)?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab
from oracle import aa_oracle as O
torch.manual_seed(0)
kw = dict(in_channels=2, capacity=64, latent_dim=32, c_mults=[2, 4, 8, 16, 32], strides=[2, 2, 2, 2, 2])
enc_o = O.SoundStreamXLEncoderOracle(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in kw.items()}).eval()
x = torch.rand(1, 2, 8192) - 0.5
with torch.no_grad():
    ref = enc_o(x)
print("ref", ref.shape)
for dt in ("fp32_cuda_cores", "fp32", "bf16"):
    try:
        enc = aab.SoundStreamXLEncoder(**kw, compute_dtype=dt)
        enc.load_oracle_weights(enc_o)
        enc = enc.cuda()
        y = enc(x.cuda())
        torch.cuda.synchronize()
        rel = float(torch.linalg.vector_norm(y.cpu().double() - ref.double()) / torch.linalg.vector_norm(ref.double()))
        xb = torch.rand(4, 2, 262144, device="cuda") - 0.5
        enc(xb); torch.cuda.synchronize()
        t0 = time.perf_counter(); enc(xb); torch.cuda.synchronize(); dt_s = time.perf_counter() - t0
        cos = torch.nn.functional.cosine_similarity(y.cpu().double().flatten(1), ref.double().flatten(1), dim=1).min().item()
        print(dt, tuple(y.shape), "rel", rel, "cos", cos, "4 x 262144:", round(dt_s * 1e3, 2), "ms")
    except Exception as e:
        print(dt, "FAILED:", repr(e)[:300])
