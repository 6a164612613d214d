"""Accuracy of the three encoder arithmetic modes against the float64 oracle (dev tool; tests/ hold the asserted bounds)."""
import sys, os, json, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab
from oracle import aa_oracle as O

torch.manual_seed(0)
enc_o = O.SoundStreamXLEncoderOracle().eval()
enc64 = copy.deepcopy(enc_o).double()
g = torch.Generator().manual_seed(21)
x = torch.rand(2, 2, 16384, generator=g) - 0.5
with torch.no_grad():
    ref = torch.tanh(enc64(x.double()))
    ref32 = O.dvae_encode_it(enc_o, x)
rel = lambda a, b: ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()
out = {"oracle_fp32_cpu": rel(ref32, ref)}
for mode in ("fp32_cuda_cores", "tf32x3", "bf16"):
    dv = aab.DVAEWrapper(debug=False, compute_dtype=mode)
    dv.model.load_oracle_weights(enc_o)
    dv = dv.cuda()
    out[mode] = rel(dv.encode(x.cuda()), ref)
print(json.dumps(out, indent=1))
