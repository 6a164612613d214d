"""One bf16 encoder forward (after warm-up) for ncu launch lists / captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab
B = int(os.environ.get("B", 64)); N = int(os.environ.get("N", 131072))
dv = aab.DVAEWrapper(debug=False, compute_dtype=os.environ.get("DTYPE", "bf16")).cuda()
x = torch.rand(B, 2, N, device="cuda") - 0.5
for _ in range(int(os.environ.get("REPS", 2))):
    y = dv.encode(x)
torch.cuda.synchronize()
print("ok", y.shape)
