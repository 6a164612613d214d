"""One mixer training step (B x 2^16, 2 stems, bf16 encoder) after a warm-up step, for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab
from audio_algebra_b200.training import MixerTrainer
B = int(os.environ.get("B", 512)); N = int(os.environ.get("N", 65536))
dv = aab.DVAEWrapper(debug=False, compute_dtype="bf16").cuda()
torch.manual_seed(2)
aa = aab.AudioAlgebra(64, 64).cuda()
tr = MixerTrainer(dv.model, aa, total_steps=100)
stems = [torch.rand(B, 2, N, device="cuda") - 0.5 for _ in range(2)]
for _ in range(int(os.environ.get("REPS", 2))):
    out = tr.step(stems, [1.4630, -0.5718])
torch.cuda.synchronize()
print("ok", float(out["loss"]))
