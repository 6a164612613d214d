"""VICReg covariance loss forward / backward time at the training size (dev tool; AA_COV_TC=0 selects the CUDA-core kernels)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab
B = int(os.environ.get("B", 512)); C = 64; T = 512
z = torch.randn(B, C, T, device="cuda")
def run():
    zc = z.clone().requires_grad_(True)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); l = aab.vicreg_cov_loss(zc); e[1].record(); l.backward(); e[2].record(); torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), float(l)
for _ in range(3): run()
r = sorted(run() for _ in range(9))[4]
print(json.dumps({"AA_COV_TC": os.environ.get("AA_COV_TC", "1"), "B": B, "fwd_ms": r[0], "bwd_ms": r[1], "loss": r[2]}))
