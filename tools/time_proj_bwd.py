"""Projector half backward at the training size: time and accuracy against float64 autograd (dev tool; AA_PROJ_BWD_TC=0 = CUDA-core kernel)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab
def rel_l2(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
B = int(os.environ.get("B", 512)); T = int(os.environ.get("T", 512))
torch.manual_seed(1)
aa = aab.AudioAlgebra(64, 64).cuda()
y = (torch.randn(B, 64, T, device="cuda") * 0.7).requires_grad_(True)
gz = torch.randn(B, 64, T, device="cuda")
def run():
    aa.zero_grad(); y.grad = None
    z = aa.encode(y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); (z * gz).sum().backward(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for _ in range(3): run()
ms = sorted(run() for _ in range(7))[3]
# float64 reference (plain torch)
ws = [(getattr(aa.encoder[i].lin, "weight").detach().double().requires_grad_(True), aa.encoder[i].lin.bias.detach().double().requires_grad_(True)) for i in range(4)]
yd = y.detach().double().requires_grad_(True)
h = yd.transpose(1, 2)
for i, (w, b) in enumerate(ws):
    u = h @ w.T + b
    h = h + (torch.nn.functional.gelu(u) if i < 3 else u)
zd = yd + h.transpose(1, 2)
(zd * gz.double()).sum().backward()
errs = {"gx": rel_l2(y.grad, yd.grad)}
for i, (w, b) in enumerate(ws):
    errs[f"gw{i}"] = rel_l2(aa.encoder[i].lin.weight.grad, w.grad); errs[f"gb{i}"] = rel_l2(aa.encoder[i].lin.bias.grad, b.grad)
print(json.dumps({"AA_PROJ_BWD_TC": os.environ.get("AA_PROJ_BWD_TC", "1"), "B": B, "T": T, "bwd_ms_incl_mul_sum": ms, "max_rel_err": max(errs.values()), "errs": errs}))
