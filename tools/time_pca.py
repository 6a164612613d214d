import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab
from audio_algebra_b200.pca import RunningCovariance
ys = torch.tanh(torch.randn(256, 64, 512, device="cuda"))
rc = RunningCovariance(64, "cuda")
for _ in range(3): rc.update(ys)
torch.cuda.synchronize(); ts=[]
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc.update(ys); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
print("pca update us", sorted(ts)[5])
