"""BASELINE.json config 5: bf16 encode sweep, batch 64 .. 4096 chunks of [2, 2^17] per GPU (writes a JSON summary)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab

N = 131072
dv = aab.DVAEWrapper(debug=False, compute_dtype="bf16").cuda()
res = []
for B in [int(b) for b in os.environ.get("BS", "64,128,256,512,1024,2048,4096").split(",")]:
    x = torch.rand(B, 2, N, device="cuda") - 0.5
    for _ in range(2):
        y = dv.encode(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = dv.encode(x); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[1]
    res.append(dict(batch=B, ms=ms, tflops=B * 68.17 / ms, audio_s_per_s=B * N / 48000 / (ms * 1e-3),
                    peak_mem_GB=torch.cuda.max_memory_allocated() / 1e9))
    del x, y
    torch.cuda.empty_cache()
print(json.dumps({"workload": "DVAEWrapper(bf16).encode on [B,2,131072], 1 x B200", "points": res}, indent=1))
