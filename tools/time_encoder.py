"""Device-side timing of the conv encoder (dev tool): fp32 CUDA-core path vs bf16 tcgen05 path."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab

N = int(os.environ.get("N", 131072))
GF_PER_CHUNK = 68.17 * N / 131072
res = {}
for dtype, batches in (("bf16", [int(b) for b in os.environ.get("BS", "8,64,256").split(",")]), ("tf32x3", [int(b) for b in os.environ.get("BS_TF", "8,64").split(",")]), ("fp32_cuda_cores", [2])):
    dv = aab.DVAEWrapper(debug=False, compute_dtype=dtype).cuda()
    for B in batches:
        x = torch.rand(B, 2, N, device="cuda") - 0.5
        for _ in range(2):
            y = dv.encode(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); y = dv.encode(x); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        res[f"{dtype}_B{B}"] = dict(ms=ms, tflops=B * GF_PER_CHUNK / ms, audio_s_per_s=B * N / 48000 / (ms * 1e-3))
        del x, y
print(json.dumps(res, indent=1))
