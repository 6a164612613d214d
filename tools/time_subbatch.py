"""Encoder time vs sub-batch size (AA_ENC_SUB_SAMPLES is read once per process: one subprocess per point).  Dev tool."""
import json, os, subprocess, sys
CODE = r'''
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
import audio_algebra_b200 as aab
B, N, dt = int(os.environ["B"]), int(os.environ["N"]), os.environ.get("DTYPE", "bf16")
dv = aab.DVAEWrapper(debug=False, compute_dtype=dt).cuda()
x = torch.rand(B, 2, N, device="cuda") - 0.5
for _ in range(2): y = dv.encode(x)
torch.cuda.synchronize()
ts = []
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = dv.encode(x); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); ms = ts[len(ts) // 2]
print(json.dumps({"ms": round(ms, 3), "tflops": round(B * 68.17 * N / 131072 / ms, 1), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}))
'''
for B, N in ((512, 131072), (512, 65536)):
    for sb in (16, 32, 48, 64, 96, 128, 512):
        env = dict(os.environ, B=str(B), N=str(N), AA_ENC_SUB_SAMPLES=str(sb * N))
        r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=120)
        print(B, N, "sub-batch", sb, r.stdout.strip() or r.stderr[-300:], flush=True)
