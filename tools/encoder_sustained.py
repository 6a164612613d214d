"""Is the B >= 256 encoder rate (705-760 TFLOP/s) lower than the B = 64 rate (800) because of batch size, or because the GPU
cannot hold the burst clock for more than a few ms?  Times 40 back-to-back B = 64 passes (one CUDA-event pair each) and
samples the SM clock: if the per-pass time climbs to the large-batch per-chunk time, it is the sustained-power clock."""
import sys, os, json, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_algebra_b200 as aab
dv = aab.DVAEWrapper(debug=False, compute_dtype="bf16").cuda()
x = torch.rand(64, 2, 131072, device="cuda") - 0.5
for _ in range(2):
    dv.encode(x)
torch.cuda.synchronize()
time.sleep(2.0)   # let the clocks / power state relax
clk = []
stop = False
def sample():
    while not stop:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
        clk.append((time.time(), r.stdout.strip()))
th = threading.Thread(target=sample); th.start()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(40)]
t0 = time.time()
for a, b in ev:
    a.record(); dv.encode(x); b.record()
torch.cuda.synchronize()
t1 = time.time()
stop = True; th.join()
ms = [a.elapsed_time(b) for a, b in ev]
print(json.dumps({"per_pass_ms": [round(v, 3) for v in ms], "tflops_first3": [round(64 * 68.17 / v, 1) for v in ms[:3]],
                  "tflops_last3": [round(64 * 68.17 / v, 1) for v in ms[-3:]], "wall_ms": round(1e3 * (t1 - t0), 1),
                  "clock_samples_during": [s for t, s in clk if t0 <= t <= t1][:12]}))
