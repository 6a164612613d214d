import torch, time, subprocess
x = torch.empty(256, 2, 131072).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); t = time.perf_counter() - t0
print("H2D pinned 268MB: %.2f ms -> %.1f GB/s" % (t * 1e3, x.numel() * 4 / t / 1e9))
o = torch.empty(256, 2, 128, 257, device="cuda"); oh = torch.empty(o.shape).pin_memory()
for _ in range(2): oh.copy_(o, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); oh.copy_(o, non_blocking=True); torch.cuda.synchronize(); t = time.perf_counter() - t0
print("D2H pinned 67MB: %.2f ms -> %.1f GB/s" % (t * 1e3, o.numel() * 4 / t / 1e9))
t0 = time.perf_counter(); z = torch.empty(256, 2, 128, 257).pin_memory(); t = time.perf_counter() - t0
print("pin alloc 67MB: %.2f ms" % (t * 1e3))
t0 = time.perf_counter(); z = torch.empty((256, 2, 128, 257), pin_memory=True); t = time.perf_counter() - t0
print("pin alloc (cached?) 67MB: %.2f ms" % (t * 1e3))
for q in ["clocks_event_reasons.active,clocks_event_reasons.sw_power_cap", "clocks_throttle_reasons.active,clocks_throttle_reasons.sw_power_cap", "clocks.sm,clocks.max.sm,power.draw"]:
    r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
    print(q, "->", r.returncode, r.stdout.strip()[:100], r.stderr.strip()[:100])
