#!/usr/bin/env python3
"""Benchmark of the audio-algebra hot path on B200 (BASELINE.json metric: audio-seconds encoded per
second; STFT+mel HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512)
on a batch of 256 x [2, 131072] fp32 synthetic 48 kHz stereo chunks PER GPU (weak scaling: chunks are
independent, ranks share nothing, no data-path collective).  One step = one pass of the front-end over
the rank's batch.
  value      whole-job audio-seconds per second, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public API with HOST (pinned) buffers: H2D + kernel + D2H per step
  roofline   the mel kernel against the MEASURED HBM copy bandwidth (algorithmic bytes / kernel time)
  cpu_baseline  the oracle's library-call restatement of the reference (torch.stft + filterbank matmul)
             on this box's host cores, on the same 256-chunk workload
  dp         (every N, top-level copies: encoder_audio_s_per_s, train_step_ms, allreduce_us, pca_ms, dp_parity_ok)
             BASELINE configs[3]/[4]: per-rank bf16 encoder, MixerTrainer.step with the NCCL gradient all-reduce, the
             all-reduce alone, sharded PCA accumulation; cross-rank parity is asserted (mismatch => rc != 0)
`--impl reference` times that CPU path as its own arm (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, N_FFT, HOP, N_MELS = 48000, 2048, 512, 128
CHUNK, BATCH = 131072, 256                      # samples per chunk, chunks per GPU
ALG_BYTES = BATCH * 2 * CHUNK * 4 + BATCH * 2 * N_MELS * (1 + CHUNK // HOP) * 4   # 335 806 464
AUDIO_S_PER_STEP = BATCH * CHUNK / SR           # 699.05 audio-seconds per rank per step
METRIC, UNIT = "audio_seconds_encoded_per_second", "audio-s/s"
WORKLOAD = "configs[1]: MelSpectrogramAE n_fft=2048 hop=512 n_mels=128 on 256 x [2,131072] f32 chunks per GPU"


def synth(batch, seed, device):
    "tonal + noise 48 kHz stereo chunks in [-1, 1] (SURVEY.md section 8d), generated on `device`"
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.arange(CHUNK, device=device, dtype=torch.float32) / SR
    f = 55.0 + (8000.0 - 55.0) * torch.rand(batch, 2, 1, generator=g, device=device)
    ph = 6.283185307179586 * torch.rand(batch, 2, 1, generator=g, device=device)
    x = 0.5 * torch.sin(6.283185307179586 * f * t + ph)
    x += 0.1 * torch.randn(batch, 2, CHUNK, generator=g, device=device)
    return x.clamp_(-1, 1)


def numa_node_count():
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    "DRAM bytes per launch of the mel kernel from the committed ncu capture (profiles/), or None"
    p = os.path.join(ROOT, "profiles", "stft_mel_summary.json")
    try:
        return int(json.load(open(p))["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            probe = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(gpu_index)],
                                   capture_output=True, text=True, timeout=20)
            if probe.returncode != 0 or "not a valid field" in (probe.stdout + probe.stderr).lower():
                self.Q = self.Q.replace("clocks_event_reasons", "clocks_throttle_reasons")
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
            t0 = time.time()   # nvidia-smi needs 0.3 .. 3 s to print its first line (longer with 8 GPUs behind it): wait for it,
            while time.time() - t0 < 8.0 and os.path.getsize(self.f.name) == 0:   # so that the timed region is really sampled
                time.sleep(0.05)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_reference_rate(sample_chunks, reps):
    """The reference's CPU path (oracle restatement of torchaudio MelSpectrogram: torch.stft + |.|^2 + fb matmul)
    on `sample_chunks` chunks of the workload; returns (audio-s/s best of reps, cores, seconds per rep list)."""
    import torch
    from oracle import aa_oracle as O
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    x = synth(sample_chunks, 1234, "cpu")
    O.mel_spectrogram_library_f32(x[:2], SR, N_FFT, HOP, N_MELS)  # warm-up (FFT plans, thread pool)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        O.mel_spectrogram_library_f32(x, SR, N_FFT, HOP, N_MELS)
        ts.append(time.perf_counter() - t0)
    return sample_chunks * CHUNK / SR / min(ts), cores, ts


def measured_bf16_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 2250.0, 2250.0, "nominal dense bf16 (no MEASURED_PEAKS.json)"


def run_extras(dev):
    """Secondary measurements of the same hot path (BASELINE.json configs 2 variants, 3, 5) on one GPU, reported under
    `extras` of the JSON line: STFT power / complex variants against the HBM roofline, the bf16 tcgen05 conv encoder
    against the measured bf16 tensor peak, one data-parallel mixer training step."""
    import torch
    import audio_algebra_b200 as aab
    hbm, _ = measured_peaks()
    burst, sustained, src = measured_bf16_peaks()
    out = {}

    def timed(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    x = synth(BATCH, 4321, dev)
    for name, cls in (("stft_power", aab.MagSpectrogramAE), ("stft_complex", aab.SpectrogramAE)):
        m = cls(n_fft=N_FFT, hop_length=HOP)
        res = {}
        ms = timed(lambda: res.__setitem__("o", m.encode(x)), 10)
        nbytes = x.numel() * 4 + res["o"].numel() * res["o"].element_size()
        out[name] = {"ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / ms / 1e6, "hbm_frac": nbytes / ms / 1e6 / hbm,
                     "audio_s_per_s": AUDIO_S_PER_STEP / (ms * 1e-3)}
        del res
    # the reference's own default geometry (n_fft = 1024, hop = 256, given_models.py:259-264) runs on the warp kernel
    m = aab.MelSpectrogramAE(sample_rate=SR)
    res = {}
    ms = timed(lambda: res.__setitem__("o", m.encode(x)), 10)
    nbytes = x.numel() * 4 + res["o"].numel() * 4
    out["stft_mel_n1024_h256"] = {"ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / ms / 1e6, "hbm_frac": nbytes / ms / 1e6 / hbm,
                                  "audio_s_per_s": AUDIO_S_PER_STEP / (ms * 1e-3), "note": "MelSpectrogramAE() reference defaults, stft_v3_kernel<1024> (two frames per warp item)"}
    del res
    del x
    # ---- conv encoder, bf16 tcgen05 path (config 5 encode sweep point): 68.17 GFLOP per 2^17-sample chunk ----
    dvb = aab.DVAEWrapper(debug=False, compute_dtype="bf16").cuda()
    for B in (64, 1024):               # more points of the config-5 encode sweep (B = 256 per rank is in the `dp` section)
        xe = synth(B, 99, dev)
        ms = timed(lambda: dvb.encode(xe), 5 if B <= 256 else 3)
        tf = B * 68.17 / ms            # GFLOP / ms = TFLOP/s
        out[f"encoder_bf16_B{B}"] = {"ms": ms, "tflops": tf, "frac_of_bf16_burst_peak": tf / burst, "frac_of_bf16_sustained_peak": tf / sustained,
                                     "peak_source": src, "audio_s_per_s": B * CHUNK / SR / (ms * 1e-3), "chunk_samples": CHUNK,
                                     "note": "SoundStreamXL-style encoder restated per SURVEY.md Appendix A (random init); bf16 operands, fp32 accumulate"}
        del xe
    # ---- conv encoder at fp32 accuracy (compute_dtype="fp32", the wrapper default): 3xTF32 tcgen05 kernels ----
    dvf = aab.DVAEWrapper(debug=False, compute_dtype="fp32").cuda()
    xe = synth(64, 98, dev)
    ms = timed(lambda: dvf.encode(xe), 3)
    out["encoder_fp32_B64"] = {"ms": ms, "tflops_fp32_equivalent": 64 * 68.17 / ms, "tf32_mma_tflops": 3 * 64 * 68.17 / ms,
                               "audio_s_per_s": 64 * CHUNK / SR / (ms * 1e-3), "chunk_samples": CHUNK,
                               "note": "fp32 operands split in two TF32 parts, 3 MMAs per K step, accumulator chains of 24 MMAs summed in registers; "
                                       "1.4e-6 relative L2 vs the float64 oracle (the CUDA-core fp32 kernel: 1.4e-6 at 9.8 TFLOP/s)"}
    del xe, dvf
    # ---- bulk encode loop end to end (xae_dataset.ipynb cell 50): host dataset -> pinned staging -> H2D / encode / D2H overlapped ----
    nb_tot, nb = 512, 128
    data_h = torch.empty(nb_tot, 2, CHUNK, dtype=torch.float32, pin_memory=True)
    data_h.copy_(synth(nb_tot, 5, dev))
    reps_h = torch.empty(nb_tot, 64, CHUNK // 128, dtype=torch.float32, pin_memory=True)
    aab.encode_all(dvb, data_h[:nb], batch_size=nb, out=reps_h[:nb])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    aab.encode_all(dvb, data_h, batch_size=nb, out=reps_h)
    dt = time.perf_counter() - t0
    out["encoder_bf16_bulk_e2e"] = {"ms": 1e3 * dt, "chunks": nb_tot, "batch": nb, "audio_s_per_s": nb_tot * CHUNK / SR / dt,
                                    "h2d_bytes": data_h.numel() * 4, "d2h_bytes": reps_h.numel() * 4,
                                    "note": "encode_all(DVAEWrapper bf16, pinned host dataset [512,2,131072]) wall clock incl. H2D and D2H (three streams)"}
    del data_h, reps_h
    Nm = 65536
    g = torch.Generator(device=dev).manual_seed(7)
    # ---- config 4: effects step (4 encodes as one 4B batch + projector + guesses + losses, fwd + bwd) and PCA accumulation ----
    Be = 256
    batch = {k: torch.rand(Be, 2, Nm, generator=g, device=dev) - 0.5 for k in ("a1", "b1", "a2", "b2")}
    aa2 = aab.aa_effects.AudioAlgebra(64, 64).cuda()

    def effects_step():
        for p_ in aa2.parameters():
            p_.grad = None
        arch = aab.aa_effects.do_mixing(batch, dvb.model, aa2, dev)
        aab.aa_effects.effects_losses(arch)["loss"].backward()

    # ---- round-2 tensor-core contractions of the training step at its sizes (B = 512, [64, 512] latents) ----
    zt = torch.randn(512, 64, 512, generator=g, device=dev)
    gt = torch.randn(512, 64, 512, generator=g, device=dev)
    aa3 = aab.AudioAlgebra(64, 64).cuda()
    yv = zt.clone().requires_grad_(True)
    fwd_ms = timed(lambda: aa3.encode(zt), 5)

    def proj_fb():
        yv.grad = None
        for p_ in aa3.parameters():
            p_.grad = None
        (aa3.encode(yv) * gt).sum().backward()

    fb_ms = timed(proj_fb, 5)
    out["projector_half"] = {"fwd_ms": fwd_ms, "fwd_plus_bwd_ms": fb_ms, "tokens": 512 * 512,
                             "note": "one projector half (4 EmbedBlocks) on [512,64,512]: proj_fwd_tc_kernel (3xTF32) / proj_bwd_tc_kernel (bf16x3), "
                                     "fwd+bwd includes the elementwise product and sum of the test loss"}
    zc = zt.clone().requires_grad_(True)

    def cov_fb():
        zc.grad = None
        aab.vicreg_cov_loss(zc).backward()

    out["vicreg_cov"] = {"fwd_ms": timed(lambda: aab.vicreg_cov_loss(zt), 5), "fwd_plus_bwd_ms": timed(cov_fb, 5), "b": 512, "d": 32768,
                         "note": "Gram identity on tcgen05: gram_bb_tc_kernel + cov_bwd_tc_kernel (3-term TF32 split, MN-major Xc operand)"}
    del zt, gt, yv, zc
    # ---- decoder round trip: SpectrogramAE x -> encode -> decode on 64 chunks ----
    xr = synth(64, 77, dev)
    sp = aab.SpectrogramAE(n_fft=N_FFT, hop_length=HOP)
    rr = {}
    rt_ms = timed(lambda: rr.__setitem__("o", sp.decode(sp.encode(xr))), 3, warm=1)
    out["spectrogram_roundtrip"] = {"ms": rt_ms, "chunks": 64, "rel_l2_error": float((rr["o"] - xr).norm() / xr.norm()),
                                    "note": "SpectrogramAE.encode -> decode (aa_istft_f32): the reference's 'perfect reconstruction' decoder"}
    del xr, rr
    ms = timed(effects_step, 3, warm=1)
    out["effects_step"] = {"ms": ms, "batch": Be, "chunk_samples": Nm, "audio_s_per_s": 4 * Be * Nm / SR / (ms * 1e-3),
                           "note": "train_aa_effects step: a1,b1,a2,b2 encoded as one 4B batch, projector enc/dec, effect guesses, 4 loss terms, backward"}
    return out


def run_dp(dev, world, rank, steps=3):
    """BASELINE.json configs[3]/[4] under the driver's clock, at EVERY world size (weak scaling, per-rank shards):
    the bf16 encoder on 256 x 2^17 chunks per rank, MixerTrainer.step on 512 x 2^16 per rank with the real gradient all-reduce
    (train_aa_mixer_accel.py:495-545), the all-reduce alone (133 120 B, the only collective of the step) and the PCA
    accumulation + its 4 097-float all-reduce (calc_effects_pca.py:76-94).  Cross-rank parity is ASSERTED, not reported:
    after two steps the replicated parameters must be bit-identical on all ranks, the all-reduced gradient must equal the
    mean of the per-rank gradients, the reduced PCA numerator must be bit-identical; a mismatch raises (rc != 0)."""
    import torch
    import torch.distributed as dist
    import audio_algebra_b200 as aab
    from audio_algebra_b200.training import MixerTrainer
    from audio_algebra_b200.parallel import allreduce_mean_
    from audio_algebra_b200.pca import RunningCovariance
    out = {}

    def rank_max_ms(fn, reps, warm):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([sorted(ts)[len(ts) // 2]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- encoder sweep point (config 5): bf16 tcgen05 encoder, 256 chunks of 2^17 samples per rank ----
    Be = 256
    torch.manual_seed(0)                                   # the SAME frozen given model on every rank (a checkpoint in real use)
    dvb = aab.DVAEWrapper(debug=False, compute_dtype="bf16").cuda()
    xe = synth(Be, 99 + rank, dev)
    ms = rank_max_ms(lambda: dvb.encode(xe), 5, 2)
    out["encoder_ms"] = ms
    out["encoder_audio_s_per_s"] = world * Be * CHUNK / SR / (ms * 1e-3)
    out["encoder_tflops_per_gpu"] = Be * 68.17 / ms
    del xe
    # ---- data-parallel mixer training step (config 3 / 5): 2 stems of 512 x 2^16 per rank ----
    Bm, Nm = 512, 65536
    torch.manual_seed(2)                                   # same projector on every rank (train_aa_mixer_accel.py:51,479)
    aa = aab.AudioAlgebra(64, 64).cuda()
    tr = MixerTrainer(dvb.model, aa, total_steps=100)
    g = torch.Generator(device=dev).manual_seed(7 + rank)  # a different data shard per rank
    stems = [torch.rand(Bm, 2, Nm, generator=g, device=dev) - 0.5 for _ in range(2)]
    faders = [1.4630, -0.5718]
    ms = rank_max_ms(lambda: tr.step(stems, faders), steps, 1)
    out["train_step_ms"] = ms
    out["train_audio_s_per_s"] = world * Bm * Nm / SR / (ms * 1e-3)
    # parity: two more steps, keeping each rank's pre-all-reduce gradient of the last one
    tr.keep_local_grad = True
    tr.step(stems, faders)
    tr.step(stems, faders)
    torch.cuda.synchronize()
    ok = True
    detail = {}
    if world > 1:
        gp = [torch.empty_like(tr.flat) for _ in range(world)]
        dist.all_gather(gp, tr.flat)
        same_params = all(torch.equal(gp[0], q) for q in gp[1:])
        gl = [torch.empty_like(tr.local_grad) for _ in range(world)]
        dist.all_gather(gl, tr.local_grad)
        mean = torch.stack([q.double() for q in gl]).mean(0)
        rel = float(torch.linalg.vector_norm(tr.flat_grad.double() - mean) / torch.linalg.vector_norm(mean))
        gr = [torch.empty_like(tr.flat_grad) for _ in range(world)]
        dist.all_gather(gr, tr.flat_grad)
        same_grads = all(torch.equal(gr[0], q) for q in gr[1:])
        shards_differ = not torch.equal(gl[0], gl[1])     # the ranks really worked on different data
        detail = {"params_bit_identical": same_params, "reduced_grad_bit_identical": same_grads,
                  "reduced_grad_vs_mean_of_shards_rel_l2": rel, "local_grads_differ": shards_differ}
        ok = same_params and same_grads and rel < 1e-5 and shards_differ
    # ---- the collective alone: mean all-reduce of the flat 33 280-float gradient ----
    n_ar = 50
    buf = tr.flat_grad.clone()
    for _ in range(5):
        allreduce_mean_(buf)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_ar):
        allreduce_mean_(buf)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n_ar], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["allreduce_us"] = 1e3 * float(t.item())
    out["allreduce_bytes"] = buf.numel() * 4
    del stems, tr, buf
    # ---- PCA accumulation (config 4): 256 x [64, 512] latents per rank + the 4 097-float sum all-reduce ----
    gy = torch.Generator(device=dev).manual_seed(31 + rank)
    ys = torch.tanh(torch.randn(256, 64, Nm // 128, generator=gy, device=dev))
    rc = RunningCovariance(64, dev)
    out["pca_ms"] = rank_max_ms(lambda: rc.update(ys), 10, 2)
    out["pca_GBps_per_gpu"] = ys.numel() * 4 / out["pca_ms"] / 1e6
    rc2 = RunningCovariance(64, dev).update(ys)
    local_num = rc2.cov_numerator.clone()
    rc2.all_reduce()
    torch.cuda.synchronize()
    if world > 1:
        gn = [torch.empty_like(local_num) for _ in range(world)]
        dist.all_gather(gn, local_num)
        tot = torch.stack([q.double() for q in gn]).sum(0)
        rel = float(torch.linalg.vector_norm(rc2.cov_numerator.double() - tot) / torch.linalg.vector_norm(tot))
        gr = [torch.empty_like(local_num) for _ in range(world)]
        dist.all_gather(gr, rc2.cov_numerator)
        same = all(torch.equal(gr[0], q) for q in gr[1:])
        cnt_ok = int(rc2.count.item()) == world * 256 * (Nm // 128)
        detail.update({"pca_numerator_bit_identical": same, "pca_sum_rel_l2": rel, "pca_count_ok": cnt_ok})
        ok = ok and same and rel < 1e-5 and cnt_ok
    out["dp_parity_ok"] = bool(ok)
    out["dp_parity"] = detail
    if not ok:
        raise SystemExit(f"bench.py: cross-rank parity FAILED on rank {rank}: {detail}")
    return out


def host_copy_ceiling(dev, world, h2d_bytes, d2h_bytes, reps=3):
    """What the host side can feed: every rank copies the e2e step's bytes pinned host -> device and device -> pinned host
    CONCURRENTLY (two streams, no kernel); returns seconds per step, max over ranks.  e2e cannot beat this."""
    import torch
    import torch.distributed as dist
    hi = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
    ho = torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True)
    di = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    do = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = None
    for i in range(reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            di.copy_(hi, non_blocking=True)
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i > 0:
            best = dt if best is None else min(best, dt)
    t = torch.tensor([best], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = BATCH          # the stated workload: all 256 chunks of configs[1] per step (about 0.3 s of CPU per step)
    for _ in range(max(min(args.warmup, 2), 1)):
        cpu_reference_rate(16, 1)
    import torch
    from oracle import aa_oracle as O
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    x = synth(sample, 1234, "cpu")
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        O.mel_spectrogram_library_f32(x, SR, N_FFT, HOP, N_MELS)
        ts.append(time.perf_counter() - t0)
    mean_t = sum(ts) / len(ts)
    val = sample * CHUNK / SR / mean_t
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * mean_t, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "chunk_samples": CHUNK, "sample": f"all {sample} chunks of the workload per step",
                   "path": "oracle restatement of the reference's CPU code path (torchaudio MelSpectrogram = torch.stft + "
                           "abs^2 + filterbank matmul), all host threads"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} chunks x {args.steps} steps, torch {torch.__version__} CPU, {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import audio_algebra_b200 as aab
    from audio_algebra_b200._lib import lib
    numa_cpus = aab.parallel.bind_to_gpu_numa_node(local) if (world > 1 and not os.environ.get("AA_NO_NUMA_BIND")) else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mel = aab.MelSpectrogramAE(sample_rate=SR, n_fft=N_FFT, hop_length=HOP)
    x = synth(BATCH, 1234 + rank, dev)                       # 268 MB: larger than the 126 MB L2
    sampler = ClockSampler(local) if rank == 0 else None     # returns once nvidia-smi has produced its first sample
    out = None
    for _ in range(max(args.warmup, 3)):
        out = mel.encode(x)
    barrier()

    # ---- device-resident throughput: K steps, one CUDA-event pair per step on the launch stream ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = lib.aa_launch_count()
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_all0.record()
    for a, b in ev:
        a.record()
        out = mel.encode(x)
        b.record()
    e_all1.record()
    barrier()
    launches = lib.aa_launch_count() - l0
    total_ms = e_all0.elapsed_time(e_all1)
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    t_max = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    total_ms_max = float(t_max.item())

    # ---- end to end through the public API with host buffers (H2D + kernel + D2H inside the timed region) ----
    xh = x.cpu().pin_memory()
    e2e_steps = max(2, min(args.steps, 5))
    # preallocated (pinned) result buffer, like the reference's bulk-encode loop, in the layout the reference's own mel tensors
    # have: a [.., n_mels, T] view of a [.., T, n_mels] buffer (torchaudio MelScale returns matmul(spec^T, fb)^T)
    oh = torch.empty(tuple(out.shape[:-2]) + (out.shape[-1], out.shape[-2]), dtype=torch.float32, pin_memory=True).transpose(-1, -2)
    mel.encode(xh, out=oh)                                   # warm-up allocates staging
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mel.encode(xh, out=oh)                               # H2D + kernel + D2H, synchronous on return
    torch.cuda.synchronize()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = world * AUDIO_S_PER_STEP / float(t_e2e.item())
    assert oh.shape == out.shape
    h2d_b, d2h_b = BATCH * 2 * CHUNK * 4, BATCH * 2 * N_MELS * (1 + CHUNK // HOP) * 4
    t_copy = host_copy_ceiling(dev, world, h2d_b, d2h_b)
    copy_ceiling = world * AUDIO_S_PER_STEP / t_copy
    del xh, oh
    dp = None
    if not args.no_dp:
        dp = run_dp(dev, world, rank)                        # raises on a cross-rank parity failure
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = measured_peaks()
        k_ms = sum(kernel_ms) / len(kernel_ms)
        achieved = ALG_BYTES / (k_ms * 1e-3) / 1e9
        cpu_val, cores, ts = cpu_reference_rate(BATCH, 3) if world == 1 else (None, None, None)
        line = {
            "metric": METRIC, "value": world * args.steps * AUDIO_S_PER_STEP / (total_ms_max * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "chunk_samples": CHUNK, "sharding": "batch of chunks, no collective",
                       "l2": "inputs (268 MB per step) exceed the 126 MB L2; no extra flush"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": BATCH * 2 * CHUNK * 4,
                    "d2h_bytes_per_step": BATCH * 2 * N_MELS * (1 + CHUNK // HOP) * 4,
                    "api": "MelSpectrogramAE.encode(pinned CPU tensor, out=pinned) -> aa_stft_mel_tf_f32_host (chunked, copies overlapped)",
                    "host_copy_ceiling": copy_ceiling, "frac_of_host_copy_ceiling": e2e_val / copy_ceiling,
                    "host_copy_GBps_per_rank": (h2d_b + d2h_b) / t_copy / 1e9,
                    "host_copy_note": "every rank copies the step's H2D and D2H bytes concurrently from / to pinned memory, no kernel; max over ranks",
                    "numa_nodes": numa_node_count(), "numa_bound_cpus": (len(numa_cpus) if numa_cpus else None)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "kernel": "stft2048_v3_kernel<MEL>", "kernel_us": 1e3 * k_ms,
                         "algorithmic_bytes": ALG_BYTES, "peak_source": peak_src,
                         # SURVEY.md 8d asks for the FP32 view next to the (judged) HBM view: 131 584 frame-channels x
                         # (56 kFLOP real FFT + 3 k power + 4 k sparse mel) = 8.29 GFLOP against 148 SM x 128 lanes x 2 x 1.965 GHz
                         "fp32_tflops": 8.29e9 * (BATCH / 256) / (k_ms * 1e-3) / 1e12, "fp32_peak_tflops": 74.46,
                         "fp32_frac": 8.29e9 * (BATCH / 256) / (k_ms * 1e-3) / 1e12 / 74.46,
                         "note": "FP32-issue / latency bound (12 warps per SM at 168 registers), not HBM bound: see DESIGN.md 4.1 and profiles/stft_mel_summary.json"},
            "clocks": clocks,
        }
        if dp is not None:
            # top-level copies of the data-parallel figures (whole job, max over ranks) next to the nested record
            for k in ("encoder_audio_s_per_s", "train_step_ms", "allreduce_us", "pca_ms", "dp_parity_ok"):
                line[k] = dp[k]
            line["dp"] = dict(dp, config={"encoder": "bf16 tcgen05 encoder, 256 x [2,131072] per rank",
                                           "train_step": "MixerTrainer.step, 2 stems of 512 x [2,65536] per rank, mean all-reduce of the flat "
                                                         "33 280-float gradient over NCCL (no collective at 1 GPU)",
                                           "pca": "RunningCovariance.update on 256 x [64,512] per rank; sum all-reduce of 4 097 floats"})
        if world == 1 and not args.no_extras:
            try:
                line["extras"] = run_extras(dev)
                # the other front-end variants next to the headline (same kernel family, same workload), as top-level keys
                line["variants"] = {k: {"ms": line["extras"][k]["ms"], "hbm_frac": line["extras"][k]["hbm_frac"]}
                                    for k in ("stft_power", "stft_complex", "stft_mel_n1024_h256") if k in line["extras"]}
            except Exception as e:   # extras never invalidate the headline line
                line["extras"] = {"error": repr(e)}
        if cpu_val is not None:
            line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"all {BATCH} chunks of the workload, best of 3 ({min(ts):.3f} s), oracle restatement of "
                                              f"the reference CPU path, torch {torch.__version__}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary (encoder / variants / training-step) measurements")
    ap.add_argument("--no-dp", action="store_true", help="skip the data-parallel section (encoder, training step + all-reduce, PCA)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)] + (["--no-extras"] if args.no_extras else []) + (["--no-dp"] if args.no_dp else [])
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
