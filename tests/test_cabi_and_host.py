"""CPU-only checks: the C-ABI library loads and exports every function include/aa_b200.h declares (no
compute call is made), the ctypes table covers the header, host-side helpers match the reference's
fixtures, and the product refuses to run without a GPU instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "aa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aa_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 25
    lib = ctypes.CDLL(os.path.join(ROOT, "audio-algebra_b200", "libaa_b200.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in aa_b200.h but not exported: {missing}"
    lib.aa_version.restype = ctypes.c_int
    assert lib.aa_version() >= 100
    lib.aa_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.aa_last_error(), bytes)


def test_ctypes_table_covers_the_header():
    import audio_algebra_b200 as aab
    bound = set(aab._lib._SIGS)
    assert set(_header_functions()) <= bound, sorted(set(_header_functions()) - bound)


def test_no_cpu_fallback():
    import audio_algebra_b200 as aab
    from audio_algebra_b200._lib import AaError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(AaError):
        aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512).encode(torch.zeros(1, 2, 8192))
    with pytest.raises(AaError):
        aab.mseloss(torch.zeros(4), torch.zeros(4))
    with pytest.raises(AaError):
        aab.AudioAlgebra(64, 64)(torch.zeros(1, 64, 8))


def test_melscale_fbanks_matches_torchaudio_fixture(golden):
    from audio_algebra_b200.given_models import melscale_fbanks
    g = golden("stft")
    assert np.array_equal(melscale_fbanks(1025, 0.0, 24000.0, 128, 48000).numpy(), g["fb_2048_48k_128"])
    assert np.array_equal(melscale_fbanks(513, 0.0, 24000.0, 128, 48000).numpy(), g["fb_1024_48k_128"])


def test_constructor_contracts():
    import audio_algebra_b200 as aab
    m = aab.MelSpectrogramAE()
    assert (m.n_fft, m.hop_length, m.n_mels, m.sample_rate) == (1024, 256, 128, 48000)  # given_models.py:259-264
    assert aab.SpectrogramAE().n_fft == 1024 and aab.MagSpectrogramAE().hop_length == 256
    assert m.next_power_of_2(55728) == 65536 and m.next_power_of_2(0) == 1
    assert tuple(m.zero_pad_po2(torch.ones(2, 5)).shape) == (2, 8)
    with pytest.raises(NotImplementedError):
        aab.MelSpectrogramAE(normalized=True)
    with pytest.raises(TypeError):
        aab.MelSpectrogramAE(bogus_kwarg=1)
    aa = aab.AudioAlgebra(dims=64, hidden_dims=64)
    assert sum(p.numel() for p in aa.parameters()) == 33280   # 33 280 parameters
    bn = aab.AudioAlgebra(dims=64, hidden_dims=64, use_bn=True)   # constructible, same extra keys as the reference's BatchNorm1d blocks
    assert "encoder.0.bn.running_mean" in bn.state_dict() and isinstance(bn.encoder[0].bn, torch.nn.BatchNorm1d)
    with pytest.raises(NotImplementedError):
        aab.EmbedBlock(4, 4, act=torch.nn.ReLU())


def test_onecycle_schedule_matches_torch():
    from audio_algebra_b200.training import onecycle_lr, onecycle_beta1
    p = torch.zeros(3, requires_grad=True)
    opt = torch.optim.Adam([p], lr=5e-4)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-3, total_steps=200)
    for step in range(199):
        assert abs(opt.param_groups[0]["lr"] - onecycle_lr(step, 200)) < 1e-12
        assert abs(opt.param_groups[0]["betas"][0] - onecycle_beta1(step, 200)) < 1e-12
        opt.step(); sched.step()


def test_shard_range_partitions_everything():
    from audio_algebra_b200.parallel import shard_range
    for n in (0, 1, 7, 256, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_bench_reference_arm_json_contract():
    "bench.py --impl reference (the CPU arm the driver runs first) prints ONE JSON line with the contract's keys"
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "audio_seconds_encoded_per_second" and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
