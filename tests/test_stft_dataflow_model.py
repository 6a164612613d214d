import numpy as np

from stft_dataflow_model import frame_fft_model, frame_pair_fft_model_1024, power_line_slot


def test_dataflow_model_matches_rfft():
    rng = np.random.default_rng(0)
    for _ in range(3):
        x = rng.standard_normal(2048)
        w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(2048) / 2048)
        X = frame_fft_model(x * w)
        ref = np.fft.rfft(x * w)
        assert np.abs(X - ref).max() < 1e-9 * np.abs(ref).max()


def test_dataflow_model_1024_two_frames_per_item():
    rng = np.random.default_rng(1)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1024) / 1024)
    x = rng.standard_normal((2, 1024)) * w
    X = frame_pair_fft_model_1024(x)
    ref = np.fft.rfft(x, axis=-1)
    assert np.abs(X - ref).max() < 1e-9 * np.abs(ref).max()


def test_power_line_slots_are_injective_and_fit_the_buffer():
    for n_fft, frames, stride in ((2048, 1, 0), (1024, 2, 529)):
        nb = n_fft // 2 + 1
        used = set()
        for fr in range(frames):
            for k in range(nb):
                s_ = fr * stride + power_line_slot(k, n_fft)
                assert s_ not in used
                used.add(s_)
        assert max(used) < 8464 // 8            # kV3Xb bytes / sizeof(float2)
        # the walk's lanes: 32 consecutive bins per lane, lanes 33 slots apart -> distinct 8-byte bank pairs within a half warp
        segs = 32 // frames
        for t in range(32):
            for half in range(2):
                banks = {((half * 16 + g) // segs * stride + 33 * ((half * 16 + g) % segs) + t) % 16 for g in range(16)}
                assert len(banks) == 16
