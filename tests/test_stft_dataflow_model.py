import numpy as np

from stft_dataflow_model import frame_fft_model


def test_dataflow_model_matches_rfft():
    rng = np.random.default_rng(0)
    for _ in range(3):
        x = rng.standard_normal(2048)
        w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(2048) / 2048)
        X = frame_fft_model(x * w)
        ref = np.fft.rfft(x * w)
        assert np.abs(X - ref).max() < 1e-9 * np.abs(ref).max()
