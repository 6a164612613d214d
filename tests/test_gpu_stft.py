"""GPU parity of the CUDA STFT / power / mel front-end (through the C ABI, via the reference-named
Python classes) against (1) fixtures produced by the reference's own code, (2) the float64 oracle on
seeded inputs, (3) size-independent properties at the full BASELINE config-2 size.
Tolerance (BASELINE.json north_star): relative L2 <= 1e-4 on spectrograms, fp32."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-4  # BASELINE.json: fp32 rel-L2 on spectrograms
T = torch.from_numpy


@pytest.fixture(scope="module")
def aab():
    import audio_algebra_b200 as aab
    return aab


def _oracle():
    from oracle import aa_oracle as O
    return O


@pytest.mark.parametrize("tag,n_fft,hop", [("2048_512", 2048, 512), ("1024_256", 1024, 256)])
def test_golden_small(aab, golden, tag, n_fft, hop):
    g = golden("stft")
    x = T(g["x_small"]).cuda()
    c = aab.SpectrogramAE(n_fft=n_fft, hop_length=hop).encode(x)
    assert c.dtype == torch.complex64 and tuple(c.shape) == g[f"complex_{tag}"].shape
    assert rel_l2(c, g[f"complex_{tag}"]) < TOL
    p = aab.MagSpectrogramAE(n_fft=n_fft, hop_length=hop).encode(x)
    assert rel_l2(p, g[f"power_{tag}"]) < TOL
    m = aab.MelSpectrogramAE(sample_rate=48000, n_fft=n_fft, hop_length=hop).encode(x)
    assert tuple(m.shape) == g[f"mel_{tag}"].shape
    assert rel_l2(m, g[f"mel_{tag}"]) < TOL


def test_golden_non_pow2_length_defaults(aab, golden):
    "given-models.ipynb cell 14 KAT: SpectrogramAE() on [2,55728] -> [2,513,257] complex64"
    g = golden("stft")
    x = T(g["x_np2"]).cuda()
    s = aab.SpectrogramAE().encode(x)
    assert tuple(s.shape) == (2, 513, 257) and s.dtype == torch.complex64
    assert rel_l2(s[:, ::4, ::4], g["complex_np2_strided"]) < TOL
    assert rel_l2(aab.MagSpectrogramAE().encode(x)[:, ::4, ::4], g["power_np2_strided"]) < TOL
    assert rel_l2(aab.MelSpectrogramAE().encode(x), g["mel_np2"]) < TOL
    # same signal through the fast n_fft=2048 path (edge tiles see the zero-padded tail)
    O = _oracle()
    for cls, fn in [(aab.MelSpectrogramAE, lambda v: O.mel_spectrogram(v, 48000, 2048, 512)),
                    (aab.MagSpectrogramAE, lambda v: O.power_spectrogram(v, 2048, 512)),
                    (aab.SpectrogramAE, lambda v: O.stft_complex(v, 2048, 512))]:
        out = cls(n_fft=2048, hop_length=512).encode(x)
        assert rel_l2(out, fn(x.cpu())) < TOL


def test_golden_odd_rows_unaligned_length(aab, golden):
    g = golden("stft")
    x = T(g["x_odd"]).cuda()  # [3,1,5001]: odd row count, n_in % 4 != 0
    assert rel_l2(aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512).encode(x), g["mel_odd_2048_512"]) < TOL
    assert rel_l2(aab.MagSpectrogramAE().encode(x)[:, :, ::4, :], g["power_odd_1024_256"]) < TOL
    O = _oracle()
    assert rel_l2(aab.SpectrogramAE(n_fft=2048, hop_length=512).encode(x), O.stft_complex(x.cpu(), 2048, 512)) < TOL


def test_golden_headline_chunk(aab, golden):
    from oracle.make_golden import synth
    g = golden("stft")
    x = synth((1, 2, 131072), int(g["x_big_seed"][0])).cuda()
    m = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512).encode(x)
    assert tuple(m.shape) == (1, 2, 128, 257)
    assert rel_l2(m, g["mel_big_2048_512"]) < TOL


def test_magdphase(aab, golden):
    g = golden("stft")
    out = aab.MagDPhaseSpectrogramAE().encode(T(g["x_mdp"]).cuda()).cpu().double()
    ref = T(g["magdphase"]).double()
    c = ref.shape[0] // 2
    assert tuple(out.shape) == tuple(ref.shape)
    assert rel_l2(out[:c], ref[:c]) < TOL
    d = (out[c:] - ref[c:] + np.pi) % (2 * np.pi) - np.pi
    strong = ref[:c] > 1e-2 * ref[:c].max()
    assert d[strong].abs().max() < 2e-3


@pytest.mark.parametrize("key,kw", [("magdphase_use_cos", dict(use_cos=True)), ("magdphase_debug", dict(debug=True))])
def test_magdphase_use_cos_and_debug_branches(aab, golden, key, kw):
    g = golden("stft")
    out = aab.MagDPhaseSpectrogramAE(**kw).encode(T(g["x_mdp"]).cuda()).cpu().double()
    ref = T(g[key]).double()
    c = ref.shape[0] // 2
    assert tuple(out.shape) == tuple(ref.shape) and rel_l2(out[:c], ref[:c]) < TOL
    strong = ref[:c] > 1e-2 * ref[:c].max()
    if "use_cos" in kw:   # acos is ill-conditioned near +-1: compare the cosines on bins that carry energy; column 0 holds theta
        assert (torch.cos(out[c:])[..., 1:][strong[..., 1:]] - torch.cos(ref[c:])[..., 1:][strong[..., 1:]]).abs().max() < 2e-3
        d0 = (out[c:, :, 0] - ref[c:, :, 0] + np.pi) % (2 * np.pi) - np.pi
        assert d0[strong[..., 0]].abs().max() < 2e-3
    else:
        d = (out[c:] - ref[c:] + np.pi) % (2 * np.pi) - np.pi
        assert d[strong].abs().max() < 2e-3
        assert out[c:].min() >= 0          # debug: phases wrapped into [0, 2 pi) before differencing, differences wrapped too


@pytest.mark.parametrize("n_fft,hop,n", [(2048, 512, 16384), (2048, 256, 8192), (2048, 1024, 8192), (2048, 500, 9000),
                                         (512, 128, 4096), (4096, 1024, 16384), (256, 64, 1000), (2048, 2048, 8192)])
def test_oracle_sweep(aab, n_fft, hop, n):
    "fast path (n_fft=2048, hop%4==0, hop<=1024) and generic path against the float64 oracle"
    O = _oracle()
    g = torch.Generator().manual_seed(n_fft + hop)
    x = (torch.rand(3, 2, n, generator=g) * 2 - 1)
    xc = x.cuda()
    assert rel_l2(aab.SpectrogramAE(n_fft=n_fft, hop_length=hop).encode(xc), O.stft_complex(x, n_fft, hop)) < TOL
    assert rel_l2(aab.MagSpectrogramAE(n_fft=n_fft, hop_length=hop).encode(xc), O.power_spectrogram(x, n_fft, hop)) < TOL
    assert rel_l2(aab.MelSpectrogramAE(sample_rate=48000, n_fft=n_fft, hop_length=hop, n_mels=64).encode(xc),
                  O.mel_spectrogram(x, 48000, n_fft, hop, n_mels=64)) < TOL


def test_custom_window_uses_the_table_path(aab):
    "a non-Hann window_fn (torchaudio kwarg) must not go through the Hann-synthesising fast kernel"
    O = _oracle()
    g = torch.Generator().manual_seed(9)
    x = torch.rand(2, 2, 8192, generator=g) - 0.5
    w = torch.hamming_window(2048)
    out = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512, window_fn=torch.hamming_window).encode(x.cuda())
    assert rel_l2(out, O.mel_spectrogram(x, 48000, 2048, 512, window=w.double())) < TOL


def test_center_false_and_no_zero_pad(aab):
    O = _oracle()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 2, 6000, generator=g) - 0.5
    m = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512, center=False)
    m.zero_pad = False
    ref = O.mel_spectrogram(x, 48000, 2048, 512, center=False, zero_pad=False)
    out = m.encode(x.cuda())
    assert tuple(out.shape) == tuple(ref.shape)
    assert rel_l2(out, ref) < TOL


def test_cpu_input_goes_through_the_device_and_comes_back(aab, golden):
    g = golden("stft")
    x = T(g["x_small"])  # CPU tensor: H2D + kernel + D2H (the e2e path of bench.py)
    m = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512).encode(x)
    assert m.device.type == "cpu" and rel_l2(m, g["mel_2048_512"]) < TOL
    s = aab.SpectrogramAE(n_fft=2048, hop_length=512).encode(x)
    assert s.device.type == "cpu" and rel_l2(s, g["complex_2048_512"]) < TOL
    big = torch.rand(150, 2, 8192) - 0.5   # several chunks of the pipelined host path, odd chunk tail
    O = _oracle()
    assert rel_l2(aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512).encode(big),
                  O.mel_spectrogram(big, 48000, 2048, 512)) < TOL


def test_empty_batch_and_errors(aab):
    m = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512)
    out = m.encode(torch.zeros(0, 2, 8192, device="cuda"))
    assert tuple(out.shape) == (0, 2, 128, 17)
    from audio_algebra_b200._lib import AaError
    with pytest.raises(AaError):
        m.encode(torch.zeros(1, 2, 512, device="cuda"))  # reflect pad needs > n_fft/2 samples
    with pytest.raises(AaError):
        aab.SpectrogramAE(n_fft=1000)  # not a power of two
    with pytest.raises(ValueError):
        m.decode(torch.zeros(1))   # needs [..., 128, T]
    with pytest.raises(NotImplementedError):
        aab.DVAEWrapper(debug=False).decode(torch.zeros(1, 64, 8))   # the diffusion sampler is out of scope


def test_full_size_properties(aab):
    """BASELINE config 2 geometry: [256,2,131072] -> [256,2,128,257].  The oracle is checked on a
    row subset; the whole tensor through size-independent properties: hop-shift equivariance of
    interior frames, Parseval (sum of one-sided power = n_fft * sum (x w)^2) and scaling."""
    O = _oracle()
    B, n = 256, 131072
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = (torch.rand(B, 2, n, device="cuda", generator=g) - 0.5)
    mel = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512)
    m = mel.encode(x)
    assert tuple(m.shape) == (B, 2, 128, 257) and torch.isfinite(m).all()
    for b in (0, 97, 255):
        assert rel_l2(m[b], O.mel_spectrogram(x[b].cpu(), 48000, 2048, 512)) < TOL
    # shift by one hop: frame t of the shifted signal == frame t+1 of the original (interior frames)
    xs = torch.roll(x, shifts=-512, dims=-1)
    ms = mel.encode(xs)
    assert rel_l2(ms[..., 2:250], m[..., 3:251]) < 1e-5
    # mel is quadratic in the signal
    assert rel_l2(mel.encode(2.0 * x), 4.0 * m) < 1e-6
    # Parseval on the power spectrogram of a few rows (full rows would need 1 GB)
    p = aab.MagSpectrogramAE(n_fft=2048, hop_length=512).encode(x[:8])
    w = torch.hann_window(2048, periodic=True, device="cuda")
    fr = torch.nn.functional.pad(x[:8].reshape(16, 1, n), (1024, 1024), mode="reflect").reshape(8, 2, -1).unfold(-1, 2048, 512)
    energy = ((fr * w) ** 2).sum(-1) * 2048            # [8,2,257]
    onesided = 2 * p.sum(dim=-2) - p[..., 0, :] - p[..., -1, :]
    assert rel_l2(onesided, energy) < 1e-5


def test_mel_filterbank_variants(aab):
    """The fused n_fft=2048 mel epilogue (per-filter weight tables built from the caller's filterbank) against the float64
    product power @ fb for HTK / Slaney scales and norms and other filter counts, on inputs with a large dynamic range
    across bins, filter by filter (a global norm would hide a wrong quiet filter next to the 440 Hz peak)."""
    O = _oracle()
    g = torch.Generator().manual_seed(77)
    t = torch.arange(16384) / 48000.0
    x = 0.8 * torch.sin(2 * np.pi * 440.0 * t)[None, None] + 1e-3 * (torch.rand(3, 2, 16384, generator=g) - 0.5)
    p = O.power_spectrogram(x, 2048, 512)                                         # float64 [3,2,1025,T]
    for kw in (dict(), dict(norm="slaney"), dict(mel_scale="slaney"), dict(n_mels=80, f_min=30.0, f_max=16000.0),
               dict(n_mels=31), dict(n_mels=200)):
        m = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512, **kw)
        ref = torch.einsum("bcft,fm->bcmt", p, m.fb.double())
        got = m.encode(x.cuda()).cpu().double()
        assert rel_l2(got, ref) < TOL, kw
        err = (got - ref).abs().amax(dim=(0, 1, 3)) / ref.abs().amax(dim=(0, 1, 3)).clamp_min(1e-30)
        assert err.max().item() < 1e-4, (kw, err.argmax().item(), err.max().item())
    # odd row count, non-power-of-two length (zero_pad_po2 tail + reflect edges)
    x2 = torch.rand(3, 1, 20000, generator=g) - 0.5
    m = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512)
    assert rel_l2(m.encode(x2.cuda()), O.mel_spectrogram(x2, 48000, 2048, 512)) < TOL


@pytest.mark.parametrize("n_fft,hop,shape", [(64, 16, (2, 2, 1000)), (128, 32, (3, 1, 777)), (256, 100, (1, 2, 5000)), (512, 128, (5, 1, 4096)),
                                             (1024, 256, (3, 2, 20000)), (1024, 200, (1, 3, 9001)), (1024, 1024, (2, 2, 8192))])
def test_warp_kernel_small_n_fft(aab, n_fft, hop, shape):
    """n_fft <= 1024 (incl. the reference defaults 1024 / 256, given_models.py:151-160) runs on the warp kernel (register FFT x
    shuffle FFT across lanes, row pairs packed): every R = n_fft / 64, odd row counts, odd lengths (unaligned rows, zero_pad_po2
    tails, reflect edges), odd hops, a non-Hann window and center=False, against the float64 oracle."""
    O = _oracle()
    g = torch.Generator().manual_seed(n_fft * 7 + hop)
    x = torch.rand(*shape, generator=g) * 2 - 1
    xc = x.cuda()
    assert rel_l2(aab.SpectrogramAE(n_fft=n_fft, hop_length=hop).encode(xc), O.stft_complex(x, n_fft, hop)) < TOL
    assert rel_l2(aab.MagSpectrogramAE(n_fft=n_fft, hop_length=hop).encode(xc), O.power_spectrogram(x, n_fft, hop)) < TOL
    if n_fft >= 256:
        assert rel_l2(aab.MelSpectrogramAE(sample_rate=48000, n_fft=n_fft, hop_length=hop, n_mels=40).encode(xc),
                      O.mel_spectrogram(x, 48000, n_fft, hop, n_mels=40)) < TOL
    w = torch.hamming_window(n_fft)
    out = aab.MagSpectrogramAE(n_fft=n_fft, hop_length=hop, window_fn=torch.hamming_window).encode(xc)
    assert rel_l2(out, O.power_spectrogram(x, n_fft, hop, window=w.double())) < TOL
    m = aab.SpectrogramAE(n_fft=n_fft, hop_length=hop, center=False)
    m.zero_pad = False
    ref = O.stft_complex(x, n_fft, hop, center=False, zero_pad=False)
    out = m.encode(xc)
    assert tuple(out.shape) == tuple(ref.shape) and rel_l2(out, ref) < TOL


@pytest.mark.parametrize("n_fft,hop,shape", [(2048, 512, (3, 2, 16384)), (1024, 256, (2, 3, 5000)), (256, 64, (1, 1, 4096)),
                                             (4096, 1024, (2, 2, 16384)), (2048, 500, (1, 3, 9000))])
def test_output_layouts_agree_and_match_the_reference_strides(aab, n_fft, hop, shape):
    """Complex / power spectrograms are written frequency-minor (aa_stft_*_tf_f32) and returned as transposed views: the strides
    torch.stft / torchaudio give the reference's tensors.  The contiguous [.., F, T] variant (freq_major=True, used by the
    MagDPhase epilogue) must hold the same values: tile kernel, warp kernel and generic kernel, odd row counts included."""
    import torchaudio
    g = torch.Generator().manual_seed(n_fft + hop)
    x = (torch.rand(*shape, generator=g) - 0.5).cuda()
    for cls, mode, power in ((aab.SpectrogramAE, "complex", None), (aab.MagSpectrogramAE, "power", 2)):
        m = cls(n_fft=n_fft, hop_length=hop)
        y = m.encode(x)
        yc = m._run(x, mode, freq_major=True)
        assert yc.is_contiguous() and tuple(y.shape) == tuple(yc.shape)
        # two kernels since round 2 (registers -> global for the frequency-minor layout, staged tiles for [F, T]): same values up to fp32 rounding
        assert rel_l2(torch.view_as_real(y) if power is None else y, torch.view_as_real(yc) if power is None else yc) < 2e-6
        m.zero_pad = False
        ref = torchaudio.transforms.Spectrogram(n_fft=n_fft, hop_length=hop, power=power)(x.cpu())
        y2 = m.encode(x)
        assert tuple(y2.shape) == tuple(ref.shape) and y2.stride()[-2:] == ref.stride()[-2:] == (1, n_fft // 2 + 1)
        assert rel_l2(torch.view_as_real(y2) if power is None else y2, torch.view_as_real(ref) if power is None else ref) < 1e-4


@pytest.mark.parametrize("hop,n,rows", [(512, 131072, (2, 2)), (512, 20000, (3, 1)), (256, 16384, (1, 2)), (1024, 40000, (2, 2)), (500, 9000, (1, 2))])
def test_v2_kernel_banded_mel_and_v1_kernels_agree_with_the_oracle(aab, monkeypatch, hop, n, rows):
    """The decoupled-warp n_fft = 2048 kernel (stft2048_v2_kernel) serves the frequency-minor complex / power outputs by default
    and the mel output when AA_STFT_V2_MEL=1; the v1 tile kernel serves the rest.  Both paths, every mode, against the float64
    oracle: multi-tile rows (131072 samples = 22 tiles of 12 frames), odd row counts (no partner row), non-power-of-two lengths
    (zero_pad_po2 tail inside a frame), hop = 1024 (sample ring depth 1) and a hop that is not a multiple of 4 (generic kernel)."""
    from oracle import aa_oracle as O
    g = torch.Generator().manual_seed(hop + n)
    x = torch.rand(*rows, n, generator=g) - 0.5
    ref = O.mel_spectrogram(x, 48000, 2048, hop)
    mel = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=hop)
    y3 = mel.encode(x.cuda())                     # stft2048_v3_kernel: warp-private walk, [row][frame][mel] buffer
    assert tuple(y3.shape) == tuple(ref.shape) and rel_l2(y3, ref) < 1e-5, "mel, v3 kernel"
    if hop % 4 == 0:
        assert y3.stride()[-2:] == (1, 128), "the v3 kernel returns torchaudio's own layout (a transposed view)"
    for v2_mel in ("0", "1"):                     # contiguous [.., mel, T] output: v1 tile kernel / v2 kernel with the banded walk
        monkeypatch.setenv("AA_STFT_V2_MEL", v2_mel)
        y = mel.encode(x.cuda(), freq_major=True)
        assert tuple(y.shape) == tuple(ref.shape) and y.is_contiguous()
        assert rel_l2(y, ref) < 1e-5, f"mel, AA_STFT_V2_MEL={v2_mel}"
        assert rel_l2(y, y3) < 2e-6
    refp = O.power_spectrogram(x, 2048, hop) if hasattr(O, "power_spectrogram") else None
    for v3, v2 in (("1", "1"), ("0", "1"), ("0", "0")):
        monkeypatch.setenv("AA_STFT_V3", v3)
        monkeypatch.setenv("AA_STFT_V2", v2)
        yp = aab.MagSpectrogramAE(n_fft=2048, hop_length=hop).encode(x.cuda())
        yc = aab.SpectrogramAE(n_fft=2048, hop_length=hop).encode(x.cuda())
        assert rel_l2(yc.abs() ** 2, yp) < 1e-5
        if refp is not None:
            assert rel_l2(yp, refp) < 1e-5, f"power, AA_STFT_V3={v3} AA_STFT_V2={v2}"


def test_mel_layout_matches_torchaudio(aab):
    """torchaudio's MelScale returns matmul(spec^T, fb)^T: the reference's mel tensors are [.., n_mels, T] VIEWS of a
    [.., T, n_mels] buffer.  The n_fft = 2048 front-end reproduces shape, values and strides; host input with a preallocated
    result in either layout (the reference's bulk loop preallocates) goes through the matching C entry point."""
    import torchaudio
    g = torch.Generator().manual_seed(77)
    x = torch.rand(5, 2, 24000, generator=g) - 0.5
    ref = torchaudio.transforms.MelSpectrogram(sample_rate=48000, n_fft=2048, hop_length=512)(x)
    mel = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512)
    mel.zero_pad = False
    y = mel.encode(x.cuda())
    assert tuple(y.shape) == tuple(ref.shape) and y.stride() == ref.stride() and rel_l2(y, ref) < TOL
    yh = mel.encode(x)                                     # host in, host out (pinned), same layout
    assert yh.device.type == "cpu" and yh.stride() == ref.stride() and rel_l2(yh, ref) < TOL
    out_c = torch.empty(tuple(ref.shape), pin_memory=True)                                   # contiguous [.., mel, T]
    out_t = torch.empty(5, 2, ref.shape[-1], 128, pin_memory=True).transpose(-1, -2)         # torchaudio's layout
    assert mel.encode(x, out=out_c) is out_c and rel_l2(out_c, ref) < TOL
    assert mel.encode(x, out=out_t) is out_t and rel_l2(out_t, ref) < TOL
    assert torch.equal(out_t, yh)


@pytest.mark.parametrize("hop,n,rows", [(256, 131072, (2, 2)), (256, 20000, (3, 1)), (128, 16384, (1, 2)), (512, 40000, (2, 2)), (250, 9000, (1, 3)),
                                        (255, 9000, (1, 2))])
def test_v3_kernel_n1024_two_frames_per_item(aab, monkeypatch, hop, n, rows):
    """stft_v3_kernel<1024>: the reference-default transform size, two consecutive frames per warp item.  Against the float64
    oracle and the older kernels: odd frame counts (the second frame of the last item does not exist), odd row counts, lengths
    that are not powers of two (zero_pad_po2 tail inside a frame), chunk edges, a hop that is even but not a multiple of 4 and
    an odd hop (not eligible: warp kernel, contiguous layout)."""
    from oracle import aa_oracle as O
    g = torch.Generator().manual_seed(1024 + hop + n)
    x = torch.rand(*rows, n, generator=g) - 0.5
    ref = O.mel_spectrogram(x, 48000, 1024, hop)
    mel = aab.MelSpectrogramAE(sample_rate=48000, n_fft=1024, hop_length=hop)
    y3 = mel.encode(x.cuda())
    assert tuple(y3.shape) == tuple(ref.shape) and rel_l2(y3, ref) < 1e-5
    if hop % 2 == 0:
        assert y3.stride()[-2:] == (1, 128), "v3 kernel: torchaudio's own layout (a transposed view)"
    yc = mel.encode(x.cuda(), freq_major=True)            # warp kernel, contiguous [.., mel, T]
    assert yc.is_contiguous() and rel_l2(yc, y3) < 2e-6
    refp = O.power_spectrogram(x, 1024, hop)
    for v3 in ("1", "0"):
        monkeypatch.setenv("AA_STFT_V3", v3)
        yp = aab.MagSpectrogramAE(n_fft=1024, hop_length=hop).encode(x.cuda())
        ycx = aab.SpectrogramAE(n_fft=1024, hop_length=hop).encode(x.cuda())
        assert rel_l2(yp, refp) < 1e-5, f"power, AA_STFT_V3={v3}"
        assert rel_l2(ycx.abs() ** 2, yp) < 1e-5


# ---------------------------------------------------------------- decoders
@pytest.mark.parametrize("tag,n_fft,hop", [("1024_256", 1024, 256), ("2048_512", 2048, 512)])
def test_decoders_match_the_reference_decoders(golden, tag, n_fft, hop):
    """SpectrogramAE.decode (aa_istft_f32), MagSpectrogramAE.decode / MelSpectrogramAE.decode (GriffinLim on the CUDA STFT and
    inverse-STFT kernels, aa_griffinlim_update_c64, aa_inverse_mel_f32) against outputs of the reference's own decoders
    (tests/golden/decoders.npz); GriffinLim with the reproducible start (ones) and 8 iterations as in the fixture."""
    import audio_algebra_b200 as aab
    g = golden("decoders")
    x = torch.from_numpy(g["x"]).cuda()
    m = aab.SpectrogramAE(n_fft=n_fft, hop_length=hop)
    z = torch.from_numpy(g[f"rand_spec_{tag}"]).cuda()
    assert rel_l2(m._istft(z), g[f"rand_istft_{tag}"]) < 5e-6                       # contiguous [F][T] input
    spec = m.encode(x)                                                               # torch.stft's own (transposed-view) layout
    rec = m.decode(spec)
    assert rec.shape == x.shape and rel_l2(rec, g[f"roundtrip_{tag}"]) < 5e-6 and rel_l2(rec, x) < 5e-6
    reps, recons = m(x)                                                              # GivenModelClass.forward -> (reps, recons)
    assert rel_l2(recons, x) < 5e-6
    mm = aab.MagSpectrogramAE(n_fft=n_fft, hop_length=hop)
    p = mm.encode(x)
    gl = mm.decode(p, rand_init=False, n_iter=8)
    assert gl.shape == x.shape and rel_l2(gl, g[f"gl8_{tag}"]) < 2e-4
    me = aab.MelSpectrogramAE(sample_rate=48000, n_fft=n_fft, hop_length=hop)
    mel = me.encode(x)
    inv = me.inverse_melscale(mel)
    assert rel_l2(inv, g[f"invmel_{tag}"]) < 5e-5
    assert rel_l2(me.decode(mel, rand_init=False, n_iter=8), g[f"mel_gl8_{tag}"]) < 2e-2


def test_griffinlim_default_random_start_reduces_the_spectral_error():
    "T.GriffinLim defaults (random phase start, 32 iterations, momentum 0.99): the rebuilt magnitude approaches the given one"
    import audio_algebra_b200 as aab
    torch.manual_seed(0)
    t = torch.arange(16384, device="cuda") / 48000.0
    x = (0.5 * torch.sin(2 * torch.pi * 440.0 * t) + 0.25 * torch.sin(2 * torch.pi * 1234.0 * t))[None, None].repeat(2, 2, 1)
    mm = aab.MagSpectrogramAE(n_fft=1024, hop_length=256)
    p = mm.encode(x)
    errs = []
    for n_iter in (1, 32):
        torch.manual_seed(1)
        w = mm.decode(p, n_iter=n_iter)
        assert w.shape == x.shape
        errs.append(float((mm.encode(w).sqrt() - p.sqrt()).norm() / p.sqrt().norm()))
    assert errs[1] < 0.5 * errs[0] and errs[1] < 0.2, errs


def test_magdphase_decode_matches_the_reference(golden):
    "MagDPhaseSpectrogramAE.decode (aa_magdphase_decode_f32 + aa_istft_f32) against the reference's own output; round trip; init variants run"
    import audio_algebra_b200 as aab
    g = golden("decoders")
    reps = torch.from_numpy(g["mdp_reps"]).cuda()
    md = aab.MagDPhaseSpectrogramAE(n_fft=1024, hop_length=256)
    x0 = torch.from_numpy(g["x"][0]).cuda()
    md.encode(x0)   # sets orig_shape for match_sizes (the encode parity has its own test: phases wrap, so it is not compared elementwise here)
    rec = md.decode(reps)
    assert rec.shape == x0.shape and rel_l2(rec, g["mdp_decode"]) < 2e-4 and rel_l2(rec, x0) < 2e-4
    for init in ("zero", "rand"):
        w = aab.MagDPhaseSpectrogramAE(n_fft=1024, hop_length=256, init=init)
        w.encode(x0)
        out = w.decode(reps)
        assert out.shape == x0.shape and bool(torch.isfinite(out).all())
