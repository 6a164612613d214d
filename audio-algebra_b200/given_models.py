"""Given-model ("f: audio -> y") encoders of the reference's audio_algebra/given_models.py, with the
same class names, constructor arguments and encode() contract, computed by libaa_b200 on a B200.

Mirrors /root/reference/audio_algebra/given_models.py:
    GivenModelClass :58-145, SpectrogramAE :149-168, MagSpectrogramAE :171-189,
    MagDPhaseSpectrogramAE :192-254, MelSpectrogramAE :257-283.
Only the encode side is on the hot path; decode() raises.
"""
import ctypes as C
import math
import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check, ptr, stream_ptr

__all__ = ['encode_all', 'GivenModelClass', 'SpectrogramAE', 'MagSpectrogramAE', 'MagDPhaseSpectrogramAE', 'MelSpectrogramAE', 'DVAEWrapper',
           'StackedDiffAEWrapper']


class GivenModelClass(nn.Module):
    "Same optional 'shorthand' structure as the reference's GivenModelClass (given_models.py:58-82)"

    def __init__(self, zero_pad=True, make_sizes_match=True,
                 ckpt_info={'ckpt_path': '', 'ckpt_url': '', 'ckpt_hash': '', 'gdrive_path': ''}, **kwargs):
        super().__init__()
        self.make_sizes_match, self.orig_shape, self.zero_pad, self.ckpt_info = make_sizes_match, None, zero_pad, ckpt_info
        self.name = self.__class__.__name__
        self.ckpt_dir = os.path.expanduser('~/checkpoints')  # the reference creates it eagerly (:69-70); we do not

    def setup(self, gdrive=True):
        "Setup can include things such as downloading checkpoints (none needed for the STFT family)"
        pass

    def encode(self, waveform: torch.Tensor, **kwargs) -> torch.Tensor:
        return None

    def decode(self, reps: torch.Tensor, **kwargs) -> torch.Tensor:
        raise NotImplementedError(f"{self.name}.decode: the generative decoders (diffusion sample() loops) are outside the accelerated path")

    def forward(self, waveform: torch.Tensor):
        "given_models.py:79-82: (reps, recons)"
        reps = self.encode(waveform)
        return (reps, self.decode(reps))

    def match_sizes(self, recon) -> torch.Tensor:
        "given_models.py:123-133, kept on the tensor's own device"
        if self.make_sizes_match and (self.orig_shape is not None) and (recon.shape != self.orig_shape):
            if recon.shape[-1] > self.orig_shape[-1]:
                recon = recon[..., :self.orig_shape[-1]]
            else:
                recon2 = torch.zeros(self.orig_shape, device=recon.device, dtype=recon.dtype)
                recon2[..., :recon.shape[-1]] = recon
                recon = recon2
            assert recon.shape == self.orig_shape, \
                f"Did not succeed in making size match. recon.shape ({recon.shape}) != self.orig_shape ({self.orig_shape})"
        return recon

    def next_power_of_2(self, x: int) -> int:
        return 1 if x == 0 else 2 ** (x - 1).bit_length()

    def zero_pad_po2(self, x):
        """given_models.py:139-145.  The CUDA front-ends never call this: the pad is folded into the
        kernel's load.  Kept for API compatibility; stays on x's device."""
        new_shape = list(x.shape)
        new_shape[-1] = self.next_power_of_2(new_shape[-1])
        new_x = torch.zeros(new_shape, device=x.device, dtype=x.dtype)
        new_x[..., :x.shape[-1]] = x
        return new_x


# ---------------------------------------------------------------------------------------------------
# mel filterbank (host side, once per module): torchaudio.functional.melscale_fbanks restated
# ---------------------------------------------------------------------------------------------------

def _hz_to_mel(freq, mel_scale):
    if mel_scale == "htk":
        return 2595.0 * math.log10(1.0 + (freq / 700.0))
    f_min, f_sp = 0.0, 200.0 / 3
    mels = (freq - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = math.log(6.4) / 27.0
    if freq >= min_log_hz:
        mels = min_log_mel + math.log(freq / min_log_hz) / logstep
    return mels


def _mel_to_hz(mels, mel_scale):
    if mel_scale == "htk":
        return 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    f_min, f_sp = 0.0, 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = math.log(6.4) / 27.0
    log_t = mels >= min_log_mel
    freqs[log_t] = min_log_hz * torch.exp(logstep * (mels[log_t] - min_log_mel))
    return freqs


def melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, norm=None, mel_scale="htk"):
    "[n_freqs, n_mels] triangular filterbank, same formula and fp32 arithmetic as torchaudio"
    if norm is not None and norm != "slaney":
        raise ValueError('norm must be one of None or "slaney"')
    if mel_scale not in ("htk", "slaney"):
        raise ValueError('mel_scale should be one of "htk" or "slaney".')
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel(f_min, mel_scale), _hz_to_mel(f_max, mel_scale), n_mels + 2)
    f_pts = _mel_to_hz(m_pts, mel_scale)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    zero = torch.zeros(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(zero, torch.min(down_slopes, up_slopes))
    if norm == "slaney":
        enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
        fb = fb * enorm.unsqueeze(0)
    return fb


class _StftPlan:
    "RAII holder of an AaStftPlan (one per module per device)"

    def __init__(self, n_fft, hop, center, window, n_mels=0, sample_rate=48000.0, f_min=0.0, f_max=24000.0, fb=None):
        self.handle = C.c_void_p()
        win = None if window is None else window.detach().to("cpu", torch.float32).contiguous()
        fbc = None if fb is None else fb.detach().to("cpu", torch.float32).contiguous()
        check(lib.aa_stft_plan_create(C.byref(self.handle), int(n_fft), int(hop), int(bool(center)),
                                      None if win is None else C.c_void_p(win.data_ptr()),
                                      int(n_mels), float(sample_rate), float(f_min), float(f_max),
                                      None if fbc is None else C.c_void_p(fbc.data_ptr())))

    def __del__(self):
        try:
            if self.handle:
                lib.aa_stft_plan_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    def out_shape(self, n_in, zero_pad):
        n_pad, n_frames = C.c_int64(), C.c_int64()
        check(lib.aa_stft_out_shape(self.handle, int(n_in), int(bool(zero_pad)), C.byref(n_pad), C.byref(n_frames)))
        return n_pad.value, n_frames.value

    def mel_tf_supported(self):
        return bool(lib.aa_stft_mel_tf_supported(self.handle))


class _StftFrontEnd(GivenModelClass):
    """Shared plumbing of the three STFT encoders.  torchaudio kwargs understood: win_length,
    window_fn, wkwargs (and for mel: n_mels, f_min, f_max, norm, mel_scale).  Options whose arithmetic
    the kernels do not implement (pad != 0, normalized, pad_mode != 'reflect', onesided=False,
    power not in the class's fixed value) raise instead of silently differing."""

    def __init__(self, n_fft, hop_length, center, kwargs, mel=None):
        super().__init__()
        self.n_fft, self.hop_length, self.center = int(n_fft), int(hop_length), bool(center)
        if self.n_fft < 64 or self.n_fft > 8192 or (self.n_fft & (self.n_fft - 1)) != 0:
            raise _lib.AaError(f"n_fft={n_fft}: the CUDA STFT supports powers of two in [64, 8192]")
        if self.hop_length < 1:
            raise _lib.AaError(f"hop_length={hop_length} must be >= 1")
        kw = dict(kwargs)
        win_length = kw.pop("win_length", None) or self.n_fft
        window_fn = kw.pop("window_fn", torch.hann_window)
        wkwargs = kw.pop("wkwargs", None)
        if kw.pop("pad", 0) != 0:
            raise NotImplementedError("pad != 0 is not supported by the CUDA STFT front-end")
        if kw.pop("normalized", False) not in (False, None):
            raise NotImplementedError("normalized spectrograms are not supported by the CUDA STFT front-end")
        if kw.pop("pad_mode", "reflect") != "reflect":
            raise NotImplementedError("only pad_mode='reflect' is supported")
        if kw.pop("onesided", True) is not True:
            raise NotImplementedError("only onesided=True is supported")
        window = window_fn(win_length) if wkwargs is None else window_fn(win_length, **wkwargs)
        if win_length < self.n_fft:  # torch.stft centres a short window inside n_fft
            left = (self.n_fft - win_length) // 2
            w = torch.zeros(self.n_fft)
            w[left:left + win_length] = window
            window = w
        elif win_length > self.n_fft:
            raise ValueError("win_length must be <= n_fft")
        self.register_buffer("window", window.float(), persistent=False)
        self._mel = None
        if mel is not None:
            sample_rate = mel["sample_rate"]
            self.n_mels = int(kw.pop("n_mels", 128))
            self.f_min = float(kw.pop("f_min", 0.0))
            f_max = kw.pop("f_max", None)
            self.f_max = float(f_max) if f_max is not None else float(sample_rate // 2)
            if kw.pop("power", 2.0) != 2.0:
                raise NotImplementedError("MelSpectrogramAE computes the power (=2) mel spectrogram only")
            norm, mel_scale = kw.pop("norm", None), kw.pop("mel_scale", "htk")
            fb = melscale_fbanks(self.n_fft // 2 + 1, self.f_min, self.f_max, self.n_mels, sample_rate, norm, mel_scale)
            self.register_buffer("fb", fb, persistent=False)
            self._mel = dict(sample_rate=sample_rate)
        if kw:
            raise TypeError(f"unsupported keyword arguments for {self.__class__.__name__}: {sorted(kw)}")
        self._plans = {}

    def _plan(self, device_index):
        if device_index not in self._plans:
            with torch.cuda.device(device_index):
                if self._mel is None:
                    self._plans[device_index] = _StftPlan(self.n_fft, self.hop_length, self.center, self.window)
                else:
                    self._plans[device_index] = _StftPlan(self.n_fft, self.hop_length, self.center, self.window,
                                                          self.n_mels, self._mel["sample_rate"], self.f_min, self.f_max,
                                                          self.fb)
        return self._plans[device_index]

    def _run(self, waveform, mode, out=None, freq_major=False):
        """waveform [..., N] float32 -> [..., F|n_mels, T].  Spectrograms come back as the transposed view of a
        [..., T, F|n_mels] buffer -- exactly what torch.stft / torchaudio.transforms.Spectrogram / MelScale return (same shape,
        values AND strides as the reference's tensors), and the layout the kernels write at full sector width (mel: when the plan
        supports it, i.e. n_fft = 2048, Hann, triangular bank); freq_major=True asks for a contiguous [..., F, T] tensor instead.  CUDA input: stream-ordered kernel on the
        current stream.  CPU input: H2D, kernel, D2H (result returned on the CPU, like the reference
        keeps the input's device) -- the mel variant pipelines the copies in chunks; `out` may be a
        preallocated (ideally pinned) CPU tensor, as in the reference's bulk-encode loop
        (xae_dataset.ipynb cell 50 writes into a preallocated `reps` array)."""
        self.orig_shape = waveform.shape
        if waveform.dtype != torch.float32:
            waveform = waveform.float()
        on_cpu = not waveform.is_cuda
        dev = _lib.ensure_device(None if on_cpu else waveform.device)
        lead, n_in = waveform.shape[:-1], waveform.shape[-1]
        rows = int(math.prod(lead)) if len(lead) else 1
        plan = self._plan(dev)
        n_pad, n_frames = plan.out_shape(n_in, self.zero_pad)
        bins = self.n_mels if mode == "mel" else self.n_fft // 2 + 1
        with torch.cuda.device(dev):
            mel_tf = mode == "mel" and not freq_major and plan.mel_tf_supported()
            if on_cpu and mode == "mel":
                x = waveform.contiguous()
                shape = (*lead, bins, n_frames)
                if out is None:
                    out = (torch.empty((*lead, n_frames, bins), dtype=torch.float32, pin_memory=True).transpose(-1, -2) if mel_tf
                           else torch.empty(shape, dtype=torch.float32, pin_memory=True))
                elif tuple(out.shape) != shape or out.dtype != torch.float32 or out.is_cuda:
                    raise ValueError(f"out must be a float32 CPU tensor of shape {shape}")
                if out.is_contiguous():
                    check(lib.aa_stft_mel_f32_host(plan.handle, ptr(x), rows, n_in, int(self.zero_pad), ptr(out), 64))
                elif mel_tf and out.transpose(-1, -2).is_contiguous():   # the reference's own layout: a [.., mel, T] view of [.., T, mel]
                    check(lib.aa_stft_mel_tf_f32_host(plan.handle, ptr(x), rows, n_in, int(self.zero_pad), ptr(out), 64))
                else:
                    raise ValueError("out must be contiguous, or the transposed view of a contiguous [..., frames, n_mels] tensor")
                return out
            x = waveform.to(f"cuda:{dev}", non_blocking=True).contiguous()
            if mode in ("complex", "power") and not freq_major:
                out = torch.empty((*lead, n_frames, bins), dtype=torch.complex64 if mode == "complex" else torch.float32, device=x.device)
                fn = lib.aa_stft_complex_tf_f32 if mode == "complex" else lib.aa_stft_power_tf_f32
                check(fn(plan.handle, ptr(x), rows, n_in, int(self.zero_pad), ptr(out), stream_ptr()))
                out = out.transpose(-1, -2)
            elif mode == "complex":
                out = torch.empty((*lead, bins, n_frames), dtype=torch.complex64, device=x.device)
                check(lib.aa_stft_complex_f32(plan.handle, ptr(x), rows, n_in, int(self.zero_pad), ptr(out), stream_ptr()))
            elif mode == "power":
                out = torch.empty((*lead, bins, n_frames), dtype=torch.float32, device=x.device)
                check(lib.aa_stft_power_f32(plan.handle, ptr(x), rows, n_in, int(self.zero_pad), ptr(out), stream_ptr()))
            elif mel_tf:
                out = torch.empty((*lead, n_frames, bins), dtype=torch.float32, device=x.device)
                check(lib.aa_stft_mel_tf_f32(plan.handle, ptr(x), rows, n_in, int(self.zero_pad), ptr(out), stream_ptr()))
                out = out.transpose(-1, -2)
            else:
                out = torch.empty((*lead, bins, n_frames), dtype=torch.float32, device=x.device)
                check(lib.aa_stft_mel_f32(plan.handle, ptr(x), rows, n_in, int(self.zero_pad), ptr(out), stream_ptr()))
        return out.cpu() if on_cpu else out


    # ---- decoder side (round-trip demos) -------------------------------------------------------------
    def _istft(self, spec: torch.Tensor) -> torch.Tensor:
        """torch.istft / T.InverseSpectrogram(n_fft, hop, center, window) of a complex [..., F, T] tensor (any strides: the
        transposed view encode() returns is read in place) -> [..., hop * (T - 1)] float32 on spec's device."""
        if not spec.is_complex():
            raise TypeError("expected a complex spectrogram")
        if self.n_fft > 4096:
            raise NotImplementedError("the inverse STFT supports n_fft <= 4096")
        on_cpu = not spec.is_cuda
        dev = _lib.ensure_device(None if on_cpu else spec.device)
        with torch.cuda.device(dev):
            z = spec.to(f"cuda:{dev}").to(torch.complex64)
            lead, (f, t) = z.shape[:-2], z.shape[-2:]
            assert f == self.n_fft // 2 + 1, f"expected {self.n_fft // 2 + 1} frequency bins, got {f}"
            rows = int(math.prod(lead)) if len(lead) else 1
            z3 = z.reshape(rows, f, t) if z.dim() != 3 else z     # views keep the [T][F] memory of torch.stft's own layout
            if rows > 1 and z3.stride(0) == 0:
                z3 = z3.contiguous()
            if not self.center and t * self.hop_length < 1:
                raise ValueError("empty output")
            out_len = self.hop_length * (t - 1) if self.center else self.n_fft + self.hop_length * (t - 1)
            if out_len < 1:
                raise ValueError("istft needs at least two frames when center=True")
            out = torch.empty((rows, out_len), dtype=torch.float32, device=z.device)
            ws = torch.empty(int(lib.aa_istft_workspace_floats(rows, self.n_fft, t)), dtype=torch.float32, device=z.device)
            win = self.window.to(z.device)
            check(lib.aa_istft_f32(ptr(z3), rows, self.n_fft, self.hop_length, int(self.center), t, z3.stride(0), z3.stride(1), z3.stride(2),
                                   ptr(win), ptr(out), out_len, ptr(ws), stream_ptr()))
            out = out.reshape(*lead, out_len)
        return out.cpu() if on_cpu else out

    def _griffinlim(self, specgram: torch.Tensor, power=2.0, n_iter=32, momentum=0.99, rand_init=True, init_angles=None) -> torch.Tensor:
        """torchaudio.functional.griffinlim (T.GriffinLim defaults: power 2, 32 iterations, momentum 0.99, random initial phase)
        on the CUDA STFT / inverse-STFT kernels.  init_angles (complex [..., F, T]) replaces the random start (tests)."""
        on_cpu = not specgram.is_cuda
        dev = _lib.ensure_device(None if on_cpu else specgram.device)
        with torch.cuda.device(dev):
            sg = specgram.to(f"cuda:{dev}").float()
            lead, (f, t) = sg.shape[:-2], sg.shape[-2:]
            rows = int(math.prod(lead)) if len(lead) else 1
            mag = sg.reshape(rows, f, t).pow(1.0 / power).transpose(1, 2).contiguous()          # [rows][T][F]: the STFT kernels' layout
            if init_angles is not None:
                ang = init_angles.to(sg.device).to(torch.complex64).reshape(rows, f, t).transpose(1, 2).contiguous()
            elif rand_init:
                ang = torch.rand((rows, f, t), dtype=torch.complex64, device=sg.device).transpose(1, 2).contiguous()
            else:
                ang = torch.ones((rows, t, f), dtype=torch.complex64, device=sg.device)
            prod = (mag * ang).contiguous()
            tprev = torch.zeros_like(prod)
            mom = float(momentum) / (1.0 + float(momentum))
            saved = (self.zero_pad, self.orig_shape)
            self.zero_pad = False
            try:
                for it in range(n_iter):
                    inverse = self._istft(prod.transpose(1, 2))
                    rebuilt = self._run(inverse, "complex")                   # transposed view of a contiguous [rows][T][F] buffer
                    rb = rebuilt.transpose(-1, -2)
                    assert rb.is_contiguous() and rb.shape == prod.shape
                    check(lib.aa_griffinlim_update_c64(ptr(rb), ptr(tprev), ptr(mag), ptr(prod), prod.numel(), mom, int(it == 0), stream_ptr()))
                wav = self._istft(prod.transpose(1, 2))
            finally:
                self.zero_pad, self.orig_shape = saved
            wav = wav.reshape(*lead, wav.shape[-1])
        return wav.cpu() if on_cpu else wav


class SpectrogramAE(_StftFrontEnd):
    "Raw (complex) spectrogram (given_models.py:149-168); encode -> complex64 [..., n_fft/2+1, frames]"

    def __init__(self, n_fft=1024, hop_length=256, center=True, **kwargs):
        super().__init__(n_fft, hop_length, center, kwargs)

    def encode(self, waveform: torch.Tensor, **kwargs) -> torch.Tensor:
        return self._run(waveform, "complex")

    def decode(self, reps: torch.Tensor, **kwargs) -> torch.Tensor:
        "given_models.py:166-168: InverseSpectrogram -- perfect reconstruction (aa_istft_f32)"
        return self.match_sizes(self._istft(reps))


class MagSpectrogramAE(_StftFrontEnd):
    "Power spectrogram |X|^2 (given_models.py:171-189; torchaudio power=2)"

    def __init__(self, n_fft=1024, hop_length=256, center=True, **kwargs):
        super().__init__(n_fft, hop_length, center, kwargs)

    def encode(self, waveform: torch.Tensor, **kwargs) -> torch.Tensor:
        return self._run(waveform, "power")

    def decode(self, reps: torch.Tensor, **kwargs) -> torch.Tensor:
        "given_models.py:187-189: GriffinLim *guesses* at the phase (kwargs: n_iter, momentum, rand_init, init_angles)"
        return self.match_sizes(self._griffinlim(reps, **kwargs))


class MagDPhaseSpectrogramAE(_StftFrontEnd):
    """Magnitude + phase-change spectrogram (given_models.py:192-254): both the phase-difference branch and the use_cos
    (acos of the normalised dot product) branch, plus the `debug` phase wrap.  Like the reference, encode() is defined for
    unbatched [c, N] input and returns [2c, F, T]."""

    def __init__(self, n_fft=1024, hop_length=256, center=True, init='true', use_cos=False, debug=False,
                 cheat=False, **kwargs):
        super().__init__(n_fft, hop_length, center, kwargs)
        self.use_cos, self.cheat, self.debug, self.init = use_cos, cheat, debug, init
        self.pi = 3.141592653589

    def encode(self, waveform: torch.Tensor, **kwargs) -> torch.Tensor:
        if waveform.dim() != 2:
            raise ValueError("MagDPhaseSpectrogramAE.encode expects unbatched [channels, samples] input "
                             "(the reference indexes dtheta[:,:,0] and concatenates on dim 0)")
        on_cpu = not waveform.is_cuda
        spec = self._run(waveform.cuda() if on_cpu else waveform, "complex", freq_major=True)   # aa_magdphase_f32 reads [c][F][T]
        c, f, t = spec.shape
        out = torch.empty((2 * c, f, t), dtype=torch.float32, device=spec.device)
        with torch.cuda.device(spec.device):
            check(lib.aa_magdphase_ex_f32(ptr(spec), c, f, t, int(bool(self.use_cos)), int(bool(self.debug)), ptr(out), stream_ptr()))
        return out.cpu() if on_cpu else out

    def decode(self, reps: torch.Tensor, **kwargs) -> torch.Tensor:
        """given_models.py:233-254: split [2c, F, T] into magnitudes and phase differences, integrate the phase along time (wrap at
        2 pi; first frame per `init`: 'true' | 'rand' | 'zero'), spec = mag * exp(i theta), then InverseSpectrogram.  `cheat`
        (reuse the encoder's stored phases) is a debugging aid of the reference and is not carried."""
        if self.cheat:
            raise NotImplementedError("cheat=True (decode with the phases stored by encode) is a debugging aid and is not supported")
        if reps.dim() != 3 or reps.shape[0] % 2:
            raise ValueError("MagDPhaseSpectrogramAE.decode expects [2 * channels, F, T]")
        on_cpu = not reps.is_cuda
        dev = _lib.ensure_device(None if on_cpu else reps.device)
        with torch.cuda.device(dev):
            r = reps.to(f"cuda:{dev}").float().contiguous()
            c, f, t = r.shape[0] // 2, r.shape[1], r.shape[2]
            mode = {"true": 0, "rand": 1}.get(self.init, 2)
            th0 = torch.rand((c, f), device=r.device) if mode == 1 else None
            spec = torch.empty((c, f, t), dtype=torch.complex64, device=r.device)
            check(lib.aa_magdphase_decode_f32(ptr(r), c, f, t, mode, None if th0 is None else ptr(th0), ptr(spec), stream_ptr()))
            wav = self.match_sizes(self._istft(spec))
        return wav.cpu() if on_cpu else wav


class MelSpectrogramAE(_StftFrontEnd):
    "Mel power spectrogram (given_models.py:257-283): torchaudio MelSpectrogram defaults (128 HTK bins)"

    def __init__(self, sample_rate=48000, n_fft=1024, hop_length=256, center=True, **kwargs):
        super().__init__(n_fft, hop_length, center, kwargs, mel=dict(sample_rate=sample_rate))
        self.sample_rate = sample_rate

    def encode(self, waveform: torch.Tensor, **kwargs) -> torch.Tensor:
        return self._run(waveform, "mel", out=kwargs.get("out"), freq_major=bool(kwargs.get("freq_major", False)))

    def _inv_mel_pinv(self):
        """T.InverseMelScale(n_stft=n_fft // 2 + 1) as the reference builds it (given_models.py:268): torchaudio DEFAULTS for
        everything else -- n_mels 128, sample_rate 16000, f_min 0, f_max 8000, HTK -- i.e. NOT the encoder's 48 kHz bank (a
        reference quirk, kept).  Its forward is relu(lstsq(fb^T, mel)) = relu(pinv(fb^T) mel) for the full-rank underdetermined
        system; the pseudo-inverse is formed once in float64."""
        if getattr(self, "_pinv", None) is None:
            fb = melscale_fbanks(self.n_fft // 2 + 1, 0.0, 8000.0, 128, 16000, None, "htk").double()      # [F, 128]
            self._pinv = torch.linalg.pinv(fb.T).float().contiguous()                                         # [F, 128]
        return self._pinv

    def inverse_melscale(self, melspec: torch.Tensor) -> torch.Tensor:
        "T.InverseMelScale.forward: [..., 128, T] -> [..., n_fft // 2 + 1, T] (aa_inverse_mel_f32)"
        on_cpu = not melspec.is_cuda
        dev = _lib.ensure_device(None if on_cpu else melspec.device)
        with torch.cuda.device(dev):
            if melspec.dim() < 2:
                raise ValueError(f"expected a mel spectrogram [..., 128, time], got shape {tuple(melspec.shape)}")
            m = melspec.to(f"cuda:{dev}").float()
            lead, (nm, t) = m.shape[:-2], m.shape[-2:]
            if nm != 128:
                raise ValueError(f"Expected an input with 128 mel bins. Found: {nm}")
            rows = int(math.prod(lead)) if len(lead) else 1
            m3 = m.reshape(rows, nm, t)
            f = self.n_fft // 2 + 1
            out = torch.empty((rows, f, t), dtype=torch.float32, device=m.device)
            P = self._inv_mel_pinv().to(m.device)
            check(lib.aa_inverse_mel_f32(ptr(P), ptr(m3), rows, nm, f, t, m3.stride(0), m3.stride(1), m3.stride(2), ptr(out), stream_ptr()))
            out = out.reshape(*lead, f, t)
        return out.cpu() if on_cpu else out

    def decode(self, melspec: torch.Tensor, **kwargs) -> torch.Tensor:
        "given_models.py:278-280: InverseMelScale then GriffinLim"
        return self.match_sizes(self._griffinlim(self.inverse_melscale(melspec), **kwargs))


class DVAEWrapper(GivenModelClass):
    """Wrapper for DiffusionDVAE (given_models.py:286-358): encode() = tanh(encoder_ema(waveform)).
    The reference also draws decoder noise of the input's size on every encode (given_models.py:320);
    it is only consumed by decode(), which is out of scope, so it is not generated."""

    def __init__(self,
                 args_dict={'num_quantizers': 0, 'sample_size': 65536, 'demo_steps': 50, 'sample_rate': 48000, 'latent_dim': 64,
                            'pqmf_bands': 1, 'ema_decay': 0.995},
                 debug=True, **kwargs):
        super().__init__()
        from .DiffusionDVAE import DiffusionDVAE

        class DictObj:
            def __init__(self, in_dict: dict):
                for key, val in in_dict.items():
                    if isinstance(val, (list, tuple)):
                        setattr(self, key, [DictObj(x) if isinstance(x, dict) else x for x in val])
                    else:
                        setattr(self, key, DictObj(val) if isinstance(val, dict) else val)

        self.global_args = DictObj(dict(args_dict, **{k: v for k, v in kwargs.items() if k == "compute_dtype"}))
        self.model = DiffusionDVAE(self.global_args)
        self.model.eval()
        self.noise = None
        self.demo_steps = self.global_args.demo_steps
        self.demo_samples = self.global_args.sample_size
        self.debug = debug
        self.ckpt_info = {'ckpt_url': 'https://drive.google.com/file/d/1C3NMdQlmOcArGt1KL7pH32KtXVCOfXKr/view?usp=sharing',
                          'ckpt_hash': '6a304c3e89ea3f7ca023f4c9accc5df8de0504595db41961cc7e8b0d07876ef5',
                          'gdrive_path': 'MyDrive/AI/checkpoints/DiffusionDVAE.ckpt',
                          'ckpt_path': '~/checkpoints/dvae_checkpoint.ckpt'}

    def setup(self, gdrive=True):
        """The reference downloads the 4 GB Lightning checkpoint here (given_models.py:340-356) and keeps
        random weights when that fails; there is no network in this build, so weights stay as initialised
        (or as loaded through load_state_dict / model.load_oracle_weights)."""
        path = os.path.expanduser(self.ckpt_info['ckpt_path'])
        if os.path.exists(path):
            from .DiffusionDVAE import load_dvae_encoder_checkpoint
            try:   # same failure policy as the reference (:351-354): report and keep the current weights
                n = load_dvae_encoder_checkpoint(self.model, path)
                if self.debug:
                    print(f"DVAEWrapper.setup: loaded encoder weights from {path}: {n}")
            except Exception as e:
                print(f"DVAEWrapper.setup: could not load {path} ({e}); going with current weights")
        elif self.debug:
            print("DVAEWrapper.setup: no checkpoint download in this build; keeping current encoder weights")

    def encode_it(self, demo_reals):
        return self.model.encode_it(demo_reals), None

    def encode(self, waveform: torch.Tensor, **kwargs) -> torch.Tensor:
        self.orig_shape = waveform.shape
        self.demo_samples = waveform.shape[-1]
        on_cpu = not waveform.is_cuda
        reps = self.model.encode_it(waveform)
        return reps.cpu() if on_cpu else reps


class StackedDiffAEWrapper(GivenModelClass):
    """Wrapper for the stacked latent diffusion autoencoder (given_models.py:361-417): encode(reals) returns the coarsest single
    stage of representations, tanh(latent_encoder(autoencoder.encode(reals))): [B, 2, N] -> [B, 32, N / 512].  Constructor
    arguments, members (first_stage_config, first_stage_autoencoder, model, latent_dim, latent_downsampling_ratio, ckpt_info) and
    the post-setup aliasing (latent_encoder = latent_encoder_ema) follow the reference; the decoders raise."""

    def __init__(self, debug=True, first_stage_config=None, ckpt_info=None, **kwargs):
        super().__init__()
        from .StackedAELatentDiffusionCond import AudioAutoencoder, LatentAudioDiffusionAutoencoder
        self.debug = debug
        self.first_stage_config = first_stage_config if first_stage_config is not None else \
            {"capacity": 64, "c_mults": [2, 4, 8, 16, 32], "strides": [2, 2, 2, 2, 2], "latent_dim": 32}
        cfg = dict(self.first_stage_config, **{k: v for k, v in kwargs.items() if k == "compute_dtype"})
        self.first_stage_autoencoder = AudioAutoencoder(**cfg).requires_grad_(False)
        self.model = LatentAudioDiffusionAutoencoder(autoencoder=self.first_stage_autoencoder)
        self.latent_dim = self.model.latent_dim
        self.latent_downsampling_ratio = self.model.latent_downsampling_ratio
        self.ckpt_info = ckpt_info if ckpt_info is not None else \
            {'ckpt_path': '~/checkpoints/stacked-diffae-more-310k.ckpt',
             'ckpt_hash': '91f33839ecb6e3c41b1e89e1a9e0de0dac2ebe1795efa034797429c202600a58',
             'ckpt_url': '', 'gdrive_path': ''}

    def encode(self, reals: torch.Tensor) -> torch.Tensor:
        self.orig_shape = reals.shape
        on_cpu = not reals.is_cuda
        dev = next(self.model.parameters()).device
        reps = self.model.encode(reals.to(dev))
        return reps.cpu() if on_cpu else reps

    def setup(self, gdrive=True):
        """The reference loads `stacked-diffae-more-310k.ckpt` with LatentAudioDiffusionAutoencoder.load_from_checkpoint (keeping random
        weights when that fails, :397-402) and then makes the EMA copies the live ones (:404-407).  No checkpoint is reachable from
        this build; the aliasing is reproduced."""
        if self.debug:
            print(f"{self.__class__.__name__}: no checkpoint download in this build; going with current weights")
        if hasattr(self.model, "latent_encoder_ema"):
            self.model.latent_encoder = self.model.latent_encoder_ema
            del self.model.latent_encoder_ema
        self.model.eval()


def _npy_shards(save_path, n, tail_shape, dtype, shard_rows):
    "file-backed output: one .npy (np.lib.format.open_memmap) or `stem_00000.npy`, ... of shard_rows rows each; returns [(lo, hi, memmap)]"
    import numpy as np
    if not shard_rows or shard_rows >= n:
        return [(0, n, np.lib.format.open_memmap(save_path, mode="w+", dtype=dtype, shape=(n,) + tail_shape))], [save_path]
    stem = save_path[:-4] if save_path.endswith(".npy") else save_path
    shards, names = [], []
    for k, lo in enumerate(range(0, n, shard_rows)):
        hi = min(lo + shard_rows, n)
        names.append(f"{stem}_{k:05d}.npy")
        shards.append((lo, hi, np.lib.format.open_memmap(names[-1], mode="w+", dtype=dtype, shape=(hi - lo,) + tail_shape)))
    return shards, names


def encode_all(given_model, data, batch_size=64, out=None, device=None, save_path=None, shard_rows=None):
    """Bulk encode loop of the reference (xae_dataset.ipynb cell 50 `encode_all`, effects_explorer.ipynb cell 36):

        reps[i:i+bs] = given_model.encode(data[i:i+bs].to(device)).cpu()     ...     np.save(reps_filename, reps_full)

    with the stages overlapped instead of run back to back: batch i+1 is copied host->device (from pinned
    staging, on a copy stream) while batch i is encoded, and the representations of batch i-1 return to the host
    on a third stream.  Encodes stay on ONE stream (a model's kernels share one workspace).  `data`: CPU tensor
    [Ntot, C, N]; `out`: optional preallocated CPU tensor [Ntot, ...] (pinned memory makes the return copy
    asynchronous); returns it.  The notebook's positional order (audio_full, batch_size, given_model, device) is accepted too.
    `save_path`: write the representations as .npy WHILE encoding (the notebook's np.save afterwards, without the second pass
    over the data): batches land in pinned staging and a writer thread moves them into np.lib.format.open_memmap files --
    one file, or `stem_00000.npy`, ... of `shard_rows` rows each; returns the list of files (np.load / np.load(mmap_mode='r'))."""
    if torch.is_tensor(given_model) or type(given_model).__module__ == "numpy":
        # the notebook's positional order: encode_all(audio_full, batch_size, given_model, device)
        audio_full, bsz, gm, dv = given_model, data, batch_size, out
        if not isinstance(gm, nn.Module) or not isinstance(bsz, int):
            raise TypeError("expected encode_all(given_model, data, ...) or the notebook's encode_all(audio_full, batch_size, given_model, device)")
        given_model, data, batch_size, out = gm, audio_full, bsz, None
        device = dv if device is None else device
    if not torch.is_tensor(data):
        data = torch.from_numpy(data).float()
    if data.is_cuda:
        raise ValueError("encode_all takes the host-resident dataset tensor; call given_model.encode for device tensors")
    dev = torch.device(device) if device is not None else next((p.device for p in given_model.parameters() if p.is_cuda),
                                                               torch.device("cuda", torch.cuda.current_device()))
    n = data.shape[0]
    data = data.float() if data.dtype != torch.float32 else data
    writer, shards, names, stage_out, wq = None, None, None, None, None
    with torch.cuda.device(dev):
        compute = torch.cuda.current_stream(dev)
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        bs = max(1, min(int(batch_size), max(n, 1)))
        pinned = None if data.is_pinned() else [torch.empty((bs,) + tuple(data.shape[1:]), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        d_in = [torch.empty((bs,) + tuple(data.shape[1:]), dtype=torch.float32, device=dev) for _ in range(2)]
        ev_h2d = [torch.cuda.Event() for _ in range(2)]      # H2D of slot s finished (its pinned buffer may be refilled)
        ev_free = [torch.cuda.Event() for _ in range(2)]     # the encode that read d_in[s] finished
        used = [False, False]
        for i, lo in enumerate(range(0, n, bs)):
            hi, s = min(lo + bs, n), i & 1
            nb = hi - lo
            if pinned is not None:
                if used[s]:
                    ev_h2d[s].synchronize()
                pinned[s][:nb].copy_(data[lo:hi])
                src = pinned[s][:nb]
            else:
                src = data[lo:hi]
            with torch.cuda.stream(s_in):
                if used[s]:
                    s_in.wait_event(ev_free[s])
                d_in[s][:nb].copy_(src, non_blocking=True)
                ev_h2d[s].record(s_in)
            compute.wait_event(ev_h2d[s])
            with torch.no_grad():
                reps = given_model.encode(d_in[s][:nb])
            ev_free[s].record(compute)
            used[s] = True
            ev_done = torch.cuda.Event()
            ev_done.record(compute)
            if save_path is not None:
                if writer is None:    # first batch fixes the representation shape: open the files, start the writer
                    import queue
                    import threading
                    shards, names = _npy_shards(save_path, n, tuple(reps.shape[1:]), "float32" if reps.dtype == torch.float32 else str(reps.dtype).split(".")[-1],
                                                shard_rows)
                    stage_out = [torch.empty((bs,) + tuple(reps.shape[1:]), dtype=reps.dtype, pin_memory=True) for _ in range(3)]
                    free_slots, wq = queue.Queue(), queue.Queue()
                    for k in range(3):
                        free_slots.put(k)

                    def _write():
                        while True:
                            item = wq.get()
                            if item is None:
                                return
                            slot, a, b_, ev = item
                            ev.synchronize()
                            arr = stage_out[slot][:b_ - a].numpy()
                            for slo, shi, mm in shards:      # a batch may straddle shard boundaries
                                c0, c1 = max(a, slo), min(b_, shi)
                                if c0 < c1:
                                    mm[c0 - slo:c1 - slo] = arr[c0 - a:c1 - a]
                            free_slots.put(slot)

                    writer = threading.Thread(target=_write, daemon=True)
                    writer.start()
                slot = free_slots.get()
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done)
                    stage_out[slot][:nb].copy_(reps, non_blocking=True)
                    ev_copied = torch.cuda.Event()
                    ev_copied.record(s_out)
                reps.record_stream(s_out)
                wq.put((slot, lo, hi, ev_copied))
                continue
            if out is None:
                out = torch.empty((n,) + tuple(reps.shape[1:]), dtype=reps.dtype, pin_memory=True)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done)
                out[lo:hi].copy_(reps, non_blocking=True)
            reps.record_stream(s_out)
        s_out.synchronize()
        compute.synchronize()
    if save_path is not None:
        if writer is not None:
            wq.put(None)
            writer.join()
            for _, _, mm in shards:
                mm.flush()
            return names
        import numpy as np   # empty dataset
        np.save(save_path, given_model.encode(data.to(dev)).cpu().numpy())
        return [save_path]
    if out is None:   # empty dataset
        out = given_model.encode(data.to(dev)).cpu()
    return out
