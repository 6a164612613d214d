"""Encoder side of the reference's DiffusionDVAE (audio_algebra/aa_mixer.py:109-202 ==
audio_algebra/DiffusionDVAE.py:98-160): constructor arguments, `encode` (eval: encoder_ema, NO tanh) and
`encode_it` (tanh(encoder_ema(x))).  The conv stack runs in libaa_b200 (aa_encoder_forward).

`SoundStreamXLEncoder` restates the third-party `autoencoders.soundstream.SoundStreamXLEncoder`
(audio-diffusion, un-pinned git dependency, source absent from the reference tree): architecture per
SURVEY.md Appendix A -- PARITY UNPINNED upstream; only the [B,2,N] -> [B,64,N/128] shape is pinned
(Destructo.ipynb cell 17).  Diffusion decoder / sampler, PQMF and Memcodes branches are out of scope and raise.
"""
import ctypes as C
import math
from copy import deepcopy

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .aa_mixer import _ptr_array, _f32c

__all__ = ['SoundStreamXLEncoder', 'DiffusionDVAE']

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_pp = C.POINTER(C.c_void_p)


class _Cfg(C.Structure):
    _fields_ = [("in_channels", _i), ("capacity", _i), ("latent_dim", _i), ("n_blocks", _i),
                ("c_mults", _i * 8), ("strides", _i * 8)]


_lib.register({
    "aa_encoder_create": (_i, [C.POINTER(_Cfg), C.POINTER(_p)]),
    "aa_encoder_destroy": (_i, [_p]),
    "aa_encoder_num_layers": (_i, [_p]),
    "aa_encoder_layer_shape": (_i, [_p, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "aa_encoder_set_weights": (_i, [_p, _i, _p, _p, _p]),
    "aa_encoder_out_length": (_i, [_p, _i64, C.POINTER(_i64)]),
    "aa_encoder_workspace_bytes": (_i64, [_p, _i64, _i64, _i]),
    "aa_encoder_forward": (_i, [_p, _pp, C.POINTER(_f), _i, _i64, _i64, _i, _i, _p, _p, _p]),
})

DTYPES = {"fp32": 0, "f32": 0, "float32": 0, "bf16": 1, "bfloat16": 1, "tf32x3": 2, "3xtf32": 2, "fp32_cuda_cores": 3}


class _Handles:
    "per-device AaEncoder handles + upload bookkeeping; deliberately NOT copied by deepcopy"

    def __init__(self):
        self.h, self.versions, self.ws = {}, {}, {}

    def __deepcopy__(self, memo):
        return _Handles()

    def __del__(self):
        try:
            for h in self.h.values():
                lib.aa_encoder_destroy(h)
        except Exception:
            pass


class _ResidualUnit(nn.Module):
    def __init__(self, c, d):
        super().__init__()
        self.conv1 = nn.Conv1d(c, c, 7, dilation=d, padding=3 * d)
        self.conv2 = nn.Conv1d(c, c, 1)


class _EncoderBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.res = nn.ModuleList([_ResidualUnit(cin, d) for d in (1, 3, 9)])
        self.down = nn.Conv1d(cin, cout, 2 * stride, stride=stride, padding=math.ceil(stride / 2))


class SoundStreamXLEncoder(nn.Module):
    """Conv1d(in->cap,k7) ELU [ResUnit(1) ELU ResUnit(3) ELU ResUnit(9) ELU Conv1d(k=2s,stride s) ELU]* Conv1d(->latent,k3).
    The nn.Conv1d children only hold parameters (default PyTorch init, state_dict, .to()); the arithmetic is
    one aa_encoder_forward call.  `compute_dtype`: 'fp32' (fp32-grade: 3xTF32 tcgen05 kernels where the layer shapes allow, else CUDA cores),
    'tf32x3' / 'fp32_cuda_cores' (force one or the other) or 'bf16' (tcgen05, cosine >= 0.999)."""

    def __init__(self, in_channels=2, capacity=32, latent_dim=64, c_mults=[2, 4, 8, 16, 32], strides=[4, 4, 2, 2, 2],
                 compute_dtype="fp32"):
        super().__init__()
        assert len(c_mults) == len(strides) <= 8
        self.in_channels, self.capacity, self.latent_dim = in_channels, capacity, latent_dim
        self.c_mults, self.strides, self.compute_dtype = list(c_mults), list(strides), compute_dtype
        cm = [1] + list(c_mults)
        self.conv_in = nn.Conv1d(in_channels, capacity, 7, padding=3)
        self.blocks = nn.ModuleList([_EncoderBlock(cm[i] * capacity, cm[i + 1] * capacity, s) for i, s in enumerate(strides)])
        self.conv_out = nn.Conv1d(cm[-1] * capacity, latent_dim, 3, padding=1)
        self._handles = _Handles()

    def flat_convs(self):
        out = [self.conv_in]
        for b in self.blocks:
            for r in b.res:
                out += [r.conv1, r.conv2]
            out.append(b.down)
        out.append(self.conv_out)
        return out

    def load_oracle_weights(self, other):
        "copy (weight, bias) pairs from any module exposing flat_weights() in the same layer order"
        with torch.no_grad():
            for conv, (w, b) in zip(self.flat_convs(), other.flat_weights()):
                conv.weight.copy_(w)
                conv.bias.copy_(b)
        return self

    def _handle(self, dev):
        H = self._handles
        if dev not in H.h:
            cfg = _Cfg(self.in_channels, self.capacity, self.latent_dim, len(self.strides),
                       (_i * 8)(*(self.c_mults + [0] * (8 - len(self.c_mults)))),
                       (_i * 8)(*(self.strides + [0] * (8 - len(self.strides)))))
            h = _p()
            check(lib.aa_encoder_create(C.byref(cfg), C.byref(h)))
            assert lib.aa_encoder_num_layers(h) == len(self.flat_convs())
            H.h[dev], H.versions[dev] = h, None
        h = H.h[dev]
        convs = self.flat_convs()
        ver = tuple((c.weight._version, c.bias._version, c.weight.data_ptr()) for c in convs)
        if H.versions[dev] != ver:
            for i, c in enumerate(convs):
                _lib.require_cuda(c.weight, "encoder weights")
                w, b = _f32c(c.weight.detach()), _f32c(c.bias.detach())
                check(lib.aa_encoder_set_weights(h, i, ptr(w), ptr(b), stream_ptr()))
            H.versions[dev] = ver
        return h

    def out_length(self, n):
        l = n
        for c in self.flat_convs():
            l = (l + 2 * c.padding[0] - c.dilation[0] * (c.kernel_size[0] - 1) - 1) // c.stride[0] + 1
        return l

    def encode_mix(self, stems, faders=None, apply_tanh=False):
        """encoder(sum_j faders[j] * stems[j]) with the fader-scaled sum fused into the first conv's load
        (aa_mixer.py:303,309).  stems: list of 1..4 tensors [B, in_channels, N]."""
        stems = [_f32c(s, "waveform") for s in stems]
        x0 = stems[0]
        assert x0.dim() == 3 and x0.shape[1] == self.in_channels, f"expected [B,{self.in_channels},N], got {tuple(x0.shape)}"
        assert all(s.shape == x0.shape for s in stems) and 1 <= len(stems) <= 4
        fl = [1.0] * len(stems) if faders is None else [float(f) for f in faders]
        b, n = x0.shape[0], x0.shape[2]
        dev = _lib.ensure_device(x0.device)
        dt = DTYPES[self.compute_dtype]
        with torch.cuda.device(dev):
            h = self._handle(dev)
            y = torch.empty((b, self.latent_dim, self.out_length(n)), dtype=torch.float32, device=x0.device)
            if b == 0:
                return y
            nbytes = int(lib.aa_encoder_workspace_bytes(h, b, n, dt))
            ws = self._handles.ws.get(dev)
            if ws is None or ws.numel() < nbytes:
                ws = torch.empty(nbytes, dtype=torch.uint8, device=x0.device)
                self._handles.ws[dev] = ws
            pa, keep = _ptr_array(stems)
            fa = (C.c_float * len(stems))(*fl)
            check(lib.aa_encoder_forward(h, pa, fa, len(stems), b, n, int(apply_tanh), dt, ptr(y), ptr(ws), stream_ptr()))
        return y

    def forward(self, x):
        squeeze = x.dim() == 2
        y = self.encode_mix([x.unsqueeze(0) if squeeze else x])
        return y.squeeze(0) if squeeze else y


class DiffusionDVAE(nn.Module):
    """global_args needs: pqmf_bands, latent_dim, num_quantizers (ema_decay etc. are accepted and ignored).
    Same members as the reference on the encode side: encoder, encoder_ema (deepcopy), pqmf_bands, quantized."""

    def __init__(self, global_args):
        super().__init__()
        self.pqmf_bands = global_args.pqmf_bands
        if self.pqmf_bands > 1:
            raise NotImplementedError("pqmf_bands > 1 (PQMF analysis front-end) is a 'next' row, not built yet")
        capacity = 32
        c_mults = [2, 4, 8, 16, 32]
        strides = [4, 4, 2, 2, 2]
        self.encoder = SoundStreamXLEncoder(in_channels=2 * global_args.pqmf_bands, capacity=capacity,
                                            latent_dim=global_args.latent_dim, c_mults=c_mults, strides=strides,
                                            compute_dtype=getattr(global_args, "compute_dtype", "fp32"))
        self.encoder_ema = deepcopy(self.encoder)
        self.num_quantizers = getattr(global_args, "num_quantizers", 0)
        self.quantized = self.num_quantizers > 0
        if self.quantized:
            raise NotImplementedError("num_quantizers > 0 (Memcodes) is a 'next' row, not built yet")
        self.ema_decay = getattr(global_args, "ema_decay", 0.995)
        self.demo_reals_shape = None

    @property
    def device(self):
        return self.encoder_ema.conv_in.weight.device

    def load_oracle_weights(self, other):
        self.encoder.load_oracle_weights(other)
        self.encoder_ema.load_oracle_weights(other)
        return self

    def encode(self, *args, **kwargs):
        "aa_mixer.py:165-168: encoder in training mode, encoder_ema otherwise; no tanh"
        if self.training:
            return self.encoder(*args, **kwargs)
        return self.encoder_ema(*args, **kwargs)

    def encode_mix(self, stems, faders):
        enc = self.encoder if self.training else self.encoder_ema
        return enc.encode_mix(stems, faders)

    def decode(self, *args, **kwargs):
        raise NotImplementedError("the diffusion decoder is outside the accelerated hot path (encode side only)")

    def encode_it(self, demo_reals):
        "aa_mixer.py:175-195: tanh(encoder_ema(x)); runs under no_grad like the reference"
        encoder_input = demo_reals.to(self.device)
        self.demo_reals_shape = demo_reals.shape
        with torch.no_grad():
            return self.encoder_ema.encode_mix([encoder_input], None, apply_tanh=True)

    def decode_it(self, *args, **kwargs):
        raise NotImplementedError("the diffusion decoder is outside the accelerated hot path (encode side only)")
