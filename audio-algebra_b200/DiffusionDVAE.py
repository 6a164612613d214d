"""Encoder side of the reference's DiffusionDVAE (audio_algebra/aa_mixer.py:109-202 ==
audio_algebra/DiffusionDVAE.py:98-160): constructor arguments, `encode` (eval: encoder_ema, NO tanh) and
`encode_it` (tanh(encoder_ema(x))).  The conv stack runs in libaa_b200 (aa_encoder_forward).

`SoundStreamXLEncoder` restates the third-party `autoencoders.soundstream.SoundStreamXLEncoder`
(audio-diffusion, un-pinned git dependency, source absent from the reference tree): architecture per
SURVEY.md Appendix A -- PARITY UNPINNED upstream; only the [B,2,N] -> [B,64,N/128] shape is pinned
(Destructo.ipynb cell 17).  `PQMF`, `Memcodes` and `ResidualMemcodes` restate the other three third-party modules on
encode_it's route (`diffusion.pqmf.PQMF`, `nwt_pytorch.Memcodes`, `dvae.residual_memcodes.ResidualMemcodes`), equally
unpinned and off at the reference defaults (pqmf_bands = 1, num_quantizers = 0).  The diffusion decoder / sampler is out
of scope and raises.
"""
import ctypes as C
import math
from copy import deepcopy

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .aa_mixer import _ptr_array, _f32c

__all__ = ['SoundStreamXLEncoder', 'DiffusionDVAE', 'PQMF', 'Memcodes', 'ResidualMemcodes', 'load_dvae_encoder_checkpoint']

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
    "aa_pqmf_analysis_f32": (_i, [_p, _i64, _i64, _p, _i, _i, _p, _p]),
    "aa_memcodes_kv_f32": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "aa_memcodes_quantize_f32": (_i, [_p, _i64, _i, _i, _i64, _p, _p, _i, _f, _p, _p, _p, _i, _p, _p]),
})

DTYPES = {"fp32": 0, "f32": 0, "float32": 0, "bf16": 1, "bfloat16": 1, "tf32x3": 2, "3xtf32": 2, "fp32_cuda_cores": 3}


class _Handles:
    "per-device AaEncoder handles + upload bookkeeping; deliberately NOT copied by deepcopy"

    def __init__(self):
        self.h, self.versions, self.ws = {}, {}, {}   # ws: one activation workspace per (device, stream)
        self.w_event = {}   # per device: recorded after the last weight upload; other streams wait on it before a forward

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
        ver = tuple((c.weight._version, c.bias._version, c.weight.data_ptr(), c.bias.data_ptr()) for c in convs)
        if H.versions[dev] != ver:
            for i, c in enumerate(convs):
                _lib.require_cuda(c.weight, "encoder weights")
                w, b = _f32c(c.weight.detach()), _f32c(c.bias.detach())
                check(lib.aa_encoder_set_weights(h, i, ptr(w), ptr(b), stream_ptr()))
            H.versions[dev] = ver
            ev = torch.cuda.Event()
            ev.record()
            H.w_event[dev] = (ev, torch.cuda.current_stream().cuda_stream)
        elif dev in H.w_event and H.w_event[dev][1] != torch.cuda.current_stream().cuda_stream:
            torch.cuda.current_stream().wait_event(H.w_event[dev][0])   # weights were uploaded on another stream
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
        if self.in_channels not in (1, 2):   # PQMF front-end (2 * bands input channels): the tensor-core first-layer kernels are
            dt = DTYPES["fp32_cuda_cores"]   # built for mono / stereo input; the generic fp32 kernel takes any layer table
        with torch.cuda.device(dev):
            h = self._handle(dev)
            y = torch.empty((b, self.latent_dim, self.out_length(n)), dtype=torch.float32, device=x0.device)
            if b == 0:
                return y
            nbytes = int(lib.aa_encoder_workspace_bytes(h, b, n, dt))
            wkey = (dev, torch.cuda.current_stream().cuda_stream)   # callers on different streams must not share activations
            ws = self._handles.ws.get(wkey)
            if ws is None or ws.numel() < nbytes:
                ws = torch.empty(nbytes, dtype=torch.uint8, device=x0.device)
                self._handles.ws[wkey] = ws
            pa, keep = _ptr_array(stems)
            fa = (C.c_float * len(stems))(*fl)
            check(lib.aa_encoder_forward(h, pa, fa, len(stems), b, n, int(apply_tanh), dt, ptr(y), ptr(ws), stream_ptr()))
        return y

    def forward(self, x):
        squeeze = x.dim() == 2
        y = self.encode_mix([x.unsqueeze(0) if squeeze else x])
        return y.squeeze(0) if squeeze else y


# ---------------------------------------------------------------------------------------------------
# PQMF analysis front-end (restated `diffusion.pqmf.PQMF`, the RAVE pseudo-QMF bank; unpinned upstream)
# ---------------------------------------------------------------------------------------------------

def _kaiser_filter(wc, atten, n=None):
    import numpy as np
    from scipy.signal import firwin, kaiserord
    n_, beta = kaiserord(atten, wc / np.pi)
    n_ = 2 * (n_ // 2) + 1
    n = n if n is not None else n_
    return firwin(n, wc, window=("kaiser", beta), scale=False, fs=2 * np.pi)


def _pqmf_prototype(atten, m, n=None):
    "low-pass prototype whose cutoff minimises the reconstruction error of the m-band cosine-modulated bank"
    import numpy as np
    from scipy.optimize import fmin

    def loss(wc):
        h = _kaiser_filter(float(np.atleast_1d(wc)[0]), atten, n)
        g = np.convolve(h, h[::-1])
        g = abs(g[g.shape[-1] // 2::2 * m][1:])
        return np.max(g)

    wc = fmin(loss, 1 / m, disp=0)[0]
    return _kaiser_filter(wc, atten, n)


def _center_pad_next_pow_2(x):
    import numpy as np
    nxt = 2 ** math.ceil(math.log2(x.shape[-1]))
    pad = nxt - x.shape[-1]
    return np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad // 2, pad // 2 + int(pad % 2))])


def pqmf_filterbank(attenuation, n_band):
    "hk [n_band][taps] float64: 2 h cos((2k+1) pi/(2M) t + (-1)^k pi/4), taps padded to a power of two (as the polyphase form wants)"
    import numpy as np
    h = _pqmf_prototype(attenuation, n_band)          # odd length, symmetric about its centre tap
    k = np.arange(n_band).reshape(-1, 1)
    n = h.shape[-1]
    t = np.arange(-(n // 2), n // 2 + 1)
    p = (-1.0) ** k * math.pi / 4
    hk = 2 * h * np.cos((2 * k + 1) * math.pi / (2 * n_band) * t + p)
    return _center_pad_next_pow_2(hk)


class PQMF(nn.Module):
    """PQMF(channels, attenuation, n_band).forward: [B, channels, N] -> [B, channels * n_band, N / n_band]; every channel is
    analysed on its own ('b c t -> (b c) 1 t', filterbank, '(b c) k t -> b (c k) t').  One aa_pqmf_analysis_f32 launch."""

    def __init__(self, channels, attenuation, n_band):
        super().__init__()
        self.channels, self.attenuation, self.n_band = channels, attenuation, n_band
        hk = torch.from_numpy(pqmf_filterbank(attenuation, n_band)).float() if n_band > 1 else torch.ones(1, 1)
        self.register_buffer("hk", hk)

    def forward(self, x):
        if self.n_band == 1:
            return x
        x = _f32c(x, "waveform")
        assert x.dim() == 3 and x.shape[1] == self.channels, f"expected [B,{self.channels},N], got {tuple(x.shape)}"
        b, c, n = x.shape
        taps = self.hk.shape[1]
        t_out = (n + 2 * (taps // 2) - taps) // self.n_band
        hk = self.hk.to(x.device)
        out = torch.empty((b, c * self.n_band, t_out), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.aa_pqmf_analysis_f32(ptr(x), b * c, n, ptr(hk), self.n_band, taps, ptr(out), stream_ptr()))
        return out


# ---------------------------------------------------------------------------------------------------
# Memcodes quantiser (restated `nwt_pytorch.Memcodes` / `dvae.residual_memcodes.ResidualMemcodes`; unpinned upstream)
# ---------------------------------------------------------------------------------------------------

class Memcodes(nn.Module):
    """Multi-head codebook lookup: queries = head slices of the input (scaled by (dim/heads)^-0.5), keys / values = grouped 1x1
    convolutions of the learned codes; eval mode picks argmax of the logits and returns that code's value.
    forward takes the reference's [B, N, dim] layout; `quantize_cf` is the channel-major entry encode_it uses (no rearrange)."""

    def __init__(self, *, dim, num_codes, heads=8, temperature=1.):
        super().__init__()
        assert dim % heads == 0, 'dimension must be divisible by number of heads'
        self.heads, self.dim, self.num_codes, self.temperature = heads, dim, num_codes, temperature
        self.scale = (dim // heads) ** -0.5
        self.codes = nn.Parameter(torch.randn(heads, num_codes, dim // heads))
        self.to_k = nn.Conv1d(dim, dim, 1, groups=heads, bias=False)
        self.to_v = nn.Conv1d(dim, dim, 1, groups=heads, bias=False)

    def get_codes(self):
        "k, v [heads, num_codes, dim/heads]"
        d = self.dim // self.heads
        codes = _f32c(self.codes.detach(), "codes")
        wk, wv = _f32c(self.to_k.weight.detach().reshape(self.dim, d)), _f32c(self.to_v.weight.detach().reshape(self.dim, d))
        k, v = torch.empty_like(codes), torch.empty_like(codes)
        with torch.cuda.device(codes.device):
            check(lib.aa_memcodes_kv_f32(ptr(codes), ptr(wk), ptr(wv), self.heads, self.num_codes, d, ptr(k), ptr(v), stream_ptr()))
        return k, v

    def quantize_cf(self, x, resid_out=None, acc=None, final_tanh=False, want_indices=True):
        "x [B, dim, N] channel-major -> (quantized [B, dim, N], indices [B, heads, N])"
        if self.training:
            raise NotImplementedError("Memcodes: the gumbel-softmax sampling path (training mode) is not built; call .eval()")
        x = _f32c(x, "embeddings")
        assert x.dim() == 3 and x.shape[1] == self.dim
        b, _, n = x.shape
        k, v = self.get_codes()
        q = torch.empty_like(x)
        idx = torch.empty((b, self.heads, n), dtype=torch.int64, device=x.device) if want_indices else None
        with torch.cuda.device(x.device):
            check(lib.aa_memcodes_quantize_f32(ptr(x), b, self.heads, self.dim // self.heads, n, ptr(k), ptr(v), self.num_codes,
                                               float(self.scale), ptr(q), None if resid_out is None else ptr(resid_out),
                                               None if acc is None else ptr(acc), int(final_tanh), None if idx is None else ptr(idx),
                                               stream_ptr()))
        return q, idx

    def forward(self, x, *, merge_output_heads=True):
        assert x.shape[-1] == self.dim
        q, idx = self.quantize_cf(x.transpose(1, 2))
        out = q.transpose(1, 2)
        if not merge_output_heads:
            out = out.reshape(out.shape[0], out.shape[1], self.heads, -1).permute(0, 2, 1, 3)
        return out, idx


class ResidualMemcodes(nn.Module):
    "residual VQ over Memcodes layers: each layer quantises what the previous ones left; outputs are summed"

    def __init__(self, *, num_quantizers, **kwargs):
        super().__init__()
        self.layers = nn.ModuleList([Memcodes(**kwargs) for _ in range(num_quantizers)])

    def quantize_cf(self, x, final_tanh=False):
        x = _f32c(x, "embeddings")
        acc = torch.zeros_like(x)
        residual, all_idx = x, []
        for i, layer in enumerate(self.layers):
            nxt = torch.empty_like(x)
            _, idx = layer.quantize_cf(residual, resid_out=nxt, acc=acc, final_tanh=final_tanh and i == len(self.layers) - 1)
            residual = nxt
            all_idx.append(idx)
        return acc, torch.stack(all_idx)

    def forward(self, x):
        q, idx = self.quantize_cf(x.transpose(1, 2))
        return q.transpose(1, 2), idx


def load_dvae_encoder_checkpoint(model, ckpt_path, map_location="cpu"):
    """Loads the encoder / encoder_ema weights of the reference's Lightning checkpoint (given_models.py:340-356 does
    `self.model.load_state_dict(ckpt['state_dict'])` on the whole DVAE) into the restated encoder.  The upstream encoder is an
    nn.Sequential, so its keys look like `encoder.layers.<i>....{weight,bias}` (or weight_g / weight_v under weight_norm) in
    layer order; they are mapped onto flat_convs() IN ORDER with strict count and shape checks -- any mismatch raises, which is
    how a wrong guess at the upstream architecture would surface.  Real-checkpoint parity is UNTESTED:
    the 4 GB checkpoint is not reachable from this build."""
    ckpt = torch.load(ckpt_path, map_location=map_location, weights_only=False)
    sd = ckpt.get("state_dict", ckpt)
    loaded = {}
    for prefix, target in (("encoder.", model.encoder), ("encoder_ema.", model.encoder_ema)):
        keys = [k for k in sd if k.startswith(prefix)]
        convs = []   # (weight, bias) in checkpoint order
        names = sorted({k.rsplit(".", 1)[0] for k in keys}, key=lambda n: [int(t) if t.isdigit() else t for t in n.split(".")])
        for nme in names:
            if nme + ".weight_g" in sd and nme + ".weight_v" in sd:     # fold weight_norm: w = g * v / ||v|| (norm over all dims but 0)
                g, v = sd[nme + ".weight_g"].float(), sd[nme + ".weight_v"].float()
                w = v * (g / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1))))
            elif nme + ".weight" in sd:
                w = sd[nme + ".weight"].float()
            else:
                continue
            if w.dim() != 3:
                continue
            convs.append((nme, w, sd.get(nme + ".bias")))
        mine = target.flat_convs()
        if len(convs) != len(mine):
            raise RuntimeError(f"{prefix}* holds {len(convs)} Conv1d layers, the restated encoder has {len(mine)}: the checkpoint's "
                               "architecture differs from SURVEY.md Appendix A")
        with torch.no_grad():
            for conv, (nme, w, b) in zip(mine, convs):
                if tuple(w.shape) != tuple(conv.weight.shape):
                    raise RuntimeError(f"{nme}: weight {tuple(w.shape)} does not fit {tuple(conv.weight.shape)}")
                if b is None or tuple(b.shape) != tuple(conv.bias.shape):
                    raise RuntimeError(f"{nme}: bias missing or of the wrong shape")
                conv.weight.copy_(w.to(conv.weight.device))
                conv.bias.copy_(b.float().to(conv.bias.device))
        loaded[prefix] = len(convs)
    return loaded


class DiffusionDVAE(nn.Module):
    """global_args needs: pqmf_bands, latent_dim, num_quantizers (ema_decay etc. are accepted and ignored).
    Same members as the reference on the encode side: encoder, encoder_ema (deepcopy), pqmf_bands, quantized."""

    def __init__(self, global_args):
        super().__init__()
        self.pqmf_bands = global_args.pqmf_bands
        if self.pqmf_bands > 1:
            self.pqmf = PQMF(2, 70, global_args.pqmf_bands)
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
            quantizer_class = ResidualMemcodes if self.num_quantizers > 1 else Memcodes
            quantizer_kwargs = {}
            if self.num_quantizers > 1:
                quantizer_kwargs["num_quantizers"] = self.num_quantizers
            self.quantizer = quantizer_class(dim=global_args.latent_dim, heads=global_args.num_heads, num_codes=global_args.codebook_size,
                                             temperature=1., **quantizer_kwargs)
            self.quantizer_ema = deepcopy(self.quantizer)
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
        """aa_mixer.py:175-195: (pqmf) -> encoder_ema -> (Memcodes) -> tanh; runs under no_grad like the reference.  Unbatched [C, N]
        input is accepted like the reference's Conv1d stack accepts it."""
        squeeze = demo_reals.dim() == 2
        encoder_input = (demo_reals.unsqueeze(0) if squeeze else demo_reals).to(self.device)
        self.demo_reals_shape = demo_reals.shape
        with torch.no_grad():
            if self.pqmf_bands > 1:
                encoder_input = self.pqmf(encoder_input)
            if not self.quantized:
                emb = self.encoder_ema.encode_mix([encoder_input], None, apply_tanh=True)
            else:
                emb = self.encoder_ema.encode_mix([encoder_input], None, apply_tanh=False)
                if isinstance(self.quantizer_ema, ResidualMemcodes):
                    emb, _ = self.quantizer_ema.quantize_cf(emb, final_tanh=True)
                else:
                    acc = torch.zeros_like(emb)
                    self.quantizer_ema.quantize_cf(emb, acc=acc, final_tanh=True, want_indices=False)
                    emb = acc
        return emb.squeeze(0) if squeeze else emb

    def decode_it(self, *args, **kwargs):
        raise NotImplementedError("the diffusion decoder is outside the accelerated hot path (encode side only)")
