"""Latent-space manipulations of the Destructo notebook (Destructo.ipynb cells 22, 48-49) as single-pass
CUDA kernels (the notebook runs one full-tensor ATen kernel per arithmetic op)."""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .aa_mixer import _f32c, _ws, _RED_WS

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_lib.register({
    "aa_latent_mul_time_f32": (_i, [_p, _p, _p, _i64, _i64, _i64, _p]),
    "aa_latent_add_flip_time_f32": (_i, [_p, _p, _i64, _i64, _i64, _p]),
    "aa_latent_zero_channels_f32": (_i, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p]),
    "aa_latent_randmix_f32": (_i, [_p, _p, _p, _f, _f, _p, _i64, _p]),
    "aa_latent_reverb_f32": (_i, [_p, _p, _p, _i64, _i64, _p]),
    "aa_effect_transfer_ex_f32": (_i, [_p, _i64, _i64, _i64, _p, _p, _i64, _i64, _p, _p]),
    "aa_latent_row_mean_diff_f32": (_i, [_p, _p, _i64, _i64, _p, _p]),
    "aa_latent_add_bcast2_f32": (_i, [_p, _i64, _i64, _i64, _p, _i64, _i64, _p, _p]),
})

SIGN_FOLD, ABSMAX_MINUS, TANH_DRIVE, FLIP_CHANNELS, FLIP_TIME = 0, 1, 2, 3, 4

__all__ = ['flip_channels', 'flip_time', 'sign_fold', 'absmax_minus', 'tanh_drive', 'effect_transfer', 'big_changes', 'wavy', 'flippy',
           'kill_half', 'reverb_time', 'call_and_response', 'hurt_drums']


def _unary(op, z, param=0.0):
    z = _f32c(z, "z")
    assert z.dim() == 3, "latents are [B, C, T]"
    out = torch.empty_like(z)
    with torch.cuda.device(z.device):
        check(lib.aa_latent_unary_f32(op, ptr(z), ptr(out), z.shape[0], z.shape[1], z.shape[2], float(param),
                                      ptr(_ws(_RED_WS, z.device)), stream_ptr()))
    return out


def flip_channels(z):
    "z.flip(dims=[1])"
    return _unary(FLIP_CHANNELS, z)


def flip_time(z):
    "z.flip(dims=[2])"
    return _unary(FLIP_TIME, z)


def sign_fold(z):
    "torch.max(z) * (torch.sign(z) - z)"
    return _unary(SIGN_FOLD, z)


def absmax_minus(z):
    "torch.max(torch.abs(z)) - z"
    return _unary(ABSMAX_MINUS, z)


def tanh_drive(z, k):
    "torch.max(z) * torch.tanh(k * z)"
    return _unary(TANH_DRIVE, z, k)


def big_changes(z):
    "z = 2*z"
    from .aa_mixer import latent_lincomb
    return latent_lincomb([z], [2.0])


def wavy(z):
    "z * torch.cos(torch.linspace(0, 4*6.28, z.shape[-1])).to(device): the modulation is built on the CPU exactly as the notebook builds it"
    z = _f32c(z, "z")
    assert z.dim() == 3, "latents are [B, C, T]"
    vec = torch.cos(torch.linspace(0, 4 * 6.28, z.shape[-1])).to(z.device)
    out = torch.empty_like(z)
    with torch.cuda.device(z.device):
        check(lib.aa_latent_mul_time_f32(ptr(z), ptr(vec), ptr(out), z.shape[0], z.shape[1], z.shape[2], stream_ptr()))
    return out


def flippy(z):
    "z.clone() + torch.flip(z, [-1])"
    z = _f32c(z, "z")
    assert z.dim() == 3, "latents are [B, C, T]"
    out = torch.empty_like(z)
    with torch.cuda.device(z.device):
        check(lib.aa_latent_add_flip_time_f32(ptr(z), ptr(out), z.shape[0], z.shape[1], z.shape[2], stream_ptr()))
    return out


def kill_half(z, start=33, stop=-1):
    "z[:, 33:-1, :] = 0. (returns a new tensor; python slice semantics for start / stop)"
    z = _f32c(z, "z")
    assert z.dim() == 3, "latents are [B, C, T]"
    c0, c1, _ = slice(start, stop).indices(z.shape[1])
    out = torch.empty_like(z)
    with torch.cuda.device(z.device):
        check(lib.aa_latent_zero_channels_f32(ptr(z), ptr(out), z.shape[0], z.shape[1], z.shape[2], c0, max(c1, c0), stream_ptr()))
    return out


def reverb_time(z, reverb_time):
    """for i in range(T): z = z + math.exp(-i/reverb_time) * F.pad(z, (i+1, 0, 0, 0, 0, 0))[:, :, :T]  -- the notebook's T full-tensor
    passes as one kernel (a block per (b, c) row, the row stays in shared memory)."""
    z = _f32c(z, "z")
    assert z.dim() == 3, "latents are [B, C, T]"
    if reverb_time == 0:
        return z.clone()
    t = z.shape[-1]
    coef = torch.tensor([math.exp(-i / reverb_time) for i in range(t)], dtype=torch.float64).to(torch.float32).to(z.device)
    out = torch.empty_like(z)
    with torch.cuda.device(z.device):
        check(lib.aa_latent_reverb_f32(ptr(z), ptr(coef), ptr(out), z.shape[0] * z.shape[1], t, stream_ptr()))
    return out


def _randmix(x, z, a, b, u=None, generator=None):
    x, z = _f32c(x, "x"), _f32c(z, "z")
    assert x.shape == z.shape
    if u is None:   # torch.rand_like on the tensor's device: the RNG the notebook itself draws from
        u = torch.rand(z.shape, dtype=z.dtype, device=z.device, generator=generator) if generator is not None else torch.rand_like(z)
    u = _f32c(u, "u")
    assert u.shape == z.shape
    out = torch.empty_like(z)
    with torch.cuda.device(z.device):
        check(lib.aa_latent_randmix_f32(ptr(x), ptr(z), ptr(u), float(a), float(b), ptr(out), z.numel(), stream_ptr()))
    return out


def call_and_response(z, rand_fac=0.5, u=None, generator=None):
    "z = -z + rand_fac*z*(2*torch.rand_like(z)-1)"
    return _randmix(z, z, -1.0, rand_fac, u, generator)


def hurt_drums(embeddings, z, rand_fac=0.5, u=None, generator=None):
    "z = (1-rand_fac)*embeddings + rand_fac*z*(2*torch.rand_like(z)-1)"
    return _randmix(embeddings, z, 1.0 - rand_fac, rand_fac, u, generator)


def effect_transfer(embeddings, wet_emb, dry_emb, time_avg=False):
    """Destructo.ipynb cells 48-49: diff = wet_emb - dry_emb; z = embeddings + (diff padded / truncated to the embedding length).mean(0),
    or with time_avg=True: z = embeddings + diff.mean(-1) under torch's broadcasting rules (which accept only shapes where the
    [Bw, C] mean lines up with the trailing [C, T] axes -- otherwise the same RuntimeError the notebook raises)."""
    e, w, d = _f32c(embeddings), _f32c(wet_emb), _f32c(dry_emb)
    assert e.dim() == 3 and w.dim() == 3 and w.shape == d.shape and (time_avg or e.shape[1] == w.shape[1])
    out = torch.empty_like(e)
    with torch.cuda.device(e.device):
        if time_avg:
            bw, c, td = w.shape
            torch.broadcast_shapes(tuple(e.shape), (bw, c))      # raises like `z + diff` would
            dm = torch.empty((bw, c), dtype=torch.float32, device=e.device)
            check(lib.aa_latent_row_mean_diff_f32(ptr(w), ptr(d), bw * c, td, ptr(dm), stream_ptr()))
            check(lib.aa_latent_add_bcast2_f32(ptr(e), e.shape[0], e.shape[1], e.shape[2], ptr(dm), bw, c, ptr(out), stream_ptr()))
        elif e.shape[2] == w.shape[2]:
            ct = e[0].numel()
            check(lib.aa_effect_transfer_f32(ptr(e), e.shape[0], ptr(w), ptr(d), w.shape[0], ct, ptr(out), stream_ptr()))
        else:
            check(lib.aa_effect_transfer_ex_f32(ptr(e), e.shape[0], e.shape[1], e.shape[2], ptr(w), ptr(d), w.shape[0], w.shape[2],
                                                ptr(out), stream_ptr()))
    return out
