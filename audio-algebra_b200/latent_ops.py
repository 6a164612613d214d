"""Latent-space manipulations of the Destructo notebook (Destructo.ipynb cells 22, 48-49) as single-pass
CUDA kernels (the notebook runs one full-tensor ATen kernel per arithmetic op)."""
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .aa_mixer import _f32c, _ws, _RED_WS

SIGN_FOLD, ABSMAX_MINUS, TANH_DRIVE, FLIP_CHANNELS, FLIP_TIME = 0, 1, 2, 3, 4

__all__ = ['flip_channels', 'flip_time', 'sign_fold', 'absmax_minus', 'tanh_drive', 'effect_transfer']


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


def effect_transfer(embeddings, wet_emb, dry_emb):
    "cells 48-49: diff = (wet_emb - dry_emb).mean(0); z = embeddings + diff"
    e, w, d = _f32c(embeddings), _f32c(wet_emb), _f32c(dry_emb)
    assert w.shape == d.shape and e.shape[1:] == w.shape[1:]
    out = torch.empty_like(e)
    ct = e[0].numel()
    with torch.cuda.device(e.device):
        check(lib.aa_effect_transfer_f32(ptr(e), e.shape[0], ptr(w), ptr(d), w.shape[0], ct, ptr(out), stream_ptr()))
    return out
