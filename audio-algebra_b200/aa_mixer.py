"""AA core of the reference's audio_algebra/aa_mixer.py with the same names and call contracts:
EmbedBlock, AudioAlgebra, get_stems_faders, do_mixing, mseloss, vicreg_var_loss, vicreg_cov_loss,
off_diagonal -- computed by libaa_b200 kernels (fused projector halves, Gram-identity covariance loss,
deterministic two-stage reductions) with torch.autograd.Function wrappers for training.

Mirrors /root/reference/audio_algebra/aa_mixer.py:205-364 (and the copies in aa_effects.py:51-162,
train_aa_mixer_accel.py:274-460).  Tensors must live on a B200; there is no CPU path.
"""
import ctypes as C
import random

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib
from ._lib import lib, check, ptr, stream_ptr

__all__ = ['EmbedBlock', 'AudioAlgebra', 'get_stems_faders', 'do_mixing', 'mseloss', 'vicreg_var_loss',
           'vicreg_cov_loss', 'off_diagonal', 'latent_lincomb', 'mixer_loss_fused', 'effects_loss_fused']

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_pp = C.POINTER(C.c_void_p)
_lib.register({
    "aa_latent_lincomb_f32": (_i, [_i, _pp, C.POINTER(_f), _p, _i64, _p]),
    "aa_latent_unary_f32": (_i, [_i, _p, _p, _i64, _i64, _i64, _f, _p, _p]),
    "aa_effect_transfer_f32": (_i, [_p, _i64, _p, _p, _i64, _i64, _p, _p]),
    "aa_reduce_workspace_floats": (_i64, []),
    "aa_mse_fwd_f32": (_i, [_p, _p, _i64, _p, _p, _p]),
    "aa_mse_bwd_f32": (_i, [_p, _p, _i64, _p, _f, _p, _p, _i, _p]),
    "aa_vicreg_var_fwd_f32": (_i, [_p, _i64, _i64, _f, _f, _i, _p, _p, _p, _p]),
    "aa_vicreg_var_bwd_f32": (_i, [_p, _p, _i64, _i64, _f, _f, _i, _p, _f, _p, _i, _p]),
    "aa_cov_loss_workspace_floats": (_i64, [_i64, _i64]),
    "aa_vicreg_cov_fwd_f32": (_i, [_p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "aa_vicreg_cov_bwd_f32": (_i, [_p, _p, _p, _i64, _i64, _p, _f, _p, _i, _p]),
    "aa_projector_half_fwd_f32": (_i, [_pp, _pp, _i, _i, _i, _p, _i64, _i64, _p, _p]),
    "aa_projector_bwd_workspace_floats": (_i64, []),
    "aa_projector_half_bwd_f32": (_i, [_pp, _pp, _i, _i, _i, _p, _p, _i64, _i64, _p, _i, _pp, _pp, _i, _f, _p, _p]),
    "aa_cov_workspace_floats": (_i64, [_i64]),
    "aa_cov_accumulate_f32": (_i, [_p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "aa_adam_step_f32": (_i, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _i64, _p]),
    "aa_embed_block_fwd_f32": (_i, [_p, _p, _p, _i64, _i, _i, _i, _i, _p, _p, _p]),
    "aa_embed_block_bwd_f32": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "aa_batchnorm_fwd_f32": (_i, [_p, _i64, _i, _p, _p, _p, _p, _i, _f, _f, _p, _p, _p]),
    "aa_batchnorm_bwd_f32": (_i, [_p, _p, _i64, _i, _p, _p, _i, _p, _p, _p, _p]),
    "aa_mixer_loss_saved_floats": (_i64, [_i64, _i64]),
    "aa_effects_loss_saved_floats": (_i64, [_i64, _i64]),
    "aa_fused_loss_workspace_floats": (_i64, [_i64, _i64]),
    "aa_mixer_loss_fwd_f32": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i, _f, _f, _p, _p, _p, _p]),
    "aa_mixer_loss_bwd_f32": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i, _f, _f, _p, _p, _p, _p, _p, _p, _p]),
    "aa_effects_loss_fwd_f32": (_i, [_pp, _pp, _pp, _i64, _i64, _i64, _f, _f, _p, _p, _p, _p]),
    "aa_effects_loss_bwd_f32": (_i, [_pp, _pp, _pp, _i64, _i64, _i64, _f, _f, _p, _p, _pp, _pp, _p]),
})

_RED_WS = int(lib.aa_reduce_workspace_floats())


def _f32c(t: Tensor, what="tensor") -> Tensor:
    _lib.require_cuda(t, what)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])
    return C.cast(arr, _pp), arr  # keep `arr` alive while the call runs


def _ws(n, device):
    return torch.empty(int(n), dtype=torch.float32, device=device)


# ---------------------------------------------------------------------------------------------------
# latent algebra
# ---------------------------------------------------------------------------------------------------

def latent_lincomb(zs, coeffs, out=None) -> Tensor:
    "sum_j coeffs[j] * zs[j] in one pass (zsum, effect guesses zb2 - zb1 + za1, z + diff, ...)"
    zs = [_f32c(z) for z in zs]
    assert len(zs) >= 1 and len(zs) == len(coeffs)
    n = zs[0].numel()
    assert all(z.numel() == n for z in zs)
    if out is None:
        out = torch.empty_like(zs[0])
    cs = [float(c) for c in coeffs]
    with torch.cuda.device(zs[0].device):
        first = True
        while zs:   # the kernel takes up to 8 terms per pass; more stems (the reference allows any maxstems) accumulate into `out`
            take = 8 if first else 7
            terms, tc = (zs[:take], cs[:take]) if first else ([out] + zs[:take], [1.0] + cs[:take])
            zs, cs = zs[take:], cs[take:]
            pa, keep = _ptr_array(terms)
            ca = (C.c_float * len(terms))(*tc)
            check(lib.aa_latent_lincomb_f32(len(terms), pa, ca, ptr(out), n, stream_ptr()))
            first = False
    return out


class _LinComb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coeffs, *zs):
        ctx.coeffs = coeffs
        return latent_lincomb([z.detach() for z in zs], coeffs)

    @staticmethod
    def backward(ctx, g):
        return (None,) + tuple(g * c if c != 1.0 else g for c in ctx.coeffs)


def _lincomb_ad(zs, coeffs):
    return _LinComb.apply(tuple(float(c) for c in coeffs), *zs)


# ---------------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------------

class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _f32c(a, "mseloss input"), _f32c(b, "mseloss target")
        assert a.shape == b.shape, f"mseloss: shapes differ {a.shape} vs {b.shape}"
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            check(lib.aa_mse_fwd_f32(ptr(a), ptr(b), a.numel(), ptr(loss), ptr(_ws(_RED_WS, a.device)), stream_ptr()))
        ctx.save_for_backward(a, b)
        return loss

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.contiguous().float()
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(a.device):
            check(lib.aa_mse_bwd_f32(ptr(a), ptr(b), a.numel(), ptr(g), 1.0,
                                     None if ga is None else ptr(ga), None if gb is None else ptr(gb), 0, stream_ptr()))
        return ga, gb


def mseloss(a: Tensor, b: Tensor) -> Tensor:
    "nn.MSELoss() (aa_mixer.py:344): mean((a-b)^2)"
    return _MSE.apply(a, b)


class _VarLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, gamma, eps, l2):
        z = _f32c(z, "z")
        b, d = z.shape[0], z.numel() // z.shape[0]
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        stats = torch.empty((2, d), dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            check(lib.aa_vicreg_var_fwd_f32(ptr(z), b, d, float(gamma), float(eps), int(l2), ptr(loss), ptr(stats),
                                            ptr(_ws(max(_RED_WS, (d + 255) // 256), z.device)), stream_ptr()))
        ctx.save_for_backward(z, stats)
        ctx.cfg = (b, d, float(gamma), float(eps), int(l2))
        return loss

    @staticmethod
    def backward(ctx, g):
        z, stats = ctx.saved_tensors
        b, d, gamma, eps, l2 = ctx.cfg
        gz = torch.empty_like(z)
        g = g.contiguous().float()
        with torch.cuda.device(z.device):
            check(lib.aa_vicreg_var_bwd_f32(ptr(z), ptr(stats), b, d, gamma, eps, l2, ptr(g), 1.0, ptr(gz), 0, stream_ptr()))
        return gz, None, None, None


def vicreg_var_loss(z: Tensor, gamma=1, eps=1e-4) -> Tensor:
    "aa_mixer.py:351-353: mean(relu(gamma - sqrt(z.var(dim=0) + eps)))"
    return _VarLoss.apply(z, gamma, eps, False)


def vicreg_var_loss_l2(z: Tensor, gamma=1, eps=1e-4) -> Tensor:
    "train_aa_effects.py:42-44: mean(relu(gamma - std)**2)"
    return _VarLoss.apply(z, gamma, eps, True)


def off_diagonal(x: Tensor) -> Tensor:
    "aa_mixer.py:355-358 (kept for API compatibility; the CUDA covariance loss never builds the matrix)"
    n, m = x.shape
    assert n == m
    return x.flatten()[:-1].view(n - 1, n + 1)[:, 1:].flatten()


class _CovLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        z = _f32c(z, "z")
        b, d = z.shape[0], z.numel() // z.shape[0]
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        stats = torch.empty((2, d), dtype=torch.float32, device=z.device)
        gram = torch.empty((b, b), dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            ws = _ws(lib.aa_cov_loss_workspace_floats(b, d), z.device)
            check(lib.aa_vicreg_cov_fwd_f32(ptr(z), b, d, None, ptr(stats), ptr(gram), ptr(loss), ptr(ws), stream_ptr()))
        ctx.save_for_backward(z, stats, gram)
        return loss

    @staticmethod
    def backward(ctx, g):
        z, stats, gram = ctx.saved_tensors
        b, d = z.shape[0], z.numel() // z.shape[0]
        gz = torch.empty_like(z)
        g = g.contiguous().float()
        with torch.cuda.device(z.device):
            check(lib.aa_vicreg_cov_bwd_f32(ptr(z), ptr(stats), ptr(gram), b, d, ptr(g), 1.0, ptr(gz), 0, stream_ptr()))
        return gz


def _as_bct(t: Tensor, what):
    t = _f32c(t, what)
    assert t.dim() >= 2, f"{what}: expected [B, ...], got {tuple(t.shape)}"
    return t


class _MixerLossFused(torch.autograd.Function):
    """aa_mixer_loss_fwd/bwd_f32: every term of train_aa_mixer_accel.py:504-517 in one C call each way."""

    @staticmethod
    def forward(ctx, zsum, zmix, y, yrecon, ymix, ymix_recon, hinge_l2, gamma, eps):
        ts = [_as_bct(v, n) for v, n in ((zsum, "zsum"), (zmix, "zmix"), (y, "y"), (yrecon, "yrecon"), (ymix, "ymix"), (ymix_recon, "ymix_recon"))]
        assert all(v.shape == ts[0].shape for v in ts), "mixer loss: all six tensors must have one shape"
        b, d = ts[0].shape[0], ts[0].numel() // ts[0].shape[0]
        dev = ts[0].device
        losses = torch.empty(5, dtype=torch.float32, device=dev)
        saved = torch.empty(int(lib.aa_mixer_loss_saved_floats(b, d)), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = _ws(int(lib.aa_fused_loss_workspace_floats(b, d)), dev)
            check(lib.aa_mixer_loss_fwd_f32(*[ptr(v) for v in ts], b, 1, d, int(hinge_l2), float(gamma), float(eps), ptr(losses), ptr(saved),
                                            ptr(ws), stream_ptr()))
        ctx.save_for_backward(*ts, saved)
        ctx.cfg = (b, d, int(hinge_l2), float(gamma), float(eps))
        terms = losses[1:].clone()
        ctx.mark_non_differentiable(terms)
        return losses[0], terms

    @staticmethod
    def backward(ctx, g, _g_terms):
        *ts, saved = ctx.saved_tensors
        b, d, l2, gamma, eps = ctx.cfg
        g = g.contiguous().float()
        outs = [torch.empty_like(ts[0]) for _ in range(4)]
        with torch.cuda.device(ts[0].device):
            check(lib.aa_mixer_loss_bwd_f32(*[ptr(v) for v in ts], b, 1, d, l2, gamma, eps, ptr(saved), ptr(g), *[ptr(o) for o in outs],
                                            stream_ptr()))
        return outs[0], outs[1], None, outs[2], None, outs[3], None, None, None


def mixer_loss_fused(zsum, zmix, y, yrecon, ymix, ymix_recon, hinge_l2=False, gamma=1.0, eps=1e-4):
    """All loss terms of the mixer training step (train_aa_mixer_accel.py:504-517) through the fused C entry points:
    returns the dict {'loss', 'mix_loss', 'var_loss', 'cov_loss', 'aa_recon_loss'}; only 'loss' carries gradient
    (to zsum, zmix, yrecon, ymix_recon -- y and ymix come from the frozen given model)."""
    loss, terms = _MixerLossFused.apply(zsum, zmix, y, yrecon, ymix, ymix_recon, hinge_l2, gamma, eps)
    return {'loss': loss, 'mix_loss': terms[0], 'var_loss': terms[1], 'cov_loss': terms[2], 'aa_recon_loss': terms[3]}


class _EffectsLossFused(torch.autograd.Function):
    """aa_effects_loss_fwd/bwd_f32: every term of train_aa_effects.py:66-82 in one C call each way."""

    @staticmethod
    def forward(ctx, gamma, eps, *tensors):   # za1, zb1, za2, zb2, y x 4, yrecon x 4
        ts = [_as_bct(v, "effects loss input") for v in tensors]
        assert len(ts) == 12 and all(v.shape == ts[0].shape for v in ts)
        b, d = ts[0].shape[0], ts[0].numel() // ts[0].shape[0]
        dev = ts[0].device
        losses = torch.empty(5, dtype=torch.float32, device=dev)
        saved = torch.empty(int(lib.aa_effects_loss_saved_floats(b, d)), dtype=torch.float32, device=dev)
        zp, k1 = _ptr_array(ts[0:4]); yp, k2 = _ptr_array(ts[4:8]); rp, k3 = _ptr_array(ts[8:12])
        with torch.cuda.device(dev):
            ws = _ws(int(lib.aa_fused_loss_workspace_floats(b, d)), dev)
            check(lib.aa_effects_loss_fwd_f32(zp, yp, rp, b, 1, d, float(gamma), float(eps), ptr(losses), ptr(saved), ptr(ws), stream_ptr()))
        ctx.save_for_backward(*ts, saved)
        ctx.cfg = (b, d, float(gamma), float(eps))
        terms = losses[1:].clone()
        ctx.mark_non_differentiable(terms)
        return losses[0], terms

    @staticmethod
    def backward(ctx, g, _g_terms):
        *ts, saved = ctx.saved_tensors
        b, d, gamma, eps = ctx.cfg
        g = g.contiguous().float()
        gz = [torch.empty_like(ts[0]) for _ in range(4)]
        gr = [torch.empty_like(ts[0]) for _ in range(4)]
        zp, k1 = _ptr_array(ts[0:4]); yp, k2 = _ptr_array(ts[4:8]); rp, k3 = _ptr_array(ts[8:12])
        gzp, k4 = _ptr_array(gz); grp, k5 = _ptr_array(gr)
        with torch.cuda.device(ts[0].device):
            check(lib.aa_effects_loss_bwd_f32(zp, yp, rp, b, 1, d, gamma, eps, ptr(saved), ptr(g), gzp, grp, stream_ptr()))
        return (None, None) + tuple(gz) + (None,) * 4 + tuple(gr)


def effects_loss_fused(zs, ys, yrecons, gamma=1.0, eps=1e-4):
    """All loss terms of the effects training step (train_aa_effects.py:66-82, L2-hinge variance loss) through the fused C entry
    points; zs = [za1, zb1, za2, zb2], ys / yrecons likewise.  Only 'loss' carries gradient (to zs and yrecons)."""
    loss, terms = _EffectsLossFused.apply(gamma, eps, *zs, *ys, *yrecons)
    return {'loss': loss, 'mix_loss': terms[0], 'var_loss': terms[1], 'cov_loss': terms[2], 'aa_recon_loss': terms[3]}


def vicreg_cov_loss(z: Tensor) -> Tensor:
    """aa_mixer.py:360-364: sum of squared off-diagonal entries of cov(z as [(c t), b]) / (c t).
    Computed through the B x B Gram matrix (identical scalar, no (C T)^2 matrix)."""
    return _CovLoss.apply(z)


# ---------------------------------------------------------------------------------------------------
# projector
# ---------------------------------------------------------------------------------------------------

class _ProjHalf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, resid, dims, hidden, *wb):
        x = _f32c(x, "projector input")
        assert x.dim() == 3 and x.shape[1] == dims, f"expected [B,{dims},T], got {tuple(x.shape)}"
        ws, bs = [_f32c(w) for w in wb[:4]], [_f32c(b) for b in wb[4:]]
        out = torch.empty_like(x)
        wp, k1 = _ptr_array(ws)
        bp, k2 = _ptr_array(bs)
        with torch.cuda.device(x.device):
            check(lib.aa_projector_half_fwd_f32(wp, bp, dims, hidden, int(resid), ptr(x), x.shape[0], x.shape[2], ptr(out),
                                                stream_ptr()))
        ctx.save_for_backward(x, *ws, *bs)
        ctx.cfg = (bool(resid), dims, hidden)
        return out

    @staticmethod
    def backward(ctx, g):
        x, *wb = ctx.saved_tensors
        ws, bs = wb[:4], wb[4:]
        resid, dims, hidden = ctx.cfg
        g = _f32c(g)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gws, gbs = [torch.empty_like(w) for w in ws], [torch.empty_like(b) for b in bs]
        wp, k1 = _ptr_array(ws)
        bp, k2 = _ptr_array(bs)
        gwp, k3 = _ptr_array(gws)
        gbp, k4 = _ptr_array(gbs)
        with torch.cuda.device(x.device):
            wsb = _ws(lib.aa_projector_bwd_workspace_floats(), x.device)
            check(lib.aa_projector_half_bwd_f32(wp, bp, dims, hidden, int(resid), ptr(x), ptr(g), x.shape[0], x.shape[2],
                                                None if gx is None else ptr(gx), 0, gwp, gbp, 0, 1.0, ptr(wsb), stream_ptr()))
        return (gx, None, None, None, *gws, *gbs)


class _EmbedFn(torch.autograd.Function):
    "one EmbedBlock on [n_tok, din] rows: Linear (+ exact-erf GELU) (+ residual) -> aa_embed_block_{fwd,bwd}_f32"

    @staticmethod
    def forward(ctx, x, w, b, act, resid):
        x, w = _f32c(x, "EmbedBlock input"), _f32c(w)
        b = None if b is None else _f32c(b)
        n_tok, din, dout = x.shape[0], w.shape[1], w.shape[0]
        y = torch.empty((n_tok, dout), dtype=torch.float32, device=x.device)
        pre = torch.empty_like(y)
        with torch.cuda.device(x.device):
            check(lib.aa_embed_block_fwd_f32(ptr(x), ptr(w), None if b is None else ptr(b), n_tok, din, dout, int(act), int(resid),
                                             ptr(y), ptr(pre), stream_ptr()))
        ctx.save_for_backward(x, w, pre)
        ctx.cfg = (bool(act), bool(resid), b is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, pre = ctx.saved_tensors
        act, resid, has_b = ctx.cfg
        gy = _f32c(gy)
        n_tok, din, dout = x.shape[0], w.shape[1], w.shape[0]
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        need_w = ctx.needs_input_grad[1] or (has_b and ctx.needs_input_grad[2])
        gw = torch.empty_like(w) if need_w else None
        gb = torch.empty((dout,), dtype=torch.float32, device=x.device) if (need_w and has_b) else None
        with torch.cuda.device(x.device):
            check(lib.aa_embed_block_bwd_f32(ptr(x), ptr(w), ptr(pre), ptr(gy), n_tok, din, dout, int(act), int(resid),
                                             None if gx is None else ptr(gx), None if gw is None else ptr(gw),
                                             None if gb is None else ptr(gb), ptr(_ws(max(1, n_tok * dout), x.device)), stream_ptr()))
        return gx, gw, gb, None, None


class _BatchNormFn(torch.autograd.Function):
    "nn.BatchNorm1d on [N, C] (training: batch statistics + running-statistics update; eval: running statistics)"

    @staticmethod
    def forward(ctx, x, gamma, beta, run_mean, run_var, training, momentum, eps):
        x = _f32c(x, "BatchNorm input")
        n, c = x.shape
        y = torch.empty_like(x)
        save = torch.empty((2, c), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.aa_batchnorm_fwd_f32(ptr(x), n, c, None if gamma is None else ptr(gamma), None if beta is None else ptr(beta),
                                           None if run_mean is None else ptr(run_mean), None if run_var is None else ptr(run_var),
                                           int(training), float(momentum), float(eps), ptr(y), ptr(save), stream_ptr()))
        ctx.save_for_backward(x, gamma, save)
        ctx.training = bool(training)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, save = ctx.saved_tensors
        gy = _f32c(gy)
        n, c = x.shape
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gg = torch.empty((c,), dtype=torch.float32, device=x.device)
        gb = torch.empty((c,), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.aa_batchnorm_bwd_f32(ptr(x), ptr(gy), n, c, None if gamma is None else ptr(gamma), ptr(save), int(ctx.training),
                                           None if gx is None else ptr(gx), ptr(gg), ptr(gb), stream_ptr()))
        return gx, (gg if gamma is not None else None), (gb if gamma is not None else None), None, None, None, None, None


class EmbedBlock(nn.Module):
    """The reference's EmbedBlock (aa_mixer.py:205-221): `lin` = nn.Linear(in, out); y = lin(x); y = act(y); y = bn(y) (use_bn);
    x + y iff resid and in == out.  Callable on its own ([..., in_dims] rows -> aa_embed_block_fwd_f32, autograd included);
    inside AudioAlgebra four consecutive blocks run as ONE fused kernel per half instead.  act must be nn.GELU() (exact erf)
    or None.  use_bn: nn.BatchNorm1d(out_dims) applied to 2-D [N, out_dims] activations, the only rank for which the
    reference's placement of the layer is well formed (with [B, T, C] input BatchNorm1d would read T as the channel axis)."""

    def __init__(self, in_dims: int, out_dims: int, act=nn.GELU(), resid=True, use_bn=False, requires_grad=True, **kwargs) -> None:
        super().__init__()
        if act is not None and not (isinstance(act, nn.GELU) and getattr(act, "approximate", "none") == "none"):
            raise NotImplementedError("EmbedBlock implements act=nn.GELU() (exact erf) or act=None")
        self.in_dims, self.out_dims, self.act, self.resid = in_dims, out_dims, act, resid
        self.lin = nn.Linear(in_dims, out_dims, **kwargs)
        self.bn = nn.BatchNorm1d(out_dims) if use_bn else None   # parameter / running-statistics holder; arithmetic in _BatchNormFn
        if requires_grad == False:  # noqa: E712  (reference spelling)
            self.lin.weight.requires_grad = False
            self.lin.bias.requires_grad = False

    def forward(self, xin: Tensor) -> Tensor:
        _lib.require_cuda(xin, "EmbedBlock input")
        assert xin.shape[-1] == self.in_dims, f"expected [..., {self.in_dims}], got {tuple(xin.shape)}"
        add_resid = self.resid and self.in_dims == self.out_dims
        rows = xin.reshape(-1, self.in_dims)
        if self.bn is None:
            y = _EmbedFn.apply(rows, self.lin.weight, self.lin.bias, self.act is not None, add_resid)
            return y.reshape(*xin.shape[:-1], self.out_dims)
        if xin.dim() != 2:
            raise ValueError(f"EmbedBlock(use_bn=True): BatchNorm1d({self.out_dims}) needs [N, {self.out_dims}] activations, got input "
                             f"{tuple(xin.shape)} (the reference's BatchNorm1d would take axis 1 of a 3-D tensor as channels)")
        y = _EmbedFn.apply(rows, self.lin.weight, self.lin.bias, self.act is not None, False)
        bn = self.bn
        training = bn.training or bn.running_mean is None
        if bn.training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
        y = _BatchNormFn.apply(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, momentum, bn.eps)
        return _lincomb_ad([xin, y], [1.0, 1.0]) if add_resid else y


class AudioAlgebra(nn.Module):
    """Main AudioAlgebra model (aa_mixer.py:224-267): encode(x) = x + enc(x^T)^T, decode likewise,
    forward -> (z, decode(z)); x is [B, dims, T].  state_dict keys match the reference
    (encoder.{0..3}.lin.{weight,bias}, decoder.{0..3}.lin.{weight,bias})."""

    def __init__(self, dims=32, hidden_dims=64, act=nn.GELU(), use_bn=False, resid=True, block=EmbedBlock, trivial=False):
        super().__init__()
        if block is not EmbedBlock:
            raise NotImplementedError("custom block classes are not supported by the fused projector")
        if dims > 64 or hidden_dims > 64:
            raise NotImplementedError("the fused projector supports dims, hidden_dims <= 64")
        self.resid, self.trivial, self.dims, self.hidden_dims = resid, trivial, dims, hidden_dims
        self.encoder = nn.Sequential(
            block(dims, hidden_dims, act=act, use_bn=use_bn, resid=resid),
            block(hidden_dims, hidden_dims, act=act, use_bn=use_bn, resid=resid),
            block(hidden_dims, hidden_dims, act=act, use_bn=use_bn, resid=resid),
            block(hidden_dims, dims, act=None, use_bn=use_bn, resid=resid),
        )
        self.decoder = nn.Sequential(
            block(dims, hidden_dims, act=act, use_bn=use_bn, resid=resid),
            block(hidden_dims, hidden_dims, act=act, use_bn=use_bn, resid=resid),
            block(hidden_dims, hidden_dims, act=act, use_bn=use_bn, resid=resid),
            block(hidden_dims, dims, act=None, use_bn=use_bn, resid=resid),
        )

    def _half(self, seq, xin):
        if any(blk.bn is not None for blk in seq):   # use_bn: block by block, exactly the reference's data flow (aa_mixer.py:252-254)
            x = seq(xin.transpose(1, 2)).transpose(1, 2)
            return _lincomb_ad([x, xin], [1.0, 1.0]) if self.resid else x
        ws = [blk.lin.weight for blk in seq]
        bs = [blk.lin.bias for blk in seq]
        return _ProjHalf.apply(xin, self.resid, self.dims, self.hidden_dims, *ws, *bs)

    def encode(self, xin):
        if self.trivial:
            return xin
        return self._half(self.encoder, xin)

    def decode(self, xin):
        if self.trivial:
            return xin
        return self._half(self.decoder, xin)

    def forward(self, x):
        xprime = self.encode(x)
        xprimeprime = self.decode(xprime)
        return xprime, xprimeprime


# ---------------------------------------------------------------------------------------------------
# stems / faders / mixing
# ---------------------------------------------------------------------------------------------------

def get_stems_faders(batch, dl_iter, dl, maxstems=2, unity_gain=False, debug=False):
    "aa_mixer.py:270-292, same RNG recipe (python `random` for nstems, torch CPU RNG for the faders)"
    nstems = random.randint(2, maxstems)
    if debug:
        print("maxstems, nstems =", maxstems, nstems)
    device = batch.device
    faders = torch.sgn(2 * torch.rand(nstems) - 1)
    if not unity_gain:
        faders += 0.5 * torch.tanh(2 * (2 * torch.rand(nstems) - 1))
    stems = [batch]
    for i in range(nstems - 1):
        try:
            next_stem = next(dl_iter).to(device)
        except StopIteration:
            dl_iter = iter(dl)
            next_stem = next(dl_iter).to(device)
        if debug:
            print("  next_stem.shape = ", next_stem.shape)
        stems.append(next_stem)
    fdev = faders.to(device)
    fdev._aa_host = faders.tolist()   # read by do_mixing instead of a device -> host copy of the same numbers
    return stems, fdev, dl_iter


class _LazyArchive(dict):
    """The archive dict of do_mixing with its two waveform-sized entries ('fadedstems', 'mix') materialised on first access: the
    training step never reads them (train_aa_mixer_accel.py:504-517 uses ymix / ymix_recon / yrecons), the demo code does."""

    def __init__(self, eager, lazy):
        super().__init__(eager)
        self._lazy = dict(lazy)

    def __getitem__(self, k):
        if not dict.__contains__(self, k) and k in self._lazy:
            dict.__setitem__(self, k, self._lazy.pop(k)())
        return dict.__getitem__(self, k)

    def __contains__(self, k):
        return dict.__contains__(self, k) or k in self._lazy

    def keys(self):
        return list(dict.keys(self)) + list(self._lazy.keys())

    def get(self, k, default=None):
        return self[k] if k in self else default


def do_mixing(stems, faders, given_model, aa_model, device, debug=False, **kwargs):
    """aa_mixer.py:295-327.  Same outputs (zsum, zmix, archive).  The reference re-encodes the running
    mix after every stem and keeps only the last result; here the mix is encoded once after the loop
    (identical zmix / archive['ymix']).  Given models that expose `encode_mix` (the conv encoder) take the fader scaling and the
    stem sum inside their first layer's load, so `fadedstem` / `mix` are not written to HBM unless the archive entry is read."""
    zs, ys, yrecons = [], [], []
    host = getattr(faders, "_aa_host", None)   # get_stems_faders keeps the host copy it drew the faders from: no device sync here
    fl = [float(f) for f in (host if host is not None else (faders.detach().cpu().tolist() if torch.is_tensor(faders) else faders))]
    stems = [s.to(device) for s in stems]
    fused = hasattr(given_model, "encode_mix") and len(stems) <= 4
    fadedstems = None if fused else []
    for s, f in zip(stems, fl):
        with torch.no_grad():
            if fused:
                y = given_model.encode_mix([s], [f])
            else:
                fadedstem = latent_lincomb([s], [f])
                fadedstems.append(fadedstem)
                y = given_model.encode(fadedstem)
        z, y_recon = aa_model(y)
        yrecons.append(y_recon); zs.append(z); ys.append(y)
    n = len(zs)
    zsum = _lincomb_ad(zs, [1.0] * n) if n > 1 else zs[0]
    with torch.no_grad():
        if fused:
            mix = None
            ymix = given_model.encode_mix(stems[:n], fl[:n])
        else:
            mix = latent_lincomb(stems[:n], fl[:n])
            ymix = given_model.encode(mix)
        ysum = latent_lincomb(ys, [1.0] * n) if n > 1 else ys[0]
    zmix, ymix_recon = aa_model(ymix)
    eager = {'zs': zs, 'ys': ys, 'ymix': ymix, 'ymix_recon': ymix_recon, 'yrecons': yrecons, 'ysum': ysum}
    if fused:
        archive = _LazyArchive(eager, {'mix': lambda: latent_lincomb(stems[:n], fl[:n]),
                                       'fadedstems': lambda: [latent_lincomb([s], [f]) for s, f in zip(stems, fl)]})
    else:
        archive = dict(eager, mix=mix, fadedstems=fadedstems)
    return zsum, zmix, archive
