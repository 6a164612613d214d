"""Encode side of the reference's second given-model family:

    StackedDiffAEWrapper.encode        audio_algebra/given_models.py:361-385
    LatentAudioDiffusionAutoencoder    audio_algebra/StackedAELatentDiffusionCond.py:183-227

    encode(reals [B, 2, N]) = tanh( latent_encoder( autoencoder.encode(reals) ) )      [12, 2, 262144] -> [12, 32, 512]

`autoencoder` = AudioAutoencoder(capacity 64, c_mults [2,4,8,16,32], strides [2]*5, latent_dim 32): a SoundStreamXL encoder
(the same layer table as the DVAE's, other widths / strides) followed by tanh -- it runs on the existing aa_encoder_* entry
points.  `latent_encoder` = Encoder1d(in 32, out 32, channels 128, multipliers [1,2,4,8,8], factors [2]*4, num_blocks [8]*4):
a 1x1 input conv, four stages of (strided k = 2f+1 conv, eight GroupNorm-SiLU-Conv3 ResNet blocks), a 1x1 output conv ->
aa_conv1d_f32 / aa_groupnorm_act_f32.

Both classes are THIRD-PARTY in the reference (`autoencoders.models.AudioAutoencoder` from the audio-diffusion fork,
`audio_encoders_pytorch.Encoder1d`), their source is not in the reference tree and their git dependencies are un-pinned:
the layer structure below is restated from memory of the upstream packages -- PARITY UNPINNED, exactly like the DVAE encoder
(SURVEY.md Appendix A).  What the reference itself pins is the shape [12, 2, 262144] -> [12, 32, 512]
(StackedDiffAE.ipynb cell 13).  The reference's two `print(torch.std(...))` calls inside encode (:223, :225) force a device
sync each and are not reproduced.  decode / the diffusion sampler are out of scope and raise.
"""
import ctypes as C
from copy import deepcopy

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .aa_mixer import _f32c
from .DiffusionDVAE import SoundStreamXLEncoder

__all__ = ['AudioAutoencoder', 'Encoder1d', 'LatentAudioDiffusionAutoencoder']

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_lib.register({
    "aa_conv1d_f32": (_i, [_p, _i64, _i, _i64, _p, _p, _i, _i, _i, _i, _i, _p, _i, _p, _p]),
    "aa_groupnorm_act_f32": (_i, [_p, _i64, _i, _i64, _p, _p, _i, _f, _i, _p, _p]),
})

ACT_NONE, ACT_ELU, ACT_TANH = 0, 1, 2


def conv1d(x, conv: nn.Conv1d, res=None, act=ACT_NONE):
    "act(conv(x) + bias (+ res)) on [B, C, L] fp32 through aa_conv1d_f32 (`conv` only holds the parameters and the geometry)"
    x = _f32c(x, "activation")
    w, b = _f32c(conv.weight.detach()), _f32c(conv.bias.detach())
    k, s, d, pd = conv.kernel_size[0], conv.stride[0], conv.dilation[0], conv.padding[0]
    bsz, cin, lin = x.shape
    assert cin == conv.in_channels
    lout = (lin + 2 * pd - d * (k - 1) - 1) // s + 1
    out = torch.empty((bsz, conv.out_channels, lout), dtype=torch.float32, device=x.device)
    if res is not None:
        res = _f32c(res)
        assert res.shape == out.shape
    with torch.cuda.device(x.device):
        check(lib.aa_conv1d_f32(ptr(x), bsz, cin, lin, ptr(w), ptr(b), conv.out_channels, k, s, d, pd,
                                None if res is None else ptr(res), int(act), ptr(out), stream_ptr()))
    return out


def groupnorm_silu(x, gn: nn.GroupNorm, silu=True):
    x = _f32c(x, "activation")
    bsz, c, l = x.shape
    out = torch.empty_like(x)
    g = None if gn.weight is None else _f32c(gn.weight.detach())
    b = None if gn.bias is None else _f32c(gn.bias.detach())
    with torch.cuda.device(x.device):
        check(lib.aa_groupnorm_act_f32(ptr(x), bsz, c, l, None if g is None else ptr(g), None if b is None else ptr(b),
                                       gn.num_groups, float(gn.eps), int(silu), ptr(out), stream_ptr()))
    return out


class AudioAutoencoder(nn.Module):
    """First stage: `encode(audio) = tanh(SoundStreamXLEncoder(audio))` (the decoder half is out of scope).  Members follow the
    upstream class: encoder, latent_dim, downsampling_ratio."""

    def __init__(self, capacity=64, c_mults=[2, 4, 8, 16, 32], strides=[2, 2, 2, 2, 2], latent_dim=32, compute_dtype="fp32", **kwargs):
        super().__init__()
        self.latent_dim = latent_dim
        self.downsampling_ratio = 1
        for s in strides:
            self.downsampling_ratio *= s
        self.encoder = SoundStreamXLEncoder(in_channels=2, capacity=capacity, latent_dim=latent_dim, c_mults=c_mults, strides=strides,
                                            compute_dtype=compute_dtype)

    def encode(self, audio, skip_bottleneck=False):
        return self.encoder.encode_mix([audio], None, apply_tanh=not skip_bottleneck)

    def decode(self, latents):
        raise NotImplementedError("AudioAutoencoder.decode is outside the accelerated hot path (encode side only)")


class ConvBlock1d(nn.Module):
    "GroupNorm -> SiLU -> Conv1d(k = 3, padding 1)"

    def __init__(self, in_channels, out_channels, *, kernel_size=3, stride=1, padding=1, dilation=1, num_groups=8):
        super().__init__()
        self.groupnorm = nn.GroupNorm(num_groups=num_groups, num_channels=in_channels)
        self.activation = nn.SiLU()
        self.project = nn.Conv1d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation)

    def forward(self, x, res=None):
        return conv1d(groupnorm_silu(x, self.groupnorm), self.project, res=res)


class ResnetBlock1d(nn.Module):
    "x -> ConvBlock1d -> ConvBlock1d, plus x (through a 1x1 conv when the widths differ); the sum is fused into the second conv"

    def __init__(self, in_channels, out_channels, *, num_groups=8):
        super().__init__()
        self.block1 = ConvBlock1d(in_channels, out_channels, num_groups=num_groups)
        self.block2 = ConvBlock1d(out_channels, out_channels, num_groups=num_groups)
        self.to_out = nn.Conv1d(in_channels, out_channels, 1) if in_channels != out_channels else nn.Identity()

    def forward(self, x):
        skip = x if isinstance(self.to_out, nn.Identity) else conv1d(x, self.to_out)
        return self.block2(self.block1(x), res=skip)


class DownsampleBlock1d(nn.Module):
    "Conv1d(k = 2 f + 1, stride f, padding f), then num_layers ResNet blocks at the new width"

    def __init__(self, in_channels, out_channels, *, factor, num_groups, num_layers):
        super().__init__()
        self.downsample = nn.Conv1d(in_channels, out_channels, kernel_size=factor * 2 + 1, stride=factor, padding=factor)
        self.blocks = nn.ModuleList([ResnetBlock1d(out_channels, out_channels, num_groups=num_groups) for _ in range(num_layers)])

    def forward(self, x):
        x = conv1d(x, self.downsample)
        for blk in self.blocks:
            x = blk(x)
        return x


class Encoder1d(nn.Module):
    def __init__(self, in_channels, channels, multipliers, factors, num_blocks, patch_size=1, resnet_groups=8, out_channels=None):
        super().__init__()
        assert patch_size == 1, "patch_size > 1 is not used by the reference"
        self.num_layers = len(multipliers) - 1
        assert len(factors) == self.num_layers and len(num_blocks) == self.num_layers
        self.downsample_factor = patch_size
        for f in factors:
            self.downsample_factor *= f
        self.out_channels = out_channels if out_channels is not None else channels * multipliers[-1]
        self.to_in = nn.Conv1d(in_channels, channels * multipliers[0], 1)
        self.downsamples = nn.ModuleList([
            DownsampleBlock1d(channels * multipliers[i], channels * multipliers[i + 1], factor=factors[i], num_groups=resnet_groups,
                              num_layers=num_blocks[i]) for i in range(self.num_layers)])
        self.to_out = nn.Conv1d(channels * multipliers[-1], out_channels, 1) if out_channels is not None else nn.Identity()

    def forward(self, x, act_out=ACT_NONE):
        x = conv1d(x, self.to_in)
        for d in self.downsamples:
            x = d(x)
        if isinstance(self.to_out, nn.Identity):
            return torch.tanh(x) if act_out == ACT_TANH else x
        return conv1d(x, self.to_out, act=act_out)


class LatentAudioDiffusionAutoencoder(nn.Module):
    """Encode side of StackedAELatentDiffusionCond.py:183-227.  Same members as the reference where the encode path touches
    them: latent_dim, second_stage_latent_dim, latent_downsampling_ratio, downsampling_ratio, latent_encoder (+ _ema), autoencoder."""

    def __init__(self, autoencoder: AudioAutoencoder):
        super().__init__()
        self.latent_dim = autoencoder.latent_dim
        self.second_stage_latent_dim = 32
        factors = [2, 2, 2, 2]
        self.latent_downsampling_ratio = 16
        self.downsampling_ratio = autoencoder.downsampling_ratio * self.latent_downsampling_ratio
        self.latent_encoder = Encoder1d(in_channels=self.latent_dim, out_channels=self.second_stage_latent_dim, channels=128,
                                        multipliers=[1, 2, 4, 8, 8], factors=factors, num_blocks=[8, 8, 8, 8])
        self.latent_encoder_ema = deepcopy(self.latent_encoder)
        self.latent_encoder_ema.requires_grad_(False)
        self.autoencoder = autoencoder
        self.autoencoder.requires_grad_(False).eval()

    def encode(self, reals):
        with torch.no_grad():
            first_stage_latents = self.autoencoder.encode(reals)
            return self.latent_encoder(first_stage_latents, act_out=ACT_TANH)   # tanh fused into the output conv

    def decode(self, *args, **kwargs):
        raise NotImplementedError("the latent diffusion decoder is outside the accelerated hot path (encode side only)")
