"""Mixer training step of the reference's train_aa_mixer_accel.py:463-550 on B200: same loss recipe,
Adam(lr=5e-4) + OneCycleLR(max_lr=1e-3) semantics (including OneCycleLR's default beta1 cycling), one
flat fp32 parameter / gradient buffer, fused Adam kernel, and a *real* gradient all-reduce over
torch.distributed (the reference's accelerate script bypasses DDP.forward and never synchronises
gradients -- SURVEY.md section 5)."""
import math

import torch

from ._lib import lib, check, ptr, stream_ptr
from .parallel import allreduce_mean_
from .aa_mixer import (AudioAlgebra, do_mixing, get_stems_faders, mseloss, vicreg_var_loss, vicreg_cov_loss, mixer_loss_fused)  # noqa: F401

__all__ = ['FlatAdam', 'onecycle_lr', 'onecycle_beta1', 'mixer_losses', 'MixerTrainer', 'save_aa_checkpoint', 'load_aa_checkpoint']


def _cos(a, b, pct):
    return b + (a - b) / 2.0 * (math.cos(math.pi * pct) + 1)


def onecycle_lr(step, total_steps, max_lr=1e-3, pct_start=0.3, div_factor=25.0, final_div_factor=1e4):
    "lr used by optimiser step `step` (0-based) under torch's OneCycleLR defaults"
    initial, min_lr = max_lr / div_factor, max_lr / div_factor / final_div_factor
    up_end, down_end = float(pct_start * total_steps) - 1, total_steps - 1
    if step <= up_end:
        return _cos(initial, max_lr, step / up_end)
    return _cos(max_lr, min_lr, (step - up_end) / (down_end - up_end))


def onecycle_beta1(step, total_steps, pct_start=0.3, base_momentum=0.85, max_momentum=0.95):
    up_end, down_end = float(pct_start * total_steps) - 1, total_steps - 1
    if step <= up_end:
        return _cos(max_momentum, base_momentum, step / up_end)
    return _cos(base_momentum, max_momentum, (step - up_end) / (down_end - up_end))


class FlatAdam:
    "torch.optim.Adam + OneCycleLR on one flat fp32 CUDA buffer, one kernel per step"

    def __init__(self, flat_params, lr=5e-4, max_lr=1e-3, total_steps=None, betas=(0.9, 0.999), eps=1e-8):
        assert flat_params.is_cuda and flat_params.dtype == torch.float32 and flat_params.is_contiguous()
        self.params = flat_params
        self.m = torch.zeros_like(flat_params)
        self.v = torch.zeros_like(flat_params)
        self.lr, self.max_lr, self.total_steps, self.betas, self.eps = lr, max_lr, total_steps, betas, eps
        self.t = 0

    def current_lr(self):
        return self.lr if self.total_steps is None else onecycle_lr(self.t, self.total_steps, self.max_lr)

    def state_dict(self):
        "moments, step counter (= OneCycle position) and hyper-parameters: enough to resume training bit-exactly"
        return {'m': self.m.detach().clone(), 'v': self.v.detach().clone(), 't': int(self.t), 'lr': self.lr, 'max_lr': self.max_lr,
                'total_steps': self.total_steps, 'betas': tuple(self.betas), 'eps': self.eps}

    def load_state_dict(self, sd):
        assert sd['m'].numel() == self.m.numel() and sd['v'].numel() == self.v.numel(), "optimizer state does not match the parameters"
        self.m.copy_(sd['m'].to(self.m.device).reshape_as(self.m))
        self.v.copy_(sd['v'].to(self.v.device).reshape_as(self.v))
        self.t = int(sd['t'])
        self.lr, self.max_lr, self.total_steps = sd['lr'], sd['max_lr'], sd['total_steps']
        self.betas, self.eps = tuple(sd['betas']), sd['eps']

    def step(self, flat_grads):
        lr = self.current_lr()
        b1 = self.betas[0] if self.total_steps is None else onecycle_beta1(self.t, self.total_steps)
        self.t += 1
        with torch.cuda.device(self.params.device):
            check(lib.aa_adam_step_f32(ptr(self.params), ptr(flat_grads.contiguous()), ptr(self.m), ptr(self.v),
                                       self.params.numel(), lr, b1, self.betas[1], self.eps, self.t, stream_ptr()))


def mixer_losses(zsum, zmix, y, yrecon, ymix, ymix_recon, fused=True):
    """train_aa_mixer_accel.py:504-517.  fused (default): one aa_mixer_loss_fwd_f32 / _bwd_f32 call each way; fused=False assembles
    the same terms from the standalone loss Functions (the tests compare the two)."""
    if fused:
        return mixer_loss_fused(zsum, zmix, y, yrecon, ymix, ymix_recon)
    mix_loss = mseloss(zsum, zmix)
    var_loss = (vicreg_var_loss(zsum) + vicreg_var_loss(zmix)) / 2
    cov_loss = (vicreg_cov_loss(zsum) + vicreg_cov_loss(zmix)) / 2
    aa_recon_loss = mseloss(y, yrecon) + mseloss(ymix, ymix_recon)
    loss = mix_loss + var_loss + cov_loss + aa_recon_loss
    return {'loss': loss, 'mix_loss': mix_loss, 'var_loss': var_loss, 'cov_loss': cov_loss, 'aa_recon_loss': aa_recon_loss}


class MixerTrainer:
    """One data-parallel rank of the mixer training loop.  Parameters of `aa_model` are re-pointed into a
    flat buffer so that the gradient all-reduce is ONE 133 KB message and Adam is one kernel."""

    def __init__(self, given_model, aa_model, total_steps, lr=5e-4, max_lr=1e-3, process_group=None):
        import torch.distributed as dist
        self.given_model, self.aa_model, self.group = given_model, aa_model, process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        params = [p for p in aa_model.parameters()]
        self.flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        self.flat_grad = torch.zeros_like(self.flat)
        off = 0
        for p in params:  # parameters and their .grad become views of the flat buffers
            n = p.numel()
            p.data = self.flat[off:off + n].view_as(p)
            p.grad = self.flat_grad[off:off + n].view_as(p)
            off += n
        self.opt = FlatAdam(self.flat, lr=lr, max_lr=max_lr, total_steps=total_steps)
        self.keep_local_grad = False   # tests / bench: keep this rank's gradient as it was BEFORE the all-reduce
        self.local_grad = None

    def step(self, stems, faders, batch=None):
        """stems: list of [B,2,N] device tensors (stems[0] doubles as `batch` of the reference loop);
        returns the dict of (detached, on-device) loss terms -- no host sync."""
        self.flat_grad.zero_()
        device = stems[0].device
        zsum, zmix, archive = do_mixing(stems, faders, self.given_model, self.aa_model, device)
        with torch.no_grad():   # the reference re-encodes the un-faded batch (= stems[0]) for the recon term
            y = self.given_model.encode(stems[0] if batch is None else batch)
        z, yrecon = self.aa_model(y)
        losses = mixer_losses(zsum, zmix, y, yrecon, archive['ymix'], archive['ymix_recon'])
        losses['loss'].backward()
        if self.keep_local_grad:
            self.local_grad = self.flat_grad.clone()
        allreduce_mean_(self.flat_grad, self.group)
        self.opt.step(self.flat_grad)
        return {k: v.detach() for k, v in losses.items()}

    def state_dict(self):
        return {'aa_model': self.aa_model.state_dict(), 'opt': self.opt.state_dict()}

    def load_state_dict(self, sd):
        "parameters are copied INTO the flat buffer (the views installed by __init__ stay valid)"
        with torch.no_grad():
            own = self.aa_model.state_dict()
            for k, v in sd['aa_model'].items():
                own[k].copy_(v.to(own[k].device))
        self.opt.load_state_dict(sd['opt'])


def save_aa_checkpoint(aa_model, path, trainer=None, **extra):
    """The `save_aa_checkpoint` the reference calls but never defines in the script (train_aa_mixer_accel.py:550; the notebook
    version is a torch.save of the module).  Writes a plain dict: 'state_dict' under the reference's key names
    (encoder.{0..3}.lin.{weight,bias}, decoder....) so that it loads into the reference's AudioAlgebra with load_state_dict,
    plus -- when a trainer is given -- the flat Adam moments, step counter and OneCycle position for exact resume."""
    ckpt = {'state_dict': {k: v.detach().cpu().clone() for k, v in aa_model.state_dict().items()},
            'dims': getattr(aa_model, 'dims', None), 'hidden_dims': getattr(aa_model, 'hidden_dims', None)}
    if trainer is not None:
        ckpt['opt'] = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in trainer.opt.state_dict().items()}
    ckpt.update(extra)
    torch.save(ckpt, path)
    return path


def load_aa_checkpoint(aa_model, path, trainer=None, map_location='cpu'):
    "inverse of save_aa_checkpoint; also accepts a bare state_dict or a pickled module (the notebook's torch.save(model))"
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    if isinstance(ckpt, torch.nn.Module):
        ckpt = {'state_dict': ckpt.state_dict()}
    sd = ckpt.get('state_dict', ckpt)
    if trainer is not None:
        trainer.load_state_dict({'aa_model': sd, 'opt': ckpt['opt']} if 'opt' in ckpt else {'aa_model': sd, 'opt': trainer.opt.state_dict()})
    else:
        with torch.no_grad():
            own = aa_model.state_dict()
            missing = set(own) - set(sd)
            assert not missing, f"checkpoint lacks {sorted(missing)}"
            for k in own:
                own[k].copy_(sd[k].to(own[k].device))
    return ckpt
