"""Effects flavour of the AA core: reference audio_algebra/aa_effects.py (do_mixing :116-125, the same
EmbedBlock / AudioAlgebra / losses re-exported) plus the effects training-step algebra of
train_aa_effects.py:58-98 (L2-hinge variance loss :42-46)."""
import torch

from .aa_mixer import (EmbedBlock, AudioAlgebra, mseloss, vicreg_var_loss, vicreg_var_loss_l2, vicreg_cov_loss,  # noqa: F401
                       off_diagonal, _lincomb_ad, effects_loss_fused)

__all__ = ['EmbedBlock', 'AudioAlgebra', 'do_mixing', 'mseloss', 'vicreg_var_loss', 'vicreg_var_loss_l2',
           'vicreg_cov_loss', 'off_diagonal', 'effects_guesses', 'effects_losses']


def do_mixing(batch, given_model, aa_model, device, debug=False):
    """aa_effects.py:116-125: ys = f(x), zs = h(y), yrecons = h^-1(z) for x in (a1, b1, a2, b2).
    The four inputs are encoded as ONE 4B batch (same values, one pass over the encoder weights)."""
    xs = [batch[k].to(device) for k in ("a1", "b1", "a2", "b2")]
    b = xs[0].shape[0]
    if all(x.shape == xs[0].shape for x in xs):
        y_all = given_model.encode(torch.cat(xs, dim=0))
        ys = list(torch.split(y_all, b, dim=0))
    else:
        ys = [given_model.encode(x) for x in xs]
    ys = [y.contiguous() for y in ys]
    zs = [aa_model.encode(y) for y in ys]
    yrecons = [aa_model.decode(z) for z in zs]
    return {'ys': ys, 'zs': zs, 'yrecons': yrecons}


def effects_guesses(za1, zb1, za2, zb2):
    "train_aa_effects.py:70-71: za2_guess = zb2 - zb1 + za1 ; zb2_guess = za2 - za1 + zb1 (one pass each)"
    return _lincomb_ad([zb2, zb1, za1], [1.0, -1.0, 1.0]), _lincomb_ad([za2, za1, zb1], [1.0, -1.0, 1.0])


def effects_losses(archive, fused=True):
    """loss terms of AAEffectsModule.training_step (train_aa_effects.py:66-82).  fused (default): one aa_effects_loss_fwd_f32 /
    _bwd_f32 call each way; fused=False assembles the same terms from the standalone loss Functions (the tests compare the two)."""
    if fused:
        return effects_loss_fused([z.float() for z in archive["zs"]], [y.float() for y in archive["ys"]],
                                  [r.float() for r in archive["yrecons"]])
    za1, zb1, za2, zb2 = [z.float() for z in archive["zs"]]
    za2_guess, zb2_guess = effects_guesses(za1, zb1, za2, zb2)
    mix_loss = (mseloss(za2_guess, za2) + mseloss(zb2_guess, zb2)) / 2
    var_loss = (vicreg_var_loss_l2(za1) + vicreg_var_loss_l2(za2) + vicreg_var_loss_l2(zb1) + vicreg_var_loss_l2(zb2)) / 4
    cov_loss = (vicreg_cov_loss(za1) + vicreg_cov_loss(za2) + vicreg_cov_loss(zb1) + vicreg_cov_loss(zb2)) / 4
    aa_recon_loss = mseloss(archive["yrecons"][0].float(), archive["ys"][0].float())
    for i in range(1, 4):
        aa_recon_loss = aa_recon_loss + mseloss(archive["yrecons"][i].float(), archive["ys"][i].float())
    loss = mix_loss + var_loss + cov_loss + aa_recon_loss
    return {'loss': loss, 'mix_loss': mix_loss, 'var_loss': var_loss, 'cov_loss': cov_loss, 'aa_recon_loss': aa_recon_loss}
