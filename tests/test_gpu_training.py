"""GPU parity of the data-parallel mixer training step (train_aa_mixer_accel.py:495-545) and of
the effects step (train_aa_effects.py:58-98) against the oracle: the same step restated with torch autograd in float64 on
the CPU -- restated encoder (no tanh: DiffusionDVAE.encode, aa_mixer.py:165-168), projector, the four loss terms, Adam with
the OneCycleLR lr / beta1 schedule.  Checked per step: every loss term, the flat gradient, the updated parameters."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

FADERS = [1.4630, -0.5718]   # aa-mixer-toy.ipynb cell 39


def _x(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) - 0.5


@pytest.fixture(scope="module")
def setup():
    import audio_algebra_b200 as aab
    from oracle import aa_oracle as O
    torch.manual_seed(0)
    enc_o = O.SoundStreamXLEncoderOracle().eval()
    dv = aab.DVAEWrapper(debug=False)
    dv.model.load_oracle_weights(enc_o)
    return aab, O, enc_o, dv.cuda()


def _oracle_params(O, sd):
    "float64 leaf tensors in AudioAlgebra.parameters() order (encoder.0.lin.weight, encoder.0.lin.bias, ...)"
    names = [f"{half}.{i}.lin.{wb}" for half in ("encoder", "decoder") for i in range(4) for wb in ("weight", "bias")]
    return names, [sd[n].double().clone().requires_grad_(True) for n in names]


def _oracle_forward(O, names, params, y):
    sd = dict(zip(names, params))
    ew = [sd[f"encoder.{i}.lin.weight"] for i in range(4)]
    eb = [sd[f"encoder.{i}.lin.bias"] for i in range(4)]
    dw = [sd[f"decoder.{i}.lin.weight"] for i in range(4)]
    db = [sd[f"decoder.{i}.lin.bias"] for i in range(4)]
    return O.projector_forward(y, ew, eb, dw, db)


def test_mixer_training_steps_match_oracle(setup):
    aab, O, enc_o, dv = setup
    from audio_algebra_b200.training import MixerTrainer
    sd = O.init_projector_state_dict(64, 64, seed=2)
    aa = aab.AudioAlgebra(64, 64)
    aa.load_state_dict(sd)
    aa = aa.cuda()
    total_steps = 10
    trainer = MixerTrainer(dv.model, aa, total_steps=total_steps)
    names, P = _oracle_params(O, sd)
    m = [torch.zeros_like(p) for p in P]
    v = [torch.zeros_like(p) for p in P]
    for step in range(3):
        stems = [_x((4, 2, 2048), 100 + 2 * step), _x((4, 2, 2048), 101 + 2 * step)]
        out = trainer.step([s.cuda() for s in stems], FADERS)
        # ---- oracle step ----
        with torch.no_grad():
            ys = [O.dvae_encode(enc_o, f * s).double() for s, f in zip(stems, FADERS)]
            ymix = O.dvae_encode(enc_o, FADERS[0] * stems[0] + FADERS[1] * stems[1]).double()
            y = O.dvae_encode(enc_o, stems[0]).double()
        zs = [_oracle_forward(O, names, P, yy)[0] for yy in ys]
        zmix, ymix_recon = _oracle_forward(O, names, P, ymix)
        z, yrecon = _oracle_forward(O, names, P, y)
        Lo = O.mixer_losses(zs[0] + zs[1], zmix, y, yrecon, ymix, ymix_recon)
        for p in P:
            p.grad = None
        Lo["loss"].backward()
        for k in ("mix_loss", "var_loss", "cov_loss", "aa_recon_loss", "loss"):
            assert abs(out[k].item() - Lo[k].item()) < 2e-3 * max(abs(Lo[k].item()), 1e-3), (step, k, out[k].item(), Lo[k].item())
        g_ref = torch.cat([p.grad.reshape(-1) for p in P])
        assert rel_l2(trainer.flat_grad, g_ref) < 5e-3, (step, rel_l2(trainer.flat_grad, g_ref))
        lr, b1 = O.onecycle_lr(step, total_steps), O.onecycle_beta1(step, total_steps)
        with torch.no_grad():
            before = torch.cat([p.reshape(-1) for p in P]).clone()
            for i, p in enumerate(P):
                pn, m[i], v[i] = O.adam_step(p.detach(), p.grad, m[i], v[i], step + 1, lr, b1=b1)
                p.copy_(pn)
            after = torch.cat([p.reshape(-1) for p in P])
        # Adam's first steps are sign-like (|update| ~ lr): compare the update itself, loosely, and the parameters tightly
        assert rel_l2(trainer.flat.cpu().double() - before, after - before) < 0.1, step
        assert rel_l2(trainer.flat, after) < 1e-4, step
    # the module's parameters are views of the flat buffer: state_dict reflects the trained values
    sd_new = aa.state_dict()
    assert rel_l2(sd_new["encoder.0.lin.weight"], P[0]) < 1e-4


def test_effects_step_matches_oracle(setup):
    aab, O, enc_o, dv = setup
    sd = O.init_projector_state_dict(64, 64, seed=2)
    aa = aab.aa_effects.AudioAlgebra(64, 64)
    aa.load_state_dict(sd)
    aa = aa.cuda()
    batch = {k: _x((3, 2, 2048), 200 + i) for i, k in enumerate(("a1", "b1", "a2", "b2"))}
    arch = aab.aa_effects.do_mixing(batch, dv.model, aa, "cuda")
    L = aab.aa_effects.effects_losses(arch)
    L["loss"].backward()
    names, P = _oracle_params(O, sd)
    with torch.no_grad():
        ys = [O.dvae_encode(enc_o, batch[k]).double() for k in ("a1", "b1", "a2", "b2")]
    zs, yrecons = zip(*[_oracle_forward(O, names, P, y) for y in ys])
    Lo = O.effects_losses(list(zs), ys, list(yrecons))
    Lo["loss"].backward()
    for k in ("mix_loss", "var_loss", "cov_loss", "aa_recon_loss", "loss"):
        assert abs(L[k].item() - Lo[k].item()) < 2e-3 * max(abs(Lo[k].item()), 1e-3), (k, L[k].item(), Lo[k].item())
    g = torch.cat([p.grad.reshape(-1) for p in aa.parameters()])
    g_ref = torch.cat([p.grad.reshape(-1) for p in P])
    assert rel_l2(g, g_ref) < 5e-3
    for i in range(4):
        assert rel_l2(arch["ys"][i], ys[i]) < 1e-3 and rel_l2(arch["zs"][i], zs[i]) < 1e-3


def test_fused_loss_entry_points_match_the_assembled_losses():
    """aa_mixer_loss_fwd/bwd_f32 and aa_effects_loss_fwd/bwd_f32 (one C call per direction) against the same terms
    assembled from the standalone loss Functions: every log_dict entry and every gradient."""
    import audio_algebra_b200 as aab
    from audio_algebra_b200.training import mixer_losses
    g = torch.Generator(device="cuda").manual_seed(11)
    b, c, t = 96, 64, 32

    def mk():
        return (0.8 * torch.randn(b, c, t, device="cuda", generator=g) + 0.1)

    base = [mk() for _ in range(6)]
    res = {}
    for fused in (True, False):
        zsum, zmix, y, yrecon, ymix, ymix_recon = [v.clone().requires_grad_(i in (0, 1, 3, 5)) for i, v in enumerate(base)]
        L = mixer_losses(zsum, zmix, y, yrecon, ymix, ymix_recon, fused=fused)
        (1.7 * L["loss"]).backward()
        res[fused] = ({k: float(v.detach()) for k, v in L.items()}, [zsum.grad, zmix.grad, yrecon.grad, ymix_recon.grad])
    for k in res[True][0]:
        assert abs(res[True][0][k] - res[False][0][k]) <= 2e-6 * abs(res[False][0][k]), k
    for ga, gb in zip(res[True][1], res[False][1]):
        assert rel_l2(ga, gb) < 2e-6
    base = [mk() for _ in range(12)]
    res = {}
    for fused in (True, False):
        ts = [v.clone().requires_grad_(i < 4 or i >= 8) for i, v in enumerate(base)]
        L = aab.aa_effects.effects_losses({"zs": ts[0:4], "ys": ts[4:8], "yrecons": ts[8:12]}, fused=fused)
        (0.6 * L["loss"]).backward()
        res[fused] = ({k: float(v.detach()) for k, v in L.items()}, [v.grad for v in ts[0:4] + ts[8:12]])
    for k in res[True][0]:
        assert abs(res[True][0][k] - res[False][0][k]) <= 2e-6 * abs(res[False][0][k]), k
    for ga, gb in zip(res[True][1], res[False][1]):
        assert rel_l2(ga, gb) < 2e-6
