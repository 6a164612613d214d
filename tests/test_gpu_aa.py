"""GPU parity of the AA core (projector, latent algebra, losses, mixing, PCA, Adam) against fixtures
from the reference's own code (tests/golden) and the float64 oracle.  Tolerances: fp32 kernels,
relative L2 <= 1e-3 on embeddings (BASELINE.json); most checks are much tighter."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
T = torch.from_numpy


@pytest.fixture(scope="module")
def aab():
    import audio_algebra_b200 as aab
    return aab


def _oracle():
    from oracle import aa_oracle as O
    return O


def _load_aa(aab, g, prefix="sd.", dims=64, hidden=64):
    m = aab.AudioAlgebra(dims=dims, hidden_dims=hidden)
    m.load_state_dict({k[len(prefix):]: T(v) for k, v in g.items() if k.startswith(prefix)})
    return m.cuda()


def test_projector_forward_golden(aab, golden):
    g = golden("projector")
    aa = _load_aa(aab, g)
    y = T(g["y"]).cuda()
    z, yr = aa(y)
    assert rel_l2(z, g["z"]) < 1e-5 and rel_l2(yr, g["y_recon"]) < 1e-5
    assert rel_l2(aa.encode(y), g["z_encode"]) < 1e-5
    assert rel_l2(aa.decode(y), g["y_decode_of_y"]) < 1e-5
    # state_dict layout is the reference's
    assert set(aa.state_dict()) == {k[3:] for k in g if k.startswith("sd.")}


def test_projector_toy_dims(aab, golden):
    g = golden("projector")
    aa = _load_aa(aab, g, "toy_sd.", 2, 16)
    z, yr = aa(T(g["toy_y"]).cuda())
    assert rel_l2(z, g["toy_z"]) < 1e-5 and rel_l2(yr, g["toy_y_recon"]) < 1e-5


def test_projector_backward_golden(aab, golden):
    g = golden("projector")
    aa = _load_aa(aab, g)
    y = T(g["y"]).cuda().requires_grad_(True)
    z, yr = aa(y)
    ((z * T(g["gz"]).cuda()).sum() + (yr * T(g["gyr"]).cuda()).sum()).backward()
    assert rel_l2(y.grad, g["grad_y"]) < 1e-4
    for k, p in aa.named_parameters():
        assert rel_l2(p.grad, g["grad." + k]) < 1e-4, k


def test_projector_ragged_and_large(aab):
    "T not a multiple of the 32-token tile; many tiles per CTA; compared with the float64 oracle"
    O = _oracle()
    sd = O.init_projector_state_dict(64, 64, seed=2)
    aa = aab.AudioAlgebra(64, 64)
    aa.load_state_dict(sd)
    aa = aa.cuda()
    ew, eb, dw, db = O.projector_params_from_state_dict(sd)
    g = torch.Generator().manual_seed(3)
    for shape in [(1, 64, 1), (2, 64, 37), (40, 64, 515)]:
        y = torch.randn(*shape, generator=g)
        z, yr = aa(y.cuda())
        zo, yro = O.projector_forward(y.double(), ew, eb, dw, db)
        assert rel_l2(z, zo) < 1e-5 and rel_l2(yr, yro) < 1e-5
    # gradient check at a ragged size
    y = torch.randn(5, 64, 70, generator=g)
    gz = torch.randn(5, 64, 70, generator=g)
    yc = y.cuda().requires_grad_(True)
    aa.zero_grad()
    z, yr = aa(yc)
    (z * gz.cuda()).sum().backward()
    sdd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    ew, eb, dw, db = O.projector_params_from_state_dict(sdd)
    yd = y.double().requires_grad_(True)
    zo, _ = O.projector_forward(yd, ew, eb, dw, db)
    (zo * gz.double()).sum().backward()
    assert rel_l2(yc.grad, yd.grad) < 1e-4
    for k, p in aa.named_parameters():
        if k.startswith("encoder"):
            assert rel_l2(p.grad, sdd[k].grad) < 1e-4, k
    assert aa.decoder[0].lin.weight.grad is None or float(aa.decoder[0].lin.weight.grad.abs().max()) == 0.0


def test_losses_golden(aab, golden):
    g = golden("losses")
    za, zb = T(g["za"]).cuda(), T(g["zb"]).cuda()
    assert abs(aab.mseloss(za, zb).item() - g["mse"]) < 1e-5 * abs(g["mse"])
    assert abs(aab.vicreg_var_loss(za).item() - g["var_a"]) < 1e-4 * max(abs(g["var_a"]), 1e-3)
    assert abs(aab.vicreg_var_loss(0.2 * zb).item() - g["var_b_small"]) < 1e-5 * abs(g["var_b_small"])
    assert abs(aab.vicreg_var_loss_l2(0.2 * zb).item() - g["var_b_small_l2"]) < 1e-5 * abs(g["var_b_small_l2"])
    assert abs(aab.vicreg_cov_loss(za).item() - g["cov_a"]) < 1e-4 * abs(g["cov_a"])
    assert abs(aab.vicreg_cov_loss(zb).item() - g["cov_b"]) < 1e-4 * abs(g["cov_b"])


def test_loss_gradients_golden(aab, golden):
    g = golden("losses")
    za, zb = T(g["za"]).cuda(), T(g["zb"]).cuda()
    t = (0.2 * zb).clone().requires_grad_(True)
    aab.vicreg_var_loss(t).backward()
    assert rel_l2(t.grad, g["grad_var"]) < 1e-4
    t = za.clone().requires_grad_(True)
    aab.vicreg_cov_loss(t).backward()
    assert rel_l2(t.grad, g["grad_cov"]) < 1e-4
    a = za.clone().requires_grad_(True)
    aab.mseloss(a, zb).backward()
    assert rel_l2(a.grad, g["grad_mse_a"]) < 1e-5


@pytest.mark.parametrize("b,t", [(160, 640), (9, 1000), (3, 200)])
def test_projector_backward_tcgen05_many_tiles(aab, b, t):
    """csrc/proj_bwd_tc.cu (bf16x3 pieces, K-major + MN-major views of one SWIZZLE_128B copy, weight-gradient accumulators in TMEM):
    several tiles per CTA incl. the periodic accumulator drain (160 x 640: 800 tiles on 148 CTAs), ragged last tiles, against a
    float64 torch restatement of the four blocks; both halves (encode and decode) are exercised."""
    torch.manual_seed(b + t)
    aa = aab.AudioAlgebra(64, 64).cuda()
    y = (0.7 * torch.randn(b, 64, t, device="cuda")).requires_grad_(True)
    gz, gy = torch.randn(b, 64, t, device="cuda"), torch.randn(b, 64, t, device="cuda")
    z, yr = aa(y)
    ((z * gz).sum() + (yr * gy).sum()).backward()

    def half(seq, xin):
        ws = [(blk.lin.weight.detach().double().requires_grad_(True), blk.lin.bias.detach().double().requires_grad_(True)) for blk in seq]
        h = xin.transpose(1, 2)
        for i, (w, bb) in enumerate(ws):
            u = h @ w.T + bb
            h = h + (torch.nn.functional.gelu(u) if i < 3 else u)
        return xin + h.transpose(1, 2), ws

    yd = y.detach().double().requires_grad_(True)
    zd, we = half(aa.encoder, yd)
    yrd, wd = half(aa.decoder, zd)
    ((zd * gz.double()).sum() + (yrd * gy.double()).sum()).backward()
    assert rel_l2(z, zd) < 1e-5 and rel_l2(yr, yrd) < 1e-5
    assert rel_l2(y.grad, yd.grad) < 1e-5
    for seq, ws in ((aa.encoder, we), (aa.decoder, wd)):
        for blk, (w, bb) in zip(seq, ws):
            assert rel_l2(blk.lin.weight.grad, w.grad) < 2e-5 and rel_l2(blk.lin.bias.grad, bb.grad) < 1e-5


@pytest.mark.parametrize("b,c,t", [(2, 3, 5), (16, 64, 16), (33, 8, 40), (96, 64, 64)])
def test_cov_and_var_vs_naive_oracle(aab, b, c, t):
    "Gram-identity covariance loss == the reference's materialised (C T)^2 covariance (small T only)"
    O = _oracle()
    g = torch.Generator().manual_seed(b * 1000 + t)
    mix = torch.randn(c * t, c * t, generator=g) / (c * t) ** 0.5
    z = (torch.randn(b, c * t, generator=g) @ mix).reshape(b, c, t) + 0.3   # correlated features, non-zero mean
    zd = z.double().requires_grad_(True)
    lo = O.vicreg_cov_loss(zd)
    lo.backward()
    zc = z.cuda().requires_grad_(True)
    lc = aab.vicreg_cov_loss(zc)
    lc.backward()
    assert abs(lc.item() - lo.item()) < 2e-4 * abs(lo.item())
    assert rel_l2(zc.grad, zd.grad) < 2e-4
    for l2 in (False, True):
        zd2 = (0.5 * z).double().requires_grad_(True)
        lo2 = O.vicreg_var_loss(zd2, l2_hinge=l2)
        lo2.backward()
        zc2 = (0.5 * z).cuda().requires_grad_(True)
        lc2 = (aab.vicreg_var_loss_l2 if l2 else aab.vicreg_var_loss)(zc2)
        lc2.backward()
        assert abs(lc2.item() - lo2.item()) < 1e-5 * max(abs(lo2.item()), 1e-3)
        assert rel_l2(zc2.grad, zd2.grad) < 1e-4


def test_cov_loss_training_size_is_cheap_and_consistent(aab):
    """[B,64,512] (the reference would build a 32768^2 = 4 GiB covariance): check the scalar against an
    independent float64 Gram evaluation in torch and the gradient through a directional derivative."""
    g = torch.Generator(device="cuda").manual_seed(0)
    z = torch.randn(64, 64, 512, device="cuda", generator=g)
    loss = aab.vicreg_cov_loss(z)
    x = z.reshape(64, -1).double()
    xc = x - x.mean(0, keepdim=True)
    gram = xc @ xc.T
    s = (xc * xc).sum(0)
    ref = ((gram ** 2).sum() - (s ** 2).sum()) / (63 ** 2) / x.shape[1]
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item())
    zc = z.clone().requires_grad_(True)
    aab.vicreg_cov_loss(zc).backward()
    d = torch.randn_like(z)
    eps = 1e-2
    fd = (aab.vicreg_cov_loss(z + eps * d).item() - aab.vicreg_cov_loss(z - eps * d).item()) / (2 * eps)
    an = (zc.grad * d).sum().item()
    assert abs(fd - an) < 2e-2 * max(abs(an), 1e-6)


@pytest.mark.parametrize("b,d", [(512, 32768), (200, 4096), (130, 2048), (64, 1028), (256, 5000)])
def test_cov_loss_tcgen05_gram_and_backward(aab, b, d):
    """B x B x D Gram (split over D) and the G Xc backward on tcgen05 (csrc/cov_tc.cu: 3-term TF32 split, the backward's Xc operand
    MN-major in the SWIZZLE_128B / 32-byte-atom layout): loss and gradient against the float64 closed form, incl. batch sizes that
    are not tile multiples (zero-padded rows), b % 4 != 0 (forward on the tensor core, backward on the CUDA-core kernel) and a
    feature count that is not a chunk multiple."""
    g = torch.Generator(device="cuda").manual_seed(b + d)
    z = torch.randn(b, d, device="cuda", generator=g) * torch.linspace(0.2, 2.0, d, device="cuda") + 0.5
    z = z + 0.6 * torch.roll(z, 1, dims=1)   # correlated neighbouring columns
    zc = z.clone().reshape(b, 1, d).requires_grad_(True)
    loss = aab.vicreg_cov_loss(zc)
    loss.backward()
    x = z.double()
    xc = x - x.mean(0, keepdim=True)
    gram = xc @ xc.T
    s = (xc * xc).sum(0)
    k = 1.0 / ((b - 1) ** 2 * d)
    ref = ((gram ** 2).sum() - (s ** 2).sum()) * k
    gref = 4.0 * k * (gram @ xc - xc * s)
    assert abs(loss.item() - ref.item()) < 2e-5 * abs(ref.item()), (loss.item(), ref.item())
    assert rel_l2(zc.grad.reshape(b, d), gref) < 2e-5


def test_latent_ops(aab):
    O = _oracle()
    from audio_algebra_b200 import latent_ops as L
    g = torch.Generator().manual_seed(4)
    z = torch.tanh(torch.randn(3, 64, 50, generator=g))
    zc = z.cuda()
    assert torch.equal(L.flip_channels(zc).cpu(), O.destructo_flip_channels(z))
    assert torch.equal(L.flip_time(zc).cpu(), z.flip(dims=[2]))
    assert rel_l2(L.sign_fold(zc), O.destructo_sign_fold(z.double())) < 1e-6
    assert rel_l2(L.absmax_minus(zc), O.destructo_absmax_minus(z.double())) < 1e-6
    assert rel_l2(L.tanh_drive(zc, 3.0), O.destructo_tanh_drive(z.double(), 3.0)) < 1e-5
    wet, dry = torch.randn(5, 64, 50, generator=g), torch.randn(5, 64, 50, generator=g)
    assert rel_l2(L.effect_transfer(zc, wet.cuda(), dry.cuda()), O.effect_transfer(z.double(), wet.double(), dry.double())) < 1e-6
    zs = [torch.randn(2, 64, 33, generator=g) for _ in range(3)]
    out = aab.latent_lincomb([t.cuda() for t in zs], [1.0, -1.0, 1.0])
    assert rel_l2(out, zs[0].double() - zs[1].double() + zs[2].double()) < 1e-6
    odd = [torch.randn(7, generator=g) for _ in range(2)]   # unaligned / tiny
    assert rel_l2(aab.latent_lincomb([t.cuda()[1:] for t in odd], [2.0, 0.5]), 2 * odd[0][1:].double() + 0.5 * odd[1][1:].double()) < 1e-6


def test_destructo_cell22_ops(aab):
    "the remaining Destructo.ipynb cell 22 operations against the oracle's restatement of the cell"
    O = _oracle()
    from audio_algebra_b200 import latent_ops as L
    g = torch.Generator().manual_seed(9)
    z = torch.tanh(torch.randn(3, 64, 80, generator=g))
    zc = z.cuda()
    assert torch.equal(L.wavy(zc).cpu(), O.destructo_wavy(z))                    # fp32 modulation built on the CPU, one multiply
    assert torch.equal(L.flippy(zc).cpu(), O.destructo_flippy(z))
    assert torch.equal(L.kill_half(zc).cpu(), O.destructo_kill_half(z))
    assert torch.equal(L.big_changes(zc).cpu(), 2 * z)
    for rt in (2.5, 40.0):
        assert torch.equal(L.reverb_time(zc, rt).cpu(), O.destructo_reverb(z, rt))  # same unfused fp32 operations in the same order
    assert torch.equal(L.reverb_time(zc, 0).cpu(), z)
    zl = torch.tanh(torch.randn(1, 64, 1024, generator=g))                       # config-1 latent length
    assert rel_l2(L.reverb_time(zl.cuda(), 100.0), O.destructo_reverb(zl.double(), 100.0)) < 1e-5
    u = torch.rand(z.shape, generator=g)
    assert torch.equal(L.call_and_response(zc, 0.5, u=u.cuda()).cpu(), O.destructo_call_and_response(z, u, 0.5))
    emb = torch.tanh(torch.randn(3, 64, 80, generator=g))
    assert rel_l2(L.hurt_drums(emb.cuda(), zc, 0.3, u=u.cuda()), O.destructo_hurt_drums(emb.double(), z.double(), u.double(), 0.3)) < 1e-6
    gen = torch.Generator(device="cuda").manual_seed(1)
    a = L.call_and_response(zc, 0.5, generator=gen)
    gen.manual_seed(1)
    b = L.call_and_response(zc, 0.5, generator=gen)
    assert torch.equal(a, b) and not torch.equal(a, -zc)


def test_effect_transfer_cell49_variants(aab):
    "cell 49: left zero-padding / truncation of the wet-dry difference to the embedding length, and the time_avg branch"
    O = _oracle()
    from audio_algebra_b200 import latent_ops as L
    g = torch.Generator().manual_seed(10)
    z = torch.tanh(torch.randn(4, 64, 50, generator=g))
    for td in (50, 37, 64):
        wet, dry = torch.randn(5, 64, td, generator=g), torch.randn(5, 64, td, generator=g)
        assert rel_l2(L.effect_transfer(z.cuda(), wet.cuda(), dry.cuda()), O.effect_transfer(z.double(), wet.double(), dry.double())) < 1e-6
    wet, dry = torch.randn(5, 64, 50, generator=g), torch.randn(5, 64, 50, generator=g)
    with pytest.raises(RuntimeError):                                            # [5, 64] vs [4, 64, 50]: the notebook's own broadcasting error
        L.effect_transfer(z.cuda(), wet.cuda(), dry.cuda(), time_avg=True)
    z2 = torch.tanh(torch.randn(3, 8, 16, generator=g))                          # a shape for which `z + diff.mean(-1)` is defined
    wet2, dry2 = torch.randn(8, 16, 21, generator=g), torch.randn(8, 16, 21, generator=g)
    assert rel_l2(L.effect_transfer(z2.cuda(), wet2.cuda(), dry2.cuda(), time_avg=True),
                  O.effect_transfer(z2.double(), wet2.double(), dry2.double(), time_avg=True)) < 1e-6


def test_embed_block_standalone_golden(aab, golden):
    "EmbedBlock.forward on its own (aa_mixer.py:216-221) against the reference's outputs, gradients against fp64 autograd"
    O = _oracle()
    g = golden("projector")
    blk = aab.EmbedBlock(64, 64)
    blk.load_state_dict({"lin.weight": T(g["blk_w"]), "lin.bias": T(g["blk_b"])})
    blk = blk.cuda()
    x = T(g["blk_x"]).cuda().requires_grad_(True)
    y = blk(x)
    assert tuple(y.shape) == (3, 7, 64) and rel_l2(y, g["blk_y"]) < 1e-5
    gy = torch.randn(3, 7, 64, generator=torch.Generator().manual_seed(1))
    y.backward(gy.cuda())
    xd = T(g["blk_x"]).double().requires_grad_(True)
    wd, bd = T(g["blk_w"]).double().requires_grad_(True), T(g["blk_b"]).double().requires_grad_(True)
    O.embed_block(xd, wd, bd, act=True, resid=True).backward(gy.double())
    assert rel_l2(x.grad, xd.grad) < 1e-5 and rel_l2(blk.lin.weight.grad, wd.grad) < 1e-5 and rel_l2(blk.lin.bias.grad, bd.grad) < 1e-5
    blk2 = aab.EmbedBlock(16, 24, act=None)
    blk2.load_state_dict({"lin.weight": T(g["blk2_w"]), "lin.bias": T(g["blk2_b"])})
    assert rel_l2(blk2.cuda()(T(g["blk2_x"]).cuda()), g["blk2_y"]) < 1e-5
    # use_bn: BatchNorm1d on [N, C] rows, training (batch statistics, running statistics updated) then eval
    blk3 = aab.EmbedBlock(8, 8, use_bn=True)
    blk3.load_state_dict({"lin.weight": T(g["blk3_w"]), "lin.bias": T(g["blk3_b"]), "bn.weight": T(g["blk3_bn_w"]), "bn.bias": T(g["blk3_bn_b"])},
                         strict=False)
    blk3 = blk3.cuda().train()
    x3 = T(g["blk3_x"]).cuda().requires_grad_(True)
    y3 = blk3(x3)
    assert rel_l2(y3, g["blk3_y_train"]) < 1e-5
    assert rel_l2(blk3.bn.running_mean, g["blk3_run_mean"]) < 1e-5 and rel_l2(blk3.bn.running_var, g["blk3_run_var"]) < 1e-5
    y3.square().sum().backward()
    ref = torch.nn.Sequential()   # fp64 torch restatement of the block for the gradient check
    lin, bn = torch.nn.Linear(8, 8).double(), torch.nn.BatchNorm1d(8).double()
    lin.load_state_dict({"weight": T(g["blk3_w"]).double(), "bias": T(g["blk3_b"]).double()})
    bn.weight.data, bn.bias.data = T(g["blk3_bn_w"]).double(), T(g["blk3_bn_b"]).double()
    xr = T(g["blk3_x"]).double().requires_grad_(True)
    (xr + bn(torch.nn.functional.gelu(lin(xr)))).square().sum().backward()
    assert rel_l2(x3.grad, xr.grad) < 1e-4 and rel_l2(blk3.lin.weight.grad, lin.weight.grad) < 1e-4
    assert rel_l2(blk3.bn.weight.grad, bn.weight.grad) < 1e-4 and rel_l2(blk3.bn.bias.grad, bn.bias.grad) < 1e-4
    assert rel_l2(blk3.eval()(T(g["blk3_x"]).cuda()), g["blk3_y_eval"]) < 1e-5
    with pytest.raises(ValueError):
        blk3(torch.zeros(2, 3, 8, device="cuda"))
    with pytest.raises(ValueError):                       # AudioAlgebra(use_bn=True) walks the blocks one by one and hits the same wall
        aab.AudioAlgebra(8, 8, use_bn=True).cuda()(torch.zeros(2, 8, 5, device="cuda"))


class _ToyGiven(torch.nn.Module):
    def __init__(self, w):
        super().__init__()
        self.register_buffer("w", w)

    def encode(self, x):
        return torch.tanh(torch.nn.functional.conv1d(x, self.w, stride=64))


def test_do_mixing_mixer_golden(aab, golden):
    g, gp = golden("mixing"), golden("projector")
    aa = _load_aa(aab, gp)
    toy = _ToyGiven(T(g["toy_w"])).cuda()
    stems = [T(g["stem0"]).cuda(), T(g["stem1"]).cuda()]
    zsum, zmix, arch = aab.do_mixing(stems, T(g["faders"]).cuda(), toy, aa, "cuda")
    assert rel_l2(zsum, g["zsum"]) < 1e-4 and rel_l2(zmix, g["zmix"]) < 1e-4
    assert rel_l2(arch["ymix"], g["ymix"]) < 1e-4 and rel_l2(arch["ymix_recon"], g["ymix_recon"]) < 1e-4
    assert rel_l2(arch["mix"], g["mix"]) < 1e-6 and rel_l2(arch["ysum"], g["ysum"]) < 1e-4
    for i in range(2):
        assert rel_l2(arch["zs"][i], g[f"zs{i}"]) < 1e-4 and rel_l2(arch["yrecons"][i], g[f"yrecons{i}"]) < 1e-4
    # loss terms exactly as train_aa_mixer_accel.py:504-517
    y = toy.encode(stems[0])
    z, yrecon = aa(y)
    L_mix = aab.mseloss(zsum, zmix)
    L_var = (aab.vicreg_var_loss(zsum) + aab.vicreg_var_loss(zmix)) / 2
    L_cov = (aab.vicreg_cov_loss(zsum) + aab.vicreg_cov_loss(zmix)) / 2
    L_rec = aab.mseloss(y, yrecon) + aab.mseloss(arch["ymix"], arch["ymix_recon"])
    for v, k in [(L_mix, "L_mix"), (L_var, "L_var"), (L_cov, "L_cov"), (L_rec, "L_rec")]:
        assert abs(v.item() - g[k]) < 2e-4 * max(abs(g[k]), 1e-3), k
    (L_mix + L_var + L_cov + L_rec).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in aa.parameters())


def test_do_mixing_effects_golden(aab, golden):
    g, gp = golden("mixing"), golden("projector")
    aa = _load_aa(aab, gp)
    toy = _ToyGiven(T(g["toy_w"])).cuda()
    batch = {k: T(g["fx_" + k]) for k in ("a1", "b1", "a2", "b2")}
    arch = aab.aa_effects.do_mixing(batch, toy, aa, "cuda")
    for i in range(4):
        assert rel_l2(arch["ys"][i], g[f"fx_ys{i}"]) < 1e-4
        assert rel_l2(arch["zs"][i], g[f"fx_zs{i}"]) < 1e-4
        assert rel_l2(arch["yrecons"][i], g[f"fx_yrecons{i}"]) < 1e-4
    O = _oracle()
    L = aab.aa_effects.effects_losses(arch)
    Lo = O.effects_losses([T(g[f"fx_zs{i}"]).double() for i in range(4)], [T(g[f"fx_ys{i}"]).double() for i in range(4)],
                          [T(g[f"fx_yrecons{i}"]).double() for i in range(4)])
    for k in ("mix_loss", "var_loss", "cov_loss", "aa_recon_loss", "loss"):
        assert abs(L[k].item() - Lo[k].item()) < 3e-4 * max(abs(Lo[k].item()), 1e-3), k
    L["loss"].backward()


def test_get_stems_faders_recipe(aab, golden):
    import random
    g = golden("mixing")
    random.seed(0)
    torch.manual_seed(0)
    b0, b1 = T(g["stem0"]).cuda(), T(g["stem1"])
    stems, faders, it = aab.get_stems_faders(b0, iter([b1]), [b1], maxstems=2)
    assert len(stems) == 2 and stems[1].is_cuda and faders.is_cuda
    assert np.allclose(faders.cpu().numpy(), g["faders_seed0"], atol=1e-6)


def test_pca_accumulation_golden(aab, golden):
    g = golden("pca")
    rc = aab.pca.RunningCovariance(64, "cuda")
    for bi in range(2):
        rc.update(T(g[f"ys{bi}"]).cuda())
    assert int(rc.count.item()) == int(g["npoints"][0])
    assert rel_l2(rc.cov_numerator, g["cov_numerator"]) < 1e-4
    assert rel_l2(rc.eigenvalues(), g["lambdas"]) < 1e-3
    # ragged: C not a multiple of 64, T not a multiple of the tile
    O = _oracle()
    y = torch.tanh(torch.randn(3, 10, 45, generator=torch.Generator().manual_seed(1)))
    rc2 = aab.pca.RunningCovariance(10, "cuda").update(y.cuda())
    num, n = O.pca_cov_numerator(y.double())
    assert rel_l2(rc2.cov_numerator, num) < 1e-5 and int(rc2.count.item()) == n


def test_pca_tensor_core_path_matches_oracle_and_cuda_core_path(aab, monkeypatch):
    """C = 64: one pass, the 64 x 64 rank-n update on tcgen05 (hi / lo TF32 split, csrc/gram_tc.cu).  Against the float64 oracle
    and the two-pass CUDA-core kernel: a large mean (the pivot keeps the raw moments conditioned), a T that is not a multiple of
    the 64-point tile, several updates (accumulation), more tiles than CTAs."""
    O = _oracle()
    g = torch.Generator().manual_seed(5)
    for shape, shift in (((3, 64, 512), 0.0), ((5, 64, 100), 3.0), ((300, 64, 128), -0.5), ((1, 64, 4), 0.1)):
        y = torch.tanh(torch.randn(*shape, generator=g)) + shift
        num, n = O.pca_cov_numerator(y.double())
        rc = aab.pca.RunningCovariance(64, "cuda").update(y.cuda()).update(y.cuda())
        assert int(rc.count.item()) == 2 * n
        assert rel_l2(rc.cov_numerator, 2 * num) < 1e-5, shape
        monkeypatch.setenv("AA_PCA_CUDA_CORES", "1")
        rc1 = aab.pca.RunningCovariance(64, "cuda").update(y.cuda()).update(y.cuda())
        monkeypatch.delenv("AA_PCA_CUDA_CORES")
        assert rel_l2(rc.cov_numerator, rc1.cov_numerator) < 1e-5


def test_adam_matches_torch(aab):
    from audio_algebra_b200.training import FlatAdam
    O = _oracle()
    torch.manual_seed(0)
    p0 = torch.randn(33280)
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=5e-4)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-3, total_steps=50)
    mine = FlatAdam(p0.clone().cuda(), lr=5e-4, max_lr=1e-3, total_steps=50)
    for step in range(20):
        gr = torch.randn(33280)
        pt.grad = gr.clone()
        opt.step(); sched.step()
        mine.step(gr.cuda())
    assert rel_l2(mine.params, pt.detach()) < 1e-6


def test_projector_forward_tensor_core_path_is_fp32_accurate(aab):
    """T' >= 64 with the standard 64 -> 64 residual projector runs the forward on tcgen05 (kind::tf32 with a 3-term hi/lo split,
    csrc/proj_tc.cu).  Against the float64 oracle it must be as accurate as fp32 arithmetic (the gate is 1e-4; a single-pass
    TF32 product would sit at ~1e-3), for ragged token counts, both halves, and large-magnitude inputs."""
    O = _oracle()
    sd = O.init_projector_state_dict(64, 64, seed=2)
    aa = aab.AudioAlgebra(64, 64)
    aa.load_state_dict(sd)
    aa = aa.cuda()
    ew, eb, dw, db = O.projector_params_from_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    for shape, scale in [((3, 64, 200), 1.0), ((2, 64, 128), 4.0), ((5, 64, 1024), 0.3), ((1, 64, 64), 1.0), ((2, 64, 65), 1.0)]:
        y = scale * torch.randn(*shape, generator=g)
        z, yr = aa(y.cuda())
        zo, yro = O.projector_forward(y.double(), ew, eb, dw, db)
        assert rel_l2(z, zo) < 2e-6 and rel_l2(yr, yro) < 2e-6, (shape, rel_l2(z, zo), rel_l2(yr, yro))
        # elementwise, relative to the tensor's scale: no token / channel may be off (row or panel mix-ups)
        assert (z.cpu().double() - zo).abs().max().item() < 2e-5 * zo.abs().max().item()
    # gradients still flow through the (CUDA-core) backward and match the oracle's autograd
    y = torch.randn(2, 64, 192, generator=g)
    yc = y.cuda().requires_grad_(True)
    z, yr = aa(yc)
    (z.square().mean() + yr.square().mean()).backward()
    yd = y.double().requires_grad_(True)
    P = [p.clone().requires_grad_(True) for p in ew + eb + dw + db]
    zo, yro = O.projector_forward(yd, P[0:4], P[4:8], P[8:12], P[12:16])
    (zo.square().mean() + yro.square().mean()).backward()
    assert rel_l2(yc.grad, yd.grad) < 1e-4
    assert rel_l2(aa.encoder[0].lin.weight.grad, P[0].grad) < 1e-4


def test_projector_backward_cabi_flags(aab):
    "aa_projector_half_bwd_f32 on the tcgen05 path: gscale, accumulate_gx / accumulate_gw, and gx = NULL behave as the header says"
    from audio_algebra_b200._lib import lib, check, ptr, stream_ptr
    from audio_algebra_b200.aa_mixer import _ptr_array
    torch.manual_seed(5)
    aa = aab.AudioAlgebra(64, 64).cuda()
    ws = [blk.lin.weight.detach().contiguous() for blk in aa.encoder]
    bs = [blk.lin.bias.detach().contiguous() for blk in aa.encoder]
    wp, k1 = _ptr_array(ws)
    bp, k2 = _ptr_array(bs)
    b, t = 6, 300
    x, gout = torch.randn(b, 64, t, device="cuda"), torch.randn(b, 64, t, device="cuda")
    wsb = torch.empty(int(lib.aa_projector_bwd_workspace_floats()), device="cuda")

    def run(gx, gws, gbs, acc_gx, acc_gw, gscale):
        gwp, k3 = _ptr_array(gws)
        gbp, k4 = _ptr_array(gbs)
        check(lib.aa_projector_half_bwd_f32(wp, bp, 64, 64, 1, ptr(x), ptr(gout), b, t, None if gx is None else ptr(gx), acc_gx, gwp, gbp, acc_gw,
                                            gscale, ptr(wsb), stream_ptr()))
        torch.cuda.synchronize()

    gx0 = torch.empty_like(x)
    gw0, gb0 = [torch.empty_like(w) for w in ws], [torch.empty_like(v) for v in bs]
    run(gx0, gw0, gb0, 0, 0, 1.0)
    gx1 = torch.full_like(x, 0.25)
    gw1, gb1 = [torch.full_like(w, -1.5) for w in ws], [torch.full_like(v, 2.0) for v in bs]
    run(gx1, gw1, gb1, 1, 1, 0.5)
    assert rel_l2(gx1, gx0 + 0.25) < 1e-6            # gscale applies to the parameter gradients only
    for i in range(4):
        assert rel_l2(gw1[i], 0.5 * gw0[i] - 1.5) < 1e-6 and rel_l2(gb1[i], 0.5 * gb0[i] + 2.0) < 1e-6
    gw2, gb2 = [torch.empty_like(w) for w in ws], [torch.empty_like(v) for v in bs]
    run(None, gw2, gb2, 0, 0, 1.0)
    for i in range(4):
        assert torch.equal(gw2[i], gw0[i]) and torch.equal(gb2[i], gb0[i])   # deterministic, and independent of gx


def test_lincomb_more_than_eight_terms_and_host_faders(aab):
    "latent_lincomb chunks beyond 8 terms (the reference allows any maxstems); get_stems_faders hands do_mixing its host copy of the faders"
    g = torch.Generator().manual_seed(8)
    zs = [torch.randn(3, 64, 40, generator=g) for _ in range(19)]
    cs = [float(c) for c in torch.randn(19, generator=g)]
    ref = sum(c * z.double() for c, z in zip(cs, zs))
    assert rel_l2(aab.latent_lincomb([z.cuda() for z in zs], cs), ref) < 1e-6
    batch = torch.rand(4, 2, 1024, device="cuda")
    dl = [torch.rand(4, 2, 1024) for _ in range(3)]
    stems, faders, it = aab.get_stems_faders(batch, iter(dl), dl, maxstems=3)
    assert faders.is_cuda and faders._aa_host == faders.cpu().tolist()
