"""GPU parity of StackedDiffAEWrapper.encode (given_models.py:361-385, StackedAELatentDiffusionCond.py:221-227)
against the oracle restatement with shared seeded weights.  Both networks are third-party in the reference (PARITY UNPINNED upstream);
the reference pins the shape [12, 2, 262144] -> [12, 32, 512] (StackedDiffAE.ipynb cell 13)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import audio_algebra_b200 as aab
    from oracle import aa_oracle as O
    torch.manual_seed(0)
    orc = O.StackedDiffAEEncodeOracle().eval()
    m = aab.StackedDiffAEWrapper(debug=False)
    m.first_stage_autoencoder.encoder.load_oracle_weights(orc.first)
    m.model.latent_encoder.load_state_dict(orc.second.state_dict())
    m.model.latent_encoder_ema.load_state_dict(orc.second.state_dict())
    return aab, O, orc, m.cuda()


def _x(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) - 0.5


def test_generic_conv_and_groupnorm_layers(setup):
    "the two C-ABI layers the second stage is made of, against torch: strided k5 / k3 / 1x1 convs with residual, GroupNorm(8) + SiLU"
    aab, O, orc, m = setup
    from audio_algebra_b200.StackedAELatentDiffusionCond import conv1d, groupnorm_silu, ACT_TANH
    torch.manual_seed(1)
    for cin, cout, k, s, pd, l in [(32, 128, 1, 1, 0, 100), (128, 256, 5, 2, 2, 101), (256, 256, 3, 1, 1, 77), (24, 40, 5, 2, 2, 64)]:
        conv = torch.nn.Conv1d(cin, cout, k, stride=s, padding=pd)
        x = torch.randn(2, cin, l)
        ref = conv.double()(x.double())
        res = torch.randn_like(ref).float()
        y = conv1d(x.cuda(), conv.float().cuda(), res=res.cuda(), act=ACT_TANH)
        assert tuple(y.shape) == tuple(ref.shape) and rel_l2(y, torch.tanh(ref + res.double())) < 1e-5
    gn = torch.nn.GroupNorm(8, 256)
    gn.weight.data, gn.bias.data = torch.rand(256) + 0.5, torch.randn(256) * 0.1
    x = torch.randn(3, 256, 50) * 2 + 0.3
    ref = torch.nn.functional.silu(gn.double()(x.double()))
    assert rel_l2(groupnorm_silu(x.cuda(), gn.float().cuda()), ref) < 1e-5


def test_encode_matches_the_restated_oracle(setup):
    aab, O, orc, m = setup
    x = _x((2, 2, 16384), 3)
    with torch.no_grad():
        ref = orc(x)
    y = m.encode(x.cuda())
    assert tuple(y.shape) == tuple(ref.shape) == (2, 32, 32) and y.dtype == torch.float32
    assert rel_l2(y, ref) < 1e-3                                  # BASELINE.json: fp32 mode, embeddings
    first = m.first_stage_autoencoder.encode(x.cuda())            # first-stage latents on their own: tanh(SoundStreamXL)
    with torch.no_grad():
        assert rel_l2(first, torch.tanh(orc.first(x))) < 1e-3
    assert m.encode(x).device.type == "cpu"                       # CPU in -> CPU out, like the other wrappers
    ragged = _x((1, 2, 10000), 4)                                 # length that is not a multiple of the total stride (512)
    with torch.no_grad():
        assert rel_l2(m.encode(ragged.cuda()), orc(ragged)) < 1e-3


def test_reference_shape_kat_and_setup(setup):
    aab, O, orc, m = setup
    y = m.encode(torch.zeros(12, 2, 262144, device="cuda"))
    assert tuple(y.shape) == (12, 32, 512)                        # StackedDiffAE.ipynb cell 13
    assert m.latent_dim == 32 and m.latent_downsampling_ratio == 16 and m.model.downsampling_ratio == 512
    w = aab.StackedDiffAEWrapper(debug=False)
    w.setup()                                                     # EMA copies become the live modules (given_models.py:404-407)
    assert not hasattr(w.model, "latent_encoder_ema") and not w.model.training
    with pytest.raises(NotImplementedError):
        w.decode(torch.zeros(1, 32, 4))


def test_first_stage_on_tcgen05_bf16(setup):
    """compute_dtype="bf16": the capacity-64 / stride-2 SoundStreamXL first stage (1.6 TFLOP per 2^18-sample chunk, 87 % of the encode)
    runs on the tcgen05 implicit-GEMM kernels (K up to 7168: 112 K chunks per tile); BASELINE.json bf16 gate: cosine >= 0.999."""
    aab, O, orc, m = setup
    mb = aab.StackedDiffAEWrapper(debug=False, compute_dtype="bf16")
    mb.first_stage_autoencoder.encoder.load_oracle_weights(orc.first)
    mb.model.latent_encoder.load_state_dict(orc.second.state_dict())
    mb = mb.cuda()
    x = _x((2, 2, 16384), 5)
    with torch.no_grad():
        ref1, ref = torch.tanh(orc.first(x)), orc(x)
    cos1 = torch.nn.functional.cosine_similarity(mb.first_stage_autoencoder.encode(x.cuda()).cpu().flatten(1).double(), ref1.flatten(1).double(), dim=1)
    assert cos1.min().item() >= 0.999
    cos = torch.nn.functional.cosine_similarity(mb.encode(x.cuda()).cpu().flatten(1).double(), ref.flatten(1).double(), dim=1)
    assert cos.min().item() >= 0.999
