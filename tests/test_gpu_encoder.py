"""GPU parity of the conv encoder (DiffusionDVAE.encode / encode_it, DVAEWrapper.encode) against the
oracle restatement with shared seeded random-init weights.  NOTE: the oracle itself is a restatement of a
third-party module whose source is not in the reference tree (PARITY UNPINNED upstream);
the reference pins only the shape [6,2,65536] -> [6,64,512] (Destructo.ipynb cell 17).
Tolerances (BASELINE.json): fp32 relative L2 <= 1e-3 on embeddings; bf16 cosine >= 0.999 per embedding."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import audio_algebra_b200 as aab
    from oracle import aa_oracle as O
    torch.manual_seed(0)
    enc_o = O.SoundStreamXLEncoderOracle().eval()
    dv = aab.DVAEWrapper(debug=False)
    dv.model.load_oracle_weights(enc_o)
    return aab, O, enc_o, dv.cuda()


def _x(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) - 0.5


def test_destructo_shape_kat(setup):
    aab, O, enc_o, dv = setup
    y = dv.encode(torch.zeros(6, 2, 65536, device="cuda"))
    assert tuple(y.shape) == (6, 64, 512) and y.dtype == torch.float32   # Destructo.ipynb cell 17


def test_fp32_encode_it_and_encode(setup):
    aab, O, enc_o, dv = setup
    x = _x((2, 2, 16384), 1)
    y = dv.encode(x.cuda())                              # tanh(encoder_ema(x))
    assert rel_l2(y, O.dvae_encode_it(enc_o, x)) < 1e-3
    dv.model.eval()
    y2 = dv.model.encode(x.cuda())                       # no tanh
    assert rel_l2(y2, O.dvae_encode(enc_o, x)) < 1e-3
    assert rel_l2(torch.tanh(y2), y) < 1e-6
    assert dv.encode(x).device.type == "cpu"             # CPU in -> CPU out


def test_fp32_ragged_lengths_and_batch(setup):
    aab, O, enc_o, dv = setup
    for shape, seed in [((1, 2, 128), 2), ((3, 2, 5000), 3), ((1, 2, 131072), 4)]:
        x = _x(shape, seed)
        ref = O.dvae_encode_it(enc_o, x)
        y = dv.encode(x.cuda())
        assert tuple(y.shape) == tuple(ref.shape)
        assert rel_l2(y, ref) < 1e-3
    assert tuple(dv.encode(torch.zeros(0, 2, 4096, device="cuda")).shape) == (0, 64, 32)


def test_fused_fader_mix(setup):
    "encode(sum f_i s_i) with the scale-and-sum folded into the first conv's load (aa_mixer.py:303,309)"
    aab, O, enc_o, dv = setup
    s0, s1 = _x((2, 2, 8192), 5), _x((2, 2, 8192), 6)
    f = [1.4630, -0.5718]
    y = dv.model.encode_mix([s0.cuda(), s1.cuda()], f)
    assert rel_l2(y, O.dvae_encode(enc_o, f[0] * s0 + f[1] * s1)) < 1e-3


def test_config1_destructo_chain(setup):
    "BASELINE config 1: one 2^17 stereo chunk -> encoder -> tanh -> latent add/subtract + Destructo ops"
    aab, O, enc_o, dv = setup
    from audio_algebra_b200 import latent_ops as L
    x, wet, dry = _x((1, 2, 131072), 7), _x((1, 2, 131072), 8), _x((1, 2, 131072), 9)
    z, zw, zd = [dv.encode(t.cuda()) for t in (x, wet, dry)]
    zo, zwo, zdo = [O.dvae_encode_it(enc_o, t).double() for t in (x, wet, dry)]
    assert tuple(z.shape) == (1, 64, 1024)
    assert rel_l2(L.effect_transfer(z, zw, zd), O.effect_transfer(zo, zwo, zdo)) < 1e-3
    assert rel_l2(L.sign_fold(L.flip_channels(z)), O.destructo_sign_fold(O.destructo_flip_channels(zo))) < 1e-3


def test_weight_update_is_picked_up(setup):
    aab, O, enc_o, dv = setup
    import copy
    x = _x((1, 2, 4096), 10)
    y0 = dv.encode(x.cuda())
    dv2 = copy.deepcopy(dv)
    with torch.no_grad():
        dv2.model.encoder_ema.conv_out.bias.add_(0.25)
    y1 = dv2.encode(x.cuda())
    assert rel_l2(torch.atanh(y1.clamp(-0.999, 0.999)) - 0.25, torch.atanh(y0.clamp(-0.999, 0.999))) < 1e-3
    assert rel_l2(dv.encode(x.cuda()), y0) < 1e-7


def test_bf16_mode_cosine(setup):
    aab, O, enc_o, dv = setup
    dvb = aab.DVAEWrapper(debug=False, compute_dtype="bf16")
    dvb.model.load_oracle_weights(enc_o)
    dvb = dvb.cuda()
    for shape, seed in [((2, 2, 16384), 11), ((3, 2, 5000), 12), ((1, 2, 131072), 13)]:
        x = _x(shape, seed)
        y = dvb.encode(x.cuda()).cpu().double()
        ref = O.dvae_encode_it(enc_o, x).double()
        assert tuple(y.shape) == tuple(ref.shape)
        cos = torch.nn.functional.cosine_similarity(y.flatten(1), ref.flatten(1), dim=1)
        assert cos.min().item() >= 0.999, cos
    s0, s1 = _x((2, 2, 8192), 5), _x((2, 2, 8192), 6)
    f = [1.4630, -0.5718]
    y = dvb.model.encode_mix([s0.cuda(), s1.cuda()], f).cpu().double()
    ref = O.dvae_encode(enc_o, f[0] * s0 + f[1] * s1).double()
    cos = torch.nn.functional.cosine_similarity(y.flatten(1), ref.flatten(1), dim=1)
    assert cos.min().item() >= 0.999, cos


def test_tf32x3_mode_is_fp32_grade(setup):
    """3xTF32 tensor-core path (conv_tf32.cuh): fp32 operands as hi + lo TF32 parts, three MMAs per K step, short accumulator
    chains.  compute_dtype='fp32' (the fixture) selects it for this layer table; here it is forced, and compared with the
    CUDA-core fp32 kernel: same tolerance against the oracle, and agreement between the two far inside it."""
    aab, O, enc_o, _ = setup
    dvt = aab.DVAEWrapper(debug=False, compute_dtype="tf32x3")
    dvt.model.load_oracle_weights(enc_o)
    dvt = dvt.cuda()
    dv = aab.DVAEWrapper(debug=False, compute_dtype="fp32_cuda_cores")
    dv.model.load_oracle_weights(enc_o)
    dv = dv.cuda()
    for shape, seed in [((2, 2, 16384), 21), ((3, 2, 5000), 22), ((1, 2, 128), 23), ((1, 2, 131072), 24), ((5, 2, 3001), 25)]:
        x = _x(shape, seed)
        y = dvt.encode(x.cuda())
        ref = O.dvae_encode_it(enc_o, x)
        assert tuple(y.shape) == tuple(ref.shape)
        assert rel_l2(y, ref) < 1e-3, (shape, rel_l2(y, ref))
        y32 = dv.encode(x.cuda())
        assert rel_l2(y32, ref) < 1e-3, (shape, rel_l2(y32, ref))
        assert rel_l2(y, y32) < 1e-5, (shape, rel_l2(y, y32))
    s0, s1 = _x((2, 2, 8192), 5), _x((2, 2, 8192), 6)
    f = [1.4630, -0.5718]
    y = dvt.model.encode_mix([s0.cuda(), s1.cuda()], f)
    assert rel_l2(y, O.dvae_encode(enc_o, f[0] * s0 + f[1] * s1)) < 1e-3
    assert tuple(dvt.encode(torch.zeros(0, 2, 4096, device="cuda")).shape) == (0, 64, 32)


def test_fp32_falls_back_to_cuda_cores_for_other_shapes(setup):
    "capacity 16 / stride 3 have no tensor-core shape: compute_dtype='fp32' runs the CUDA-core kernel; 'tf32x3' reports an error"
    aab, O, enc_o, _ = setup
    from audio_algebra_b200.DiffusionDVAE import SoundStreamXLEncoder
    torch.manual_seed(3)
    enc = SoundStreamXLEncoder(in_channels=2, capacity=16, latent_dim=32, c_mults=[2, 4], strides=[3, 2]).cuda()
    ref = O.SoundStreamXLEncoderOracle(in_channels=2, capacity=16, latent_dim=32, c_mults=[2, 4], strides=[3, 2]).eval()
    enc.load_oracle_weights(ref)
    x = _x((2, 2, 3000), 31)
    y = enc(x.cuda())
    with torch.no_grad():
        yr = ref(x)
    assert tuple(y.shape) == tuple(yr.shape) and rel_l2(y, yr) < 1e-3
    enc.compute_dtype = "tf32x3"
    with pytest.raises(RuntimeError):
        enc(x.cuda())


def test_bf16_fused_residual_units_match_layerwise(setup):
    """The fused ResidualUnit kernel (conv_ru.cuh: k7 -> ELU -> k1 -> +x -> ELU in one tcgen05 kernel, C = 32 / 64) against
    the layer-by-layer tcgen05 path (AA_NO_RU_FUSION=1 at handle creation): same bf16 operands, so the two agree to
    accumulation-order / bf16 rounding noise -- a tap, dilation or row-offset mistake would show as O(1) error."""
    import os
    aab, O, enc_o, dv = setup

    def make(no_fusion):
        if no_fusion:
            os.environ["AA_NO_RU_FUSION"] = "1"
        try:
            w = aab.DVAEWrapper(debug=False, compute_dtype="bf16")
            w.model.load_oracle_weights(enc_o)
            w = w.cuda()
            w.encode(torch.zeros(1, 2, 1024, device="cuda"))   # the handle reads the switch at its first bf16 forward
        finally:
            os.environ.pop("AA_NO_RU_FUSION", None)
        return w

    fused, layerwise = make(False), make(True)
    os.environ["AA_RU128"] = "1"          # also fuse the C = 128 units (streamed k7 weights; off by default: slower)
    try:
        fused128 = make(False)
    finally:
        os.environ.pop("AA_RU128", None)
    for shape, seed in [((2, 2, 16384), 21), ((3, 2, 5000), 22), ((1, 2, 131072), 23), ((5, 2, 640), 24), ((1, 2, 128), 25)]:
        x = _x(shape, seed).cuda()
        a, b, c = fused.encode(x), layerwise.encode(x), fused128.encode(x)
        assert tuple(a.shape) == tuple(b.shape) == tuple(c.shape)
        assert rel_l2(a, b.cpu()) < 1.5e-2, (shape, rel_l2(a, b.cpu()))
        assert rel_l2(c, b.cpu()) < 1.5e-2, (shape, rel_l2(c, b.cpu()))


@pytest.mark.parametrize("capacity,stride", [(32, 4), (64, 2)])
def test_bf16_fused_units_are_bit_identical_to_layerwise_in_exact_arithmetic(capacity, stride):
    """Fused ResidualUnit kernels (C = 32 and C = 64, dilations 1 / 3 / 9) against the layer-wise tcgen05 path with operands chosen so
    that NOTHING rounds: every conv output channel copies ONE (input channel, tap) with weight 1 (the down-conv: 1/8), inputs are
    integers 0..3, biases 0 -> all activations are small non-negative dyadic numbers (ELU is the identity on them, bf16 holds them
    exactly, fp32 sums of them are exact in any order).  The two paths must then agree BIT FOR BIT; a wrong tap, dilation or row
    offset at a tile seam moves a value.  Lengths cover many 128-row tiles plus ragged tails."""
    import os
    import audio_algebra_b200 as aab

    def build(no_fusion):
        if no_fusion:
            os.environ["AA_NO_RU_FUSION"] = "1"
        try:
            enc = aab.SoundStreamXLEncoder(in_channels=2, capacity=capacity, latent_dim=64, c_mults=[2], strides=[stride], compute_dtype="bf16")
            with torch.no_grad():
                for li, conv in enumerate(enc.flat_convs()):
                    cout, cin, k = conv.weight.shape
                    w = torch.zeros(cout, cin, k)
                    o = torch.arange(cout)
                    w[o, (o * 5 + li) % cin, (o + li) % k] = 0.125 if conv.stride[0] > 1 else 1.0
                    conv.weight.copy_(w)
                    conv.bias.zero_()
            enc = enc.cuda()
            enc(torch.zeros(1, 2, 512, device="cuda"))      # the handle reads the switch at its first bf16 forward
        finally:
            os.environ.pop("AA_NO_RU_FUSION", None)
        return enc

    fused, layerwise = build(False), build(True)
    for shape, seed in [((2, 2, 131072), 1), ((3, 2, 5000), 2), ((1, 2, 128 * 37 + 5), 3)]:
        g = torch.Generator().manual_seed(seed)
        x = torch.randint(0, 4, shape, generator=g).float()
        a, b = fused(x.cuda()), layerwise(x.cuda())
        assert torch.equal(a, b), (shape, (a - b).abs().max().item())
        assert a.abs().max() > 0                             # not a trivially empty signal path
        if shape[2] == 5000:                                 # and the exact path equals the float64 convolution of the same weights
            ref = x.double()
            import torch.nn.functional as F
            convs = layerwise.flat_convs()
            ref = F.elu(F.conv1d(ref, convs[0].weight.double().cpu(), padding=3))
            for r in range(3):
                c1, c2 = convs[1 + 2 * r], convs[2 + 2 * r]
                h = F.elu(F.conv1d(ref, c1.weight.double().cpu(), padding=c1.padding[0], dilation=c1.dilation[0]))
                ref = F.elu(ref + F.conv1d(h, c2.weight.double().cpu()))
            ref = F.elu(F.conv1d(ref, convs[7].weight.double().cpu(), stride=stride, padding=convs[7].padding[0]))
            ref = F.conv1d(ref, convs[8].weight.double().cpu(), padding=1)
            assert torch.equal(a.cpu().double(), ref)


def test_fused_fader_mix_three_and_four_stems(setup):
    "layer 0 sums up to four fader-scaled stems in its load (get_stems_faders maxstems > 2): fp32 and bf16 paths"
    aab, O, enc_o, dv = setup
    stems = [_x((2, 2, 6000), 30 + i) for i in range(4)]
    f = [1.4630, -0.5718, 0.8112, -1.2031]
    dvb = aab.DVAEWrapper(debug=False, compute_dtype="bf16")
    dvb.model.load_oracle_weights(enc_o)
    dvb = dvb.cuda()
    for n in (3, 4):
        ref = O.dvae_encode(enc_o, sum(fi * si for fi, si in zip(f[:n], stems[:n])))
        y = dv.model.encode_mix([s.cuda() for s in stems[:n]], f[:n])
        assert rel_l2(y, ref) < 1e-3, n
        yb = dvb.model.encode_mix([s.cuda() for s in stems[:n]], f[:n]).cpu().double()
        cos = torch.nn.functional.cosine_similarity(yb.flatten(1), ref.double().flatten(1), dim=1)
        assert cos.min().item() >= 0.999, (n, cos)


def test_encode_all_bulk_loop_matches_batch_by_batch(setup):
    """xae_dataset.ipynb cell 50: reps[i:i+bs] = given_model.encode(data[i:i+bs].to(device)).cpu() -- the pipelined loop
    (H2D / encode / D2H on three streams, pinned staging) returns exactly what the sequential loop returns, for a ragged
    last batch, pageable and pinned inputs, a preallocated output, and for an STFT given model as well."""
    aab, O, enc_o, dv = setup
    data = _x((11, 2, 4096), 40)
    ref = torch.cat([dv.encode(data[i:i + 4].cuda()).cpu() for i in range(0, 11, 4)], dim=0)
    out = aab.encode_all(dv, data, batch_size=4)
    assert out.device.type == "cpu" and torch.equal(out, ref)
    pre = torch.empty_like(ref).pin_memory()
    out2 = aab.encode_all(dv, data.pin_memory(), batch_size=5, out=pre)
    assert out2 is pre and rel_l2(out2, ref) < 1e-6
    mel = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512)
    refm = mel.encode(data.cuda()).cpu()
    assert torch.equal(aab.encode_all(mel, data, batch_size=3), refm)
    assert rel_l2(ref[:4], O.dvae_encode_it(enc_o, data[:4])) < 1e-3


def test_sub_batched_walk_matches_single_pass(setup, tmp_path):
    """Large batches are walked in sub-batches through the whole layer stack (encoder.cu sub_batch); forced small here through
    AA_ENC_SUB_SAMPLES in a subprocess, ragged last sub-batch included: bit-identical embeddings, both tensor-core modes."""
    import os, subprocess, sys
    code = r"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import audio_algebra_b200 as aab
torch.manual_seed(0)
for mode in ("bf16", "fp32"):
    dv = aab.DVAEWrapper(debug=False, compute_dtype=mode).cuda()
    x = torch.rand(7, 2, 4096, device="cuda") - 0.5
    torch.save(dv.encode(x).cpu(), os.environ["OUT"] + mode)
"""
    outs = {}
    for tag, sub in (("one", str(1 << 30)), ("sub", str(3 * 4096))):
        env = dict(os.environ, AA_ENC_SUB_SAMPLES=sub, OUT=str(tmp_path / f"{tag}_"))
        subprocess.run([sys.executable, "-c", code], env=env, check=True, timeout=300,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        outs[tag] = {m: torch.load(str(tmp_path / f"{tag}_{m}")) for m in ("bf16", "fp32")}
    for m in ("bf16", "fp32"):
        assert tuple(outs["one"][m].shape) == (7, 64, 32)
        assert torch.equal(outs["one"][m], outs["sub"][m]), m


def test_full_size_batch_modes_agree(setup):
    """BASELINE-size chunks (2^17 samples), a batch that the fp32 mode walks in two sub-batches (64 + ragged 16): the fp32-grade and
    bf16 tensor-core modes agree to cosine >= 0.999 per embedding on every chunk, and the first and last chunks match the oracle."""
    aab, O, enc_o, dv = setup
    dvb = aab.DVAEWrapper(debug=False, compute_dtype="bf16")
    dvb.model.load_oracle_weights(enc_o)
    dvb = dvb.cuda()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(80, 2, 131072, generator=g, device="cuda") - 0.5
    y32 = dv.encode(x)
    yb = dvb.encode(x)
    assert tuple(y32.shape) == (80, 64, 1024)
    cos = torch.nn.functional.cosine_similarity(y32.flatten(1).double(), yb.flatten(1).double(), dim=1)
    assert cos.min().item() >= 0.999, cos.min().item()
    for i in (0, 79):
        ref = O.dvae_encode_it(enc_o, x[i:i + 1].cpu())
        assert rel_l2(y32[i:i + 1], ref) < 1e-3, i


@pytest.mark.parametrize("in_ch,c_mults,strides,latent", [(1, [2, 4], [4, 2], 32), (2, [1, 2, 4], [2, 2, 4], 96), (4, [2], [2], 64)])
def test_other_encoder_configs_on_the_tensor_core_paths(setup, in_ch, c_mults, strides, latent):
    "mono / 4-channel (PQMF-style) inputs, other depths, strides and latent widths: fp32-grade and bf16 modes against the oracle"
    aab, O, _, _ = setup
    from audio_algebra_b200.DiffusionDVAE import SoundStreamXLEncoder
    torch.manual_seed(11)
    ref = O.SoundStreamXLEncoderOracle(in_channels=in_ch, capacity=32, latent_dim=latent, c_mults=c_mults, strides=strides).eval()
    x = _x((3, in_ch, 6000), 41)
    with torch.no_grad():
        yr = ref(x)
    for mode in ("tf32x3", "bf16"):
        enc = SoundStreamXLEncoder(in_channels=in_ch, capacity=32, latent_dim=latent, c_mults=c_mults, strides=strides, compute_dtype=mode)
        enc.load_oracle_weights(ref)
        y = enc.cuda()(x.cuda())
        assert tuple(y.shape) == tuple(yr.shape)
        if mode == "tf32x3":
            assert rel_l2(y, yr) < 1e-3, (mode, rel_l2(y, yr))
        else:
            cos = torch.nn.functional.cosine_similarity(y.flatten(1).double().cpu(), yr.flatten(1).double(), dim=1)
            assert cos.min().item() >= 0.999, (mode, cos)


def test_do_mixing_uses_the_fused_encode_and_keeps_the_archive_contract(setup):
    "aa_mixer.do_mixing with the conv encoder: fader scaling / stem sum inside the first conv, 'mix' / 'fadedstems' built on first access"
    aab, O, enc_o, dv = setup
    torch.manual_seed(2)
    aa = aab.AudioAlgebra(64, 64).cuda()
    stems = [_x((2, 2, 4096), 31).cuda(), _x((2, 2, 4096), 32).cuda()]
    faders = [1.4630, -0.5718]
    dv.model.eval()
    zsum, zmix, archive = aab.do_mixing(stems, faders, dv.model, aa, "cuda")
    assert set(archive.keys()) == {'zs', 'mix', 'ys', 'ymix', 'ymix_recon', 'fadedstems', 'yrecons', 'ysum'} and 'mix' in archive
    mix = faders[0] * stems[0] + faders[1] * stems[1]
    assert rel_l2(archive['mix'], mix) < 1e-6 and rel_l2(archive['fadedstems'][1], faders[1] * stems[1]) < 1e-6
    assert rel_l2(archive['ymix'], O.dvae_encode(enc_o, mix.cpu())) < 1e-3          # DiffusionDVAE.encode: no tanh
    assert rel_l2(archive['ys'][0], O.dvae_encode(enc_o, (faders[0] * stems[0]).cpu())) < 1e-3
    assert rel_l2(zsum, archive['zs'][0] + archive['zs'][1]) < 1e-6 and archive.get('nope') is None


def test_encode_all_notebook_signature_and_npy_writer(setup, tmp_path):
    "xae_dataset.ipynb cell 50: encode_all(audio_full, batch_size, given_model, device) + np.save -- here written while encoding, optionally sharded"
    import numpy as np
    aab, O, enc_o, dv = setup
    data = _x((11, 2, 4096), 21)
    ref = aab.encode_all(dv, data, batch_size=4)
    assert torch.equal(aab.encode_all(data, 4, dv, torch.device("cuda")), ref)          # the notebook's positional order
    assert torch.equal(aab.encode_all(data.numpy(), 4, dv, "cuda"), ref)                 # ... which also takes the numpy array it loads
    one = aab.encode_all(dv, data, batch_size=4, save_path=str(tmp_path / "reps.npy"))
    assert one == [str(tmp_path / "reps.npy")] and np.array_equal(np.load(one[0]), ref.numpy())
    many = aab.encode_all(dv, data, batch_size=4, save_path=str(tmp_path / "sh.npy"), shard_rows=5)   # batches straddle shard boundaries
    assert [p.split("/")[-1] for p in many] == ["sh_00000.npy", "sh_00001.npy", "sh_00002.npy"]
    assert np.array_equal(np.concatenate([np.load(p) for p in many]), ref.numpy())
    assert [np.load(p, mmap_mode="r").shape[0] for p in many] == [5, 5, 1]


# ---- dormant branches of encode_it (aa_mixer.py:178-179, 189-192): PQMF front-end, Memcodes quantiser ------------------------------
class _Args:
    def __init__(self, **kw):
        self.__dict__.update(kw)


@pytest.mark.parametrize("bands,n", [(2, 4096), (4, 8192), (8, 8192), (4, 5000)])
def test_pqmf_analysis_kernel(bands, n):
    "aa_pqmf_analysis_f32 against the oracle's conv1d restatement, the same filterbank on both sides; the bank is designed twice"
    import audio_algebra_b200 as aab
    from audio_algebra_b200.DiffusionDVAE import PQMF
    from oracle import aa_oracle as O
    pq = PQMF(2, 70, bands).cuda()
    hk_o = O.pqmf_filterbank(70, bands)
    assert rel_l2(pq.hk, hk_o) < 1e-6                      # product-side and oracle-side designs agree
    x = _x((3, 2, n), bands)
    y = pq(x.cuda())
    ref = O.pqmf_analysis(x.double(), hk_o)
    assert tuple(y.shape) == tuple(ref.shape) == (3, 2 * bands, n // bands)
    assert rel_l2(y, ref) < 1e-5


def test_encode_it_with_pqmf_front_end():
    "pqmf_bands = 4: [B, 2, N] -> PQMF -> encoder with 8 input channels (generic fp32 conv kernel) -> tanh"
    import audio_algebra_b200 as aab
    from oracle import aa_oracle as O
    torch.manual_seed(3)
    enc_o = O.SoundStreamXLEncoderOracle(in_channels=8).eval()
    m = aab.DiffusionDVAE(_Args(pqmf_bands=4, latent_dim=64, num_quantizers=0)).eval()
    m.load_oracle_weights(enc_o)
    m = m.cuda()
    x = _x((2, 2, 16384), 5)
    y = m.encode_it(x.cuda())
    ref = torch.tanh(enc_o(O.pqmf_analysis(x, O.pqmf_filterbank(70, 4).float())))
    assert tuple(y.shape) == tuple(ref.shape) == (2, 64, 16384 // 4 // 128)
    assert rel_l2(y, ref) < 1e-3


@pytest.mark.parametrize("heads,codes_n,nq", [(8, 1024, 1), (4, 64, 1), (8, 256, 3)])
def test_memcodes_quantiser(heads, codes_n, nq):
    "Memcodes / ResidualMemcodes lookup (eval path) against the oracle restatement: same indices, same values"
    from audio_algebra_b200.DiffusionDVAE import Memcodes, ResidualMemcodes
    from oracle import aa_oracle as O
    torch.manual_seed(heads + nq)
    q = (ResidualMemcodes(num_quantizers=nq, dim=64, heads=heads, num_codes=codes_n) if nq > 1 else Memcodes(dim=64, heads=heads, num_codes=codes_n))
    q = q.cuda().eval()
    x = torch.randn(3, 64, 70, generator=torch.Generator().manual_seed(1))
    layers = list(q.layers) if nq > 1 else [q]
    params = [(l.codes.detach().cpu().double(), l.to_k.weight.detach().cpu().double().reshape(64, -1), l.to_v.weight.detach().cpu().double().reshape(64, -1))
              for l in layers]
    out, idx = q.quantize_cf(x.cuda())
    if nq > 1:
        ref, ridx = O.residual_memcodes_eval(x.double(), params, heads)
        same = (idx.cpu() == ridx).float().mean().item()
    else:
        ref, ridx = O.memcodes_eval(x.double(), *params[0], heads)
        same = (idx.cpu() == ridx).float().mean().item()
    assert same > 0.995                                    # fp32 vs fp64 logits may flip a near tie
    if same == 1.0:
        assert rel_l2(out, ref) < 1e-5
    # reference layout: [B, N, dim] in, ([B, N, dim], indices) out
    o2, i2 = q(x.cuda().transpose(1, 2))
    assert tuple(o2.shape) == (3, 70, 64) and torch.equal(o2.transpose(1, 2), out)
    with pytest.raises(NotImplementedError):
        q.train()(x.cuda().transpose(1, 2))


def test_encode_it_quantised_branch():
    "num_quantizers = 1 / 2: tanh(quantizer_ema(encoder_ema(x))), rearranges folded into index math"
    import audio_algebra_b200 as aab
    from oracle import aa_oracle as O
    for nq in (1, 2):
        torch.manual_seed(11)
        enc_o = O.SoundStreamXLEncoderOracle().eval()
        m = aab.DiffusionDVAE(_Args(pqmf_bands=1, latent_dim=64, num_quantizers=nq, num_heads=8, codebook_size=128)).eval()
        m.load_oracle_weights(enc_o)
        m = m.cuda()
        x = _x((2, 2, 8192), 6)
        y = m.encode_it(x.cuda())
        emb = m.encoder_ema(x.cuda())                      # quantise the SAME fp32 embeddings on both sides (argmax is discontinuous)
        layers = list(m.quantizer_ema.layers) if nq > 1 else [m.quantizer_ema]
        params = [(l.codes.detach().cpu().double(), l.to_k.weight.detach().cpu().double().reshape(64, -1),
                   l.to_v.weight.detach().cpu().double().reshape(64, -1)) for l in layers]
        ref = (O.residual_memcodes_eval(emb.cpu().double(), params, 8) if nq > 1 else O.memcodes_eval(emb.cpu().double(), *params[0], 8))[0]
        assert tuple(y.shape) == (2, 64, 64)
        bad = ((y.cpu().double() - torch.tanh(ref)).abs() > 1e-4).float().mean().item()
        assert bad < 0.01                                  # a flipped near-tie changes 8 of 64 channels at one position
        assert y.abs().max() <= 1.0
    assert m.encode_it(x[0].cuda()).shape == (64, 64)      # unbatched [C, N] input, like the reference's Conv1d stack accepts


def test_dvae_checkpoint_loader(tmp_path):
    "Lightning-style state_dict (encoder.layers.*, encoder_ema.layers.*; plain and weight-normed convs) -> flat_convs() in order"
    import audio_algebra_b200 as aab
    from audio_algebra_b200.DiffusionDVAE import load_dvae_encoder_checkpoint
    from oracle import aa_oracle as O
    torch.manual_seed(4)
    enc_o = O.SoundStreamXLEncoderOracle().eval()
    sd = {}
    for i, (w, b) in enumerate(enc_o.flat_weights()):
        for pre in ("encoder", "encoder_ema"):
            if i % 2 == 0:
                sd[f"{pre}.layers.{i}.weight"], sd[f"{pre}.layers.{i}.bias"] = w.detach().clone(), b.detach().clone()
            else:                                          # weight_norm parametrisation of the same weight
                g = w.detach().flatten(1).norm(dim=1).view(-1, 1, 1)
                sd[f"{pre}.layers.{i}.weight_g"], sd[f"{pre}.layers.{i}.weight_v"] = g, w.detach() * 3.0
                sd[f"{pre}.layers.{i}.bias"] = b.detach().clone()
    sd["diffusion.net.0.weight"] = torch.zeros(4, 4, 3)   # other members of the DVAE are ignored
    path = tmp_path / "dvae.ckpt"
    torch.save({"state_dict": sd}, path)
    dv = aab.DVAEWrapper(debug=False)
    n_l = len(enc_o.flat_weights())
    assert load_dvae_encoder_checkpoint(dv.model, str(path)) == {"encoder.": n_l, "encoder_ema.": n_l}
    dv = dv.cuda()
    x = _x((1, 2, 4096), 8)
    assert rel_l2(dv.encode(x.cuda()), O.dvae_encode_it(enc_o, x)) < 1e-3
    del sd["encoder.layers.5.bias"]
    torch.save({"state_dict": sd}, path)
    with pytest.raises(RuntimeError):
        load_dvae_encoder_checkpoint(aab.DVAEWrapper(debug=False).model, str(path))


def test_encode_on_two_streams_does_not_share_activations(setup):
    """Two CUDA streams driving the same encoder concurrently (one activation workspace per (device, stream); the stream that did
    not upload the weights waits on the upload's event): results equal the sequential ones bit for bit."""
    aab, O, enc_o, dv = setup
    xa, xb = _x((6, 2, 32768), 31).cuda(), _x((6, 2, 32768), 32).cuda()
    ya, yb = dv.encode(xa).clone(), dv.encode(xb).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = {}
    for _ in range(3):   # several rounds: the kernels of the two streams overlap on the device
        with torch.cuda.stream(s1):
            outs["a"] = dv.encode(xa)
        with torch.cuda.stream(s2):
            outs["b"] = dv.encode(xb)
    torch.cuda.synchronize()
    assert torch.equal(outs["a"], ya) and torch.equal(outs["b"], yb)


def test_cta_pair_layers_are_bit_identical_to_the_single_cta_kernel(setup):
    """conv_tc2_kernel (2-CTA clusters, tcgen05.mma.cta_group::2 with M = 256, each CTA loading half of the weight box) against
    conv_tc_kernel (AA_TC_CG2=0 at handle creation) on the wide layers: the same products accumulated in the same K order, so the
    embeddings must agree bit for bit -- incl. odd row-tile counts (ghost tile of the pair) and ragged lengths."""
    import os
    aab, O, enc_o, dv = setup

    def make(cg2):
        os.environ["AA_TC_CG2"] = "1" if cg2 else "0"
        os.environ["AA_TC_HALO"] = "0"   # the halo mode of conv_tc_kernel accumulates chunk-major, not tap-major: other rounding
        try:
            w = aab.DVAEWrapper(debug=False, compute_dtype="bf16")
            w.model.load_oracle_weights(enc_o)
            w = w.cuda()
            w.encode(torch.zeros(1, 2, 1024, device="cuda"))   # the handle reads the switch at its first bf16 forward
        finally:
            os.environ.pop("AA_TC_CG2", None)
            os.environ.pop("AA_TC_HALO", None)
        return w

    pair, single = make(True), make(False)
    for shape, seed in [((4, 2, 131072), 41), ((3, 2, 128 * 128 * 3), 42), ((2, 2, 50000), 43), ((1, 2, 4096 * 5 + 77), 44), ((7, 2, 16384), 45)]:
        x = _x(shape, seed).cuda()
        a, b = pair.encode(x), single.encode(x)
        assert torch.equal(a, b), (shape, float((a - b).abs().max()))


def test_halo_box_layers_match_the_per_tap_boxes(setup):
    """conv_tc_kernel's halo mode (ONE [128 + 6 d rows x 64 channels] activation box per channel chunk, the seven taps as row-shifted
    descriptor views of it; used by the k7 layers that stay on one CTA per tile -- C = 128 by default, every wide k7 layer with
    AA_TC_CG2=0) against the per-tap boxes (AA_TC_HALO=0).  (1) Operands for which nothing rounds (one unit weight per output
    channel, small integer inputs, zero biases): bit-identical through the whole 37-layer encoder, dilations 1 / 3 / 9, ragged
    lengths, many tile seams.  (2) Oracle weights: the two accumulation orders agree to bf16 rounding."""
    import os
    aab, O, enc_o, dv = setup

    def make(halo, cg2, exact):
        os.environ["AA_TC_HALO"] = "1" if halo else "0"
        os.environ["AA_TC_CG2"] = "1" if cg2 else "0"
        try:
            w = aab.DVAEWrapper(debug=False, compute_dtype="bf16")
            if exact:
                with torch.no_grad():
                    for enc in (w.model.encoder, w.model.encoder_ema):
                        for li, conv in enumerate(enc.flat_convs()):
                            cout, cin, k = conv.weight.shape
                            wt = torch.zeros(cout, cin, k)
                            o = torch.arange(cout)
                            wt[o, (o * 5 + li) % cin, (o + li) % k] = 0.125 if conv.stride[0] > 1 else 1.0
                            conv.weight.copy_(wt)
                            conv.bias.zero_()
            else:
                w.model.load_oracle_weights(enc_o)
            w = w.cuda()
            w.encode(torch.zeros(1, 2, 1024, device="cuda"))   # the handle reads the switches at its first bf16 forward
        finally:
            os.environ.pop("AA_TC_HALO", None)
            os.environ.pop("AA_TC_CG2", None)
        return w

    for cg2 in (True, False):   # with the pair kernel (its own halo ring, both CTAs' boxes on the leader's barrier) and without
        on, off = make(True, cg2, True), make(False, cg2, True)
        for shape, seed in [((2, 2, 131072), 51), ((3, 2, 128 * 256 * 3 + 1300), 52), ((1, 2, 70000), 53)]:
            g = torch.Generator().manual_seed(seed)
            x = torch.randint(0, 4, shape, generator=g).float().cuda()
            a, b = on.encode(x), off.encode(x)
            assert torch.equal(a, b), (cg2, shape, float((a - b).abs().max()))
            assert a.abs().max() > 0
        on, off = make(True, cg2, False), make(False, cg2, False)
        for shape, seed in [((4, 2, 131072), 61), ((2, 2, 50000), 62)]:
            x = _x(shape, seed).cuda()
            a, b = on.encode(x), off.encode(x)
            assert rel_l2(a.cpu(), b.cpu()) < 5e-3, (cg2, shape, rel_l2(a.cpu(), b.cpu()))
