"""Out-of-bounds write check without compute-sanitizer (the tool is closed on this pool: profiles/sanitizer_r02_unavailable.txt).
Every output buffer handed to the C ABI here sits between two guard bands filled with a canary bit pattern; after the kernel the
bands must be intact and the payload fully overwritten (no canary left inside).  Shapes are chosen ragged: odd row counts, lengths
that are not multiples of the tile sizes, frame counts that leave partial tiles."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
CANARY = -1.2345678e30


def _guarded(numel, dtype=torch.float32, guard=4096):
    buf = torch.full((numel + 2 * guard,), CANARY, dtype=torch.float32, device="cuda") if dtype == torch.float32 else None
    return buf, buf[guard:guard + numel], guard


def _check(buf, numel, guard, what, payload_written=True):
    assert torch.all(buf[:guard] == CANARY).item(), f"{what}: wrote BEFORE the buffer"
    assert torch.all(buf[guard + numel:] == CANARY).item(), f"{what}: wrote PAST the buffer"
    if payload_written:
        assert not torch.any(buf[guard:guard + numel] == CANARY).item(), f"{what}: left part of the output unwritten"


@pytest.mark.parametrize("rows,n,hop", [(3, 20000, 512), (2, 131072, 512), (1, 9000, 256), (5, 40000, 1024)])
def test_stft_outputs_stay_inside_their_buffers(monkeypatch, rows, n, hop):
    import audio_algebra_b200 as aab
    from audio_algebra_b200._lib import lib, check, ptr, stream_ptr
    x = (torch.rand(rows, n, device="cuda") - 0.5).contiguous()
    for v2 in ("1", "0"):
        monkeypatch.setenv("AA_STFT_V2", v2)
        monkeypatch.setenv("AA_STFT_V2_MEL", v2)
        m = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=hop)
        plan = m._plan(torch.cuda.current_device())
        n_pad, frames = plan.out_shape(n, True)
        for what, fn, per_row in (("mel", lib.aa_stft_mel_f32, 128 * frames), ("power_tf", lib.aa_stft_power_tf_f32, 1025 * frames),
                                  ("complex_tf", lib.aa_stft_complex_tf_f32, 2 * 1025 * frames), ("power", lib.aa_stft_power_f32, 1025 * frames),
                                  ("complex", lib.aa_stft_complex_f32, 2 * 1025 * frames)):
            buf, out, g = _guarded(rows * per_row)
            check(fn(plan.handle, ptr(x), rows, n, 1, ptr(out), stream_ptr()))
            torch.cuda.synchronize()
            _check(buf, rows * per_row, g, f"{what} v2={v2} rows={rows} n={n} hop={hop}")
    m1 = aab.MelSpectrogramAE(sample_rate=48000)              # reference defaults n_fft 1024 / hop 256: warp kernel
    plan = m1._plan(torch.cuda.current_device())
    _, frames = plan.out_shape(n, True)
    buf, out, g = _guarded(rows * 128 * frames)
    check(lib.aa_stft_mel_f32(plan.handle, ptr(x), rows, n, 1, ptr(out), stream_ptr()))
    torch.cuda.synchronize()
    _check(buf, rows * 128 * frames, g, "mel n_fft=1024")


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_encoder_output_stays_inside_its_buffer(dtype):
    import audio_algebra_b200 as aab
    from audio_algebra_b200._lib import lib, check, ptr, stream_ptr
    from audio_algebra_b200.DiffusionDVAE import DTYPES
    torch.manual_seed(0)
    enc = aab.SoundStreamXLEncoder(compute_dtype=dtype).cuda()
    for b, n in [(3, 5000), (1, 128 * 129 + 7), (2, 640)]:
        x = (torch.rand(b, 2, n, device="cuda") - 0.5).contiguous()
        t_out = enc.out_length(n)
        h = enc._handle(torch.cuda.current_device())
        nbytes = int(lib.aa_encoder_workspace_bytes(h, b, n, DTYPES[dtype]))
        ws_guard = 1 << 16
        ws = torch.full((nbytes + 2 * ws_guard,), 0x5A, dtype=torch.uint8, device="cuda")
        buf, y, g = _guarded(b * 64 * t_out)
        arr = (C.c_void_p * 1)(x.data_ptr())
        fad = (C.c_float * 1)(1.0)
        check(lib.aa_encoder_forward(h, C.cast(arr, C.POINTER(C.c_void_p)), fad, 1, b, n, 1, DTYPES[dtype], ptr(y),
                                     C.c_void_p(ws.data_ptr() + ws_guard), stream_ptr()))
        torch.cuda.synchronize()
        _check(buf, b * 64 * t_out, g, f"encoder {dtype} b={b} n={n}")
        assert torch.all(ws[:ws_guard] == 0x5A).item() and torch.all(ws[-ws_guard:] == 0x5A).item(), f"encoder {dtype}: workspace overrun (b={b}, n={n})"


def test_projector_and_loss_outputs_stay_inside_their_buffers():
    import audio_algebra_b200 as aab
    from audio_algebra_b200._lib import lib, check, ptr, stream_ptr
    from audio_algebra_b200.aa_mixer import _ptr_array
    torch.manual_seed(2)
    aa = aab.AudioAlgebra(64, 64).cuda()
    for b, t in [(3, 70), (1, 1000), (5, 129)]:
        x = torch.randn(b, 64, t, device="cuda")
        ws = [blk.lin.weight.detach().contiguous() for blk in aa.encoder]
        bs = [blk.lin.bias.detach().contiguous() for blk in aa.encoder]
        wp, k1 = _ptr_array(ws)
        bp, k2 = _ptr_array(bs)
        buf, out, g = _guarded(x.numel())
        check(lib.aa_projector_half_fwd_f32(wp, bp, 64, 64, 1, ptr(x), b, t, ptr(out), stream_ptr()))
        torch.cuda.synchronize()
        _check(buf, x.numel(), g, f"projector fwd b={b} t={t}")
        buf, out, g = _guarded(x.numel())
        arr, keep = _ptr_array([x, x])
        ca = (C.c_float * 2)(0.5, 0.25)
        check(lib.aa_latent_lincomb_f32(2, arr, ca, ptr(out), x.numel(), stream_ptr()))
        torch.cuda.synchronize()
        _check(buf, x.numel(), g, f"lincomb b={b} t={t}")


def test_tcgen05_backward_kernels_stay_inside_their_buffers():
    """Round-2 tensor-core kernels: projector backward (gx, four weight / bias gradients, the per-CTA partial workspace) and the
    covariance loss Gram / backward (gram, grad_z, the split-K workspace) at ragged shapes."""
    import audio_algebra_b200 as aab
    from audio_algebra_b200._lib import lib, check, ptr, stream_ptr
    from audio_algebra_b200.aa_mixer import _ptr_array
    torch.manual_seed(3)
    aa = aab.AudioAlgebra(64, 64).cuda()
    ws = [blk.lin.weight.detach().contiguous() for blk in aa.encoder]
    bs = [blk.lin.bias.detach().contiguous() for blk in aa.encoder]
    wp, k1 = _ptr_array(ws)
    bp, k2 = _ptr_array(bs)
    nws = int(lib.aa_projector_bwd_workspace_floats())
    for b, t in [(3, 70), (2, 1000), (150, 129)]:
        x, gout = torch.randn(b, 64, t, device="cuda"), torch.randn(b, 64, t, device="cuda")
        gbuf, gx, g = _guarded(x.numel())
        wsbuf, wsp, gw_ = _guarded(nws)
        gws = [_guarded(64 * 64) for _ in range(4)]
        gbs = [_guarded(64) for _ in range(4)]
        gwp, k3 = _ptr_array([q[1] for q in gws])
        gbp, k4 = _ptr_array([q[1] for q in gbs])
        check(lib.aa_projector_half_bwd_f32(wp, bp, 64, 64, 1, ptr(x), ptr(gout), b, t, ptr(gx), 0, gwp, gbp, 0, 1.0, ptr(wsp), stream_ptr()))
        torch.cuda.synchronize()
        _check(gbuf, x.numel(), g, f"projector bwd gx b={b} t={t}")
        _check(wsbuf, nws, gw_, f"projector bwd workspace b={b} t={t}", payload_written=False)
        for i in range(4):
            _check(gws[i][0], 64 * 64, gws[i][2], f"projector bwd gw{i}")
            _check(gbs[i][0], 64, gbs[i][2], f"projector bwd gb{i}")
    for b, d in [(130, 2052), (64, 1024), (200, 4096)]:
        z = torch.randn(b, d, device="cuda")
        nws = int(lib.aa_cov_loss_workspace_floats(b, d))
        wsbuf, wsp, gw_ = _guarded(nws)
        sbuf, stats, gs = _guarded(2 * d)
        grbuf, gram, gg = _guarded(b * b)
        lbuf, loss, gl = _guarded(1)
        check(lib.aa_vicreg_cov_fwd_f32(ptr(z), b, d, None, ptr(stats), ptr(gram), ptr(loss), ptr(wsp), stream_ptr()))
        torch.cuda.synchronize()
        _check(wsbuf, nws, gw_, f"cov fwd workspace b={b} d={d}", payload_written=False)
        _check(sbuf, 2 * d, gs, f"cov fwd stats b={b} d={d}")
        _check(grbuf, b * b, gg, f"cov fwd gram b={b} d={d}")
        _check(lbuf, 1, gl, "cov fwd loss")
        gzbuf, gz, gz_ = _guarded(b * d)
        check(lib.aa_vicreg_cov_bwd_f32(ptr(z), ptr(stats), ptr(gram), b, d, None, 1.0, ptr(gz), 0, stream_ptr()))
        torch.cuda.synchronize()
        _check(gzbuf, b * d, gz_, f"cov bwd grad_z b={b} d={d}")
