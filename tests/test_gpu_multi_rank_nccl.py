"""Two-rank NCCL run of the data-parallel pieces on real GPUs (skipped on a single-GPU box; run with `gpurun --gpus 2`):
MixerTrainer.step with per-rank batch shards (mean all-reduce of the flat 33 280-float gradient, replicated parameters stay
bit-identical), the PCA numerator sum, and batch-sharded mel encoding.  The expected values come from one process that
evaluates both shards itself (per-rank VICReg statistics = DDP semantics)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FADERS = [1.4630, -0.5718]


def _data(seed, shape):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) - 0.5


def _make(aab, device):
    from oracle import aa_oracle as O
    torch.manual_seed(0)
    enc_o = O.SoundStreamXLEncoderOracle().eval()
    dv = aab.DVAEWrapper(debug=False)
    dv.model.load_oracle_weights(enc_o)
    aa = aab.AudioAlgebra(64, 64)
    aa.load_state_dict(O.init_projector_state_dict(64, 64, seed=2))
    return dv.to(device), aa.to(device)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import audio_algebra_b200 as aab
    from audio_algebra_b200.training import MixerTrainer
    from audio_algebra_b200.pca import RunningCovariance
    from audio_algebra_b200 import parallel as P
    dv, aa = _make(aab, dev)
    tr = MixerTrainer(dv.model, aa, total_steps=10)
    full = [_data(11, (4, 2, 2048)), _data(12, (4, 2, 2048))]
    stems = [P.shard_batch(s).to(dev) for s in full]              # contiguous batch shard of this rank
    tr.step(stems, FADERS)
    grad = tr.flat_grad.detach().cpu().clone()
    tr.step(stems, FADERS)
    params = tr.flat.detach().cpu().clone()
    ys = torch.tanh(_data(13, (6, 64, 32)))
    rc = RunningCovariance(64, dev).update(P.shard_batch(ys).to(dev)).all_reduce()
    mel = aab.MelSpectrogramAE(sample_rate=48000, n_fft=2048, hop_length=512).encode(P.shard_batch(_data(14, (4, 2, 8192))).to(dev))
    q.put((rank, grad, params, rc.cov_numerator.cpu(), float(rc.count.item()), mel.cpu()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl():
    import audio_algebra_b200 as aab
    from audio_algebra_b200.training import MixerTrainer
    from audio_algebra_b200.parallel import shard_range
    from oracle import aa_oracle as O
    import socket
    with socket.socket() as sk:          # a free port (a fixed one can still be in TIME_WAIT from an earlier run)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # replicated state: gradient after the all-reduce and parameters after two steps are bit-identical on both ranks
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    # expected gradient: mean over ranks of the per-shard gradients, computed here shard by shard on one GPU
    full = [_data(11, (4, 2, 2048)), _data(12, (4, 2, 2048))]
    grads = []
    for r in range(world):
        dv, aa = _make(aab, "cuda")
        tr = MixerTrainer(dv.model, aa, total_steps=10)
        lo, hi = shard_range(4, r, world)
        tr.step([s[lo:hi].cuda() for s in full], FADERS)
        grads.append(tr.flat_grad.detach().cpu().clone())
    assert rel_l2(res[0][1], (grads[0] + grads[1]) / 2) < 1e-5
    # PCA: sum of per-shard scatters and counts
    ys = torch.tanh(_data(13, (6, 64, 32)))
    num_ref, n_ref = None, 0
    for r in range(world):
        lo, hi = shard_range(6, r, world)
        c, n = O.pca_cov_numerator(ys[lo:hi].double())
        num_ref = c if num_ref is None else num_ref + c
        n_ref += n
    assert rel_l2(res[0][3], num_ref) < 1e-4 and int(res[0][4]) == n_ref and torch.equal(res[0][3], res[1][3])
    # batch-sharded mel = the rows of the full-batch result
    x = _data(14, (4, 2, 8192))
    ref = O.mel_spectrogram(x, 48000, 2048, 512)
    assert rel_l2(torch.cat([res[0][5], res[1][5]], dim=0), ref) < 1e-4
