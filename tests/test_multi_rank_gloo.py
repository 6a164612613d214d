"""World-size-2 gloo run (CPU) of the multi-rank host logic: batch sharding, DDP-style gradient mean,
PCA numerator sum.  The reference semantics: var/cov statistics are per rank;
the N-rank gradient is the MEAN over ranks of per-shard gradients; parameters start identical."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audio_algebra_b200 import parallel as P
    from oracle import aa_oracle as O
    torch.manual_seed(0)
    full = torch.randn(6, 64, 8, dtype=torch.float64)          # same on every rank
    mine = P.shard_batch(full)
    # per-rank loss on the local shard, gradient w.r.t. a shared parameter vector
    w = torch.ones(64, dtype=torch.float64, requires_grad=True)
    z = mine * w[None, :, None]
    (O.vicreg_var_loss(z) + O.vicreg_cov_loss(z)).backward()
    g = w.grad.clone()
    P.allreduce_mean_(g)
    # PCA numerator: sum over ranks of per-shard scatters
    num, n = O.pca_cov_numerator(torch.tanh(mine))
    cnt = torch.tensor([float(n)], dtype=torch.float64)
    P.allreduce_sum_(num); P.allreduce_sum_(cnt)
    q.put((rank, mine.shape[0], g, num, cnt))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    sys.path.insert(0, ROOT)
    from oracle import aa_oracle as O
    from audio_algebra_b200.parallel import shard_range
    import socket
    with socket.socket() as sk:          # a free port (a fixed one can still be in TIME_WAIT from an earlier run)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert [r[1] for r in res] == [3, 3]
    # expected: mean of per-shard gradients, identical on both ranks
    torch.manual_seed(0)
    full = torch.randn(6, 64, 8, dtype=torch.float64)
    gs, num_ref, n_ref = [], None, 0
    for r in range(world):
        lo, hi = shard_range(6, r, world)
        w = torch.ones(64, dtype=torch.float64, requires_grad=True)
        z = full[lo:hi] * w[None, :, None]
        (O.vicreg_var_loss(z) + O.vicreg_cov_loss(z)).backward()
        gs.append(w.grad)
        c, n = O.pca_cov_numerator(torch.tanh(full[lo:hi]))
        num_ref = c if num_ref is None else num_ref + c
        n_ref += n
    g_ref = (gs[0] + gs[1]) / 2
    for _, _, g, num, cnt in res:
        assert torch.allclose(g, g_ref, atol=1e-12)
        assert torch.allclose(num, num_ref, atol=1e-10) and int(cnt.item()) == n_ref
    # and it is NOT the single-process full-batch gradient (per-rank statistics)
    w = torch.ones(64, dtype=torch.float64, requires_grad=True)
    z = full * w[None, :, None]
    (O.vicreg_var_loss(z) + O.vicreg_cov_loss(z)).backward()
    assert not torch.allclose(w.grad, g_ref, atol=1e-6)
