"""Multi-GPU plumbing of the hot path: chunks are independent, so every stage shards the batch dimension
contiguously over ranks with NO data-path collective; the only exchanges are (i) the mean all-reduce of the
flat 33 280-float projector gradient in training and (ii) the sum all-reduce of the 64x64 PCA numerator
(+ point count).  Device-agnostic (NCCL on B200, gloo in the CPU tests)."""
import torch
import torch.distributed as dist

__all__ = ['world_info', 'shard_range', 'shard_batch', 'allreduce_mean_', 'allreduce_sum_', 'bind_to_gpu_numa_node']


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n, rank, world):
    "contiguous [lo, hi) of n items for `rank`; the first n % world ranks get one extra (nothing is dropped)"
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x, group=None):
    rank, world = world_info(group)
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


def allreduce_sum_(t, group=None):
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_mean_(t, group=None):
    "DDP semantics: mean over ranks of the per-rank (local-batch) gradients"
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(world)
    return t


def bind_to_gpu_numa_node(device_index):
    """One process per GPU: restrict this process to the CPUs of the NUMA node its GPU hangs off, so that pinned staging
    buffers allocated afterwards are first-touched on that node and H2D / D2H copies do not cross the socket interconnect
    (matters when all 8 ranks of a node stream from host memory at once).  Best effort: returns the CPU list, or None when
    the topology cannot be read (then nothing is changed)."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None
