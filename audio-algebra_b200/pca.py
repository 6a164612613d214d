"""Running PCA of given-model latents: the loop body of the reference's calc_effects_pca.py:76-94
(rearrange 'b d n -> d (b n)', torch.cov * (n-1), running sum, eigh, descending sort) with the
C x C scatter accumulated straight from the [B, C, T] tensor by one CUDA pass (no rearrange copy,
no cuBLAS); the 64 x 64 eigen-decomposition stays in torch (negligible).  Batch-sharded over ranks:
`all_reduce()` sums the numerator and the point count (4097 floats for C = 64)."""
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .aa_mixer import _f32c, _ws
from .parallel import allreduce_sum_

__all__ = ['sorted_eig', 'RunningCovariance']


def sorted_eig(cov):
    """calc_effects_pca.py:39-43, eigenvalues descending.  The reference reorders eigenvector ROWS
    (index_select dim 0), a latent bug that does not affect the logged eigenvalues; here the columns
    (the eigenvectors) are reordered."""
    lambdas, vs = torch.linalg.eigh(cov)
    lambdas, indices = torch.sort(lambdas, dim=0, descending=True)
    vs = torch.index_select(vs, 1, indices)
    return lambdas, vs


class RunningCovariance:
    def __init__(self, channels, device):
        self.c = int(channels)
        self.device = torch.device(device)
        _lib.ensure_device(self.device)
        self.cov_numerator = torch.zeros((self.c, self.c), dtype=torch.float32, device=self.device)
        self.count = torch.zeros((1,), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._ws = _ws(lib.aa_cov_workspace_floats(self.c), self.device)

    def update(self, ys):
        "ys [B, C, T]: adds this batch's scatter about ITS OWN mean (what torch.cov(ys)*(n-1) gives)"
        ys = _f32c(ys, "ys")
        assert ys.dim() == 3 and ys.shape[1] == self.c
        with torch.cuda.device(self.device):
            check(lib.aa_cov_accumulate_f32(ptr(ys), ys.shape[0], self.c, ys.shape[2], ptr(self.cov_numerator),
                                            ptr(self.count), ptr(self._ws), stream_ptr()))
        return self

    def all_reduce(self, group=None):
        allreduce_sum_(self.cov_numerator, group)
        allreduce_sum_(self.count, group)
        return self

    def covariance(self):
        return self.cov_numerator / (self.count.to(torch.float32) - 1)

    def eigenvalues(self):
        return sorted_eig(self.covariance())[0]
