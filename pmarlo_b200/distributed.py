"""Frame-sharded reductions: one process per GPU, ``torch.distributed`` plumbing.

Every cross-rank exchange on this path is a SUM of small partials (SURVEY.md
section 8e): column moments, the two Gram matrices, k-means centroid sums /
counts / inertia and the K x K count matrix.  Integer sums are bit-exact for any
rank count; fp64 sums are order-dependent at 1e-16 relative.  The kernels write
their partials straight into the tensors that are reduced -- no staging copy.
"""

from __future__ import annotations

import torch

__all__ = ["Comm"]


class Comm:
    """Thin view of the default process group (no-op when not initialised)."""

    def __init__(self, group=None, solo: bool = False):
        import torch.distributed as dist

        self._dist = dist
        self.group = group
        # solo: behave as a single rank even inside an initialised process group (the 1-rank reference run
        # of bench.py's multi-rank parity check)
        self.active = (not solo) and dist.is_available() and dist.is_initialized()
        self.size = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0

    def allreduce_sum(self, *tensors: torch.Tensor) -> None:
        """In-place sum over ranks of each tensor (one collective per tensor,
        issued back to back on the same stream)."""
        if not self.active or self.size == 1:
            return
        for t in tensors:
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)

    def allreduce_max(self, t: torch.Tensor) -> None:
        if not self.active or self.size == 1:
            return
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX, group=self.group)

    def broadcast(self, t: torch.Tensor, src: int = 0) -> None:
        if not self.active or self.size == 1:
            return
        self._dist.broadcast(t, src=src, group=self.group)

    def barrier(self) -> None:
        if self.active and self.size > 1:
            self._dist.barrier(group=self.group)

    def sum_int(self, value: int, device=None) -> int:
        """Sum of a host integer over ranks (metadata such as frame / pair counts)."""
        if not self.active or self.size == 1:
            return int(value)
        dev = device if device is not None else ("cuda" if self._dist.get_backend(self.group) == "nccl" else "cpu")
        t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return int(t.item())
