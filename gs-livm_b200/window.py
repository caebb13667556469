"""View-sharded keyframe window (multi-GPU), SURVEY.md section 8e.

GS-LIVM's mapping iteration renders several keyframes of the sliding window against the SAME parameter
snapshot, sums their losses and takes one optimiser step (src/liw/lioOptimization.cpp:1691-1736,
1822-1832). The renders are independent, so the window shards by view: Gaussians (and optimiser state)
are replicated on every GPU, each rank renders its share of the views forward+backward with the
single-GPU kernels, and the per-Gaussian parameter gradients (59 floats = 236 B per Gaussian at SH degree
3) are summed with ONE all-reduce over a flat buffer (NCCL over NVLink/NVSwitch) before the identical
update is applied everywhere. There is no other cross-rank traffic on the data path.

The collective is NCCL's (torch.distributed); there is no compute step whose output feeds the collective
tile by tile — the gradients are only complete when the last tile of the backward has finished — so a
fused compute+collective kernel has nothing to overlap here ("multi-GPU").
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def views_for_rank(n_views: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment of window views to ranks (view v -> rank v % world_size)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return [v for v in range(n_views) if v % world_size == rank]


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Sum (not mean: the reference sums the view losses, lioOptimization.cpp:1823-1825) in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


class WindowStep:
    """One optimisation step over a window of views, sharded over the ranks of `group`.

    render_view(view_index, accumulate_into: Tensor) must run forward + loss + backward for that view on this
    rank's device and ADD the view's flat parameter gradient into `accumulate_into` (same layout on all ranks).
    """

    def __init__(self, n_views: int, grad_numel: int, device, render_view: Callable[[int, torch.Tensor], float],
                 group=None, dtype=torch.float32):
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.views = views_for_rank(n_views, self.rank, self.world)
        self.flat = torch.zeros(grad_numel, dtype=dtype, device=device)
        self.render_view = render_view

    def run(self) -> Dict[str, float]:
        self.flat.zero_()
        loss = 0.0
        for v in self.views:
            loss += float(self.render_view(v, self.flat))
        allreduce_sum_(self.flat, self.group)
        return {"local_loss": loss, "local_views": float(len(self.views))}


def sgd_update_(params: Sequence[torch.Tensor], flat_grad: torch.Tensor, lr: float) -> None:
    """Tiny deterministic update used by the tests to check that replicas stay bit-identical. (The product keeps
    LibTorch's Adam, src/gs/gaussian.cu:396-428, which is outside the replaced path.)"""
    off = 0
    for p in params:
        n = p.numel()
        p.add_(flat_grad[off:off + n].view_as(p), alpha=-lr)
        off += n
    if off != flat_grad.numel():
        raise ValueError("flat gradient layout does not match the parameter list")
