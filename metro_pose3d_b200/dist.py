"""Data-parallel sharding of a crop batch over the GPUs of one box.

The reference has no distributed code at all (one process, '/gpu:0', src/helpers.py:58-59).  Every
crop's result depends only on that crop and the replicated weights (inference-mode BN,
src/model/architectures.py:9-11), so the path shards by batch index with no data-path collective;
the single exchange is an all-gather of the float32 [N/G, J, 3] results (<= 29 KB per rank).

One process per GPU (torchrun); ``torch.distributed`` is plumbing only.
"""
from __future__ import annotations

from typing import Callable, Tuple


def shard_bounds(n_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced shards: rank r gets [lo, hi); sizes differ by at most one crop."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedPoseEstimator:
    """``infer_fn(images_shard) -> [n_shard, J, 3]`` runs on this rank's device (MetroModel.infer on
    a GPU); ``__call__`` takes the GLOBAL batch (every rank holds or can index it), computes this
    rank's shard and returns the gathered [N, J, 3] on every rank."""

    def __init__(self, infer_fn: Callable, n_joints_out: int, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.infer_fn = infer_fn
        self.j = n_joints_out
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def local_slice(self, n_total: int) -> slice:
        lo, hi = shard_bounds(n_total, self.world, self.rank)
        return slice(lo, hi)

    def gather(self, local, n_total: int):
        """All-gather ragged shards into [n_total, J, 3] (pads to the largest shard, then trims)."""
        import torch
        if self.world == 1:
            return local
        sizes = [shard_bounds(n_total, self.world, r) for r in range(self.world)]
        cap = max(hi - lo for lo, hi in sizes)
        buf = torch.zeros((cap, self.j, 3), dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
        out = torch.empty((self.world * cap, self.j, 3), dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(out, buf, group=self.group)
        if all(hi - lo == cap for lo, hi in sizes):
            return out
        return torch.cat([out[r * cap:r * cap + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])

    def __call__(self, images_global):
        n_total = images_global.shape[0]
        local = self.infer_fn(images_global[self.local_slice(n_total)])
        return self.gather(local, n_total)
