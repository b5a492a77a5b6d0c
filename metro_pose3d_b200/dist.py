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

    def _buffers(self, local, cap):
        """Gather buffers are allocated once per (shape, device) and re-used: no allocation in the step."""
        key = (cap, local.dtype, local.device)
        if getattr(self, '_key', None) != key:
            import torch
            self._key = key
            self._pad = torch.zeros((cap, self.j, 3), dtype=local.dtype, device=local.device)
            self._out = [torch.empty((self.world * cap, self.j, 3), dtype=local.dtype, device=local.device) for _ in range(2)]
            self._turn = 0
            if local.is_cuda:
                self._stream = torch.cuda.Stream(device=local.device)
                self._ready = [torch.cuda.Event(), torch.cuda.Event()]
                self._done = [None, None]
        return self._pad, self._out

    def _collect(self, local, n_total: int, out):
        sizes = [shard_bounds(n_total, self.world, r) for r in range(self.world)]
        cap = max(hi - lo for lo, hi in sizes)
        src = local
        if local.shape[0] != cap:                       # ragged shard: pad to the largest
            self._pad[:local.shape[0]] = local
            src = self._pad
        self.dist.all_gather_into_tensor(out, src.contiguous(), group=self.group)
        if all(hi - lo == cap for lo, hi in sizes):
            return out
        import torch
        return torch.cat([out[r * cap:r * cap + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])

    def gather(self, local, n_total: int):
        """All-gather ragged shards into [n_total, J, 3] (pads to the largest shard, then trims)."""
        if self.world == 1:
            return local
        cap = max(hi - lo for lo, hi in (shard_bounds(n_total, self.world, r) for r in range(self.world)))
        _, outs = self._buffers(local, cap)
        return self._collect(local, n_total, outs[0])

    def gather_async(self, local, n_total: int):
        """The same exchange issued on a side stream, so that it runs underneath the NEXT step's kernels instead of
        holding the launch stream at a per-step rendezvous.  The caller alternates between two `local` buffers; the
        result is valid after ``wait()`` (or after the next-but-one call returns)."""
        if self.world == 1 or not local.is_cuda:
            return self.gather(local, n_total)
        import torch
        cap = max(hi - lo for lo, hi in (shard_bounds(n_total, self.world, r) for r in range(self.world)))
        _, outs = self._buffers(local, cap)
        t = self._turn
        self._turn ^= 1
        cur = torch.cuda.current_stream(local.device)
        self._ready[t].record(cur)                      # `local` is complete at this point of the launch stream
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(self._ready[t])
            res = self._collect(local, n_total, outs[t])
            ev = torch.cuda.Event()
            ev.record(self._stream)
        # the launch stream may overwrite `local`'s twin (the other buffer) freely; before THIS buffer is written
        # again (two steps from now) the gather that reads it must have finished
        if self._done[t ^ 1] is not None:
            cur.wait_event(self._done[t ^ 1])
        self._done[t] = ev
        return res

    def wait(self):
        """Joins every outstanding asynchronous gather into the current stream."""
        if self.world == 1 or getattr(self, '_done', None) is None:
            return
        import torch
        for ev in self._done:
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)

    def __call__(self, images_global):
        n_total = images_global.shape[0]
        local = self.infer_fn(images_global[self.local_slice(n_total)])
        return self.gather(local, n_total)
