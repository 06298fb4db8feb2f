"""Multi-GPU plumbing: one process per GPU (torchrun), the batch is split contiguously across ranks
and every transform is per-sample independent, so inference needs NO collective.
Training adds exactly one all-reduce per step over a single flat bucket of the trainable gradients
(0.3-36 MB: latency-bound on NVLink 5, NCCL picks its one-shot/NVLS path); loss scalars ride along."""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n samples: rank r owns [lo, hi); remainders go to the lowest ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    lo, hi = shard_range(tensors[0].shape[0], rank, world)
    return [t[lo:hi] for t in tensors]


class GradBucket:
    """All trainable gradients as ONE flat fp32 buffer -> one all-reduce (sum) -> mean."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self._flat = None

    def allreduce_mean(self, extra_scalars: Dict[str, torch.Tensor] | None = None, group=None) -> Dict[str, torch.Tensor]:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        names = sorted(extra_scalars) if extra_scalars else []
        dev = self.params[0].device
        if self._flat is None or self._flat.device != dev:
            self._flat = torch.zeros(self.numel + 64, dtype=torch.float32, device=dev)
        flat = self._flat
        views, off = [], 0
        for p in self.params:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        have = [(v, p.grad) for v, p in zip(views, self.params) if p.grad is not None]
        for v, p in zip(views, self.params):
            if p.grad is None:
                v.zero_()
        if have:   # one fused multi-tensor copy in, one out (dozens of parameters, two launches)
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        assert len(names) <= 64
        for i, k in enumerate(names):
            flat[self.numel + i] = extra_scalars[k].detach().float()
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.div_(world)
        if have:
            torch._foreach_copy_([g for _, g in have], [v for v, _ in have])
        return {k: flat[self.numel + i].clone() for i, k in enumerate(names)}


def gather_outputs(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Reassemble per-rank output shards (harness convenience, not on the timed path)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
