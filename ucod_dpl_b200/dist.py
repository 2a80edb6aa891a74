"""Data-parallel plumbing (one process per GPU, `torch.distributed`; NCCL on the GPUs, gloo in the CPU tests).

The hot path shards by image with no data-path collective: image i -> rank i mod W.  The only
exchanges are (1) the decoder-gradient all-reduce of first-stage training (ucod_dpl_b200/train.py), (2) the final
metric reduction (sums + counts) and (3) gathering the uint8 pseudo-label masks of generate_pseudo_label.py.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items: int, rank: int | None = None, world_size: int | None = None) -> range:
    """Static interleaved shard: item i belongs to rank i mod W (equal cost per item for eval / pseudo labels)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return range(rank, n_items, world_size)


def reduce_metric_sums(sums: torch.Tensor, count: int) -> tuple[torch.Tensor, int]:
    """Final metric reduction (reference: accelerator.gather_for_metrics + per-image lists, loop_UCOD_DPL.py:310,
    metric.py:19-74): every rank passes its per-measure sums (float64) and its image count; returns global sums and
    the global count.  One all-reduce of len(sums)+1 doubles."""
    buf = torch.cat([sums.to(torch.float64).flatten(), torch.tensor([float(count)], dtype=torch.float64,
                                                                    device=sums.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf[:-1].reshape(sums.shape), int(round(buf[-1].item()))


def gather_sharded_masks(local_masks: torch.Tensor, n_items: int) -> torch.Tensor | None:
    """local_masks uint8 [n_local, h, w] of the items `shard_indices(n_items)` in order -> on rank 0 the full
    [n_items, h, w] tensor in item order (other ranks: None).  Pads the shorter shards for the all_gather."""
    rank, w = world()
    if w == 1:
        return local_masks
    per = (n_items + w - 1) // w
    pad = torch.zeros((per,) + tuple(local_masks.shape[1:]), dtype=local_masks.dtype, device=local_masks.device)
    pad[: local_masks.shape[0]] = local_masks
    parts = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(parts, pad)
    if rank != 0:
        return None
    out = torch.empty((n_items,) + tuple(local_masks.shape[1:]), dtype=local_masks.dtype, device=local_masks.device)
    for r in range(w):
        idx = list(range(r, n_items, w))
        out[idx] = parts[r][: len(idx)]
    return out
