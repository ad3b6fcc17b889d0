"""Batch sharding across the GPUs of one box: one process per GPU, contiguous split of the batch of
microstructures, NO collective on the sampling path, one gather of the decoded fields at the end
(SURVEY.md section 8(e)).  A sample's slices stay together (E2D/D3D convolve across depth).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of samples owned by `rank`; earlier ranks take the remainder."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def chunk_starts(batch: int, chunk: int) -> List[int]:
    """First samples of the micro-batches a VAE pass runs over (predictor._get_session): windows of exactly `chunk`
    samples (one set of activation buffers and plans serves them all), back to back from 0; a ragged tail is covered
    by one more window that ENDS at the batch's last sample and so re-computes up to chunk-1 samples (idempotent)."""
    if batch < 1 or chunk < 1 or chunk > batch:
        raise ValueError(f"chunk_starts: need 1 <= chunk <= batch, got chunk={chunk} batch={batch}")
    starts = list(range(0, batch - chunk + 1, chunk))
    if starts[-1] + chunk < batch:
        starts.append(batch - chunk)
    return starts


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def gather_predictions(local: torch.Tensor, batch: int, group=None, dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """The final gather: returns the full (batch, ...) tensor on every rank (dst=None, all_gather) or on
    `dst` only (gather).  Ragged shards are padded to the largest shard for the collective."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    pad = pad.contiguous()
    if dst is None:
        bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
    else:
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


def predict_sharded(predictor, img, velocity_2d, noise=None, *, sampler="ddim", group=None, dst=None, **kw):
    """Run `predictor.predict_ddim` / `.predict` on this rank's shard of the batch and gather the result.
    `img`, `velocity_2d`, `noise` are the FULL batch on every rank (synthetic inputs are rank-local slices)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = img.shape[0]
    lo, hi = shard_range(B, rank, world)
    S = velocity_2d.shape[1]
    n_loc = None if noise is None else noise.reshape(B, S, *noise.shape[-3:])[lo:hi].reshape(-1, *noise.shape[-3:])
    fn = predictor.predict_ddim if sampler == "ddim" else predictor.predict
    out = fn(img[lo:hi], velocity_2d[lo:hi], noise=n_loc, **kw)
    return gather_predictions(out, B, group=group, dst=dst)
