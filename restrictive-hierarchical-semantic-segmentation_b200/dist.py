"""Batch-sharded (data-parallel) helpers: the path's only cross-rank exchange.

Every quantity on the path is per pixel or per sample (FiLM pools inside a sample, CE/Dice are
per-sample means, the confusion matrix is a pixel count), so the batch shards over ranks with
no pixel data ever leaving its GPU.  One process per GPU (torch.distributed, NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests); per step ONE all-reduce of a small
packed fp64 buffer turns the per-rank step summaries into the numbers the single-process
reference would report on the concatenated batch, and gives the factors that make the
DDP-averaged gradients equal to the single-process gradients.

Reference semantics being reproduced: CE is a mean over ALL samples of the
global batch; Dice is a mean over the samples whose dice is not NaN (Metrics/losses.py:64-66),
consistency a mean over B*H*W; confusion matrices add.
"""
from typing import List, Sequence

import torch
import torch.distributed as dist


def pack_step_summary(scalars: torch.Tensor, batch_local: int, confusion: Sequence[torch.Tensor]) -> torch.Tensor:
    """fp64 buffer: [B_local, consistency*B_local, per level (ce*B_local, dice*n_dice, n_dice, n_ce)] + confusion counts.
    `scalars` is StepOutput.scalars ([2 + 4*n]: total, consistency, then ce, dice, n_dice, n_ce per level).
    int64 counts are exact in fp64 up to 2^53."""
    n = (scalars.numel() - 2) // 4
    s = scalars.double()
    parts = [torch.tensor([float(batch_local)], dtype=torch.float64, device=s.device), (s[1] * batch_local).reshape(1)]
    for L in range(n):
        ce, dice, n_dice, n_ce = s[2 + 4 * L], s[3 + 4 * L], s[4 + 4 * L], s[5 + 4 * L]
        parts.append(torch.stack([ce * batch_local, dice * n_dice, n_dice, n_ce]))
    parts += [c.reshape(-1).double() for c in confusion]
    return torch.cat(parts)


def unpack_global(buf: torch.Tensor, n_levels: int, conf_shapes: Sequence[Sequence[int]]):
    """Inverse of pack_step_summary after the SUM all-reduce: global loss terms + confusion matrices."""
    B = buf[0]
    out = {"batch": B, "consistency": buf[1] / B, "ce": [], "dice": [], "n_dice": [], "n_ce": []}
    off = 2
    for _ in range(n_levels):
        ce_sum, dice_sum, n_dice, n_ce = buf[off:off + 4]
        out["ce"].append(ce_sum / B)
        out["dice"].append(torch.where(n_dice > 0, dice_sum / torch.clamp(n_dice, min=1.0), torch.zeros_like(dice_sum)))
        out["n_dice"].append(n_dice)
        out["n_ce"].append(n_ce)
        off += 4
    conf = []
    for shp in conf_shapes:
        k = int(shp[0]) * int(shp[1])
        conf.append(buf[off:off + k].round().to(torch.int64).view(int(shp[0]), int(shp[1])))
        off += k
    out["confusion"] = conf
    out["total"] = sum(out["ce"]) + sum(out["dice"]) + out["consistency"]
    return out


def all_reduce_step_summary(scalars, batch_local, confusion, group=None):
    """The step's single collective.  Returns the unpacked global summary (see unpack_global)."""
    buf = pack_step_summary(scalars, batch_local, confusion)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    n = (scalars.numel() - 2) // 4
    return unpack_global(buf, n, [tuple(c.shape) for c in confusion])


def all_reduce_summary(summary: torch.Tensor, n_levels: int, conf_shapes, extra: Sequence[torch.Tensor] = (), group=None):
    """Fused-step variant: `summary` is StepOutput.summary (already in the packed layout, written by
    rhseg_step_finalize).  `extra` tensors (e.g. the head / FiLM parameter gradients) ride in the same
    all-reduce.  Returns (global summary dict, reduced extras as fp64 views)."""
    flat = [summary] + [e.reshape(-1).double() for e in extra]
    buf = torch.cat(flat) if len(flat) > 1 else summary.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    out = unpack_global(buf, n_levels, conf_shapes)
    views, off = [], summary.numel()
    for e in extra:
        views.append(buf[off:off + e.numel()].view(e.shape))
        off += e.numel()
    return out, views


def dice_grad_scale(n_dice_local: torch.Tensor, n_dice_global: torch.Tensor, world_size: int) -> torch.Tensor:
    """Factor for a rank's Dice gradient so that DDP's mean over ranks equals the gradient of the
    single-process Dice (a mean over the GLOBAL count of valid samples): world * n_local / n_global.
    CE needs no correction when every rank holds the same number of samples."""
    return torch.where(n_dice_global > 0, world_size * n_dice_local / torch.clamp(n_dice_global, min=1.0),
                       torch.zeros_like(n_dice_global))


def shard_batch(batch: int, rank: int, world: int):
    """Contiguous batch slice [start, stop) of `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
