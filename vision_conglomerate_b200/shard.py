"""Image sharding across the GPUs of one box.

Every unit of work on this path is per image -- decode and NMS never cross images and the loss terms
are sums over per-image matches and cells -- so the path shards by contiguous image ranges with **no
data-path collective**; this is exactly what the reference's ``DistributedSampler`` + DDP already does
(train_det.py:83-84).  The reference does not all-reduce any loss normaliser (each rank normalises by
its local match count and DDP averages gradients), so :func:`allreduce_loss_terms` is an *optional
extension*, off by default: it makes a P-rank run report the loss of the single-GPU big-batch run by
summing the per-scale numerators and normalisers (one tiny all-reduce, 15 doubles).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch


def shard_range(n_images: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous image range of `rank`: ceil-balanced, empty ranges allowed when world > n_images."""
    per, rem = divmod(n_images, world)
    start = rank * per + min(rank, rem)
    return start, start + per + (1 if rank < rem else 0)


def shard_targets(targets: torch.Tensor, start: int, end: int) -> torch.Tensor:
    """Rows of ``targets [nt, 6]`` whose image index lies in [start, end), image index rebased to the shard
    (the per-rank ``collate_fn`` output of the reference, detection_dataset.py:81-88)."""
    if targets.numel() == 0:
        return targets.reshape(0, 6)
    img = targets[:, 0]
    sel = (img >= start) & (img < end)
    out = targets[sel].clone()
    out[:, 0] -= start
    return out


def shard_batch(tensors: Sequence[torch.Tensor], world: int, rank: int) -> list:
    s, e = shard_range(tensors[0].shape[0], world, rank)
    return [t[s:e] for t in tensors]


def allreduce_loss_terms(scalars: torch.Tensor, cells: Sequence[int], cfg: dict, group=None) -> torch.Tensor:
    """``scalars [3, 8]`` float64 per scale as produced by ``bg_loss_fwd`` on this rank's shard
    (lbox, lconf, lcls, mean_ciou, avg_pos, avg_neg, M, n_neg) and ``cells[s]`` = local cell count.
    Returns the big-batch loss: means over the *global* match / cell counts (modules/detection_loss.py:107-110
    evaluated on the concatenated batch)."""
    import torch.distributed as dist
    C_cls = float(cfg["num_classes"])
    if scalars.is_cuda:
        # two tiny kernels around the one collective: no host synchronisation, ~3 launches per step
        from . import ops
        pack = ops.loss_terms_pack(scalars, cells, int(C_cls))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)
        return ops.loss_terms_combine(pack, cfg, int(C_cls))
    # host tensors (the gloo tests of the sharding logic): the same arithmetic with torch ops
    M = scalars[:, 6]
    c = torch.as_tensor(list(cells), dtype=torch.float64, device=scalars.device)
    pack = torch.stack([scalars[:, 0] * M, scalars[:, 1] * c, scalars[:, 2] * M * C_cls, M, c], dim=1).contiguous()
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)
    Mg, cg = pack[:, 3], pack[:, 4]
    lbox = torch.where(Mg > 0, pack[:, 0] / Mg.clamp(min=1), torch.zeros_like(Mg))
    lconf = pack[:, 1] / cg
    lcls = torch.where(Mg > 0, pack[:, 2] / (Mg.clamp(min=1) * C_cls), torch.zeros_like(Mg))
    sw = torch.as_tensor(cfg.get("scale_w") or [4.0, 2.0, 1.0], dtype=torch.float64, device=scalars.device)
    return (cfg["box_w"] * (sw * lbox).sum() + cfg["conf_w"] * (sw * lconf).sum() + cfg["class_w"] * (sw * lcls).sum())
