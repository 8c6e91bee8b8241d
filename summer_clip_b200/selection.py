"""Pseudo-label selection over a train bank that is ROW-SHARDED across ranks ("Selection kernel").

`select_topk_per_label` (cache_strategy.py:48-59) picks, for every predicted class, the k most confident rows.
The row scan (confidence, label) is embarrassingly parallel over the rows, so every rank scans its own shard —
stored logits (`ops.rowconf`) or features + text classifier (`ops.rowconf_from_features`, the logits bank never
exists) — and keeps its per-class top-k (`sc_topk_per_class`).  The global answer is among those world x C x k
candidates: they are all-gathered (two small tensors, <= 8 x 16 000 entries at ImageNet scale) and merged with the
same total order the kernel uses (confidence descending — NaN largest, as torch.topk — then global row ascending),
identically on every rank.  No rank ever sees another rank's rows.

The exchange and the merge are torch plumbing (an all-gather and four stable sorts of a [C, world * k] table) and run
on whatever device the candidates live on; the scan and the local top-k are the CUDA kernels.
"""
from __future__ import annotations

import typing as tp

import torch

from . import ops


def local_candidates(conf: torch.Tensor, label: torch.Tensor, n_classes: int, k: int,
                     row_offset: int) -> tp.Tuple[torch.Tensor, torch.Tensor]:
    """This rank's per-class top-k as (confidence fp32 [C, k], GLOBAL row int64 [C, k]; -1 = no candidate).  CUDA."""
    if conf.numel() == 0:                         # an empty shard has no candidates
        return (torch.zeros((n_classes, k), dtype=torch.float32, device=conf.device),
                torch.full((n_classes, k), -1, dtype=torch.int64, device=conf.device))
    idx, _ = ops.topk_per_class(conf, label, n_classes, k)
    have = idx >= 0
    cand_conf = torch.where(have, conf.to(torch.float32)[idx.clamp_min(0)], torch.zeros((), device=conf.device))
    cand_row = torch.where(have, idx + int(row_offset), idx)
    return cand_conf, cand_row


def merge_candidates(cand_conf: torch.Tensor, cand_row: torch.Tensor, k: int) -> torch.Tensor:
    """[W, C, k] candidates of W shards -> int64 [C, k]: per class the k best by (confidence descending with NaN
    largest, row ascending), -1 padded.  Any device."""
    W, C, kk = cand_conf.shape
    conf = cand_conf.permute(1, 0, 2).reshape(C, W * kk).to(torch.float32)
    row = cand_row.permute(1, 0, 2).reshape(C, W * kk)
    none = row < 0

    def by(key: torch.Tensor, descending: bool = False):
        nonlocal conf, row, none
        order = key.argsort(dim=1, descending=descending, stable=True)
        conf, row, none = conf.gather(1, order), row.gather(1, order), none.gather(1, order)

    # stable sorts from the least to the most significant key: row ascending, confidence descending, NaN before
    # everything (sc::float_order_key puts NaN above +inf), missing candidates last
    by(row)
    by(torch.where(torch.isnan(conf), torch.full_like(conf, float("inf")), conf), descending=True)
    by(torch.isnan(conf).to(torch.int8), descending=True)
    by(none.to(torch.int8))
    row = torch.where(none, torch.full_like(row, -1), row)[:, :k]
    if row.shape[1] < k:
        row = torch.cat([row, row.new_full((C, k - row.shape[1]), -1)], dim=1)
    return row.contiguous()


def exchange_and_merge(cand_conf: torch.Tensor, cand_row: torch.Tensor, k: int, group=None) -> torch.Tensor:
    """All-gather every rank's [C, k] candidates and merge them; the same int64 [C, k] on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    shape = (world,) + tuple(cand_conf.shape)
    confs = torch.empty(world * cand_conf.numel(), dtype=torch.float32, device=cand_conf.device)
    rows = torch.empty(world * cand_row.numel(), dtype=torch.int64, device=cand_row.device)
    dist.all_gather_into_tensor(confs, cand_conf.to(torch.float32).reshape(-1).contiguous(), group=group)   # flat: gloo too
    dist.all_gather_into_tensor(rows, cand_row.reshape(-1).contiguous(), group=group)
    confs, rows = confs.view(shape), rows.view(shape)
    return merge_candidates(confs, rows, k)


def row_offsets(n_local: int, device, group=None) -> tp.List[int]:
    """Global index of the first row of every rank's shard (shards are consecutive row ranges in rank order)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    sizes[rank] = n_local
    dist.all_reduce(sizes, group=group)
    sizes = sizes.tolist()
    return [int(sum(sizes[:r])) for r in range(world + 1)]


def select_topk_per_label_sharded(conf: torch.Tensor, label: torch.Tensor, n_classes: int, k: int, group=None,
                                  row_offset: tp.Optional[int] = None) -> torch.Tensor:
    """`select_topk_per_label` over the union of all ranks' rows: LongTensor of GLOBAL row indices, classes
    ascending, most confident first within a class — the reference's order, the same on every rank.  `conf`, `label`
    cover this rank's shard; `row_offset` (default: from the ranks' shard sizes) is its first global row."""
    import torch.distributed as dist
    if row_offset is None:
        row_offset = row_offsets(conf.numel(), conf.device, group)[dist.get_rank(group)]
    cand_conf, cand_row = local_candidates(conf, label, n_classes, k, row_offset)
    flat = exchange_and_merge(cand_conf, cand_row, k, group).reshape(-1)
    return flat[flat >= 0]


def topk_select_sharded(image_outs, topk: int, prob_scale: tp.Optional[float] = None, group=None,
                        row_offset: tp.Optional[int] = None) -> torch.Tensor:
    """TopKStrategy.select (prob_scale None, cache_strategy.py:67-70) / TopKProbStrategy.select (prob_scale = the
    strategy's scale, :79-81) for a logits bank whose rows are sharded over the ranks.  `image_outs`: this rank's
    [N_local, C] logits, or a `LazyLogitsBank` over its feature shard."""
    from .clip_searcher.cache_strategy import _rowconf
    if prob_scale is None:
        conf, label = _rowconf(image_outs, prob=False)
    else:
        conf, label = _rowconf(image_outs, scale=prob_scale, prob=True)
    return select_topk_per_label_sharded(conf, label, image_outs.shape[1], topk, group, row_offset)
