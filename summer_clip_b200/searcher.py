"""ClipSearcher — the resident engine behind the strategy classes, bench.py and the multi-GPU path.

It keeps the (selected, normalised, K-major) key bank and the transposed cache values on the
device, streams query banks through normalise -> zero-shot logits -> fused attention -> epilogue,
and — when a torch.distributed process group is given — shards the KEY bank across ranks: every
rank scores all queries against its slice of the keys and the partial O tiles are combined with one
NCCL reduce-scatter (the Tip weights exp(beta(A-1)) <= 1 need no running maximum, so the (m, l, O)
log-sum-exp merge degenerates to a sum).  The temperature-softmax mode
(`set_cache(softmax_normalize=True)`) keeps a running row maximum and merges real (m, l, O) triples.

Reference call sites this engine implements: image_attention.py:48-70 (build_cache), :80-83
(compute_clip_logits), :106-117 (weights, values, `@`, alpha sweep, accuracy).
"""
from __future__ import annotations

import os
import typing as tp

import torch

from . import ops


def shard_range(n: int, rank: int, world: int, align: int = 128) -> tp.Tuple[int, int]:
    """Contiguous key range of `rank`: shards are multiples of `align` keys (the kernel's key tile), the
    last rank takes the ragged tail."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(rank * per, n)
    hi = min(lo + per, n)
    return lo, hi


def query_slice(n_queries: int, rank: int, world: int) -> tp.Tuple[int, int]:
    """Rows of the merged result that `rank` finishes (zero-shot logits, epilogue) after the exchange."""
    per = -(-n_queries // world)
    lo = min(rank * per, n_queries)
    return lo, min(lo + per, n_queries)


def exchange_partials(part: torch.Tensor, group: tp.Any) -> tp.Tuple[torch.Tensor, int, int]:
    """Key-sharded ranks each hold a partial O_r [Nq, C] over their keys.  Sum them over ranks and leave rank r
    with the rows of ITS query slice only: one reduce-scatter moves (world-1)/world of a tile per rank (an
    all-gather would move world-1 whole tiles) and the zero-shot logits and the epilogue then cost 1/world per
    rank instead of being repeated on every rank.  Returns (rows [hi - lo, C], lo, hi)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nq = part.shape[0]
    lo, hi = query_slice(nq, rank, world)
    per = -(-nq // world)
    if dist.get_backend(group) == "nccl":
        src = part.contiguous()
        if per * world != nq:                              # ragged query count: pad the tile with zero rows
            src = torch.cat([src, src.new_zeros((per * world - nq, src.shape[1]))])
        out = torch.empty((per, src.shape[1]), dtype=src.dtype, device=src.device)
        dist.reduce_scatter_tensor(out, src, group=group)
        return out[: hi - lo], lo, hi
    # backends without reduce-scatter (gloo: the CPU tests of this host logic): all-reduce, then slice
    full = part.clone()
    dist.all_reduce(full, group=group)
    return full[lo:hi], lo, hi


class _PeerTiles:
    """Partial-tile slots in symmetric memory (torch.distributed._symmetric_memory: one allocation per rank, mapped
    into every peer over NVLink): rank r writes its partial tile of query block b into ITS slot b, and after a
    device-side barrier every rank sums the rows of its query slice straight out of all ranks' slots
    (ops.merge_peer_parts) — the key-sharded exchange without a collective kernel."""

    def __init__(self, group: tp.Any, device: torch.device, n_slots: int, rows: int, cols: int) -> None:
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.shape = (n_slots, rows, cols)
        self.buf = symm_mem.empty(self.shape, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group=group.group_name)
        self.views = [self.hdl.get_buffer(r, self.shape, torch.float32) for r in range(dist.get_world_size(group))]

    def barrier(self, channel: int) -> None:
        self.hdl.barrier(channel=channel)


class ClipSearcher:
    def __init__(self, device: tp.Union[str, torch.device] = "cuda", op_dtype: tp.Optional[torch.dtype] = None,
                 group: tp.Optional[tp.Any] = None, shard: str = "keys") -> None:
        """`group`: a torch.distributed process group (one rank per GPU).  `shard`: what the ranks split —
        "keys" (each rank keeps a contiguous slice of the key bank and scores every query against it; one
        reduce-scatter per beta; the choice for large banks and small-batch latency) or "queries" (each rank keeps
        the WHOLE bank and scores its slice of the queries; no data-path collective at all, only the predictions
        are all-gathered and the counters all-reduced; the choice for many queries against a small cache such as
        Tip-Adapter's 16-shot one)."""
        if shard not in ("keys", "queries"):
            raise ValueError("shard must be 'keys' or 'queries'")
        self.shard = shard
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ops._lib.SummerClipError("ClipSearcher needs a CUDA device: the CLIP-search path has no CPU fallback")
        self.op_dtype = ops._op(op_dtype, allow_e4m3=True)
        self.group = group
        if group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        self.text: tp.Optional[torch.Tensor] = None          # [D, C] fp32
        self.text_split: tp.Optional[tp.Tuple[torch.Tensor, torch.Tensor]] = None
        self.k_norm: tp.Optional[torch.Tensor] = None        # [Nk_local, D_pad]
        self.vt: tp.Optional[torch.Tensor] = None            # [C_pad, Nk_pad] (dense values)
        self.hard_bank: tp.Optional[ops.HardBank] = None     # one-hot values: label-sorted key bank (replaces k_norm / vt)
        self.n_keys = 0                                      # local keys
        self.n_keys_global = 0
        self.n_classes = 0
        self.rowsum_col: tp.Optional[int] = None
        self.softmax = False                                 # temperature-softmax mode: `betas` are temperatures
        self.gpu_launches = 0                                # kernels of ours launched (bench bookkeeping)
        self.events: tp.Optional[list] = None                # bench.py: set to [] to collect (name, start, end) CUDA events
        self.trace = bool(os.environ.get("SUMMER_CLIP_B200_TRACE"))   # finer marks inside the key-sharded pipeline
        # key-sharded exchange: "p2p" = peers' partial tiles read in place over NVLink (symmetric memory), "nccl" =
        # reduce-scatter; None = p2p when the process group supports it (decided at the first sharded search)
        self.exchange: tp.Optional[str] = os.environ.get("SUMMER_CLIP_B200_EXCHANGE") or None
        self._peer_tiles: tp.Dict[tuple, _PeerTiles] = {}

    # ------------------------------------------------------------------ bank
    def set_text(self, text_features: torch.Tensor) -> None:
        """Zero-shot classifier T [D, C] (eval_clip.zeroshot_classifier output; an input of this path)."""
        self.text = text_features.to(self.device, non_blocking=True).float().contiguous()
        self.text_split = ops.text_split(self.text)           # split-fp16 rows of T^T for the tensor-core GEMM

    def set_cache(self, cache_image_features: torch.Tensor, cache_image_outs: tp.Optional[torch.Tensor],
                  idx: tp.Optional[torch.Tensor] = None, *, feature_major: bool = True,
                  softmax_scale: tp.Optional[float] = None, labels: tp.Optional[torch.Tensor] = None,
                  n_classes: tp.Optional[int] = None, softmax_normalize: bool = False, local_shard: bool = False) -> None:
        """Build the resident cache: K[:, idx] normalised/cast (image_attention.py:54-55 +
        cache_weights_strategy.py:20) and V = f(L[idx]) (cache_value_strategy.py).  With a process group the
        selected keys are sharded contiguously across ranks.  `labels` (gold, already per selected key)
        replaces argmax(L) (one-hot values only).
        `softmax_normalize=True` selects the temperature-softmax mode (north-star extension, no reference
        implementation): `search(betas=...)` then computes softmax_k(beta * A) @ V with a running row maximum
        instead of the Tip-Adapter weights exp(-beta (1 - A)) @ V.
        `local_shard=True` (key-sharded groups): the tensors given ARE this rank's key shard already (a bank too
        large to exist on one rank, or generated / loaded per rank); nothing is sliced."""
        if labels is not None and softmax_scale is not None:
            raise ValueError("set_cache: `labels` replace the argmax of one-hot values; softmax values are built from "
                             "cache_image_outs[idx]")
        feats = cache_image_features.to(self.device, non_blocking=True)
        n_total = feats.shape[1] if feature_major else feats.shape[0]
        if idx is not None:
            idx = idx.to(self.device).to(torch.int64)
            n_sel = idx.numel()
        else:
            n_sel = n_total
        slice_here = self.world > 1 and self.shard == "keys" and not local_shard
        lo, hi = shard_range(n_sel, self.rank, self.world) if slice_here else (0, n_sel)
        if slice_here or idx is not None:
            local_idx = idx[lo:hi] if idx is not None else torch.arange(lo, hi, device=self.device)
        else:
            local_idx = None
        self.n_keys_global, self.n_keys = n_sel, hi - lo
        if local_shard and self.world > 1 and self.shard == "keys":
            import torch.distributed as dist
            total = torch.tensor([n_sel], dtype=torch.int64, device=self.device)
            dist.all_reduce(total, group=self.group)
            self.n_keys_global = int(total.item())
        if cache_image_outs is not None:
            outs = cache_image_outs.to(self.device, non_blocking=True)
            self.n_classes = outs.shape[1]
        else:
            outs = None
            assert n_classes is not None
            self.n_classes = n_classes
        self.k_norm = self.vt = self.hard_bank = None
        # decided from the arguments alone, so that every rank of a key-sharded group — including one whose shard
        # is empty — agrees on the mode and on the width of the partial tiles
        self.softmax = bool(softmax_normalize)
        hard = softmax_scale is None and ops.hard_supported(self.n_classes)
        self.rowsum_col = self.n_classes if (softmax_normalize and not hard) else None
        if self.n_keys == 0:
            return
        local_labels = labels.to(self.device)[lo:hi].contiguous() if labels is not None else None
        if hard:
            # one-hot values: W @ V is a per-class segmented row sum; sort the keys by label once and let the
            # kernel sum the exponentials per class straight out of tensor memory (sc_attn_fwd_hard).  The bank is
            # built in ONE pass over the raw features: labels -> layout -> normalised rows written to sorted places.
            labels16 = ops.hard_labels(outs, self.n_classes, idx=None if local_labels is not None else local_idx,
                                       labels=local_labels)
            self.hard_bank = ops.hard_bank_build(labels16[: self.n_keys], self.n_classes, feats, feature_major,
                                                 idx=local_idx, op_dtype=self.op_dtype)
            self.gpu_launches += 6
            return
        self.k_norm = ops.normalize_cast(feats, feature_major=feature_major, idx=local_idx, op_dtype=self.op_dtype)
        self.gpu_launches += 1
        self.vt = ops.values_prepare(outs, self.n_classes, idx=None if local_labels is not None else local_idx,
                                     labels=local_labels, softmax_scale=softmax_scale, ones_row=softmax_normalize,
                                     op_dtype=self.op_dtype)
        self.gpu_launches += 2 + int(softmax_normalize)

    def save_bank(self, directory, key: str = ""):
        """Write the resident cache as a sidecar directory (bank_io): the label-sorted bank of a one-hot cache, or the
        normalised keys + transposed values of a dense-value cache."""
        from . import bank_io
        if self.hard_bank is not None:
            return bank_io.save_hard_bank(self.hard_bank, directory, key)
        if self.k_norm is None or self.vt is None:
            raise ops._lib.SummerClipError("save_bank: no resident cache")
        if self.softmax:
            raise ops._lib.SummerClipError("save_bank: temperature-softmax caches are rebuilt from their sources")
        return bank_io.save_dense_bank(self.k_norm, self.vt, self.n_keys, self.n_classes, directory, key)

    def load_bank(self, directory, key: tp.Optional[str] = None) -> bool:
        """Make a sidecar bank resident instead of calling set_cache.  False (nothing changed) if it is absent, of
        another operand type or was built for another key."""
        from . import bank_io
        bank = bank_io.load_hard_bank(directory, self.device, key)
        if bank is not None:
            if bank.rows.dtype != self.op_dtype:
                return False
            self.hard_bank, self.k_norm, self.vt, self.rowsum_col, self.softmax = bank, None, None, None, False
            self.n_keys = self.n_keys_global = bank.n_keys
            self.n_classes = bank.n_classes
            return True
        dense = bank_io.load_dense_bank(directory, self.device, key)
        if dense is None or dense[0].dtype != self.op_dtype:
            return False
        self.k_norm, self.vt, self.n_keys, self.n_classes = dense
        self.hard_bank, self.rowsum_col, self.softmax = None, None, False
        self.n_keys_global = self.n_keys
        return True

    # ------------------------------------------------------------------ queries
    def prepare_queries(self, test_image_features: torch.Tensor, feature_major: bool = True, overlap: bool = False):
        """H2D (if needed) + normalise/cast + zero-shot logits.  Returns (Qn, Z or None[, event]).  `overlap`: the
        zero-shot GEMM (independent of the attention) is issued on the side stream FIRST — it takes a handful of SMs
        for ~10 us while the attention CTAs fill the rest — and the third return value is the event the consumer
        of Z must wait for (small batches: ~20 us less on a 0.5 ms search)."""
        q = test_image_features.to(self.device, non_blocking=True)
        z, evz = None, None
        if self.text is not None and overlap:
            main, side = torch.cuda.current_stream(self.device), self._side_stream()
            start = torch.cuda.Event()
            start.record(main)
            with torch.cuda.stream(side):
                side.wait_event(start)
                z = ops.zero_shot_logits(q, feature_major, self.text, scale=100.0, normalize=True, t_split=self.text_split)
                evz = torch.cuda.Event()
                evz.record(side)
            self.gpu_launches += 2
        qn = ops.normalize_cast(q, feature_major=feature_major, op_dtype=self.op_dtype)
        self.gpu_launches += 1
        if self.text is not None and not overlap:
            z = ops.zero_shot_logits(q, feature_major, self.text, scale=100.0, normalize=True, t_split=self.text_split)
            self.gpu_launches += 2
        return (qn, z, evz) if overlap else (qn, z)

    def local_cache_logits(self, qn: torch.Tensor, beta: float, splits: int = 0) -> torch.Tensor:
        """O_r = exp(-beta (1 - Qn Kn^T)) @ V over THIS rank's keys: fp32 [Nq, C]."""
        assert not self.softmax, "temperature-softmax caches go through local_softmax_partials"
        n_cols = self.n_classes
        nq = qn.shape[0]
        if self.n_keys > 0:
            if self.hard_bank is not None:
                if splits <= 0:
                    splits = ops.attn_hard_splits(nq, self.hard_bank.n_sorted, self.device, bank=self.hard_bank)
                part = ops.attn_fwd_hard(qn, self.hard_bank, beta, splits=splits)
            else:
                part = ops.attn_fwd(qn, self.k_norm, self.vt, self.n_keys, n_cols, beta, splits=splits, merge=True)
            self.gpu_launches += 1 + int(splits > 1)
        else:
            part = torch.zeros((nq, n_cols), dtype=torch.float32, device=self.device)
        return part

    def local_softmax_partials(self, qn: torch.Tensor, tau: float) -> tp.Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Temperature-softmax mode, THIS rank's keys: the partial triple (O [Nq, C], m [Nq], l [Nq]) with
        O[q, c] = sum_k 2^(tau log2(e) A[q, k] - m[q]) V[k, c], l[q] = sum_k 2^(...), m in base-2 exponent units —
        what `ops.merge_softmax` merges across key shards (the north star's (max, sum, O) tiles).
        One-hot values: the segmented kernel keeps a running row maximum and emits per-class log-sum-exp tiles.
        Dense values: the weights are rounded to 16 bits for GEMM-2, so the exact row maximum comes from a
        tensor-core pre-pass (`ops.attn_rowmax`) and is subtracted inside the dual-GEMM kernel."""
        nq, c = qn.shape[0], self.n_classes
        if self.n_keys == 0:
            return (torch.zeros((nq, c), dtype=torch.float32, device=self.device),
                    torch.full((nq,), float("-inf"), dtype=torch.float32, device=self.device),
                    torch.zeros((nq,), dtype=torch.float32, device=self.device))
        if self.hard_bank is not None:
            lse = ops.attn_softmax_hard(qn, self.hard_bank, tau)
            self.gpu_launches += 3
            return ops.softmax_partials(lse)
        rowmax = ops.attn_rowmax(qn, self.k_norm, self.n_keys)
        o = ops.attn_fwd(qn, self.k_norm, self.vt, self.n_keys, c + 1, tau, merge=True, row_shift=rowmax)
        self.gpu_launches += 5
        out, m, l = ops.merge_softmax(o[None, :, :c], rowmax[None], o[None, :, c].contiguous(),
                                      m_scale=float(tau) * 1.4426950408889634, normalize=False)
        return out, m, l

    def softmax_logits(self, qn: torch.Tensor, tau: float) -> torch.Tensor:
        """softmax_k(tau * Qn Kn^T) @ V over ALL keys (merged over the key shards): fp32 [Nq, C] on every rank."""
        o, m, l = self.local_softmax_partials(qn, tau)
        if self.world > 1 and self.shard == "keys":
            import torch.distributed as dist
            m_all = m.clone()
            dist.all_reduce(m_all, op=dist.ReduceOp.MAX, group=self.group)
            o, _, l = ops.merge_softmax(o[None], m[None], l[None], m_ref=m_all, normalize=False, inplace=True)
            dist.all_reduce(o, group=self.group)
            dist.all_reduce(l, group=self.group)
            m = m_all
        out, _, _ = ops.merge_softmax(o[None], m[None], l[None], normalize=True, inplace=True)
        self.gpu_launches += 1
        return out

    def local_cache_logits_many(self, qn: torch.Tensor, betas: tp.Sequence[float]) -> tp.List[torch.Tensor]:
        """`local_cache_logits` for a list of betas; a one-hot bank shares each tensor-core pass between four betas
        (ops.attn_fwd_hard_multi), every result bit-identical to its single-beta launch."""
        betas = [float(b) for b in betas]
        if self.softmax:
            return [self.softmax_logits(qn, b) for b in betas]
        if self.hard_bank is None or self.n_keys == 0 or len(betas) < 2:
            return [self.local_cache_logits(qn, b) for b in betas]
        splits = ops.attn_hard_splits(qn.shape[0], self.hard_bank.n_sorted, self.device, bank=self.hard_bank)
        outs = ops.attn_fwd_hard_multi(qn, self.hard_bank, betas, splits=splits)
        self.gpu_launches += -(-len(betas) // 4) + (len(betas) if splits > 1 else 0)
        return outs

    def cache_logits(self, qn: torch.Tensor, beta: float, splits: int = 0) -> torch.Tensor:
        """O over ALL keys for every query, on every rank (all-reduce of the per-rank partials).  `search` uses
        the cheaper reduce-scatter (`exchange_partials`) instead."""
        part = self.local_cache_logits(qn, beta, splits)
        if self.world == 1:
            return part
        import torch.distributed as dist
        dist.all_reduce(part, group=self.group)
        return part

    def search(self, test_image_features: torch.Tensor, betas: tp.Sequence[float], alphas: tp.Sequence[float],
               labels: tp.Optional[torch.Tensor] = None, feature_major: bool = True, want_logits: bool = False,
               want_pred: bool = True, want_cache_logits: bool = True, query_shard: tp.Union[bool, int] = False,
               blocks: tp.Optional[int] = None) -> tp.List[tp.Dict[str, tp.Any]]:
        """One pass of the hot path for a query bank: for every beta one fused attention launch, then one
        epilogue launch covering every alpha.  Returns one dict per beta with device tensors
        pred [na, Nq], top1/top5 [na] (if labels), logits [na, Nq, C] (if requested), cache_logits (if requested;
        otherwise the key-split tiles go to the epilogue unmerged and no [Nq, C] sum is ever written).
        Under a process group: `query_shard=True` (or the total query count) says that `test_image_features` /
        `labels` hold only THIS rank's query slice (`query_slice(nq_total, rank, world)`, e.g. copied from host by each
        rank on its own PCIe link);
        `blocks` = query blocks of the key-sharded pipeline (None: chosen from the batch size)."""
        if self.world > 1:
            return self._search_sharded(test_image_features, betas, alphas, labels, feature_major, want_logits, want_pred,
                                        query_shard, blocks)
        qn, z, evz = self.prepare_queries(test_image_features, feature_major, overlap=True)
        if labels is not None:
            labels = labels.to(self.device, non_blocking=True)
        results = []
        parts = self._local_parts_many(qn, betas, merge=want_cache_logits)
        if evz is not None:
            torch.cuda.current_stream(self.device).wait_event(evz)
        for beta, o in zip(betas, parts):
            res = ops.epilogue(z, o, alphas, labels=labels, want_logits=want_logits, want_pred=want_pred)
            self.gpu_launches += 1
            res["beta"] = float(beta)
            res["cache_logits"] = o if want_cache_logits else None
            res["clip_logits"] = z
            results.append(res)
        return results

    def _local_parts_many(self, qn: torch.Tensor, betas: tp.Sequence[float], merge: bool) -> tp.List[torch.Tensor]:
        """This rank's cache logits per beta: merged [Nq, C], or (merge=False, Tip mode) the unmerged key-split
        tiles [splits, Nq, C] that `ops.epilogue` / `ops.merge_partials` sum."""
        if merge or self.softmax or self.n_keys == 0:
            return self.local_cache_logits_many(qn, betas)
        betas = [float(b) for b in betas]
        nq = qn.shape[0]
        t0 = self._mark()
        if self.hard_bank is not None:
            splits = ops.attn_hard_splits(nq, self.hard_bank.n_sorted, self.device, bank=self.hard_bank)
            outs = ops.attn_fwd_hard_multi(qn, self.hard_bank, betas, splits=splits, merge=False)
            self.gpu_launches += -(-len(betas) // 4)
        else:       # dense values: L2-blocked key splits, summed per query chunk inside attn_fwd (bounded partial tiles)
            outs = [ops.attn_fwd(qn, self.k_norm, self.vt, self.n_keys, self.n_classes, b, merge=True) for b in betas]
            self.gpu_launches += 2 * len(betas)
        self._mark("attention", t0)
        return outs

    # device-time bookkeeping for bench.py: (name, start event, end event) triples on the launching stream
    def _mark(self, name: tp.Optional[str] = None, start: tp.Any = None):
        if self.events is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        if name is not None and start is not None:
            self.events.append((name, start, ev))
        return ev

    def capture_search(self, test_image_features: torch.Tensor, betas: tp.Sequence[float], alphas: tp.Sequence[float],
                       labels: tp.Optional[torch.Tensor] = None, feature_major: bool = True, warmup: int = 3):
        """Capture one `search` for a FIXED query-batch shape into a CUDA graph (online serving: a small batch is a
        dozen ~10 us kernels plus, when key-sharded, three small NCCL calls, so launch latency dominates).
        Returns (graph, results): copy new queries / labels INTO the given tensors (they are the graph's static
        inputs), `graph.replay()`, read the tensors in `results`.  Works on one rank and under a process group
        (NCCL collectives are captured; every rank must replay)."""
        q = test_image_features
        assert q.is_cuda and (labels is None or labels.is_cuda), "static inputs must live on the device"
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                               # allocator / NCCL warm-up outside the capture
                self.search(q, betas, alphas, labels=labels, feature_major=feature_major)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            results = self.search(q, betas, alphas, labels=labels, feature_major=feature_major)
        return graph, results

    def _search_sharded(self, test_image_features, betas, alphas, labels, feature_major, want_logits, want_pred,
                        query_shard=False, blocks=None):
        """`search` under a process group.  Every rank returns the same pred / top1 / top5; `logits`, `cache_logits`,
        `clip_logits` cover the rank's own rows only (a list of (lo, hi, tensor) pieces in `pieces`).

        shard = "queries": every rank scores its query slice against the whole bank — no data-path collective.
        shard = "keys": every rank scores ALL queries against its key shard; per query block one reduce-scatter sums
        the partial tiles and hands each rank its slice of the block, which it finishes alone (zero-shot logits,
        alpha epilogue).  The exchange and the finishing of block b run on a side stream UNDER the attention of
        block b + 1, so only the last block's exchange is exposed (at 8 GPUs the reduce-scatter of a whole
        50 000 x 1000 tile was 0.93 ms of a 12.5 ms step: profiles/r01g_bench_n8_phases.json).  Predictions are
        all-gathered per block, the accuracy counters all-reduced once."""
        import torch.distributed as dist
        dev, world, rank = self.device, self.world, self.rank
        q = test_image_features.to(dev, non_blocking=True)
        n_given = q.shape[1] if feature_major else q.shape[0]
        na = len(alphas)
        t0 = self._mark()
        qn_mine = ops.normalize_cast(q, feature_major=feature_major, op_dtype=self.op_dtype)
        self.gpu_launches += 1
        if query_shard:
            if query_shard is True:
                # the total is recovered from the ranks' row counts (one small all-reduce and a host sync; pass the
                # total query count as `query_shard=nq_total` to avoid both)
                sizes = torch.zeros(world, dtype=torch.int64, device=dev)
                sizes[rank] = n_given
                dist.all_reduce(sizes, group=self.group)
                sizes = sizes.tolist()
            else:
                sizes = [hi_ - lo_ for lo_, hi_ in (query_slice(int(query_shard), r, world) for r in range(world))]
                assert sizes[rank] == n_given, "query_shard=nq_total: this rank must hold query_slice(nq_total, rank, world)"
            nq = int(sum(sizes))
            per = max(sizes)
            my_lo = int(sum(sizes[:rank]))
        else:
            nq, per, my_lo = n_given, -(-n_given // world), None
        lab = labels.to(dev, non_blocking=True).to(torch.int32) if labels is not None else None
        results = [{"beta": float(b), "pred": None, "top1": None, "top5": None, "logits": None, "pieces": []} for b in betas]
        counts = torch.zeros((len(betas), 2, na), dtype=torch.int32, device=dev)

        if self.shard == "queries":
            lo, hi = (my_lo, my_lo + n_given) if query_shard else query_slice(nq, rank, world)
            q_mine = q if query_shard else (q[:, lo:hi] if feature_major else q[lo:hi])
            qn = qn_mine if query_shard else qn_mine[lo:hi]
            lab_mine = (lab if query_shard else lab[lo:hi].contiguous()) if lab is not None else None
            self._mark("normalize_queries", t0)
            z = None
            if self.text is not None and hi > lo:
                z = ops.zero_shot_logits(q_mine, feature_major, self.text, scale=100.0, normalize=True, t_split=self.text_split)
                self.gpu_launches += 1
            parts = self._local_parts_many(qn, betas, merge=False) if hi > lo else [None] * len(betas)
            preds = []
            for bi, o in enumerate(parts):
                mine = torch.zeros((na, per), dtype=torch.int32, device=dev)
                if o is not None:
                    r = ops.epilogue(z, o, alphas, labels=lab_mine, want_logits=want_logits, want_pred=want_pred)
                    self.gpu_launches += 1
                    if lab_mine is not None:
                        counts[bi, 0], counts[bi, 1] = r["top1"], r["top5"]
                    if want_pred:
                        mine[:, : hi - lo] = r["pred"]
                    results[bi]["pieces"].append((lo, hi, r["logits"], o, z))
                preds.append(mine)
            if want_pred:
                gathered = torch.empty((world, len(betas), na, per), dtype=torch.int32, device=dev)
                dist.all_gather_into_tensor(gathered, torch.stack(preds), group=self.group)
                for bi in range(len(betas)):
                    if query_shard:
                        results[bi]["pred"] = torch.cat([gathered[r, bi, :, : sizes[r]] for r in range(world)], dim=1)
                    else:
                        results[bi]["pred"] = gathered[:, bi].permute(1, 0, 2).reshape(na, world * per)[:, :nq].contiguous()
        else:
            # ---- key shards: all queries on every rank
            if query_shard:
                pad = qn_mine if n_given == per else torch.cat([qn_mine, qn_mine.new_zeros((per - n_given, qn_mine.shape[1]))])
                allq = torch.empty((world * per, qn_mine.shape[1]), dtype=qn_mine.dtype, device=dev)
                dist.all_gather_into_tensor(allq, pad, group=self.group)            # normalised rows over NVLink
                qn = allq[:nq] if all(sz == per for sz in sizes[:-1]) else torch.cat([allq[r * per: r * per + sizes[r]] for r in range(world)])
            else:
                qn = qn_mine
            self._mark("normalize_queries", t0)
            if blocks is None:
                blocks = 4 if (nq >= 16384 and not self.softmax) else 1
            blocks = max(1, min(int(blocks), nq))
            edges = [nq * b // blocks for b in range(blocks + 1)]
            lab_all = self._all_labels(lab, query_shard, sizes if query_shard else None, per, nq)
            main = torch.cuda.current_stream(dev)
            side = self._side_stream()
            bmax = max(edges[b + 1] - edges[b] for b in range(blocks))
            nbeta = len(betas)
            tiles = None if self.softmax else self._tiles_for("o", blocks * nbeta, bmax, self.n_classes)
            # predictions and counters travel the same way: every rank writes its rows (as exact fp32 integers, zeros
            # elsewhere) and its local counters into its own peer-mapped result tile; one summing pass over all
            # ranks' tiles at the end is the all-gather of the predictions AND the all-reduce of the counters
            rtile = self._tiles_for("r", nbeta, na, nq + 2) if tiles is not None else None
            if rtile is not None:
                rtile.buf.zero_()
            keep = []                                   # tensors the side stream still reads: alive until the final join
            preds = [torch.zeros((na, nq), dtype=torch.int32, device=dev) for _ in betas] if (want_pred and rtile is None) else None
            # zero-shot logits of MY rows of every block, up front on the side stream: the tensor-core GEMM needs whole
            # SMs (224 KB of shared memory) and would otherwise queue behind the attention CTAs of the next block
            slices = []
            for b in range(blocks):
                slo, shi = query_slice(edges[b + 1] - edges[b], rank, world)
                slices.append((edges[b] + slo, edges[b] + shi))
            start = torch.cuda.Event()
            start.record(main)
            zs: tp.List[tp.Optional[torch.Tensor]] = [None] * blocks
            with torch.cuda.stream(side):
                side.wait_event(start)
                if self.text is not None:
                    for b, (glo, ghi) in enumerate(slices):
                        if ghi <= glo:
                            continue
                        if query_shard:                     # the raw rows may live on another rank: use the normalised ones
                            zs[b] = ops.zero_shot_logits(qn[glo:ghi], False, self.text, scale=100.0, normalize=False, t_split=self.text_split)
                        else:
                            zs[b] = ops.zero_shot_logits(q[:, glo:ghi] if feature_major else q[glo:ghi], feature_major, self.text,
                                                         scale=100.0, normalize=True, t_split=self.text_split)
                        self.gpu_launches += 2
            t_last = None
            for b in range(blocks):
                b0, b1 = edges[b], edges[b + 1]
                nb = b1 - b0
                parts = self._local_parts_many(qn[b0:b1], betas, merge=False)
                t_last = self._mark()
                ready = torch.cuda.Event(enable_timing=self.events is not None)
                ready.record(main)
                keep.append(parts)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    ts = self._mark()
                    # sum of the key-split tiles: beside the next block's attention (no shared memory: co-resident)
                    merged = []
                    for bi, o in enumerate(parts):
                        if self.softmax:
                            merged.append(o)
                        elif tiles is not None:         # ... straight into my slot of the peer-mapped buffer
                            merged.append(ops.merge_partials(o if o.dim() == 3 else o[None], out=tiles.buf[b * nbeta + bi, :nb]))
                            self.gpu_launches += 1
                        else:
                            merged.append(ops.merge_partials(o) if o.dim() == 3 and o.shape[0] > 1 else (o[0] if o.dim() == 3 else o))
                            self.gpu_launches += int(o.dim() == 3 and o.shape[0] > 1)
                    keep.append(merged)
                    tsub = self._mark(f"b{b}.merge_splits", ts) if self.trace else None
                    slo, shi = query_slice(nb, rank, world)
                    bper = -(-nb // world)
                    glo, ghi = b0 + slo, b0 + shi          # my rows of this block, global numbering
                    z = zs[b]
                    lab_mine = lab_all[glo:ghi].contiguous() if lab_all is not None else None
                    if tiles is not None:
                        tiles.barrier(b)                    # every rank's slots of block b are complete
                    tsub = self._mark(f"b{b}.barrier", tsub) if self.trace else None
                    for bi, o in enumerate(merged):
                        if self.softmax:            # already merged over the key shards (log-sum-exp merge, softmax_logits)
                            o_mine = o[slo:shi]
                        elif tiles is not None:
                            o_mine = None
                            if shi > slo:
                                o_mine = ops.merge_peer_parts([v[b * nbeta + bi, slo:shi] for v in tiles.views])
                                self.gpu_launches += 1
                                tsub = self._mark(f"b{b}.peer_merge", tsub) if self.trace else None
                        else:
                            o_mine, _, _ = exchange_partials(o, self.group)
                        r = None
                        if ghi > glo:
                            r = ops.epilogue(z, o_mine, alphas, labels=lab_mine, want_logits=want_logits, want_pred=want_pred)
                            self.gpu_launches += 1
                            if lab_mine is not None:
                                counts[bi, 0] += r["top1"]
                                counts[bi, 1] += r["top5"]
                            results[bi]["pieces"].append((glo, ghi, r["logits"], o_mine, z))
                            tsub = self._mark(f"b{b}.epilogue", tsub) if self.trace else None
                        if want_pred and rtile is not None:
                            if r is not None:
                                rtile.buf[bi, :, glo:ghi] = r["pred"]          # int32 -> exact fp32
                        elif want_pred:
                            mine = torch.zeros((na, bper), dtype=torch.int32, device=dev)
                            if r is not None:
                                mine[:, : shi - slo] = r["pred"]
                            gathered = torch.empty((world, na, bper), dtype=torch.int32, device=dev)
                            dist.all_gather_into_tensor(gathered, mine, group=self.group)
                            preds[bi][:, b0:b1] = gathered.permute(1, 0, 2).reshape(na, world * bper)[:, :nb]
                            keep.append((gathered, mine))
                        keep.append((o_mine, z, r))
                    self._mark("exchange_and_finish", ts)
                    if self.events is not None:             # attention-of-block-done -> block finished, per block
                        self._mark(f"block{b}_ready_to_finished", ready)
            with torch.cuda.stream(side):
                if rtile is not None:
                    tf = self._mark()
                    if lab is not None:
                        rtile.buf[:, :, nq:] = counts.permute(0, 2, 1)          # [nbeta, na, 2] local counters
                    rtile.barrier(blocks)               # every rank's result tile is complete
                    for bi in range(nbeta):
                        total = ops.merge_peer_parts([v[bi] for v in rtile.views]).to(torch.int32)     # [na, nq + 2]
                        self.gpu_launches += 1
                        if want_pred:
                            results[bi]["pred"] = total[:, :nq]
                        if lab is not None:
                            results[bi]["top1"], results[bi]["top5"] = total[:, nq].contiguous(), total[:, nq + 1].contiguous()
                    rtile.barrier(blocks + 1)           # nobody rewrites a tile (next search) while a peer still reads it
                    self._mark("gather_results", tf)
                done = torch.cuda.Event()
                done.record(side)
            main.wait_event(done)
            self._mark("tail_after_last_attention", t_last)
            if preds is not None:
                for bi in range(nbeta):
                    results[bi]["pred"] = preds[bi]
            del keep
            if rtile is not None:
                lab = None                              # counters already summed over the ranks
        if lab is not None:
            dist.all_reduce(counts, group=self.group)
            for bi in range(len(betas)):
                results[bi]["top1"], results[bi]["top5"] = counts[bi, 0], counts[bi, 1]
        for res in results:                                   # single-piece conveniences (one block / query shards)
            if len(res["pieces"]) == 1:
                res["lo"], res["hi"], res["logits"], res["cache_logits"], res["clip_logits"] = res["pieces"][0]
        return results

    def _tiles_for(self, kind: str, n_slots: int, rows: int, cols: int) -> tp.Optional[_PeerTiles]:
        """The peer-mapped tile slots of this kind ("o" partial cache logits, "r" results) and shape, or None when the
        exchange is NCCL's.  Collective: every rank calls it with the same shape (the first call per shape allocates
        and exchanges handles; a new shape replaces the previous allocation of its kind)."""
        import torch.distributed as dist
        if self.exchange == "nccl":
            return None
        key = (kind, n_slots, rows, cols)
        if key not in self._peer_tiles:
            ok = torch.ones(1, dtype=torch.int32, device=self.device)
            tiles = None
            try:
                if dist.get_backend(self.group) != "nccl":
                    raise RuntimeError("symmetric memory needs the NCCL group of one node")
                tiles = _PeerTiles(self.group, self.device, n_slots, rows, cols)
            except Exception:  # noqa: BLE001  (no peer mapping on this system: fall back to the collective)
                if self.exchange == "p2p":
                    raise
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                self.exchange = "nccl"
                return None
            self._peer_tiles = {k: v for k, v in self._peer_tiles.items() if k[0] != kind}
            self._peer_tiles[key] = tiles
        return self._peer_tiles[key]

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device, priority=-1)
        return self._side

    def _all_labels(self, lab, query_shard, sizes, per, nq):
        """Labels of ALL queries on this rank (key shards finish rows that another rank brought in)."""
        if lab is None or not query_shard:
            return lab
        cached = getattr(self, "_lab_all_cache", None)
        if cached is not None and cached[0] is lab:
            return cached[1]
        import torch.distributed as dist
        pad = lab if lab.numel() == per else torch.cat([lab, lab.new_zeros(per - lab.numel())])
        allv = torch.empty(self.world * per, dtype=lab.dtype, device=self.device)
        dist.all_gather_into_tensor(allv, pad, group=self.group)
        out = allv[:nq] if all(sz == per for sz in sizes[:-1]) else torch.cat([allv[r * per: r * per + sizes[r]] for r in range(self.world)])
        self._lab_all_cache = (lab, out)
        return out
