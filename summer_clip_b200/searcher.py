"""ClipSearcher — the resident engine behind the strategy classes, bench.py and the multi-GPU path.

It keeps the (selected, normalised, K-major) key bank and the transposed cache values on the
device, streams query banks through normalise -> zero-shot logits -> fused attention -> epilogue,
and — when a torch.distributed process group is given — shards the KEY bank across ranks: every
rank scores all queries against its slice of the keys and the partial O tiles are combined with one
NCCL all-gather + the merge kernel (the Tip weights exp(beta(A-1)) <= 1 need no running maximum, so
the (m, l, O) log-sum-exp merge degenerates to a sum).

Reference call sites this engine implements: image_attention.py:48-70 (build_cache), :80-83
(compute_clip_logits), :106-117 (weights, values, `@`, alpha sweep, accuracy).
"""
from __future__ import annotations

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


class ClipSearcher:
    def __init__(self, device: tp.Union[str, torch.device] = "cuda", op_dtype: tp.Optional[torch.dtype] = None,
                 group: tp.Optional[tp.Any] = None) -> None:
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ops._lib.SummerClipError("ClipSearcher needs a CUDA device: the CLIP-search path has no CPU fallback")
        self.op_dtype = ops._op(op_dtype)
        self.group = group
        if group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        self.text: tp.Optional[torch.Tensor] = None          # [D, C] fp32
        self.k_norm: tp.Optional[torch.Tensor] = None        # [Nk_local, D_pad]
        self.vt: tp.Optional[torch.Tensor] = None            # [C_pad, Nk_pad] (dense values)
        self.hard_bank: tp.Optional[ops.HardBank] = None     # one-hot values: label-sorted key bank (replaces k_norm / vt)
        self.n_keys = 0                                      # local keys
        self.n_keys_global = 0
        self.n_classes = 0
        self.rowsum_col: tp.Optional[int] = None
        self.gpu_launches = 0                                # kernels of ours launched (bench bookkeeping)

    # ------------------------------------------------------------------ bank
    def set_text(self, text_features: torch.Tensor) -> None:
        """Zero-shot classifier T [D, C] (eval_clip.zeroshot_classifier output; an input of this path)."""
        self.text = text_features.to(self.device, non_blocking=True).float().contiguous()

    def set_cache(self, cache_image_features: torch.Tensor, cache_image_outs: tp.Optional[torch.Tensor],
                  idx: tp.Optional[torch.Tensor] = None, *, feature_major: bool = True,
                  softmax_scale: tp.Optional[float] = None, labels: tp.Optional[torch.Tensor] = None,
                  n_classes: tp.Optional[int] = None, softmax_normalize: bool = False) -> None:
        """Build the resident cache: K[:, idx] normalised/cast (image_attention.py:54-55 +
        cache_weights_strategy.py:20) and V = f(L[idx]) (cache_value_strategy.py).  With a process group the
        selected keys are sharded contiguously across ranks.  `labels` (gold, already per selected key)
        replaces argmax(L)."""
        feats = cache_image_features.to(self.device, non_blocking=True)
        n_total = feats.shape[1] if feature_major else feats.shape[0]
        if idx is not None:
            idx = idx.to(self.device).to(torch.int64)
            n_sel = idx.numel()
        else:
            n_sel = n_total
        lo, hi = shard_range(n_sel, self.rank, self.world)
        if self.world > 1 or idx is not None:
            local_idx = idx[lo:hi] if idx is not None else torch.arange(lo, hi, device=self.device)
        else:
            local_idx = None
        self.n_keys_global, self.n_keys = n_sel, hi - lo
        if cache_image_outs is not None:
            outs = cache_image_outs.to(self.device, non_blocking=True)
            self.n_classes = outs.shape[1]
        else:
            outs = None
            assert n_classes is not None
            self.n_classes = n_classes
        self.k_norm = self.vt = self.hard_bank = None
        if self.n_keys == 0:
            return
        self.k_norm = ops.normalize_cast(feats, feature_major=feature_major, idx=local_idx, op_dtype=self.op_dtype)
        self.gpu_launches += 1
        local_labels = labels.to(self.device)[lo:hi] if labels is not None else None
        self.rowsum_col = None
        if softmax_scale is None and not softmax_normalize and ops.hard_supported(self.n_classes):
            # one-hot values: W @ V is a per-class segmented row sum; sort the keys by label once and let the
            # kernel sum the exponentials per class straight out of tensor memory (sc_attn_fwd_hard)
            labels16 = ops.hard_labels(outs, self.n_classes, idx=None if local_labels is not None else local_idx,
                                       labels=local_labels)
            self.hard_bank = ops.hard_bank_layout(labels16[: self.n_keys], self.n_classes).gather(self.k_norm)
            self.k_norm = None
            self.gpu_launches += 1
            return
        self.vt = ops.values_prepare(outs, self.n_classes, idx=None if local_labels is not None else local_idx,
                                     labels=local_labels, softmax_scale=softmax_scale, ones_row=softmax_normalize,
                                     op_dtype=self.op_dtype)
        self.gpu_launches += 2 + int(softmax_normalize)
        self.rowsum_col = self.n_classes if softmax_normalize else None

    # ------------------------------------------------------------------ queries
    def prepare_queries(self, test_image_features: torch.Tensor, feature_major: bool = True):
        """H2D (if needed) + normalise/cast + zero-shot logits.  Returns (Qn, Z or None)."""
        q = test_image_features.to(self.device, non_blocking=True)
        qn = ops.normalize_cast(q, feature_major=feature_major, op_dtype=self.op_dtype)
        self.gpu_launches += 1
        z = None
        if self.text is not None:
            z = ops.zero_shot_logits(q, feature_major, self.text, scale=100.0, normalize=True)
            self.gpu_launches += 1
        return qn, z

    def cache_logits(self, qn: torch.Tensor, beta: float, splits: int = 0) -> torch.Tensor:
        """O = exp(-beta (1 - Qn Kn^T)) @ V over ALL keys (local keys, then the cross-rank merge):
        fp32 [Nq, C (+1 if a row-sum column was requested)]."""
        n_cols = self.n_classes + (1 if self.rowsum_col is not None else 0)
        nq = qn.shape[0]
        if self.n_keys > 0:
            if self.hard_bank is not None:
                if splits <= 0:
                    splits = ops.attn_hard_splits(nq, self.hard_bank.n_sorted, self.device)
                part = ops.attn_fwd_hard(qn, self.hard_bank, beta, splits=splits)
            else:
                if splits <= 0:
                    splits = ops.attn_splits(nq, self.n_keys, self.vt.shape[0], self.device)
                part = ops.attn_fwd(qn, self.k_norm, self.vt, self.n_keys, n_cols, beta, splits=splits, merge=True)
            self.gpu_launches += 1 + int(splits > 1)
        else:
            part = torch.zeros((nq, n_cols), dtype=torch.float32, device=self.device)
        if self.world == 1:
            return part
        import torch.distributed as dist
        gathered = torch.empty((self.world, nq, n_cols), dtype=torch.float32, device=self.device)
        dist.all_gather_into_tensor(gathered, part.contiguous(), group=self.group)
        out = ops.merge_partials(gathered)
        self.gpu_launches += 1
        return out

    def search(self, test_image_features: torch.Tensor, betas: tp.Sequence[float], alphas: tp.Sequence[float],
               labels: tp.Optional[torch.Tensor] = None, feature_major: bool = True, want_logits: bool = False,
               want_pred: bool = True) -> tp.List[tp.Dict[str, tp.Any]]:
        """One pass of the hot path for a query bank: for every beta one fused attention launch, then one
        epilogue launch covering every alpha.  Returns one dict per beta with device tensors
        pred [na, Nq], top1/top5 [na] (if labels), logits [na, Nq, C] (if requested), cache_logits."""
        qn, z = self.prepare_queries(test_image_features, feature_major)
        if labels is not None:
            labels = labels.to(self.device, non_blocking=True)
        results = []
        for beta in betas:
            o = self.cache_logits(qn, float(beta))
            rowsum = None
            if self.rowsum_col is not None:
                rowsum = o[:, self.rowsum_col].contiguous()
            res = ops.epilogue(z, o[:, : self.n_classes] if self.rowsum_col is not None else o, alphas,
                               labels=labels, rowsum=rowsum, want_logits=want_logits, want_pred=want_pred)
            self.gpu_launches += 1
            res["beta"] = float(beta)
            res["cache_logits"] = o
            res["clip_logits"] = z
            results.append(res)
        return results
