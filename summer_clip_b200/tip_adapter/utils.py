"""B200 mirror of the hot-path part of summer_clip/tip_adapter/utils.py: the training-free
Tip-Adapter head (tip_adapter.py:58-68) and the (beta, alpha) grid search `search_hp` (:99-129).

The reference recomputes Q@K, exp, @V and the zero-shot GEMM for each of the 200 x 20 grid points
(13.4 PFLOP executed at ImageNet scale).  Here the operands are cast once, each beta is one fused
attention launch and its 20 alphas are one epilogue launch with on-device accuracy counters; the
search order and the strict `>` (first best wins) are the reference's.
The CLIP encoder forward passes of build_cache_model / pre_load_features are outside this path; what those
functions do AFTER the encoder (mean over augment epochs, row normalisation, permute, one-hot, the
`keys_*.pt` / `values_*.pt` / `*_f.pt` / `*_l.pt` cache files and their load_cache / load_pre_feat switches) is
here, on encoder outputs given as tensors.
"""
from __future__ import annotations

import typing as tp

import torch

from .. import ops
from ..clip_searcher.cache_value_strategy import CacheValues


def build_cache_model(cfg, train_features: tp.Optional[torch.Tensor] = None,
                      train_labels: tp.Optional[torch.Tensor] = None) -> tp.Tuple[torch.Tensor, torch.Tensor]:
    """tip_adapter/utils.py:38-71 from the encoder outputs on: `train_features` [augment_epoch, Nk, D] (or [Nk, D])
    are the image features of the few-shot training set per augment epoch, `train_labels` [Nk] their targets.
    cfg['load_cache'] False: cache_keys = the [D, Nk] permuted view of the row-normalised epoch mean (one kernel,
    sc_mean_normalize_rows), cache_values = one_hot(labels).half(); both are saved as
    cache_dir/keys_<shots>shots.pt and values_<shots>shots.pt exactly like the reference.  True: load those files."""
    keys_path = cfg['cache_dir'] + '/keys_' + str(cfg['shots']) + "shots.pt"
    values_path = cfg['cache_dir'] + '/values_' + str(cfg['shots']) + "shots.pt"
    if cfg['load_cache'] == False:  # noqa: E712  (the reference's own test)
        if train_features is None or train_labels is None:
            raise ops._lib.SummerClipError("build_cache_model: load_cache is False and no encoder outputs were given "
                                           "(the CLIP image tower is outside this path)")
        cache_keys = ops.mean_normalize_rows(train_features).permute(1, 0)
        # the one-hot matrix exists for the cache FILE only (the format the reference and its notebooks read); the
        # attention path consumes the labels themselves (CacheValues.from_dense recognises one-hot rows)
        cache_values = torch.nn.functional.one_hot(train_labels.long()).half()
        torch.save(cache_keys, keys_path)
        torch.save(cache_values, values_path)
    else:
        cache_keys = torch.load(keys_path)
        cache_values = torch.load(values_path)
    return cache_keys, cache_values


def pre_load_features(cfg, split: str, features: tp.Optional[torch.Tensor] = None,
                      labels: tp.Optional[torch.Tensor] = None) -> tp.Tuple[torch.Tensor, torch.Tensor]:
    """tip_adapter/utils.py:74-96 from the encoder outputs on: `features` [Nq, D] un-normalised image features of
    the split, row-normalised here (:84) in their own dtype and saved as cache_dir/<split>_f.pt, <split>_l.pt."""
    f_path, l_path = cfg['cache_dir'] + "/" + split + "_f.pt", cfg['cache_dir'] + "/" + split + "_l.pt"
    if cfg['load_pre_feat'] == False:  # noqa: E712
        if features is None or labels is None:
            raise ops._lib.SummerClipError("pre_load_features: load_pre_feat is False and no encoder outputs were given")
        features = ops.mean_normalize_rows(features)
        torch.save(features, f_path)
        torch.save(labels, l_path)
    else:
        features = torch.load(f_path)
        labels = torch.load(l_path)
    return features, labels


def cls_acc(output: torch.Tensor, target: torch.Tensor, topk: int = 1) -> float:
    """tip_adapter/utils.py:10-15 — top-k accuracy in percent (k in {1, 5})."""
    res = ops.epilogue(None, output.float().contiguous(), [1.0], labels=target, want_pred=False)
    if topk == 1:
        correct = int(res["top1"][0])
    elif topk == 5:
        correct = int(res["top5"][0])
    else:
        raise NotImplementedError("cls_acc supports topk in {1, 5}")
    return 100 * correct / target.shape[0]


class TipAdapterHead:
    """Operands of the Tip-Adapter head in kernel layout.  features [Nq, D] (row-normalised by the caller,
    tip_adapter/utils.py:84), cache_keys [D, Nk] (any strides — the reference passes a permuted view,
    utils.py:61), cache_values [Nk, C] one-hot, clip_weights [D, C]."""

    def __init__(self, cache_keys: torch.Tensor, cache_values: torch.Tensor, features: torch.Tensor,
                 clip_weights: torch.Tensor, adapter: tp.Optional[torch.nn.Module] = None) -> None:
        if adapter is not None:            # Tip-Adapter-F: affinity = adapter(features), weight [Nk, D]
            self.k = ops.normalize_cast(adapter.weight.detach(), feature_major=False, normalize=False)
        else:
            self.k = ops.normalize_cast(cache_keys, feature_major=True, normalize=False)
        self.n_keys = self.k.shape[0]
        self.q = ops.normalize_cast(features, feature_major=False, normalize=False)
        self.values = cache_values if isinstance(cache_values, CacheValues) else CacheValues.from_dense(cache_values)
        self.n_classes = self.values.n_classes
        self.clip_logits = ops.zero_shot_logits(features, False, clip_weights, scale=100.0, normalize=False)

    def cache_logits(self, beta: float) -> torch.Tensor:
        """exp(-(beta - beta * features @ cache_keys)) @ cache_values."""
        if self.values.is_hard:
            return ops.attn_fwd_hard(self.q, self.values.hard_bank(self.k), beta)
        return ops.attn_fwd(self.q, self.k, self.values.vt(self.q.dtype), self.n_keys, self.n_classes, beta)

    def cache_logits_many(self, betas: tp.Sequence[float]) -> tp.List[torch.Tensor]:
        """cache_logits for a list of betas; one-hot values share the tensor-core pass between groups of 4."""
        if self.values.is_hard:
            return ops.attn_fwd_hard_multi(self.q, self.values.hard_bank(self.k), betas)
        return [self.cache_logits(b) for b in betas]

    def top1_counts_many(self, betas: tp.Sequence[float], alphas: tp.Sequence[float], labels: torch.Tensor) -> torch.Tensor:
        """[len(betas), len(alphas)] top-1 counts, in chunks of 16 betas (bounded memory)."""
        rows = []
        for s in range(0, len(betas), 8):
            if self.values.is_hard:       # unmerged key-split tiles: the epilogue sums them as it reads
                outs = ops.attn_fwd_hard_multi(self.q, self.values.hard_bank(self.k), betas[s:s + 8], merge=False)
            else:
                outs = self.cache_logits_many(betas[s:s + 8])
            for o in outs:
                rows.append(ops.epilogue(self.clip_logits, o, alphas, labels=labels, want_pred=False)["top1"])
        return torch.stack(rows)

    def logits(self, beta: float, alpha: float) -> torch.Tensor:
        """tip_logits = clip_logits + cache_logits * alpha (tip_adapter.py:68)."""
        return ops.epilogue(self.clip_logits, self.cache_logits(beta), [alpha], want_logits=True, want_pred=False)["logits"][0]

    def top1_counts(self, beta: float, alphas: tp.Sequence[float], labels: torch.Tensor) -> torch.Tensor:
        return ops.epilogue(self.clip_logits, self.cache_logits(beta), alphas, labels=labels, want_pred=False)["top1"]


def search_hp(cfg, cache_keys, cache_values, features, labels, clip_weights, adapter=None, head=None):
    """tip_adapter/utils.py:99-129.  `head` (optional): a TipAdapterHead already built from the same operands."""
    if cfg['search_hp'] is not True and cfg['search_hp'] != 1:
        return 0, 0

    def grid(axis: int) -> tp.List[float]:          # i * (scale - 0.1) / steps + 0.1, i < steps (utils.py:103-104)
        scale, steps = cfg['search_scale'][axis], cfg['search_step'][axis]
        return [i * (scale - 0.1) / steps + 0.1 for i in range(steps)]

    betas, alphas = grid(0), grid(1)
    if head is None:
        head = TipAdapterHead(cache_keys, cache_values, features, clip_weights, adapter)
    # every (beta, alpha) top-1 count on the device, ONE copy to the host; the scan below only replays the reference's
    # "strictly better keeps the first" rule and its progress lines (the messages are its user-visible output)
    correct = head.top1_counts_many(betas, alphas, labels).cpu().tolist()
    n = labels.shape[0]
    best = (0.0, 0, 0)                               # (accuracy, beta, alpha)
    for beta, row in zip(betas, correct):
        for alpha, hits in zip(alphas, row):
            acc = 100 * int(hits) / n
            if acc > best[0]:
                best = (acc, beta, alpha)
                print(f"New best setting, beta: {beta:.2f}, alpha: {alpha:.2f}; accuracy: {acc:.2f}")
    print(f"\nAfter searching, the best accuarcy: {best[0]:.2f}.\n")
    return best[1], best[2]
