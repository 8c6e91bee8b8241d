"""B200 mirror of summer_clip/clip_searcher/cache_strategy.py — same class names and signatures.

The confidence-ranked strategies (TopK*, Threshold) run on the CUDA kernels: one pass over the logits
bank for (confidence, label) and a histogram/scatter/radix-select per-class top-k that replaces the
reference's Python loop over classes (cache_strategy.py:51-57; O(N*C) compares, ~5 launches and a
host sync per class).  Ties are ordered (confidence desc, row index asc) where torch.topk is
unspecified.  The random strategies keep the reference's host-side numpy RNG stream.
"""
from __future__ import annotations

import typing as tp
from abc import ABC, abstractmethod

import numpy as np
import torch

from .. import ops
from .utils import load_labels


class LazyLogitsBank:
    """Stand-in for the `[N, C]` zero-shot logits bank `image_outs = normalise(K)^T @ T` (save_image_outs.py:25) that
    is never materialised: the confidence-ranked strategies only need (confidence, label) per
    row, which the fused GEMM + row-scan kernel produces straight from the features (`ops.rowconf_from_features`);
    rows of selected keys (cache-quality statistics, soft cache values) are computed on demand for those keys only.
    Pass it to `select` / `build_cache` wherever the reference passes `image_outs`."""

    def __init__(self, image_features: torch.Tensor, text_features: torch.Tensor) -> None:
        self.image_features, self.text_features = image_features, text_features      # [D, N], [D, C]
        self._t_split = ops.text_split(text_features.float().contiguous())
        self._conf: tp.Dict[tuple, tp.Tuple[torch.Tensor, torch.Tensor]] = {}

    @property
    def shape(self) -> tp.Tuple[int, int]:
        return (self.image_features.shape[1], self.text_features.shape[1])

    @property
    def device(self) -> torch.device:
        return self.image_features.device

    def rowconf(self, scale: float = 1.0, prob: bool = False) -> tp.Tuple[torch.Tensor, torch.Tensor]:
        key = (float(scale), bool(prob))
        if key not in self._conf:
            self._conf[key] = ops.rowconf_from_features(self.image_features, True, self.text_features, prob=prob,
                                                        prob_scale=scale, t_split=self._t_split)
        return self._conf[key]

    def __getitem__(self, idx: torch.Tensor) -> torch.Tensor:
        """L[idx]: fp32 [n_sel, C] logits of the selected keys only."""
        return ops.zero_shot_logits(self.image_features[:, idx], True, self.text_features, scale=1.0, normalize=True,
                                    t_split=self._t_split)

    def dense(self, chunk: int = 1 << 16) -> torch.Tensor:
        """The whole bank after all (strategies that index it by gold label need it): chunked tensor-core GEMM."""
        n = self.shape[0]
        return torch.cat([ops.zero_shot_logits(self.image_features[:, s:s + chunk], True, self.text_features, scale=1.0,
                                               normalize=True, t_split=self._t_split) for s in range(0, n, chunk)])


def _rowconf(image_outs, scale: float = 1.0, prob: bool = False):
    if isinstance(image_outs, LazyLogitsBank):
        return image_outs.rowconf(scale=scale, prob=prob)
    return ops.rowconf(image_outs, scale=scale, prob=prob)


Banks = tp.Tuple[torch.Tensor, torch.Tensor]      # (image features [D, n], logits [n, C])


class CacheStrategy(ABC):
    """What `cache_strategies.<name>` instantiates (the reference's interface, cache_strategy.py:10-15): the train bank
    in, the cache bank out."""

    @abstractmethod
    def transform(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> Banks:
        raise NotImplementedError


class IndexedCacheStrategy(CacheStrategy):
    """A strategy that is a row selection (cache_strategy.py:18-27): `select` names the rows, `transform` gathers
    the feature columns and the logits rows (the sweep driver calls `select` itself and fuses the gather into the
    normalise kernel; `transform` is for callers that want the reference's tensors)."""

    @abstractmethod
    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def transform(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> Banks:
        rows = self.select(image_features, image_outs)
        return image_features[:, rows], image_outs[rows]


class AllLogitsStrategy(IndexedCacheStrategy):
    """Every train image is a cache key (cache_strategy.py:30-32)."""

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        return torch.arange(image_outs.shape[0], device=image_outs.device)


class ThresholdStrategy(IndexedCacheStrategy):
    """Rows whose confidence (max softmax probability, or max logit) reaches `threshold` (cache_strategy.py:35-45)."""

    def __init__(self, threshold: float, use_softmax: bool = True) -> None:
        self.threshold, self.use_softmax = threshold, use_softmax

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        confidence, _ = _rowconf(image_outs, scale=1.0, prob=self.use_softmax)
        return torch.nonzero(confidence >= self.threshold).flatten()


def select_topk_per_label(image_labels: torch.Tensor, image_logits: torch.Tensor, topk: int,
                          n_classes: tp.Optional[int] = None) -> torch.Tensor:
    """cache_strategy.py:48-59 on the per-class top-k kernel.  n_classes defaults to max(label) + 1."""
    if n_classes is None:
        n_classes = int(image_labels.max().item()) + 1 if image_labels.numel() else 1
    return ops.select_topk_per_label(image_logits, image_labels, n_classes, topk)


class TopKStrategy(IndexedCacheStrategy):
    """Per predicted class, the `topk` rows with the largest max logit (cache_strategy.py:62-70)."""

    def __init__(self, topk: int) -> None:
        self.topk = topk

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        confidence, predicted = _rowconf(image_outs, prob=False)
        return select_topk_per_label(predicted, confidence, self.topk, image_outs.shape[1])


class TopKProbStrategy(IndexedCacheStrategy):
    """The same ranking on max softmax(scale * logits) (cache_strategy.py:73-81).  The [N, C] probabilities are
    never materialised: the row maximum of a softmax is 1 / sum_c exp(scale (l_c - l_max))."""

    def __init__(self, topk: int, scale: float) -> None:
        self.topk, self.scale = topk, scale

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        confidence, predicted = _rowconf(image_outs, scale=self.scale, prob=True)
        return select_topk_per_label(predicted, confidence, self.topk, image_outs.shape[1])


class TopKPerGoldStrategy(IndexedCacheStrategy):
    def __init__(self, topk: int, cache_dataset=None, cache_labels: tp.Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self.topk = topk
        self.cache_labels = cache_labels if cache_labels is not None else load_labels(cache_dataset)

    def _gold_logits(self, image_outs: torch.Tensor) -> tp.Tuple[torch.Tensor, torch.Tensor]:
        if isinstance(image_outs, LazyLogitsBank):
            image_outs = image_outs.dense()
        cache_labels = self.cache_labels.to(image_outs.device)
        labels_indexes = cache_labels.long().unsqueeze(dim=0).t()
        return cache_labels, image_outs.gather(1, labels_indexes).squeeze(dim=1).float()

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        cache_labels, image_logits = self._gold_logits(image_outs)
        return select_topk_per_label(cache_labels, image_logits, self.topk, image_outs.shape[1])


class TopKPerGoldProbStrategy(IndexedCacheStrategy):
    def __init__(self, topk: int, cache_dataset=None, scale: float = 1.0,
                 cache_labels: tp.Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self.scale = scale
        self.topk_strategy = TopKPerGoldStrategy(topk, cache_dataset, cache_labels)

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        # softmax(scale * l)[gold] = exp(scale * l_gold - scale * l_max) * max-prob, from two row scans
        if isinstance(image_outs, LazyLogitsBank):
            image_outs = image_outs.dense()
        max_prob, _ = ops.rowconf(image_outs, scale=self.scale, prob=True)
        max_raw, _ = ops.rowconf(image_outs, prob=False)
        cache_labels, gold_raw = self.topk_strategy._gold_logits(image_outs)
        gold_prob = torch.exp(gold_raw * self.scale - max_raw * self.scale) * max_prob
        return select_topk_per_label(cache_labels, gold_prob, self.topk_strategy.topk, image_outs.shape[1])


class GlobalRandomSampleStrategy(IndexedCacheStrategy):
    """`topk` x C rows drawn from the whole bank without replacement (cache_strategy.py:103-110).  The draw is the HOST
    numpy generator's, as in the reference, so that a seeded run (`meta.random_state`) picks the same cache."""

    def __init__(self, topk: int) -> None:
        self.topk = topk

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        n_rows, n_classes = image_outs.shape[0], image_outs.shape[1]
        drawn = np.random.choice(n_rows, size=min(self.topk * n_classes, n_rows), replace=False)
        return torch.from_numpy(drawn).to(device=image_outs.device, dtype=torch.int64)


def select_k_random_per_label(image_labels: torch.Tensor, k: int) -> torch.Tensor:
    """Up to k random rows of every label (cache_strategy.py:113-124).  Reproduces the reference's stream: labels in
    ascending order (`unique`), one `np.random.choice(n_c, min(k, n_c), replace=False)` per label, drawn positions
    mapped back through the label's rows in ascending order.  The members of all labels come from ONE stable sort of
    the label vector instead of one N-long comparison per label."""
    labels = image_labels.detach().to("cpu", torch.int64)
    order = torch.argsort(labels, stable=True)                    # rows grouped by label, ascending inside a group
    _, counts = torch.unique_consecutive(labels[order], return_counts=True)
    picked, start = [], 0
    for n_c in counts.tolist():
        members = order[start:start + n_c]
        positions = np.random.choice(n_c, size=min(k, n_c), replace=False)
        picked.append(members[torch.from_numpy(positions)])
        start += n_c
    if not picked:
        return torch.zeros(0, dtype=torch.int64, device=image_labels.device)
    return torch.cat(picked).to(image_labels.device)


class PerGoldClassRandomSampleStrategy(IndexedCacheStrategy):
    """`topk` random rows per GOLD class (cache_strategy.py:127-138); labels come from `cache_labels`
    (cache.labels_path) or, like the reference, from a `cache_dataset` of (image, label) pairs."""

    def __init__(self, topk: int, cache_dataset=None, cache_labels: tp.Optional[torch.Tensor] = None) -> None:
        self.topk = topk
        self.cache_labels = load_labels(cache_dataset) if cache_labels is None else cache_labels

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        return select_k_random_per_label(self.cache_labels, self.topk).to(image_outs.device)


class PerPredClassRandomSampleStrategy(IndexedCacheStrategy):
    """`topk` random rows per PREDICTED class (cache_strategy.py:141-149)."""

    def __init__(self, topk: int) -> None:
        self.topk = topk

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        _, predicted = _rowconf(image_outs, prob=False)
        return select_k_random_per_label(predicted, self.topk)
