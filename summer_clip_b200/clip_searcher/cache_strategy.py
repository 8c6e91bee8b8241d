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


class CacheStrategy(ABC):
    @abstractmethod
    def transform(self, image_features: torch.Tensor, image_outs: torch.Tensor) \
            -> tp.Tuple[torch.Tensor, torch.Tensor]:
        pass


class IndexedCacheStrategy(CacheStrategy):
    @abstractmethod
    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        pass

    def transform(self, image_features: torch.Tensor, image_outs: torch.Tensor) \
            -> tp.Tuple[torch.Tensor, torch.Tensor]:
        samples_inds = self.select(image_features, image_outs)
        return image_features[:, samples_inds], image_outs[samples_inds]


class AllLogitsStrategy(IndexedCacheStrategy):
    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        return torch.arange(image_outs.shape[0], device=image_outs.device)


class ThresholdStrategy(IndexedCacheStrategy):
    def __init__(self, threshold: float, use_softmax: bool = True) -> None:
        super().__init__()
        self.threshold = threshold
        self.use_softmax = use_softmax

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        max_probs, _ = _rowconf(image_outs, scale=1.0, prob=self.use_softmax)
        confidence_mask = (max_probs >= self.threshold)
        return confidence_mask.nonzero().squeeze(1)


def select_topk_per_label(image_labels: torch.Tensor, image_logits: torch.Tensor, topk: int,
                          n_classes: tp.Optional[int] = None) -> torch.Tensor:
    """cache_strategy.py:48-59 on the per-class top-k kernel.  n_classes defaults to max(label) + 1."""
    if n_classes is None:
        n_classes = int(image_labels.max().item()) + 1 if image_labels.numel() else 1
    return ops.select_topk_per_label(image_logits, image_labels, n_classes, topk)


class TopKStrategy(IndexedCacheStrategy):
    def __init__(self, topk: int) -> None:
        super().__init__()
        self.topk = topk

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        image_logits, label_preds = _rowconf(image_outs, prob=False)
        return select_topk_per_label(label_preds, image_logits, self.topk, image_outs.shape[1])


class TopKProbStrategy(IndexedCacheStrategy):
    def __init__(self, topk: int, scale: float) -> None:
        super().__init__()
        self.scale = scale
        self.topk = topk
        self.topk_strategy = TopKStrategy(topk)

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        # softmax(image_outs * scale) never materialised: its row max is 1 / sum_c exp(scale (l_c - l_max))
        image_probs, label_preds = _rowconf(image_outs, scale=self.scale, prob=True)
        return select_topk_per_label(label_preds, image_probs, self.topk, image_outs.shape[1])


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
    def __init__(self, topk: int) -> None:
        super().__init__()
        self.topk = topk

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        samples_num = self.topk * image_outs.shape[1]
        samples_num = min(samples_num, image_outs.shape[0])
        samples_ids = np.random.choice(image_outs.shape[0], size=samples_num, replace=False)
        return torch.LongTensor(samples_ids).to(image_outs.device)


def select_k_random_per_label(image_labels: torch.Tensor, k: int) -> torch.Tensor:
    """cache_strategy.py:113-124 — host numpy RNG, label order and draw order as in the reference."""
    samples_ids = []
    for label in image_labels.unique():
        label_inds = (image_labels == label).nonzero().squeeze(1)
        label_k = min(k, label_inds.shape[0])
        label_samples_inds_np = np.random.choice(label_inds.shape[0], size=label_k, replace=False)
        label_samples_inds = torch.LongTensor(label_samples_inds_np).to(label_inds.device)
        samples_ids.append(label_inds[label_samples_inds])
    return torch.cat(samples_ids)


class PerGoldClassRandomSampleStrategy(IndexedCacheStrategy):
    def __init__(self, topk: int, cache_dataset=None, cache_labels: tp.Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self.topk = topk
        self.cache_labels = cache_labels if cache_labels is not None else load_labels(cache_dataset)

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        return select_k_random_per_label(self.cache_labels, self.topk).to(image_outs.device)


class PerPredClassRandomSampleStrategy(IndexedCacheStrategy):
    def __init__(self, topk: int) -> None:
        super().__init__()
        self.topk = topk

    def select(self, image_features: torch.Tensor, image_outs: torch.Tensor) -> torch.Tensor:
        _, label_preds = _rowconf(image_outs, prob=False)
        return select_k_random_per_label(label_preds, self.topk)
