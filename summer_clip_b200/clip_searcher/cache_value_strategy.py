"""B200 mirror of summer_clip/clip_searcher/cache_value_strategy.py.

`transform(cache_outs)` returns a `CacheValues` handle (the values already in the TRANSPOSED, padded
tensor-core layout the attention kernel's GEMM-2 consumes) instead of a dense [Nk, C] tensor;
`FusedWeights @ CacheValues` is the fused kernel.  `.dense()` gives the reference-shaped tensor.
"""
from __future__ import annotations

import typing as tp
from abc import ABC, abstractmethod

import torch

from .. import ops


class CacheValues:
    def __init__(self, vt: torch.Tensor, n_keys: int, n_classes: int) -> None:
        self._vt, self.n_keys, self.n_classes = vt, int(n_keys), int(n_classes)

    @property
    def shape(self) -> tp.Tuple[int, int]:
        return (self.n_keys, self.n_classes)

    def vt(self, op_dtype: torch.dtype) -> torch.Tensor:
        if self._vt.dtype != op_dtype:
            self._vt = self._vt.to(op_dtype)
        return self._vt

    def dense(self) -> torch.Tensor:
        """[Nk, C] float tensor, what the reference's strategy would have returned."""
        return self._vt[: self.n_classes, : self.n_keys].t().float()

    @staticmethod
    def from_dense(values: torch.Tensor, op_dtype: tp.Optional[torch.dtype] = None) -> "CacheValues":
        """Arbitrary dense [Nk, C] values (e.g. Tip-Adapter's one-hot fp16 cache_values): transpose + cast +
        pad with the cast-only mode of the normalise kernel."""
        n_keys, n_classes = values.shape
        op_dtype = ops._op(op_dtype)
        c_pad, nk_pad = ops.pad_classes(n_classes), ops.pad_dim(n_keys)   # pad_dim: multiple of 64 (and of 8)
        vt = torch.zeros((c_pad, nk_pad), dtype=op_dtype, device=values.device)
        ops.normalize_cast(values, feature_major=True, normalize=False, out=vt)
        return CacheValues(vt, n_keys, n_classes)


class CacheValueStrategy(ABC):
    @abstractmethod
    def transform(self, cache_outs: torch.Tensor) -> CacheValues:
        pass


class HardCacheStrategy(CacheValueStrategy):
    """cache_value_strategy.py:13-17 — one_hot(argmax_c cache_outs)."""

    def transform(self, cache_outs: torch.Tensor, idx: tp.Optional[torch.Tensor] = None) -> CacheValues:
        vt = ops.values_prepare(cache_outs, cache_outs.shape[1], idx=idx)
        return CacheValues(vt, cache_outs.shape[0] if idx is None else idx.numel(), cache_outs.shape[1])


class SoftmaxCacheStrategy(CacheValueStrategy):
    """cache_value_strategy.py:20-28 — softmax(clip_scale * scale * cache_outs, dim=1)."""

    def __init__(self, clip_scale: float, scale: float) -> None:
        super().__init__()
        self.clip_scale = clip_scale
        self.scale = scale

    def transform(self, cache_outs: torch.Tensor, idx: tp.Optional[torch.Tensor] = None) -> CacheValues:
        vt = ops.values_prepare(cache_outs, cache_outs.shape[1], idx=idx, softmax_scale=self.clip_scale * self.scale)
        return CacheValues(vt, cache_outs.shape[0] if idx is None else idx.numel(), cache_outs.shape[1])


class GoldCacheValues(CacheValueStrategy):
    """one_hot(gold labels): cache.replace_outs_with_golds (image_attention.py:65-66) and the Tip-Adapter
    cache values (tip_adapter/utils.py:62)."""

    def __init__(self, n_classes: int) -> None:
        self.n_classes = n_classes

    def transform(self, labels: torch.Tensor) -> CacheValues:
        vt = ops.values_prepare(None, self.n_classes, labels=labels)
        return CacheValues(vt, labels.numel(), self.n_classes)
