"""B200 mirror of summer_clip/clip_searcher/cache_value_strategy.py.

`transform(cache_outs)` returns a `CacheValues` handle (the values already in the TRANSPOSED, padded
tensor-core layout the attention kernel's GEMM-2 consumes) instead of a dense [Nk, C] tensor;
`FusedWeights @ CacheValues` is the fused kernel.  `.dense()` gives the reference-shaped tensor.
"""
from __future__ import annotations

import typing as tp
import weakref
from abc import ABC, abstractmethod

import torch

from .. import ops


class CacheValues:
    """Cache values in kernel layout: the dense transposed matrix `vt` and/or — for one-hot values — the int16
    label vector `labels16` the hard-label kernel synthesises its GEMM-2 operand from (2 bytes per key instead
    of 2*C_pad).  Whichever of the two is missing is built on first use."""

    def __init__(self, vt: tp.Optional[torch.Tensor], n_keys: int, n_classes: int,
                 labels16: tp.Optional[torch.Tensor] = None) -> None:
        assert vt is not None or labels16 is not None
        self._vt, self.n_keys, self.n_classes, self.labels16 = vt, int(n_keys), int(n_classes), labels16
        self._bank: tp.Optional[ops.HardBank] = None
        self._bank_key: tp.Optional[tuple] = None

    @property
    def shape(self) -> tp.Tuple[int, int]:
        return (self.n_keys, self.n_classes)

    @property
    def is_hard(self) -> bool:
        return self.labels16 is not None

    def hard_bank(self, k_norm: torch.Tensor) -> "ops.HardBank":
        """The label-sorted copy of the normalised key bank `k_norm` for these one-hot values (built once per
        (values, bank) pair; the reference loop reuses both across betas, image_attention.py:106-109)."""
        # identity of the bank = the tensor OBJECT (weak reference) + its version; an address is not an identity
        # (the allocator hands a freed bank's address to the next tensor of the same shape)
        same = (self._bank is not None and self._bank_key is not None and self._bank_key[0]() is k_norm
                and self._bank_key[1] == k_norm._version)
        if not same:
            layout = self._bank if self._bank is not None else ops.hard_bank_layout(self.labels16[: self.n_keys], self.n_classes)
            self._bank, self._bank_key = layout.gather(k_norm), (weakref.ref(k_norm), k_norm._version)
        return self._bank

    def vt(self, op_dtype: torch.dtype) -> torch.Tensor:
        if self._vt is None:        # one-hot values from the labels (-1 pads select no class)
            self._vt = ops.values_prepare(None, self.n_classes, labels=self.labels16[: self.n_keys].to(torch.int32),
                                          op_dtype=op_dtype)
        if self._vt.dtype != op_dtype:
            self._vt = self._vt.to(op_dtype)
        return self._vt

    def dense(self) -> torch.Tensor:
        """[Nk, C] float tensor, what the reference's strategy would have returned."""
        return self.vt(ops.OP_DTYPE)[: self.n_classes, : self.n_keys].t().float()

    @staticmethod
    def from_dense(values: torch.Tensor, op_dtype: tp.Optional[torch.dtype] = None) -> "CacheValues":
        """Arbitrary dense [Nk, C] values (e.g. Tip-Adapter's one-hot fp16 cache_values): transpose + cast +
        pad with the cast-only mode of the normalise kernel."""
        n_keys, n_classes = values.shape
        if ops.hard_supported(n_classes) and n_keys > 0:
            # one-hot rows (Tip-Adapter cache_values, tip_adapter/utils.py:62) -> labels for the hard-label kernel
            rows = values if values.stride(1) == 1 else values.contiguous()
            conf, label = ops.rowconf(rows)
            if bool(((conf == 1) & (rows.float().abs().sum(1) == 1)).all()):
                return CacheValues(None, n_keys, n_classes, labels16=ops.hard_labels(None, n_classes, labels=label))
        op_dtype = ops._op(op_dtype)                                      # dense values: 16-bit operands only
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
        n_keys, n_classes = cache_outs.shape[0] if idx is None else idx.numel(), cache_outs.shape[1]
        if ops.hard_supported(n_classes):
            return CacheValues(None, n_keys, n_classes, labels16=ops.hard_labels(cache_outs, n_classes, idx=idx))
        return CacheValues(ops.values_prepare(cache_outs, n_classes, idx=idx), n_keys, n_classes)


class SoftmaxCacheStrategy(CacheValueStrategy):
    """cache_value_strategy.py:20-28 — softmax(clip_scale * scale * cache_outs, dim=1)."""

    def __init__(self, clip_scale: float, scale: float) -> None:
        self.clip_scale, self.scale = clip_scale, scale

    def transform(self, cache_outs: torch.Tensor, idx: tp.Optional[torch.Tensor] = None) -> CacheValues:
        vt = ops.values_prepare(cache_outs, cache_outs.shape[1], idx=idx, softmax_scale=self.clip_scale * self.scale)
        return CacheValues(vt, cache_outs.shape[0] if idx is None else idx.numel(), cache_outs.shape[1])


class GoldCacheValues(CacheValueStrategy):
    """one_hot(gold labels): cache.replace_outs_with_golds (image_attention.py:65-66) and the Tip-Adapter
    cache values (tip_adapter/utils.py:62)."""

    def __init__(self, n_classes: int) -> None:
        self.n_classes = n_classes

    def transform(self, labels: torch.Tensor) -> CacheValues:
        if ops.hard_supported(self.n_classes):
            return CacheValues(None, labels.numel(), self.n_classes,
                               labels16=ops.hard_labels(None, self.n_classes, labels=labels))
        return CacheValues(ops.values_prepare(None, self.n_classes, labels=labels), labels.numel(), self.n_classes)
