"""B200 mirror of summer_clip/clip_searcher/cache_weights_strategy.py.

Same class names, constructor arguments and `transform` signature.  `transform` does not
materialise the [Nq, Nk] weight matrix: it returns a lazy `FusedWeights` operand whose `@` with the
cache values launches the fused tcgen05 attention kernel, so the reference's own loop
(`cache_weights @ cache_value_strategy.transform(cache_image_outs)`, image_attention.py:107-109)
runs unmodified on top of it.
"""
from __future__ import annotations

import typing as tp
import weakref
from abc import ABC, abstractmethod

import torch

from .. import ops
from .cache_value_strategy import CacheValues


class _BankCache:
    """Normalised operand banks keyed by the identity of the source tensor: the reference loop calls
    `transform` once per beta with the same banks (image_attention.py:106-107) and would otherwise
    re-normalise them 8 times per cache."""

    def __init__(self, max_items: int = 4) -> None:
        # (key, weak reference to the SOURCE tensor, normalised bank): the key alone is not an identity — a freed
        # bank's address is routinely handed to the next tensor of the same shape by the caching allocator
        self._items: tp.List[tp.Tuple[tuple, tp.Any, torch.Tensor]] = []
        self._max = max_items

    @staticmethod
    def _key(t: torch.Tensor, feature_major: bool, dtype: torch.dtype) -> tuple:
        return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, t._version, feature_major, dtype, t.device)

    def get(self, t: torch.Tensor, feature_major: bool, dtype: torch.dtype) -> torch.Tensor:
        key = self._key(t, feature_major, dtype)
        self._items = [it for it in self._items if it[1]() is not None]          # drop banks whose source died
        for k, ref, v in self._items:
            if k == key and ref() is t:
                return v
        out = ops.normalize_cast(t, feature_major=feature_major, op_dtype=dtype)
        self._items.append((key, weakref.ref(t), out))
        if len(self._items) > self._max:
            self._items.pop(0)
        return out

    def clear(self) -> None:
        self._items.clear()


_BANKS = _BankCache()


class FusedWeights:
    """Lazy exp(-beta * (1 - Qn Kn^T)) of shape [Nq, Nk]."""

    def __init__(self, q_norm: torch.Tensor, k_norm: torch.Tensor, n_keys: int, beta: float) -> None:
        self.q_norm, self.k_norm, self.n_keys, self.beta = q_norm, k_norm, int(n_keys), float(beta)

    @property
    def shape(self) -> tp.Tuple[int, int]:
        return (self.q_norm.shape[0], self.n_keys)

    @property
    def device(self) -> torch.device:
        return self.q_norm.device

    def __matmul__(self, values: tp.Union["CacheValues", torch.Tensor]) -> torch.Tensor:
        if isinstance(values, torch.Tensor):
            values = CacheValues.from_dense(values, op_dtype=self.q_norm.dtype)
        if values.n_keys != self.n_keys:
            raise ValueError(f"weights have {self.n_keys} keys but values have {values.n_keys}")
        if values.is_hard:     # one-hot values: GEMM-2 operand synthesised on chip from the labels
            return ops.attn_fwd_hard(self.q_norm, values.hard_bank(self.k_norm), self.beta)
        vt = values.vt(self.q_norm.dtype)
        return ops.attn_fwd(self.q_norm, self.k_norm, vt, self.n_keys, values.n_classes, self.beta)

    @staticmethod
    def matmul_many(weights: tp.Sequence["FusedWeights"], values: tp.Union["CacheValues", torch.Tensor]) -> tp.List[torch.Tensor]:
        """[w @ values for w in weights] for weights that differ only in beta (the sweep of image_attention.py:
        106-109, of search_hp): with one-hot values the tensor-core pass is shared by groups of 4 betas."""
        if not weights:
            return []
        if isinstance(values, torch.Tensor):
            values = CacheValues.from_dense(values, op_dtype=weights[0].q_norm.dtype)
        w0 = weights[0]
        same_banks = all(w.q_norm is w0.q_norm and w.k_norm is w0.k_norm and w.n_keys == w0.n_keys for w in weights)
        if same_banks and values.is_hard and values.n_keys == w0.n_keys:
            return ops.attn_fwd_hard_multi(w0.q_norm, values.hard_bank(w0.k_norm), [w.beta for w in weights])
        return [w @ values for w in weights]

    def materialize(self, chunk: int = 4096) -> torch.Tensor:
        """Dense fp32 [Nq, Nk] (tests / debugging on small caches only): the kernel with V = I."""
        cols = []
        for s in range(0, self.n_keys, chunk):
            n = min(chunk, self.n_keys - s)
            lab = torch.arange(n, device=self.device, dtype=torch.int32)
            vt = ops.values_prepare(None, n, labels=lab, op_dtype=self.q_norm.dtype)
            cols.append(ops.attn_fwd(self.q_norm, self.k_norm[s:s + n], vt, n, n, self.beta))
        return torch.cat(cols, dim=1)


class CacheWeightsStrategy(ABC):
    @abstractmethod
    def transform(self, test_image_features: torch.Tensor, cache_image_features: torch.Tensor):
        """
        test_image_features: not normalized image features of the test images, [D, Nq]
        cache_image_features: not normalized image features of the selected cache images, [D, Nk]
        """


class CacheWeightsNormStrategy(CacheWeightsStrategy):
    """cache_weights_strategy.py:17-25 — column-normalise both banks, then `transform_norm`.  Here the
    normalisation is the fused normalise + transpose + cast kernel and its result is cached per bank."""

    def transform(self, test_image_features, cache_image_features):
        q = test_image_features if isinstance(test_image_features, NormalizedBank) else \
            NormalizedBank(_BANKS.get(test_image_features, True, ops.OP_DTYPE))
        k = cache_image_features if isinstance(cache_image_features, NormalizedBank) else \
            NormalizedBank(_BANKS.get(cache_image_features, True, ops.OP_DTYPE))
        return self.transform_norm(q, k)

    @abstractmethod
    def transform_norm(self, test_image_features: "NormalizedBank", cache_image_features: "NormalizedBank"):
        pass


class NormalizedBank:
    """A bank already in kernel layout: [N, D_pad] fp16/bf16 rows, L2-normalised."""

    def __init__(self, rows: torch.Tensor, n: tp.Optional[int] = None) -> None:
        self.rows = rows
        self.n = rows.shape[0] if n is None else int(n)


class TipAdapterWeightsStrategy(CacheWeightsNormStrategy):
    """cache_weights_strategy.py:28-36 — W = exp(-1 * beta * (1 - Q^T K))."""

    def __init__(self, beta: float) -> None:
        super().__init__()
        self.beta = beta

    def transform_norm(self, test_image_features: NormalizedBank, cache_image_features: NormalizedBank) -> FusedWeights:
        return FusedWeights(test_image_features.rows, cache_image_features.rows, cache_image_features.n, self.beta)
