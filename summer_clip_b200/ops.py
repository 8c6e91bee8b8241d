"""Torch-facing wrappers over the C ABI: device memory and streams come from PyTorch, every
computation is a call into libsummerclip_b200.so.  CPU tensors are rejected (no fallback)."""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (SC_BF16, SC_CONF_PROB, SC_CONF_RAW, SC_E4M3, SC_F16, SC_F32, SC_VALUES_HARD, SC_VALUES_SOFTMAX, check)

E4M3 = torch.float8_e4m3fn
_DTYPES = {torch.float16: SC_F16, torch.bfloat16: SC_BF16, torch.float32: SC_F32, E4M3: SC_E4M3}

# Tensor-core operand type of the attention path (Qn, Kn, Vt and the on-chip weights P); fp32
# accumulation either way, same tcgen05 rate.  Every operand lies in [-1, 1], where fp16 has 3 more
# mantissa bits than bf16 (DESIGN.md "Precision").  Override with SUMMER_CLIP_B200_OP_DTYPE=bf16.
# SUMMER_CLIP_B200_OP_DTYPE=e4m3 (opt-in, reduced precision): feature banks as 8-bit floats (scaled by 256), half
# the bank bytes and twice the contraction length per tensor-core instruction; only the segmented (one-hot values)
# attention kernel reads them — dense values with e4m3 banks fail loudly.
OP_DTYPE = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "f16": torch.float16, "fp16": torch.float16,
            "float16": torch.float16, "e4m3": E4M3, "fp8": E4M3}[os.environ.get("SUMMER_CLIP_B200_OP_DTYPE", "fp16").lower()]


def _op(dtype: Optional[torch.dtype], allow_e4m3: bool = False) -> torch.dtype:
    dtype = OP_DTYPE if dtype is None else dtype
    if dtype is E4M3 and not allow_e4m3:
        raise TypeError("e4m3 feature banks are read by the one-hot (segmented) attention kernel only; dense cache "
                        "values need float16 or bfloat16 operands")
    if dtype not in (torch.float16, torch.bfloat16, E4M3):
        raise TypeError(f"operand dtype must be float16, bfloat16 or float8_e4m3fn, got {dtype}")
    return dtype


def _code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; expected float16, bfloat16 or float32") from None


def _cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.SummerClipError(f"{name} must be a CUDA tensor: the CLIP-search path has no CPU fallback")
    return t


def require_cuda_device(dev, who: str) -> torch.device:
    """The device an entry point runs on: a CUDA device that exists, or a loud error (there is no CPU fallback)."""
    device = torch.device(dev or "cuda")
    if device.type != "cuda":
        raise _lib.SummerClipError(f"{who} runs on the CUDA path only (no CPU fallback): meta.device={dev!r}")
    if not torch.cuda.is_available():
        raise _lib.SummerClipError(f"{who}: no CUDA device is available and the CLIP-search path has no CPU fallback")
    return device


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def pad_dim(D: int, op_dtype: Optional[torch.dtype] = None) -> int:
    if op_dtype is E4M3:
        return int(_lib.load().sc_pad_dim_op(D, SC_E4M3))
    return int(_lib.load().sc_pad_dim(D))


def pad_queries(Nq: int) -> int:
    return int(_lib.load().sc_pad_queries(Nq))


def _tile_padded(Qn: torch.Tensor) -> torch.Tensor:
    """Qn with at least sc_pad_queries(Nq) rows ALLOCATED behind its first row (what `normalize_cast` returns, and any
    row slice that is not at the end of such a tensor); otherwise a zero-padded copy."""
    Nq, D_pad = Qn.shape
    need = (pad_queries(Nq) * D_pad + Qn.storage_offset()) * Qn.element_size()
    if Qn.is_contiguous() and Qn.untyped_storage().nbytes() >= need:
        return Qn
    buf = torch.empty((pad_queries(Nq), D_pad), dtype=Qn.dtype, device=Qn.device)
    buf[Nq:].view(torch.uint8).zero_()
    buf[:Nq].copy_(Qn)
    return buf[:Nq]


def pad_keys(Nk: int) -> int:
    return int(_lib.load().sc_pad_keys(Nk))


def pad_classes(C: int) -> int:
    return int(_lib.load().sc_pad_classes(C))


def normalize_cast(x: torch.Tensor, feature_major: bool, idx: Optional[torch.Tensor] = None,
                   normalize: bool = True, op_dtype: Optional[torch.dtype] = None,
                   out: Optional[torch.Tensor] = None, inv_norm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [D, N] if feature_major (the reference's on-disk layout, save_features.py:36) else [N, D];
    any strides.  Returns [n_out, D_pad] of op_dtype, rows L2-normalised, optionally gathered by idx.
    `out` (optional) is a preallocated contiguous [>= n_out, D_pad] buffer of op_dtype.  `inv_norm` (optional,
    fp32 [>= n_out]) receives 1 / |column| of the source (sc_transpose_norms) whether or not `normalize` is set."""
    _cuda(x, "x")
    assert x.dim() == 2
    if feature_major:
        D, N = x.shape
        stride_d, stride_n = x.stride()
    else:
        N, D = x.shape
        stride_n, stride_d = x.stride()
    if idx is not None:
        idx = _cuda(idx, "idx").to(torch.int64).contiguous()
        n_out = idx.numel()
    else:
        n_out = N
    op_dtype = _op(op_dtype if out is None else out.dtype, allow_e4m3=True)
    D_pad = pad_dim(D, op_dtype)
    if out is None:
        # whole 256-row tiles are allocated (zero padding rows) and the first n_out rows returned: the attention
        # kernels map their query operand over whole tiles (sc_pad_queries)
        rows = pad_queries(n_out)
        buf = torch.empty((rows, D_pad), dtype=op_dtype, device=x.device)
        if rows > n_out:
            buf[n_out:].view(torch.uint8).zero_()
        out = buf[:n_out]
    else:
        assert out.is_cuda and out.is_contiguous() and out.shape[1] == D_pad and out.shape[0] >= n_out
        # the library writes through the raw pointer: tell torch, so that caches keyed on the tensor's version
        # (CacheValues.hard_bank, _BankCache) notice that a preallocated bank was refilled
        torch.autograd.graph.increment_version(out)
    with torch.cuda.device(x.device):
        if inv_norm is None:
            check(_lib.load().sc_normalize_cast(_ptr(x), _code(x), D, N, stride_d, stride_n, _ptr(idx), n_out,
                                                _ptr(out), _code(out), D_pad, int(normalize), _stream()),
                  "sc_normalize_cast")
        else:
            assert inv_norm.is_cuda and inv_norm.dtype == torch.float32 and inv_norm.is_contiguous() and inv_norm.numel() >= n_out
            check(_lib.load().sc_transpose_norms(_ptr(x), _code(x), D, N, stride_d, stride_n, _ptr(idx), n_out,
                                                 _ptr(out), _code(out), D_pad, int(normalize), _ptr(inv_norm), _stream()),
                  "sc_transpose_norms")
    return out


def mean_normalize_rows(feats: torch.Tensor) -> torch.Tensor:
    """Tip-Adapter cache keys from encoder features [E, N, D] (E augment epochs) or [N, D]: row-normalised mean over
    the epochs, [N, D] of the input dtype (tip_adapter/utils.py:59-60)."""
    _cuda(feats, "feats")
    if feats.dim() == 2:
        feats = feats.unsqueeze(0)
    assert feats.dim() == 3 and feats.stride(2) == 1
    E, N, D = feats.shape
    out = torch.empty((N, D), dtype=feats.dtype, device=feats.device)
    with torch.cuda.device(feats.device):
        check(_lib.load().sc_mean_normalize_rows(_ptr(feats), _code(feats), E, N, D, feats.stride(0), feats.stride(1),
                                                 _ptr(out), D, _stream()), "sc_mean_normalize_rows")
    return out


def rowconf(L: torch.Tensor, scale: float = 1.0, prob: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """(confidence fp32 [N], predicted label int32 [N]) of a logits bank L [N, C]."""
    _cuda(L, "L")
    assert L.dim() == 2 and L.stride(1) == 1
    N, C = L.shape
    conf = torch.empty(N, dtype=torch.float32, device=L.device)
    label = torch.empty(N, dtype=torch.int32, device=L.device)
    with torch.cuda.device(L.device):
        check(_lib.load().sc_rowconf(_ptr(L), _code(L), N, C, L.stride(0) if N > 1 else C, float(scale),
                                     SC_CONF_PROB if prob else SC_CONF_RAW, _ptr(conf), _ptr(label), _stream()),
              "sc_rowconf")
    return conf, label


def rowconf_from_features(X: torch.Tensor, feature_major: bool, T: torch.Tensor, scale: float = 1.0,
                          prob: bool = False, prob_scale: float = 1.0,
                          t_split: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """`rowconf(scale * normalise(X)^T @ T, scale=prob_scale, prob=prob)` without materialising the logits bank:
    (confidence fp32 [N], predicted label int32 [N]) straight from the image features X ([D, N] if feature_major)
    and the text classifier T [D, C] — save_image_outs.py:25 fused with TopK[Prob]Strategy's row scan."""
    _cuda(X, "X"), _cuda(T, "T")
    th, tl = t_split if t_split is not None else text_split(T)
    C = th.shape[0]
    N = X.shape[1] if feature_major else X.shape[0]
    conf = torch.empty(N, dtype=torch.float32, device=X.device)
    label = torch.empty(N, dtype=torch.int32, device=X.device)
    if N == 0:
        return conf, label
    mode = SC_CONF_PROB if prob else SC_CONF_RAW
    if X.dtype == torch.float16:
        # the raw features are exact in fp16: transposed copy + 1/norm per row, two operand passes instead of the
        # three of a split (hi, lo) pair, and half the bytes written
        inv = torch.empty(N, dtype=torch.float32, device=X.device)
        xr = normalize_cast(X, feature_major, normalize=False, op_dtype=torch.float16, inv_norm=inv)
        with torch.cuda.device(X.device):
            check(_lib.load().sc_rowconf_from_rows(_ptr(xr), _ptr(inv), _ptr(th), _ptr(tl), N, C, xr.shape[1], float(scale),
                                                   float(prob_scale), mode, _ptr(conf), _ptr(label), _stream()),
                  "sc_rowconf_from_rows")
        return conf, label
    xh, xl = normalize_split(X, feature_major, normalize=True)
    with torch.cuda.device(X.device):
        check(_lib.load().sc_rowconf_from_split(_ptr(xh), _ptr(xl), _ptr(th), _ptr(tl), N, C, xh.shape[1], float(scale),
                                                float(prob_scale), mode, _ptr(conf), _ptr(label), _stream()),
              "sc_rowconf_from_split")
    return conf, label


def topk_per_class(conf: torch.Tensor, label: torch.Tensor, n_classes: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(int64 [C, k] row indices, most confident first, -1 padded; int32 [C] counts)."""
    _cuda(conf, "conf"), _cuda(label, "label")
    conf = conf.to(torch.float32).contiguous()
    label = label.to(torch.int32).contiguous()
    N = conf.numel()
    lib = _lib.load()
    ws_bytes = int(lib.sc_topk_workspace_bytes(N, n_classes))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=conf.device)
    off = (-ws.data_ptr()) % 256
    out_idx = torch.empty((n_classes, k), dtype=torch.int64, device=conf.device)
    out_count = torch.empty(n_classes, dtype=torch.int32, device=conf.device)
    with torch.cuda.device(conf.device):
        check(lib.sc_topk_per_class(_ptr(conf), _ptr(label), N, n_classes, k, _ptr(out_idx), _ptr(out_count),
                                    ctypes.c_void_p(ws.data_ptr() + off), ws_bytes, _stream()), "sc_topk_per_class")
    return out_idx, out_count


def select_topk_per_label(conf: torch.Tensor, label: torch.Tensor, n_classes: int, k: int) -> torch.Tensor:
    """Reference-shaped result of cache_strategy.py:48-59: LongTensor of selected row indices, classes
    ascending, most confident first within a class."""
    out_idx, _ = topk_per_class(conf, label, n_classes, k)
    flat = out_idx.reshape(-1)
    return flat[flat >= 0]


def values_prepare(L: Optional[torch.Tensor], n_classes: int, idx: Optional[torch.Tensor] = None,
                   labels: Optional[torch.Tensor] = None, softmax_scale: Optional[float] = None,
                   ones_row: bool = False, op_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Transposed cache values Vt [C_pad, Nk_pad] of op_dtype.  softmax_scale=None -> one-hot (argmax of L, or
    `labels` if given); otherwise softmax(softmax_scale * L, dim=1).  ones_row appends a row of ones at
    index n_classes (row sums come out of GEMM-2 as an extra class)."""
    dev = (L if L is not None else labels).device
    if L is not None:
        _cuda(L, "L")
        assert L.dim() == 2 and L.stride(1) == 1
        N, C = L.shape
        assert C == n_classes
        ld = L.stride(0) if N > 1 else C
    else:
        N, C, ld = 0, n_classes, n_classes
    if labels is not None:
        if softmax_scale is not None:
            raise ValueError("values_prepare: `labels` replace the argmax of one-hot values; softmax values are built from L[idx]")
        labels = _cuda(labels, "labels").to(torch.int32).contiguous()
        n_out = labels.numel()
    elif idx is not None:
        idx = _cuda(idx, "idx").to(torch.int64).contiguous()
        n_out = idx.numel()
    else:
        n_out = N
    C_eff = n_classes + (1 if ones_row else 0)
    C_pad, Nk_pad = pad_classes(C_eff), pad_keys(max(n_out, 1))
    Vt = torch.empty((C_pad, Nk_pad), dtype=_op(op_dtype), device=dev)
    mode = SC_VALUES_HARD if softmax_scale is None else SC_VALUES_SOFTMAX
    with torch.cuda.device(dev):
        check(_lib.load().sc_values_prepare(_ptr(L), _code(L) if L is not None else SC_F32, N, C, ld, _ptr(idx),
                                            _ptr(labels), n_out, mode,
                                            float(softmax_scale) if softmax_scale is not None else 1.0, _ptr(Vt),
                                            _code(Vt), C_pad, Nk_pad, n_classes if ones_row else -1, _stream()),
              "sc_values_prepare")
    return Vt


def hard_supported(n_classes: int) -> bool:
    """True when the hard-label attention kernel (values synthesised on chip from labels) covers n_classes."""
    if os.environ.get("SUMMER_CLIP_B200_DENSE_VALUES"):      # A/B knob: stream dense one-hot values instead
        return False
    return bool(_lib.load().sc_attn_hard_supported(int(n_classes)))


def hard_labels(L: Optional[torch.Tensor], n_classes: int, idx: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int16 [sc_pad_labels(n_out)] labels of one-hot cache values: argmax_c L[idx] (HardCacheStrategy,
    cache_value_strategy.py:14-17) or the given gold labels; -1 pads the last 128-key tile."""
    dev = (L if L is not None else labels).device
    if L is not None:
        _cuda(L, "L")
        assert L.dim() == 2 and L.stride(1) == 1 and L.shape[1] == n_classes
        N, C = L.shape
        ld = L.stride(0) if N > 1 else C
    else:
        N, C, ld = 0, n_classes, n_classes
    if labels is not None:
        labels = _cuda(labels, "labels").to(torch.int32).contiguous()
        n_out = labels.numel()
    elif idx is not None:
        idx = _cuda(idx, "idx").to(torch.int64).contiguous()
        n_out = idx.numel()
    else:
        n_out = N
    lib = _lib.load()
    n_pad = int(lib.sc_pad_labels(n_out))
    out = torch.empty(n_pad, dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        check(lib.sc_hard_labels(_ptr(L), _code(L) if L is not None else SC_F32, N, C, ld, _ptr(idx), _ptr(labels),
                                 n_out, _ptr(out), n_pad, _stream()), "sc_hard_labels")
    return out


class HardBank:
    """Label-sorted key bank for sc_attn_fwd_hard: keys of one class adjacent, every class segment padded to
    whole 16-key groups.  `perm[j]` = original index of sorted key j (-1 = padding), `gcls` int16 class per
    16-key group, `kbits` one validity bit per key (uint32 words stored as int32); `rows` (set by `gather`) =
    the permuted normalised bank."""

    def __init__(self, perm: torch.Tensor, gcls: torch.Tensor, kbits: torch.Tensor, n_sorted: int, n_keys: int,
                 n_classes: int) -> None:
        self.perm, self.gcls, self.kbits = perm, gcls, kbits
        self.n_sorted, self.n_keys, self.n_classes = int(n_sorted), int(n_keys), int(n_classes)
        self.rows: Optional[torch.Tensor] = None

    def gather(self, k_norm: torch.Tensor) -> "HardBank":
        """Permute a normalised bank [>= n_keys, D_pad] (original key order) into sorted order (sc_gather_rows)."""
        _cuda(k_norm, "k_norm")
        assert k_norm.dim() == 2 and k_norm.is_contiguous() and (k_norm.shape[1] * k_norm.element_size()) % 16 == 0
        n_out = self.perm.numel()
        rows = torch.empty((n_out, k_norm.shape[1]), dtype=k_norm.dtype, device=k_norm.device)
        with torch.cuda.device(k_norm.device):
            check(_lib.load().sc_gather_rows(_ptr(k_norm), k_norm.shape[0], k_norm.shape[1] * k_norm.element_size(),
                                             _ptr(self.perm), n_out, _ptr(rows), _stream()), "sc_gather_rows")
        self.rows = rows
        return self


def hard_bank_build(labels: torch.Tensor, n_classes: int, feats: torch.Tensor, feature_major: bool,
                    idx: Optional[torch.Tensor] = None, normalize: bool = True,
                    op_dtype: Optional[torch.dtype] = None) -> HardBank:
    """The label-sorted bank straight from the RAW feature bank in one pass over it: layout from the labels, then
    every selected key's normalised row is written directly to its sorted position (sc_hard_bank_inverse +
    sc_normalize_scatter) — no intermediate normalised bank and no gather pass (normalize_cast + HardBank.gather
    move the bank through HBM twice).  `labels`: one per selected key (int16 from hard_labels, or any int type);
    `idx`: the selected columns / rows of `feats` (None = all)."""
    _cuda(feats, "feats")
    bank = hard_bank_layout(labels, n_classes)
    n_keys = bank.n_keys
    if feature_major:
        D, N = feats.shape
        stride_d, stride_n = feats.stride()
    else:
        N, D = feats.shape
        stride_n, stride_d = feats.stride()
    if idx is not None:
        idx = _cuda(idx, "idx").to(torch.int64).contiguous()
        assert idx.numel() == n_keys
    else:
        assert N == n_keys
    op_dtype = _op(op_dtype, allow_e4m3=True)
    D_pad = pad_dim(D, op_dtype)
    n_rows = bank.perm.numel()
    rows = torch.empty((n_rows, D_pad), dtype=op_dtype, device=feats.device)
    inv = torch.empty(max(n_keys, 1), dtype=torch.int64, device=feats.device)
    lib = _lib.load()
    with torch.cuda.device(feats.device):
        check(lib.sc_hard_bank_inverse(_ptr(bank.perm), n_rows, n_keys, _ptr(inv), _ptr(rows), D_pad * rows.element_size(),
                                       _stream()), "sc_hard_bank_inverse")
        if n_keys > 0:
            check(lib.sc_normalize_scatter(_ptr(feats), _code(feats), D, N, stride_d, stride_n, _ptr(idx), n_keys, _ptr(inv),
                                           _ptr(rows), _code(rows), D_pad, int(normalize), _stream()), "sc_normalize_scatter")
    bank.rows = rows
    return bank


def hard_bank_layout(labels: torch.Tensor, n_classes: int) -> HardBank:
    """Layout of the label-sorted bank (once per cache): stable sort of the keys by label, class segments padded
    to multiples of 16, the whole bank padded to whole 256-key steps.  Labels outside [0, n_classes) select no
    class and are dropped (their one-hot row is zero).  One library call (sc_hard_bank_layout); the index
    arithmetic it is tested against, element for element, lives in tests/bank_layout_spec.py."""
    return _hard_bank_layout_cuda(_cuda(labels, "labels"), n_classes)


def _hard_bank_layout_cuda(labels: torch.Tensor, n_classes: int) -> HardBank:
    lib = _lib.load()
    dev = labels.device
    lab = labels.reshape(-1)
    n_keys = lab.numel()
    if lab.dtype != torch.int16:                       # int32 / int64 labels -> int16, invalid -> -1 (sc_hard_labels)
        lab = hard_labels(None, n_classes, labels=lab)[:n_keys]
    lab = lab.contiguous()
    cap = int(lib.sc_hard_bank_capacity(n_keys, n_classes))
    ws_bytes = int(lib.sc_hard_bank_workspace_bytes(n_keys, n_classes))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    perm = torch.empty(cap, dtype=torch.int64, device=dev)
    gcls = torch.empty(cap // 16, dtype=torch.int16, device=dev)
    kbits = torch.empty(cap // 32, dtype=torch.int32, device=dev)
    n_sorted = torch.empty(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.sc_hard_bank_layout(_ptr(lab) if n_keys else None, n_keys, n_classes, _ptr(perm), _ptr(gcls), _ptr(kbits),
                                      cap, _ptr(n_sorted), ctypes.c_void_p(ws.data_ptr() + off), ws_bytes, _stream()),
              "sc_hard_bank_layout")
    ns = int(n_sorted.item())                          # the one host sync of a cache build: sizes the bank
    steps = max(1, -(-ns // 256))
    return HardBank(perm[: steps * 256], gcls[: steps * 16], kbits[: steps * 8], ns, n_keys, n_classes)


def attn_fwd_hard(Qn: torch.Tensor, bank: HardBank, beta: float, splits: int = 0, merge: bool = True) -> torch.Tensor:
    """attn_fwd for one-hot values on a label-sorted bank (hard_bank_layout + HardBank.gather):
    fp32 [Nq, n_classes] = sum over the keys of each class of exp(beta (q.k - 1))."""
    _cuda(Qn, "Qn")
    Ks = bank.rows
    assert Ks is not None, "HardBank.gather(k_norm) must be called first"
    assert Ks.dtype == Qn.dtype and Qn.dtype in (torch.float16, torch.bfloat16, E4M3) and Ks.is_contiguous() and Qn.is_contiguous()
    Qn = _tile_padded(Qn)
    Nq, D_pad = Qn.shape
    assert Ks.shape[1] == D_pad and Ks.shape[0] >= bank.n_sorted
    n_classes = bank.n_classes
    n_sorted = max(bank.n_sorted, 1)
    if splits <= 0:
        splits = attn_hard_splits(Nq, n_sorted, Qn.device, bank=bank)
    O = torch.empty((splits, Nq, n_classes), dtype=torch.float32, device=Qn.device)      # zeroed by the library
    with torch.cuda.device(Qn.device):
        check(_lib.load().sc_attn_fwd_hard(_ptr(Qn), _ptr(Ks), _ptr(bank.gcls), _ptr(bank.kbits), _code(Qn), Nq,
                                           n_sorted, D_pad, n_classes, float(beta), splits, _ptr(O), n_classes,
                                           _stream()), "sc_attn_fwd_hard")
    if not merge:
        return O
    if splits == 1:
        return O[0]
    return merge_partials(O)


def attn_fwd_hard_multi(Qn: torch.Tensor, bank: HardBank, betas: Sequence[float], splits: int = 0,
                        merge: bool = True) -> list:
    """`attn_fwd_hard` for a LIST of betas: one tensor-core pass per group of up to 4 betas (S = Q.K^T does not
    depend on beta; only the exponentials are per beta).  Returns one fp32 [Nq, n_classes] tensor per beta
    (merge=False: the unmerged [splits, Nq, n_classes] partial tiles, which `epilogue` sums on the fly)."""
    _cuda(Qn, "Qn")
    Ks = bank.rows
    assert Ks is not None, "HardBank.gather(k_norm) must be called first"
    assert Ks.dtype == Qn.dtype and Qn.dtype in (torch.float16, torch.bfloat16, E4M3) and Ks.is_contiguous() and Qn.is_contiguous()
    Qn = _tile_padded(Qn)
    Nq, D_pad = Qn.shape
    assert Ks.shape[1] == D_pad and Ks.shape[0] >= bank.n_sorted
    n_classes, n_sorted = bank.n_classes, max(bank.n_sorted, 1)
    if splits <= 0:
        splits = attn_hard_splits(Nq, n_sorted, Qn.device, bank=bank)
    lib = _lib.load()
    outs = []
    betas = [float(b) for b in betas]
    for g0 in range(0, len(betas), 4):
        grp = betas[g0:g0 + 4]
        O = torch.empty((len(grp), splits, Nq, n_classes), dtype=torch.float32, device=Qn.device)   # zeroed by the library
        arr = (ctypes.c_float * len(grp))(*grp)
        with torch.cuda.device(Qn.device):
            check(lib.sc_attn_fwd_hard_multi(_ptr(Qn), _ptr(Ks), _ptr(bank.gcls), _ptr(bank.kbits), _code(Qn), Nq, n_sorted,
                                             D_pad, n_classes, arr, len(grp), splits, _ptr(O), n_classes, _stream()),
                  "sc_attn_fwd_hard_multi")
        for bi in range(len(grp)):
            outs.append(O[bi] if not merge else O[bi, 0] if splits == 1 else merge_partials(O[bi]))
    return outs


# A/B knob for experiments (read once): force the number of key splits of both attention kernels
_SPLITS_OVERRIDE = int(os.environ.get("SUMMER_CLIP_B200_SPLITS", "0") or 0)


def attn_hard_splits(Nq: int, Nks: int, device=None, bank: Optional[HardBank] = None) -> int:
    """Key splits of the segmented kernel.  With `bank` (gathered) the choice also charges every extra split the
    zeroing and reading back of its [Nq, n_classes] tile (sc_attn_hard_splits_for).  Deliberately independent of
    how many betas share the launch: the split count fixes the summation grouping, so a beta computed in a sweep
    stays bit-identical to the same beta computed alone."""
    if _SPLITS_OVERRIDE > 0:
        return max(1, min(_SPLITS_OVERRIDE, -(-Nks // 256)))
    sms = torch.cuda.get_device_properties(device or torch.cuda.current_device()).multi_processor_count
    if bank is not None and bank.rows is not None:
        return int(_lib.load().sc_attn_hard_splits_for(Nq, Nks, bank.rows.shape[1], _code(bank.rows), bank.n_classes, 1, sms))
    return int(_lib.load().sc_attn_hard_splits(Nq, Nks, sms))


def attn_splits(Nq: int, Nk: int, C_pad: int, device=None, D_pad: int = 0) -> int:
    """Key splits of the dense-values kernel; with `D_pad` the choice is L2-blocked (sc_attn_splits_for)."""
    if _SPLITS_OVERRIDE > 0:
        return max(1, min(_SPLITS_OVERRIDE, -(-Nk // 128)))
    sms = torch.cuda.get_device_properties(device or torch.cuda.current_device()).multi_processor_count
    if D_pad > 0:
        return int(_lib.load().sc_attn_splits_for(Nq, Nk, D_pad, C_pad, sms))
    return int(_lib.load().sc_attn_splits(Nq, Nk, C_pad, sms))


# fp32 partial tiles of one dense-values launch are bounded by chunking the queries (an L2-blocked launch has ~80 key
# splits at the ImageNet size: 16 GB of [splits, Nq, C] tiles for 50 000 x 1000, which is allowed in one piece)
_MAX_PARTIAL_BYTES = 20 << 30


def merge_partials(parts: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sum fp32 partials [n_parts, rows, cols] over the first dimension."""
    _cuda(parts, "parts")
    assert parts.dim() == 3 and parts.dtype == torch.float32 and parts.is_contiguous()
    n_parts, rows, cols = parts.shape
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.float32, device=parts.device)
    with torch.cuda.device(parts.device):
        check(_lib.load().sc_merge_partials(_ptr(parts), n_parts, rows, cols, cols, _ptr(out), out.stride(0),
                                            _stream()), "sc_merge_partials")
    return out


def merge_peer_parts(parts: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sum of equally shaped fp32 [rows, cols] tiles that may live in OTHER GPUs' memory (peer-mapped views of a
    symmetric-memory allocation): the key-sharded ranks' partial tiles, read in place over NVLink."""
    rows, cols = parts[0].shape
    for t in parts:
        assert t.is_cuda and t.dtype == torch.float32 and t.shape == (rows, cols) and t.stride(1) == 1 and \
            (rows <= 1 or t.stride(0) == parts[0].stride(0))
    dev = out.device if out is not None else torch.device("cuda", torch.cuda.current_device())
    if out is None:
        out = torch.empty((rows, cols), dtype=torch.float32, device=dev)
    arr = (ctypes.c_void_p * len(parts))(*[t.data_ptr() for t in parts])
    with torch.cuda.device(dev):
        check(_lib.load().sc_merge_peer_parts(arr, len(parts), rows, cols, parts[0].stride(0) if rows > 1 else cols, _ptr(out),
                                              out.stride(0), _stream()), "sc_merge_peer_parts")
    return out


def attn_fwd(Qn: torch.Tensor, Kn: torch.Tensor, Vt: torch.Tensor, n_keys: int, n_cols: int, beta: float,
             splits: int = 0, merge: bool = True, row_shift: Optional[torch.Tensor] = None) -> torch.Tensor:
    """O[q, c] = sum_k exp(beta (Qn[q].Kn[k] - 1)) Vt[c, k]  ->  fp32 [Nq, n_cols] (or the
    [splits, Nq, n_cols] partials when merge=False).  `row_shift` fp32 [Nq] replaces the constant 1 per query
    (sc_attn_fwd_shifted: the softmax mode passes the row maximum of `attn_rowmax`)."""
    for t, n in ((Qn, "Qn"), (Kn, "Kn"), (Vt, "Vt")):
        _cuda(t, n)
        assert t.dtype == Qn.dtype and t.dtype in (torch.float16, torch.bfloat16) and t.is_contiguous(), \
            f"{n} must be contiguous fp16/bf16 (all three of one type)"
    Nq, D_pad = Qn.shape
    assert Kn.shape[1] == D_pad and Kn.shape[0] >= n_keys
    C_pad, Nk_pad = Vt.shape
    if splits <= 0:
        splits = attn_splits(Nq, n_keys, C_pad, Qn.device, D_pad=D_pad)
    if merge and splits > 1 and splits * Nq * n_cols * 4 > _MAX_PARTIAL_BYTES:
        # query chunks of whole 128-row tiles whose partial tiles fit the bound; each chunk re-reads the bank once
        rows = _MAX_PARTIAL_BYTES // (splits * n_cols * 4)
        wave = 128 * max(1, torch.cuda.get_device_properties(Qn.device).multi_processor_count // 4)   # one wave of clusters
        rows = max(128, rows // wave * wave if rows >= wave else rows // 128 * 128)
        out = torch.empty((Nq, n_cols), dtype=torch.float32, device=Qn.device)
        for s0 in range(0, Nq, rows):
            s1 = min(Nq, s0 + rows)
            part = attn_fwd(Qn[s0:s1], Kn, Vt, n_keys, n_cols, beta, splits=splits, merge=False,
                            row_shift=row_shift[s0:s1].contiguous() if row_shift is not None else None)
            merge_partials(part, out=out[s0:s1])
        return out
    O = torch.empty((splits, Nq, n_cols), dtype=torch.float32, device=Qn.device)
    if row_shift is not None:
        row_shift = _cuda(row_shift, "row_shift")
        assert row_shift.dtype == torch.float32 and row_shift.is_contiguous() and row_shift.numel() == Nq
    with torch.cuda.device(Qn.device):
        check(_lib.load().sc_attn_fwd_shifted(_ptr(Qn), _ptr(Kn), _ptr(Vt), _code(Qn), Nq, n_keys, D_pad, n_cols, C_pad,
                                              Nk_pad, float(beta), _ptr(row_shift), splits, _ptr(O), n_cols, _stream()),
              "sc_attn_fwd")
    if not merge:
        return O
    if splits == 1:
        return O[0]
    return merge_partials(O)


def attn_rowmax(Qn: torch.Tensor, Kn: torch.Tensor, n_keys: int) -> torch.Tensor:
    """fp32 [Nq]: max_k Qn[q].Kn[k] over the first n_keys rows of Kn (tensor-core pass, nothing materialised)."""
    _cuda(Qn, "Qn"), _cuda(Kn, "Kn")
    assert Qn.dtype == Kn.dtype and Qn.is_contiguous() and Kn.is_contiguous() and Kn.shape[1] == Qn.shape[1]
    Qn = _tile_padded(Qn)
    out = torch.empty(Qn.shape[0], dtype=torch.float32, device=Qn.device)
    with torch.cuda.device(Qn.device):
        check(_lib.load().sc_attn_rowmax(_ptr(Qn), _ptr(Kn), _code(Qn), Qn.shape[0], int(n_keys), Qn.shape[1], _ptr(out),
                                         _stream()), "sc_attn_rowmax")
    return out


def attn_softmax_hard(Qn: torch.Tensor, bank: HardBank, tau: float, splits: int = 0) -> torch.Tensor:
    """Temperature-softmax attention on a label-sorted bank: fp32 [splits, Nq, n_classes] per-class base-2
    log-sum-exp of tau * Qn.Ks^T (online running maximum inside the kernel; -inf = no key of that class)."""
    _cuda(Qn, "Qn")
    Ks = bank.rows
    assert Ks is not None, "HardBank.gather(k_norm) must be called first"
    assert Ks.dtype == Qn.dtype and Ks.is_contiguous() and Qn.is_contiguous() and Ks.shape[1] == Qn.shape[1]
    Qn = _tile_padded(Qn)
    Nq, D_pad = Qn.shape
    n_classes, n_sorted = bank.n_classes, max(bank.n_sorted, 1)
    if splits <= 0:
        splits = attn_hard_splits(Nq, n_sorted, Qn.device, bank=bank)
    lse = torch.empty((splits, Nq, n_classes), dtype=torch.float32, device=Qn.device)     # filled with -inf by the library
    with torch.cuda.device(Qn.device):
        check(_lib.load().sc_attn_softmax_hard(_ptr(Qn), _ptr(Ks), _ptr(bank.gcls), _ptr(bank.kbits), _code(Qn), Nq,
                                               n_sorted, D_pad, n_classes, float(tau), splits, _ptr(lse), n_classes,
                                               _stream()), "sc_attn_softmax_hard")
    return lse


def softmax_partials(lse: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """LSE tiles [n_parts, Nq, C] -> the (O [Nq, C], m [Nq], l [Nq]) partial triple of those keys."""
    _cuda(lse, "lse")
    assert lse.dim() == 3 and lse.dtype == torch.float32 and lse.is_contiguous()
    n_parts, Nq, C = lse.shape
    O = torch.empty((Nq, C), dtype=torch.float32, device=lse.device)
    m = torch.empty(Nq, dtype=torch.float32, device=lse.device)
    l = torch.empty(Nq, dtype=torch.float32, device=lse.device)
    with torch.cuda.device(lse.device):
        check(_lib.load().sc_softmax_partials(_ptr(lse), n_parts, lse.stride(0), Nq, C, C, _ptr(O), C, _ptr(m), _ptr(l),
                                              _stream()), "sc_softmax_partials")
    return O, m, l


def merge_softmax(O_parts: torch.Tensor, m_parts: torch.Tensor, l_parts: torch.Tensor, m_scale: float = 1.0,
                  m_ref: Optional[torch.Tensor] = None, normalize: bool = True, inplace: bool = False):
    """Log-sum-exp merge of (O [P, Nq, C], m [P, Nq], l [P, Nq]) partial triples -> (out [Nq, C], M [Nq], L [Nq]);
    out is divided by L when `normalize`.  `m_ref`: a common maximum agreed between ranks."""
    for t, n in ((O_parts, "O_parts"), (m_parts, "m_parts"), (l_parts, "l_parts")):
        _cuda(t, n)
        assert t.dtype == torch.float32
    P, Nq, C = O_parts.shape
    assert O_parts.stride(2) == 1 and (P == 1 or O_parts.stride(0) >= Nq * O_parts.stride(1))
    m_parts, l_parts = m_parts.contiguous(), l_parts.contiguous()
    assert m_parts.shape == (P, Nq) and l_parts.shape == (P, Nq)
    out = O_parts[0] if (inplace and P == 1) else torch.empty((Nq, C), dtype=torch.float32, device=O_parts.device)
    M = torch.empty(Nq, dtype=torch.float32, device=O_parts.device)
    L = torch.empty(Nq, dtype=torch.float32, device=O_parts.device)
    if m_ref is not None:
        m_ref = _cuda(m_ref, "m_ref").to(torch.float32).contiguous()
    with torch.cuda.device(O_parts.device):
        check(_lib.load().sc_merge_softmax(_ptr(O_parts), _ptr(m_parts), _ptr(l_parts), P, O_parts.stride(0), Nq, Nq, C,
                                           O_parts.stride(1), float(m_scale), _ptr(m_ref), int(normalize), _ptr(out),
                                           out.stride(0), _ptr(M), _ptr(L), _stream()), "sc_merge_softmax")
    return out, M, L


def normalize_split(x: torch.Tensor, feature_major: bool, normalize: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """(hi, lo) fp16 [N, D_pad]: the (column-normalised) bank as split-fp16 rows, hi + lo = the fp32 value to 22
    bits — the operands of `gemm_split_nt`."""
    _cuda(x, "x")
    assert x.dim() == 2
    if feature_major:
        D, N = x.shape
        stride_d, stride_n = x.stride()
    else:
        N, D = x.shape
        stride_n, stride_d = x.stride()
    D_pad = pad_dim(D)
    hi = torch.empty((N, D_pad), dtype=torch.float16, device=x.device)
    lo = torch.empty((N, D_pad), dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().sc_normalize_split(_ptr(x), _code(x), D, N, stride_d, stride_n, _ptr(hi), _ptr(lo), D_pad,
                                             int(normalize), _stream()), "sc_normalize_split")
    return hi, lo


def gemm_split_nt(a: Tuple[torch.Tensor, torch.Tensor], b: Tuple[torch.Tensor, torch.Tensor], scale: float = 1.0,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Z [M, N] fp32 = scale * A @ B^T for split-fp16 operands A = (hi, lo) [M, D_pad], B = (hi, lo) [N, D_pad]
    on the tcgen05 pipeline (fp32-accurate: three passes hi.hi + hi.lo + lo.hi into one TMEM accumulator)."""
    ah, al = a
    bh, bl = b
    M, D_pad = ah.shape
    N = bh.shape[0]
    assert bh.shape[1] == D_pad and all(t.dtype == torch.float16 and t.is_contiguous() and t.is_cuda for t in (ah, al, bh, bl))
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=ah.device)
    assert out.dtype == torch.float32 and out.stride(1) == 1 and out.shape == (M, N)
    with torch.cuda.device(ah.device):
        check(_lib.load().sc_gemm_split_nt(_ptr(ah), _ptr(al), _ptr(bh), _ptr(bl), M, N, D_pad, float(scale), _ptr(out),
                                           out.stride(0), _stream()), "sc_gemm_split_nt")
    return out


def gemm_rows_nt(a: torch.Tensor, row_scale: Optional[torch.Tensor], b: Tuple[torch.Tensor, torch.Tensor], scale: float = 1.0,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Z [M, N] fp32 = scale * row_scale[:, None] * A @ B^T for fp16 rows A [M, D_pad] that need no lo part (raw fp16
    feature rows) and a split-fp16 B = (hi, lo) [N, D_pad]: two operand passes instead of gemm_split_nt's three."""
    bh, bl = b
    M, D_pad = a.shape
    N = bh.shape[0]
    assert bh.shape[1] == D_pad and all(t.dtype == torch.float16 and t.is_contiguous() and t.is_cuda for t in (a, bh, bl))
    if row_scale is not None:
        assert row_scale.is_cuda and row_scale.dtype == torch.float32 and row_scale.is_contiguous() and row_scale.numel() >= M
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    assert out.dtype == torch.float32 and out.stride(1) == 1 and out.shape == (M, N)
    with torch.cuda.device(a.device):
        check(_lib.load().sc_gemm_rows_nt(_ptr(a), _ptr(row_scale), _ptr(bh), _ptr(bl), M, N, D_pad, float(scale), _ptr(out),
                                          out.stride(0), _stream()), "sc_gemm_rows_nt")
    return out


_TWO_PASS_MIN_ROWS = 1 << 17      # zero_shot_logits: fp16 banks of at least this many rows skip the hi/lo split


def text_split(T: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Split-fp16 rows of T^T for a classifier T [D, C] (as given: not re-normalised)."""
    return normalize_split(T, feature_major=True, normalize=False)


def zero_shot_logits(X: torch.Tensor, feature_major: bool, T: torch.Tensor, scale: float = 100.0,
                     normalize: bool = True, tensor_cores: Optional[bool] = None,
                     t_split: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                     two_pass: Optional[bool] = None) -> torch.Tensor:
    """Z = scale * normalise(X)^T @ T in fp32 (image_attention.py:80-83).  T is [D, C].  Default route: split-fp16
    operands on the tensor cores (`normalize_split` + `gemm_split_nt`; `two_pass` — default for fp16 X of bank size:
    raw rows + 1/norm, `gemm_rows_nt`, two passes instead of three); `tensor_cores=False` (or
    SUMMER_CLIP_B200_ZS_SIMT=1) selects the fp32 SIMT kernel (A/B runs, cross-check in the tests).  `t_split` =
    `text_split(T)` computed once by the caller (ClipSearcher does) saves two tiny launches per call."""
    _cuda(X, "X"), _cuda(T, "T")
    if feature_major:
        D, N = X.shape
        stride_d, stride_n = X.stride()
    else:
        N, D = X.shape
        stride_n, stride_d = X.stride()
    assert T.dim() == 2 and T.shape[0] == D
    if tensor_cores is None:
        tensor_cores = not os.environ.get("SUMMER_CLIP_B200_ZS_SIMT")
    if tensor_cores and N > 0:
        ts = t_split if t_split is not None else text_split(T)
        if two_pass is None:
            # bank-sized products (the logits-bank producer, pseudo-labels) take the two-pass route; a query batch
            # keeps the three-pass split the search pipeline was tuned with (50k queries: 0.3 ms either way, and the
            # end-to-end bench measured the split's side-stream kernels a few ms kinder to the attention launch)
            two_pass = N >= _TWO_PASS_MIN_ROWS
        if X.dtype == torch.float16 and two_pass:
            # fp16 rows are exact: the raw transposed rows (+ 1/norm per row) against the split classifier, two
            # passes; rows that already are [N, D_pad] K-major go in as they are
            inv = torch.empty(N, dtype=torch.float32, device=X.device) if normalize else None
            if not normalize and not feature_major and X.is_contiguous() and D == pad_dim(D) and X.data_ptr() % 16 == 0:
                xr = X
            else:
                xr = normalize_cast(X, feature_major, normalize=False, op_dtype=torch.float16, inv_norm=inv)
            return gemm_rows_nt(xr, inv, ts, scale)
        return gemm_split_nt(normalize_split(X, feature_major, normalize), ts, scale)
    if T.stride(1) != 1:
        T = T.contiguous()
    C = T.shape[1]
    Z = torch.empty((N, C), dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        check(_lib.load().sc_zero_shot_logits(_ptr(X), _code(X), D, N, stride_d, stride_n, _ptr(T), _code(T), C,
                                              T.stride(0), float(scale), int(normalize), _ptr(Z), C, _stream()),
              "sc_zero_shot_logits")
    return Z


def epilogue(Z: Optional[torch.Tensor], O: torch.Tensor, alphas: Sequence[float],
             labels: Optional[torch.Tensor] = None, rowsum: Optional[torch.Tensor] = None,
             want_logits: bool = False, want_pred: bool = True):
    """out = Z + O * alpha for each alpha; returns dict(pred int32 [na, Nq], top1/top5 int32 [na] counts,
    logits fp32 [na, Nq, C] if requested).  O is [Nq, C], or the UNMERGED [n_parts, Nq, C] partial tiles of a
    key-split attention launch (summed on the fly, bit-identical to merge_partials first)."""
    _cuda(O, "O")
    assert O.dtype == torch.float32 and O.stride(-1) == 1 and O.dim() in (2, 3)
    n_parts, part_stride = 1, 0
    if O.dim() == 3:
        n_parts, part_stride = O.shape[0], O.stride(0)
        O = O[0]
    Nq, C = O.shape
    if Z is not None:
        _cuda(Z, "Z")
        assert Z.dtype == torch.float32 and Z.shape == O.shape and Z.stride(1) == 1
    na = len(alphas)
    dev = O.device
    arr = (ctypes.c_float * na)(*[float(a) for a in alphas])
    pred = torch.empty((na, Nq), dtype=torch.int32, device=dev) if want_pred else None
    logits = torch.empty((na, Nq, C), dtype=torch.float32, device=dev) if want_logits else None
    top1 = top5 = None
    if labels is not None:
        labels = _cuda(labels, "labels").to(torch.int32).contiguous()
        top1 = torch.zeros(na, dtype=torch.int32, device=dev)
        top5 = torch.zeros(na, dtype=torch.int32, device=dev)
    if rowsum is not None:
        rowsum = _cuda(rowsum, "rowsum").to(torch.float32).contiguous()
    with torch.cuda.device(dev):
        check(_lib.load().sc_epilogue_parts(_ptr(Z), Z.stride(0) if Z is not None else C, _ptr(O), O.stride(0),
                                            n_parts, part_stride, _ptr(rowsum), Nq, C, arr, na, _ptr(labels),
                                            _ptr(logits), _ptr(pred), _ptr(top1), _ptr(top5), _stream()),
              "sc_epilogue")
    return {"pred": pred, "top1": top1, "top5": top5, "logits": logits}
