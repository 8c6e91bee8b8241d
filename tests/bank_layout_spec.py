"""Specification of the label-sorted key bank (sc_hard_bank_layout) as plain torch index arithmetic — TEST
INFRASTRUCTURE: the CUDA layout kernel is compared with it element for element, and the CPU host-logic tests use
it where no GPU is available.  The product path (summer_clip_b200.ops.hard_bank_layout) is the kernel only."""
from __future__ import annotations

import torch

from summer_clip_b200.ops import HardBank


def hard_bank_layout_spec(labels: torch.Tensor, n_classes: int) -> HardBank:
    dev = labels.device
    lab = labels.reshape(-1).to(torch.int64)
    n_keys = lab.numel()
    valid = (lab >= 0) & (lab < n_classes)
    lab_v = torch.where(valid, lab, torch.full_like(lab, n_classes))
    order = torch.argsort(lab_v, stable=True)
    counts = torch.bincount(lab_v, minlength=n_classes + 1)[:n_classes]
    padded = (counts + 15) // 16 * 16
    seg_start = torch.cumsum(padded, 0) - padded
    cls_start = torch.cumsum(counts, 0) - counts
    n_valid, n_sorted = (int(v) for v in torch.stack([counts.sum(), padded.sum()]).tolist())
    order_v = order[:n_valid]
    lab_sorted = lab_v[order_v]
    dest = seg_start[lab_sorted] + (torch.arange(n_valid, device=dev) - cls_start[lab_sorted])
    steps = max(1, -(-n_sorted // 256))
    perm = torch.full((steps * 256,), -1, dtype=torch.int64, device=dev)
    perm[dest] = order_v
    gcls = torch.full((steps * 16,), -1, dtype=torch.int16, device=dev)
    gcls[dest // 16] = lab_sorted.to(torch.int16)
    words = ((perm >= 0).view(-1, 32).to(torch.int64) << torch.arange(32, device=dev)).sum(1)
    kbits = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    return HardBank(perm, gcls, kbits, n_sorted, n_keys, n_classes)


