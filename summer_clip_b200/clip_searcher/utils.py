"""B200 mirror of summer_clip/clip_searcher/utils.py (labels, accuracy, .npy savers)."""
from __future__ import annotations

import typing as tp
from pathlib import Path

import numpy as np
import torch

from .. import ops


def load_labels(dataset) -> torch.Tensor:
    """int32 labels of an iterable of (image, label) pairs — what the reference's `load_labels` (clip_searcher/utils.py:
    10-12) returns for the per-gold strategies' `cache_dataset`."""
    return torch.tensor([int(pair[1]) for pair in dataset], dtype=torch.int32)


def accuracy_counts(outputs: torch.Tensor, target: torch.Tensor) -> tp.Tuple[int, int]:
    """Number of rows whose target is the top-1 / within the top-5 (train_adapter.py:156-159) on the
    epilogue kernel: one pass over the logits, no [Nq, 5] top-k tensor, one small D2H copy."""
    res = ops.epilogue(None, outputs.float().contiguous(), [1.0], labels=target, want_pred=False)
    counts = torch.stack([res["top1"], res["top5"]]).cpu()
    return int(counts[0, 0]), int(counts[1, 0])


def compute_accuracy(outputs: torch.Tensor, target: torch.Tensor, topk=(1, 5)) -> tp.List[float]:
    """utils.py:15-21 — [acc@1, acc@5] in percent (only topk = (1, 5) exists on this path)."""
    if tuple(topk) != (1, 5):
        raise NotImplementedError("compute_accuracy supports topk=(1, 5), the only use on the CLIP-search path")
    c1, c5 = accuracy_counts(outputs, target)
    n = target.shape[0]
    return [100.0 * c1 / n, 100.0 * c5 / n]


class TensorsNumpySaver:
    """Numbered / named `.npy` dumps in one directory, with the reference's interface (`TensorsNumpySaver`,
    clip_searcher/utils.py:40-52: the records of the sweep carry the returned paths in `preds_path`,
    `cache_inds_path`, ...): `save_tensor` writes `<directory>/<n>.npy` with n = 0, 1, 2, ... in call order,
    `save_named_tensor` writes `<directory>/<name>.npy`; both return the absolute path."""

    def __init__(self, dir_path: tp.Union[str, Path]) -> None:
        self.directory = Path(dir_path).resolve()
        self.directory.mkdir(parents=True, exist_ok=True)
        self._saved = 0

    def save_named_tensor(self, tensor: torch.Tensor, file_name: str) -> Path:
        target = self.directory / (file_name + ".npy")
        np.save(target, tensor.detach().cpu().numpy())
        return target

    def save_tensor(self, tensor: torch.Tensor) -> Path:
        index, self._saved = self._saved, self._saved + 1
        return self.save_named_tensor(tensor, str(index))
