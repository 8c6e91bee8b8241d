"""B200 mirror of summer_clip/clip_searcher/utils.py (labels, accuracy, .npy savers)."""
from __future__ import annotations

import typing as tp
from pathlib import Path

import numpy as np
import torch

from .. import ops


def load_labels(dataset) -> torch.Tensor:
    """utils.py:10-12 — labels of a (image, label) dataset as an IntTensor."""
    labels = [label for _, label in dataset]  # type: ignore
    return torch.IntTensor(labels)


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


class FilesNamesManager:
    """utils.py:24-37."""

    def __init__(self, dir_path: Path, files_ext: str) -> None:
        self.dir_path = Path(dir_path).resolve()
        self.files_ext = files_ext
        self.dir_path.mkdir(parents=True, exist_ok=True)
        self.counter = 0

    def next_path(self) -> Path:
        new_path = self.get_path(str(self.counter))
        self.counter += 1
        return new_path

    def get_path(self, file_name: str) -> Path:
        return self.dir_path / f"{file_name}{self.files_ext}"


class TensorsNumpySaver:
    """utils.py:40-52."""

    def __init__(self, dir_path: Path) -> None:
        self.files_names_manager = FilesNamesManager(dir_path, files_ext=".npy")

    def save_tensor(self, tensor: torch.Tensor) -> Path:
        tensor_path = self.files_names_manager.next_path()
        return self.save_named_tensor(tensor, tensor_path.with_suffix("").name)

    def save_named_tensor(self, tensor: torch.Tensor, file_name: str) -> Path:
        tensor_np = tensor.cpu().numpy()
        tensor_path = self.files_names_manager.get_path(file_name)
        np.save(tensor_path, tensor_np)
        return tensor_path
