"""B200 mirror of summer_clip/tip_adapter/tip_adapter_imagenet.py — the training-free Tip-Adapter entry point.

Same trainer shape (setup_model + train_loop), same config (conf/tip_adapter_imagenet.yaml, conf/tip_adapter.yaml),
same ./caches/<dataset>/ files and the same three log lines.  `train_loop` (tip_adapter_imagenet.py:42-61) is the hot
path: zero-shot logits, the cache head at (init_beta, init_alpha), then `search_hp` over the 200 x 20 grid — on the
fused attention kernels instead of 4 000 recomputed GEMM pairs.  The CLIP towers are outside this path, so
`setup_model` (:19-40) starts from their outputs:

  load_cache / load_pre_feat True   ./caches/<dataset>/keys_<shots>shots.pt, values_<shots>shots.pt, test_f.pt,
                                    test_l.pt as the reference wrote them;
  False                             encoder outputs named by `train_features_path` ([augment_epoch, Nk, D] or
                                    [Nk, D]), `train_labels_path`, `test_features_path`, `test_labels_path`; the tail of
                                    build_cache_model / pre_load_features runs here and writes the cache files;
  clip_weights_path                 the text classifier [D, C] (tip_utils.clip_classifier output), default
                                    ./caches/<dataset>/clip_weights.pt.

    python -m summer_clip_b200.tip_adapter.tip_adapter_imagenet [CONFIG.yaml] [key=value ...]
"""
from __future__ import annotations

import os
import sys
import typing as tp
from pathlib import Path

import numpy as np
import torch

from .. import ops
from ..utils.config import Config, compose
from ..utils.log_utils import JsonLinesLogger
from . import utils as tip_utils


def _load(path: tp.Union[str, Path], device) -> torch.Tensor:
    path = Path(path)
    if path.suffix == ".npy":
        return torch.from_numpy(np.load(path)).to(device)
    return torch.load(path, map_location=device)


class TipAdapterTrainer:
    def __init__(self, cfg: tp.Mapping, run_dir: tp.Union[str, Path] = ".") -> None:
        self.cfg = cfg if isinstance(cfg, Config) else Config(cfg)
        self.run_dir = Path(run_dir)

    def setup_device(self) -> None:
        self.device = ops.require_cuda_device((self.cfg.get("meta") or {}).get("device"), "tip_adapter")

    def setup_logger(self) -> None:
        name = (self.cfg.get("exp") or {}).get("name", "tip_adapter")
        self.logger = JsonLinesLogger(name, self.run_dir / "tip_adapter.log")

    def setup_model(self) -> None:
        cache_dir = self.cfg.get("cache_dir") or os.path.join(str(self.run_dir), "caches", str(self.cfg["dataset"]))
        os.makedirs(cache_dir, exist_ok=True)
        self.cfg["cache_dir"] = cache_dir
        opt = lambda key: _load(self.cfg[key], self.device) if self.cfg.get(key) else None  # noqa: E731
        self.clip_weights = _load(self.cfg.get("clip_weights_path") or os.path.join(cache_dir, "clip_weights.pt"), self.device)
        self.cache_keys, self.cache_values = tip_utils.build_cache_model(
            self.cfg, opt("train_features_path"), opt("train_labels_path"))
        self.test_features, self.test_labels = tip_utils.pre_load_features(
            self.cfg, "test", opt("test_features_path"), opt("test_labels_path"))
        self.cache_keys, self.cache_values = self.cache_keys.to(self.device), self.cache_values.to(self.device)
        self.test_features, self.test_labels = self.test_features.to(self.device), self.test_labels.to(self.device)

    def setup(self) -> None:
        self.setup_device()
        self.setup_logger()
        self.setup_model()

    @torch.no_grad()
    def train_loop(self) -> tp.Dict[str, float]:
        head = tip_utils.TipAdapterHead(self.cache_keys, self.cache_values, self.test_features, self.clip_weights)
        # Zero-shot CLIP
        acc = tip_utils.cls_acc(head.clip_logits, self.test_labels)
        self.logger.log_info(f"**** Zero-shot CLIP's test accuracy: {acc}. ****")
        result = {"zero_shot_acc": acc}

        # Tip-Adapter
        beta, alpha = self.cfg['init_beta'], self.cfg['init_alpha']
        n = self.test_labels.shape[0]
        acc = 100 * int(head.top1_counts(beta, [alpha], self.test_labels)[0]) / n
        self.logger.log_info(f"**** Tip-Adapter's test accuracy: {acc}. ****")
        result["tip_acc"] = acc

        # Search Hyperparameters
        result["best_beta"], result["best_alpha"] = tip_utils.search_hp(
            self.cfg, self.cache_keys, self.cache_values, self.test_features, self.test_labels, self.clip_weights, head=head)
        return result


def run_trainer(trainer_cls, cfg, run_dir=".") -> TipAdapterTrainer:
    import random
    seed = int((cfg.get("meta") or {}).get("random_state", 42))
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    trainer = trainer_cls(cfg, run_dir)
    trainer.setup()
    trainer.result = trainer.train_loop()
    return trainer


def run(argv: tp.Optional[tp.Sequence[str]] = None, config_name: str = "tip_adapter_imagenet") -> TipAdapterTrainer:
    argv = list(sys.argv[1:] if argv is None else argv)
    conf_dir = Path(__file__).resolve().parent.parent / "conf"
    rest: tp.List[str] = []
    i = 0
    while i < len(argv):
        a = argv[i]
        if a == "--config-dir":
            conf_dir, i = Path(argv[i + 1]), i + 1
        elif a == "--config-name":
            config_name, i = argv[i + 1], i + 1
        elif a.endswith((".yaml", ".yml")) and "=" not in a:
            conf_dir, config_name = Path(a).resolve().parent, Path(a).stem
        else:
            rest.append(a)
        i += 1
    cfg = compose(conf_dir, config_name, rest)
    return run_trainer(TipAdapterTrainer, cfg, run_dir=(cfg.get("run_dir") or "."))


if __name__ == "__main__":
    run()
