"""B200 mirror of summer_clip/clip_searcher/image_attention.py — the CLIP-search sweep driver.

Same trainer shape (setup_* hooks + train_loop), same config keys (conf/image_attention.yaml:21-44),
same loop nest and the same JSON records in the same order (image_attention.py:89-120), but:
  * the four nested loops share work — one gather+normalise of the cache per cache strategy, one values
    build per (cache, value strategy), one fused attention launch per (cache, beta, value strategy)
    and ONE epilogue launch for all alphas with on-device accuracy counters (the reference does a
    top-k + host sync per alpha);
  * inputs the reference derives from the CLIP towers or image datasets (text classifier, labels)
    are read from files: `data.text_features_path` (T [D, C]) or `data.clip_logits_path`,
    `data.labels_path`, `cache.labels_path`.  Encoder forward passes are outside this path.

    python -m summer_clip_b200.clip_searcher.image_attention [path/to/image_attention.yaml] [key=value ...]
"""
from __future__ import annotations

import sys
import typing as tp
from pathlib import Path

import numpy as np
import torch

from .. import ops
from ..searcher import ClipSearcher
from ..utils import hydra_utils
from ..utils.config import Config, compose
from ..utils.log_utils import JsonLinesLogger
from .cache_strategy import CacheStrategy, IndexedCacheStrategy, LazyLogitsBank
from .cache_value_strategy import GoldCacheValues, HardCacheStrategy, SoftmaxCacheStrategy
from .cache_weights_strategy import FusedWeights, NormalizedBank
from .utils import TensorsNumpySaver, compute_accuracy


def _load_tensor(path: tp.Union[str, Path], device) -> torch.Tensor:
    path = Path(path)
    if path.suffix == ".npy":
        return torch.from_numpy(np.load(path)).to(device)
    return torch.load(path, map_location=device)


class ImageAttention:
    def __init__(self, cfg: tp.Mapping, run_dir: tp.Union[str, Path] = ".") -> None:
        self.cfg = cfg if isinstance(cfg, Config) else Config(cfg)
        self.run_dir = Path(run_dir)

    # ---- setup hooks, in the order of BaseTrainer.setup (utils/trainer.py:62-70)
    def setup_device(self) -> None:
        self.device = ops.require_cuda_device((self.cfg.get("meta") or {}).get("device"), "image_attention")

    def setup_logger(self) -> None:
        name = (self.cfg.get("exp") or {}).get("name", "image_attention")
        self.logger = JsonLinesLogger(name, self.run_dir / "image_attention.log")
        self.gold_labels_saver = TensorsNumpySaver(self.run_dir / "gold_labels")
        self.cache_saver = TensorsNumpySaver(self.run_dir / "cache_ids")
        self.preds_saver = TensorsNumpySaver(self.run_dir / "preds_ids")

    def setup_dataset(self) -> None:
        if not self.cfg.data.get("labels_path"):
            raise ops._lib.SummerClipError("data.labels_path is required: the dataset readers of the reference are outside "
                                           "this path, labels come from a .pt / .npy file")
        self.test_labels = _load_tensor(self.cfg.data.labels_path, self.device).to(torch.int32)
        self.cache_labels: tp.Optional[torch.Tensor] = None
        if self.cfg.cache.get("labels_path") and Path(self.cfg.cache.labels_path).exists():
            self.cache_labels = _load_tensor(self.cfg.cache.labels_path, self.device).to(torch.int32)
        if self.cfg.run_saves.save_labels:
            self.save_labels()

    def save_labels(self) -> None:
        """`run_saves.save_labels`: the gold labels next to the log, under the reference's names (image_attention.py:33-36)."""
        for name, labels in (("test_labels", self.test_labels), ("cache_labels", self.cache_labels)):
            if labels is not None:
                self.gold_labels_saver.save_named_tensor(labels, name)

    def setup_model(self) -> None:
        self.searcher = ClipSearcher(self.device)
        self.test_image_features = _load_tensor(self.cfg.data.image_features_path, self.device)
        text = None
        if self.cfg.data.get("clip_logits_path"):
            self.clip_logits = _load_tensor(self.cfg.data.clip_logits_path, self.device).float().contiguous()
        else:
            text = _load_tensor(self.cfg.data.text_features_path, self.device)
            self.clip_logits = self.compute_clip_logits(text)
        self.test_q_norm = ops.normalize_cast(self.test_image_features, feature_major=True)
        self.origin_cache_image_features = _load_tensor(self.cfg.cache.image_features_path, self.device)
        if self.cfg.cache.get("image_outs_path"):
            self.origin_cache_image_outs = _load_tensor(self.cfg.cache.image_outs_path, self.device)
        else:
            # no stored logits bank (cache.image_outs_path: null): pseudo-labels come straight from the features and
            # the text classifier through the fused GEMM + row-scan kernel; the [N, C] bank is never written
            if text is None:
                text = _load_tensor(self.cfg.data.text_features_path, self.device)
            self.origin_cache_image_outs = LazyLogitsBank(self.origin_cache_image_features, text)
        self.logger.log_info(f"original-data-size: {self.origin_cache_image_outs.shape[0]}")

    def setup(self) -> None:
        self.setup_device()
        self.setup_logger()
        self.setup_dataset()
        self.setup_model()

    # ---- the path
    def compute_clip_logits(self, test_text_features: torch.Tensor) -> torch.Tensor:
        """image_attention.py:80-83 — 100 * normalise(Q)^T @ T (fp32 kernel)."""
        return ops.zero_shot_logits(self.test_image_features, True, test_text_features, scale=100.0)

    def logits_to_preds(self, logits: torch.Tensor) -> torch.Tensor:
        return ops.epilogue(None, logits.float().contiguous(), [1.0])["pred"][0].long()

    def build_cache(self, cache_strategy: CacheStrategy, image_features: torch.Tensor, image_outs: torch.Tensor):
        """image_attention.py:48-70.  Returns (cache keys as a NormalizedBank, (image_outs, idx) or gold
        labels for the value strategies, info)."""
        if not isinstance(cache_strategy, IndexedCacheStrategy):
            feats, outs = cache_strategy.transform(image_features, image_outs)
            return NormalizedBank(ops.normalize_cast(feats, feature_major=True)), (outs, None, None), {}
        samples_inds = cache_strategy.select(image_features, image_outs)
        cache_info: tp.Dict[str, tp.Any] = dict(cache_size=int(samples_inds.numel()))
        if self.cfg.run_saves.save_cache_inds:
            cache_info["cache_inds_path"] = str(self.cache_saver.save_tensor(samples_inds))
        gold = None
        if self.cache_labels is not None:
            cache_labels = self.cache_labels[samples_inds]
            eval_top1, eval_top5 = compute_accuracy(image_outs[samples_inds], cache_labels)
            cache_info.update(dict(acc1=eval_top1, acc5=eval_top5))
            if self.cfg.cache.get("replace_outs_with_golds", False):
                gold = cache_labels
                onehot = torch.nn.functional.one_hot(cache_labels.long(), num_classes=image_outs.shape[1]).float()
                eval_top1, eval_top5 = compute_accuracy(onehot, cache_labels)
                cache_info.update(dict(acc1_replace=eval_top1, acc5_replace=eval_top5))
        k_norm = ops.normalize_cast(image_features, feature_major=True, idx=samples_inds)   # gather + norm + cast
        return NormalizedBank(k_norm), (image_outs, samples_inds, gold), cache_info

    @torch.no_grad()
    def train_loop(self) -> None:
        # the zero-shot record first (image_attention.py:90-98): same keys, same order
        clip_logits = self.clip_logits
        acc1, acc5 = compute_accuracy(clip_logits, self.test_labels)
        record: tp.Dict[str, tp.Any] = {"acc1": acc1, "acc5": acc5}
        saves = self.cfg.run_saves
        if saves.save_preds:
            record["preds_path"] = str(self.preds_saver.save_tensor(self.logits_to_preds(clip_logits)))
        if saves.save_logits:
            record["logits_path"] = str(self.preds_saver.save_tensor(clip_logits))
        record["type"] = "zero_shot"
        self.logger.log_info(record)

        alphas = [float(a) for a in self.cfg.cache.alpha]
        n_q = self.test_labels.shape[0]
        q_bank = NormalizedBank(self.test_q_norm)
        for cache_strategy_cfg in self.cfg.cache_strategies.values():
            if cache_strategy_cfg is None:            # a group switched off on the command line (`...group=null`)
                continue
            for cache_strategy, cache_strategy_params in hydra_utils.instantiate_all(
                    cache_strategy_cfg, inject=self._strategy_inject(cache_strategy_cfg)):
                k_bank, (outs, idx, gold), cache_info = self.build_cache(
                    cache_strategy, self.origin_cache_image_features, self.origin_cache_image_outs)
                self.logger.log_info(dict(**cache_info, cache_strategy=cache_strategy_params, type="cache_info"))
                value_cache: tp.Dict[int, tp.Any] = {}
                weights_grid = [(ws.transform(q_bank, k_bank), wp)
                                for ws, wp in hydra_utils.instantiate_all(self.cfg.cache_weights_strategy)]
                values_grid = list(hydra_utils.instantiate_all(self.cfg.cache_value_strategy))
                for vi, (value_strategy, value_params) in enumerate(values_grid):
                    if gold is not None:
                        value_cache[vi] = self._gold_values(value_strategy, gold, outs.shape[1])
                    elif isinstance(outs, LazyLogitsBank):
                        if isinstance(value_strategy, HardCacheStrategy):       # argmax labels from the fused row scan
                            pred = outs.rowconf()[1]
                            value_cache[vi] = GoldCacheValues(outs.shape[1]).transform(pred if idx is None else pred[idx])
                        else:
                            value_cache[vi] = value_strategy.transform(outs.dense() if idx is None else outs[idx])
                    elif isinstance(value_strategy, (HardCacheStrategy, SoftmaxCacheStrategy)):
                        value_cache[vi] = value_strategy.transform(outs, idx=idx)
                    else:
                        value_cache[vi] = value_strategy.transform(outs if idx is None else outs[idx])
                # every (weights, values) product of this cache; betas share the tensor-core pass where they can
                logits_grid = {vi: FusedWeights.matmul_many([w for w, _ in weights_grid], value_cache[vi])
                               if all(isinstance(w, FusedWeights) for w, _ in weights_grid)
                               else [w @ value_cache[vi] for w, _ in weights_grid]
                               for vi in range(len(values_grid))}
                for wi, (_, weights_params) in enumerate(weights_grid):          # record order of the reference loop
                    for vi, (_, value_params) in enumerate(values_grid):
                        cache_logits = logits_grid[vi][wi]
                        res = ops.epilogue(clip_logits, cache_logits, alphas, labels=self.test_labels,
                                           want_pred=bool(self.cfg.run_saves.save_preds))
                        top1, top5 = res["top1"].cpu().tolist(), res["top5"].cpu().tolist()   # one D2H per beta
                        for ai, alpha in enumerate(self.cfg.cache.alpha):
                            searcher_info: tp.Dict[str, tp.Any] = dict(
                                cache_strategy=cache_strategy_params, cache_value_strategy=value_params,
                                cache_weights_strategy=weights_params, alpha=alpha,
                                acc1=100.0 * top1[ai] / n_q, acc5=100.0 * top5[ai] / n_q)
                            if self.cfg.run_saves.save_preds:
                                searcher_info["preds_path"] = str(self.preds_saver.save_tensor(res["pred"][ai].long()))
                            self.logger.log_info_wandb(dict(**searcher_info, type="searcher_result"))

    def _strategy_inject(self, cfg: tp.Mapping) -> dict:
        """`cache_dataset: [${cache.dataset}]` entries (conf/cache_strategy/topk_per_gold.yaml:3-4) carry a dataset
        object in the reference, which the strategy only reads labels from (`load_labels`, clip_searcher/utils.py:
        10-12); here the gold-label strategies get the loaded label tensor as `cache_labels`, while the logged
        parameters keep the configured dataset description (hydra_utils.instantiate_all)."""
        if "cache_dataset" not in cfg:
            return {}
        if self.cache_labels is None:
            raise ops._lib.SummerClipError(f"{cfg.get('_target_')} needs gold cache labels: set cache.labels_path")
        return {"cache_dataset": ("cache_labels", self.cache_labels)}

    def _gold_values(self, value_strategy, gold: torch.Tensor, n_classes: int):
        """cache.replace_outs_with_golds (image_attention.py:65-66): the value strategy is applied to
        one_hot(gold) instead of the logits.  HardCacheStrategy: argmax(one_hot(g)) = g, the labels themselves.
        Strategies that take a row gather (SoftmaxCacheStrategy): rows `gold` of the C x C identity ARE
        one_hot(gold), so softmax(s * one_hot(gold)) comes out of the same kernel without building [Nk, C]."""
        if isinstance(value_strategy, HardCacheStrategy):
            return GoldCacheValues(n_classes).transform(gold)
        eye = torch.eye(n_classes, dtype=torch.float32, device=gold.device)
        if isinstance(value_strategy, SoftmaxCacheStrategy):
            return value_strategy.transform(eye, idx=gold.long())
        return value_strategy.transform(eye[gold.long()])


def run_trainer(trainer_cls, cfg, run_dir=".") -> "ImageAttention":
    """utils/trainer.py:125-133 (seeds as in set_random_state :113-122)."""
    import random
    seed = int((cfg.get("meta") or {}).get("random_state", 42))
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    trainer = trainer_cls(cfg, run_dir)
    trainer.setup()
    trainer.train_loop()
    return trainer


def compose_from_argv(argv: tp.Optional[tp.Sequence[str]], default_name: str) -> Config:
    """`[CONFIG.yaml | --config-dir DIR [--config-name NAME]] [key=value | group=option ...]` — the hydra command line
    of the reference's entry points: a primary config with a defaults list is composed from its directory,
    `key=value` overrides are applied after composition."""
    argv = list(sys.argv[1:] if argv is None else argv)
    conf_dir = Path(__file__).resolve().parent.parent / "conf"
    name = default_name
    rest: tp.List[str] = []
    i = 0
    while i < len(argv):
        a = argv[i]
        if a == "--config-dir":
            conf_dir, i = Path(argv[i + 1]), i + 1
        elif a == "--config-name":
            name, i = argv[i + 1], i + 1
        elif a.endswith((".yaml", ".yml")) and "=" not in a:
            conf_dir, name = Path(a).resolve().parent, Path(a).stem
        else:
            rest.append(a)
        i += 1
    return compose(conf_dir, name, rest)


def run(argv: tp.Optional[tp.Sequence[str]] = None) -> ImageAttention:
    """The reference's `image_attention.py` command line (image_attention.py:123-125)."""
    cfg = compose_from_argv(argv, "image_attention")
    return run_trainer(ImageAttention, cfg, run_dir=(cfg.get("run_dir") or "."))


if __name__ == "__main__":
    run()
