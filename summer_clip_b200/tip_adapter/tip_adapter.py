"""B200 mirror of summer_clip/tip_adapter/tip_adapter.py — the Tip-Adapter entry point for the non-ImageNet datasets.

The reference's two entry points share `train_loop` line for line (tip_adapter.py:57-76 ==
tip_adapter_imagenet.py:36-62): zero-shot accuracy, the Tip-Adapter head at (init_beta, init_alpha), `search_hp`.
They differ in the dataset plumbing above it (outside this path; `tip_adapter.py` also pre-loads a val split its
`train_loop` never reads) and in the config: `conf/tip_adapter.yaml` has search_scale [20, 10], init_beta 1,
init_alpha 3, dataset stanford_cars.  So this module is `TipAdapterTrainer` composed from `tip_adapter.yaml`.

    python -m summer_clip_b200.tip_adapter.tip_adapter [CONFIG.yaml] [key=value ...]
"""
from __future__ import annotations

import typing as tp

from .tip_adapter_imagenet import TipAdapterTrainer, run as _run, run_trainer  # noqa: F401  (the reference's names)


def run(argv: tp.Optional[tp.Sequence[str]] = None) -> TipAdapterTrainer:
    """tip_adapter.py:79-81 (`@hydra.main(config_name='tip_adapter')`)."""
    return _run(argv, config_name="tip_adapter")


if __name__ == "__main__":
    run()
