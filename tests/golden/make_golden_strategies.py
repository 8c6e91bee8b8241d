"""Golden outputs of the reference's REMAINING cache-selection strategies (cache_strategy.py:35-45,84-153):
threshold, per-gold-label top-k (raw and softmax), and the three random samplers (host numpy RNG stream).
Same recipe as make_golden.py: import the reference from /root/reference (dev container only), run it on a small
seeded bank, commit the arrays; nothing at test time reads /root/reference.

    python tests/golden/make_golden_strategies.py [/root/reference]
"""
from __future__ import annotations

import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import _install_stubs  # noqa: E402


def main(ref_root: str = "/root/reference") -> None:
    _install_stubs()
    sys.path.insert(0, ref_root)
    from summer_clip.clip_searcher import cache_strategy as ref

    g = torch.Generator().manual_seed(77)
    n, c = 900, 19
    outs = 0.25 + 0.05 * torch.randn(n, c, generator=g)                     # tie-free fp32 logits
    gold = torch.randint(0, c - 2, (n,), generator=g).int()                 # two classes have no gold sample
    dataset = [(None, int(v)) for v in gold]
    feats = torch.empty(1, n)
    out = {"image_outs": outs.numpy(), "gold_labels": gold.numpy()}
    with contextlib.redirect_stdout(io.StringIO()):
        for thr in (0.06, 0.07):
            out[f"threshold_softmax_{thr}"] = ref.ThresholdStrategy(thr, True).select(feats, outs).numpy()
        for thr in (0.3, 0.36):
            out[f"threshold_raw_{thr}"] = ref.ThresholdStrategy(thr, False).select(feats, outs).numpy()
        for k in (1, 4, 80):
            out[f"topk_per_gold_{k}"] = ref.TopKPerGoldStrategy(k, dataset).select(feats, outs).numpy()
            out[f"topk_per_gold_prob_{k}"] = ref.TopKPerGoldProbStrategy(k, dataset, 100.00000762939453).select(feats, outs).numpy()
        for k in (1, 3, 60):
            np.random.seed(42)
            out[f"global_random_{k}"] = ref.GlobalRandomSampleStrategy(k).select(feats, outs).numpy()
            np.random.seed(42)
            out[f"per_pred_random_{k}"] = ref.PerPredClassRandomSampleStrategy(k).select(feats, outs).numpy()
            np.random.seed(42)
            out[f"per_gold_random_{k}"] = ref.PerGoldClassRandomSampleStrategy(k, dataset).select(feats, outs).numpy()
    np.savez_compressed(HERE / "strategies.npz", **out)
    print("wrote", HERE / "strategies.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(*sys.argv[1:])
