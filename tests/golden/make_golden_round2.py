"""Generate tests/golden/round2.npz by running the REFERENCE's own code (imported from /root/reference, dev
container only) for the branches the first golden files do not cover:

  * cache.replace_outs_with_golds (image_attention.py:61-68) through ImageAttention.build_cache itself, followed by
    Hard / SoftmaxCacheStrategy on the replaced outs and the `@` / `Z + alpha O` lines;
  * the tail of Tip-Adapter's build_cache_model / pre_load_features (tip_adapter/utils.py:59-62, :84) on synthetic
    "encoder outputs", and TipAdapterTrainer.train_loop's three numbers (tip_adapter_imagenet.py:42-61);
  * the three *_per_gold / per_gold_class_random strategies driven through the sweep's record schema.

    python tests/golden/make_golden_round2.py [/root/reference]
"""
from __future__ import annotations

import contextlib
import io
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(HERE))

from make_golden import _install_stubs  # noqa: E402


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def main(ref_root: str = "/root/reference") -> None:
    _install_stubs()
    sys.path.insert(0, ref_root)
    from summer_clip.clip_searcher.cache_strategy import TopKStrategy
    from summer_clip.clip_searcher.cache_value_strategy import HardCacheStrategy, SoftmaxCacheStrategy
    from summer_clip.clip_searcher.cache_weights_strategy import TipAdapterWeightsStrategy
    from summer_clip.clip_searcher.image_attention import ImageAttention
    from summer_clip.clip_searcher.utils import compute_accuracy
    from summer_clip.tip_adapter import utils as tip_utils

    from oracle.clip_search_oracle import CLIP_SCALE, synthetic_banks

    torch.set_num_threads(4)
    out = {}

    # ------------------------------------------------------------------ replace_outs_with_golds
    banks = synthetic_banks(150, 333, 96, 37, seed=14, sigma=0.5, sigma_text=0.8, shared=3.0)
    Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    labels = banks["test_labels"]
    for n, v in banks.items():
        out[f"gr_{n}"] = v.numpy()
    stub = types.SimpleNamespace(cfg=_Cfg(run_saves=_Cfg(save_cache_inds=False), cache=_Cfg(replace_outs_with_golds=True)),
                                 cache_labels=banks["cache_labels"])
    with contextlib.redirect_stdout(io.StringIO()):
        Kc, outs_gold, info = ImageAttention.build_cache(stub, TopKStrategy(4), K, L)
    out["gr_info"] = np.array([info["cache_size"], info["acc1"], info["acc5"], info["acc1_replace"], info["acc5_replace"]], dtype=np.float64)
    out["gr_outs_replaced"] = outs_gold.float().numpy()
    qn = Q / Q.norm(dim=0, keepdim=True)
    Z = 100.0 * qn.t() @ T
    W = TipAdapterWeightsStrategy(5.5).transform(Q, Kc)
    for vi, vstrat in enumerate((HardCacheStrategy(), SoftmaxCacheStrategy(CLIP_SCALE, 0.1), SoftmaxCacheStrategy(CLIP_SCALE, 10.0))):
        V = vstrat.transform(outs_gold)                        # fp16 in, as build_cache returns it
        out[f"gr_values_{vi}"] = V.float().numpy()
        O = W @ V.to(W.dtype)
        out[f"gr_cache_logits_{vi}"] = O.numpy()
        out[f"gr_acc_{vi}"] = np.array([compute_accuracy(Z + O * a, labels) for a in (0.5, 1.0, 4.0)])

    # ------------------------------------------------------------------ Tip-Adapter entry point
    g = torch.Generator().manual_seed(15)
    n_cls, shots, dim, n_test, epochs = 11, 16, 64, 400, 2
    banks = synthetic_banks(n_test, shots * n_cls, dim, n_cls, seed=15, sigma=0.5, sigma_text=0.8, shared=3.0)
    base = banks["cache_image_features"].t().contiguous()                              # [Nk, D] un-normalised
    train_feats = torch.stack([base + 0.05 * torch.randn(base.shape, generator=g) for _ in range(epochs)]).half()
    train_labels = banks["cache_labels"].long()
    test_feats = banks["test_image_features"].t().contiguous().half()                  # [Nq, D] un-normalised
    test_labels = banks["test_labels"].long()
    clip_w = banks["text_features"].half()                                             # [D, C]
    # tip_adapter/utils.py:59-62, on the tensors the loop above them would have produced
    cache_keys = torch.cat([f.unsqueeze(0) for f in train_feats], dim=0).mean(dim=0)
    cache_keys /= cache_keys.norm(dim=-1, keepdim=True)
    cache_keys = cache_keys.permute(1, 0)
    cache_values = torch.nn.functional.one_hot(train_labels).half()
    # :84
    feats = test_feats.clone()
    feats /= feats.norm(dim=-1, keepdim=True)
    out.update(tip_train_features=train_feats.numpy(), tip_train_labels=train_labels.numpy(), tip_test_features=test_feats.numpy(),
               tip_test_labels=test_labels.numpy(), tip_clip_weights=clip_w.numpy(),
               tip_cache_keys=cache_keys.contiguous().numpy(), tip_cache_values=cache_values.numpy(), tip_test_f=feats.numpy())
    # tip_adapter_imagenet.py:44-54 (fp32 arithmetic on the stored fp16 tensors: half matmul is a CUDA-only path)
    f32, k32, v32, w32 = feats.float(), cache_keys.float(), cache_values.float(), clip_w.float()
    clip_logits = 100. * f32 @ w32
    acc_zs = tip_utils.cls_acc(clip_logits, test_labels)
    beta, alpha = 5.5, 1.0
    affinity = f32 @ k32
    cache_logits = ((-1) * (beta - beta * affinity)).exp() @ v32
    tip_logits = clip_logits + cache_logits * alpha
    acc_tip = tip_utils.cls_acc(tip_logits, test_labels)
    cfg = {"search_hp": True, "search_scale": [7, 3], "search_step": [20, 5]}
    with contextlib.redirect_stdout(io.StringIO()):
        best_beta, best_alpha = tip_utils.search_hp(cfg, k32, v32, f32, test_labels, w32)
    out.update(tip_acc_zero_shot=np.array(acc_zs), tip_acc=np.array(acc_tip), tip_best=np.array([best_beta, best_alpha]),
               tip_logits=tip_logits.numpy())
    np.savez_compressed(HERE / "round2.npz", **out)
    print("round2.npz", (HERE / "round2.npz").stat().st_size, "bytes", "| gold-replace info", out["gr_info"], "| tip", acc_zs, acc_tip,
          best_beta, best_alpha)


if __name__ == "__main__":
    main(*sys.argv[1:])
