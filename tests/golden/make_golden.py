"""Generate tests/golden/*.npz by running the REFERENCE's own code (imported from /root/reference,
dev container only) on small seeded synthetic banks.  The files are committed; nothing at test or
bench time reads /root/reference.

    python tests/golden/make_golden.py [/root/reference]

The reference modules on the path import `clip`, `hydra` and `omegaconf`, none of which is
installed here and none of which the numerical code uses; three stub modules stand in for them.
"""
from __future__ import annotations

import io
import contextlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))


def _install_stubs() -> None:
    clip = types.ModuleType("clip")
    hydra = types.ModuleType("hydra")
    hydra.main = lambda *a, **kw: (lambda f: f)
    hydra_utils = types.ModuleType("hydra.utils")
    hydra_utils.instantiate = lambda *a, **kw: None
    hydra.utils = hydra_utils
    omegaconf = types.ModuleType("omegaconf")
    for name in ("DictConfig", "ListConfig", "OmegaConf"):
        setattr(omegaconf, name, type(name, (), {}))
    omegaconf.open_dict = contextlib.nullcontext
    wandb = types.ModuleType("wandb")
    for name, mod in (("clip", clip), ("hydra", hydra), ("hydra.utils", hydra_utils), ("omegaconf", omegaconf)):
        sys.modules.setdefault(name, mod)
    try:
        import wandb as _w  # noqa: F401  (installed; the reference only uses it in setup_logger)
    except Exception:
        sys.modules.setdefault("wandb", wandb)


def main(ref_root: str = "/root/reference") -> None:
    _install_stubs()
    sys.path.insert(0, ref_root)
    from summer_clip.clip_searcher.cache_strategy import TopKProbStrategy, TopKStrategy, AllLogitsStrategy
    from summer_clip.clip_searcher.cache_value_strategy import HardCacheStrategy, SoftmaxCacheStrategy
    from summer_clip.clip_searcher.cache_weights_strategy import TipAdapterWeightsStrategy
    from summer_clip.clip_searcher.utils import compute_accuracy
    from summer_clip.tip_adapter import utils as tip_utils

    from oracle.clip_search_oracle import CLIP_SCALE, synthetic_banks

    torch.set_num_threads(4)

    # ------------------------------------------------------------------ selection goldens
    # 700 samples over 23 classes with 3 classes never predicted and rare classes with < k members.
    # Logits are small (std 0.02, like CLIP cosine gaps) so softmax(100 * L) is not saturated: a
    # saturated softmax yields exact ties at 1.0, where the reference's torch.topk order is arbitrary.
    banks = synthetic_banks(64, 700, 48, 23, seed=11)
    g = torch.Generator().manual_seed(110)
    outs = 0.25 + 0.02 * torch.randn(700, 23, generator=g)
    outs[:, [4, 9, 17]] -= 1.0                     # never the argmax -> empty predicted classes
    outs[:, [20, 21, 22]] -= 0.045                 # rarely the argmax -> classes with < k members
    sel = {"image_outs": outs.numpy()}
    with contextlib.redirect_stdout(io.StringIO()):
        for k in (1, 4, 16, 64):
            sel[f"topk_{k}"] = TopKStrategy(k).select(banks["cache_image_features"], outs).numpy()
            sel[f"topk_prob_{k}"] = TopKProbStrategy(k, CLIP_SCALE).select(banks["cache_image_features"], outs).numpy()
    sel["all_logits"] = AllLogitsStrategy().select(banks["cache_image_features"], outs).numpy()
    conf_raw, lab = outs.max(dim=1)
    conf_prob, lab2 = torch.softmax(outs * CLIP_SCALE, dim=1).max(dim=1)
    sel.update(conf_raw=conf_raw.numpy(), label=lab.numpy(), conf_prob=conf_prob.numpy())
    # the goldens are only meaningful if the reference's arbitrary tie-breaking was not exercised
    for name, conf in (("raw", conf_raw), ("prob", conf_prob)):
        for c in lab.unique():
            v = conf[lab == c]
            assert v.unique().numel() == v.numel(), f"tie inside class {int(c)} ({name}); change the seed"
    np.savez_compressed(HERE / "selection.npz", **sel)

    # ------------------------------------------------------------------ image-attention goldens
    banks = synthetic_banks(150, 333, 96, 37, seed=12, sigma=0.5, sigma_text=0.8, shared=3.0)
    Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    labels = banks["test_labels"]
    att = {n: v.numpy() for n, v in banks.items()}
    qn = Q / Q.norm(dim=0, keepdim=True)
    Z = 100.0 * qn.t() @ T                                         # image_attention.py:80-83
    att["clip_logits"] = Z.numpy()
    with contextlib.redirect_stdout(io.StringIO()):
        idx = TopKProbStrategy(4, CLIP_SCALE).select(K, L)
    att["cache_idx"] = idx.numpy()
    Kc, Lc = K[:, idx], L[idx]                                      # image_attention.py:54-55
    betas = [0.1, 1.0, 5.5, 11.5]
    alphas = [0.0, 0.1, 0.5, 1.0, 2.0, 3.0, 4.0]
    att["betas"], att["alphas"] = np.array(betas), np.array(alphas)
    for vi, vstrat in enumerate((HardCacheStrategy(), SoftmaxCacheStrategy(CLIP_SCALE, 0.1), SoftmaxCacheStrategy(CLIP_SCALE, 10.0))):
        V = vstrat.transform(Lc)
        att[f"values_{vi}"] = V.float().numpy()
        for bi, beta in enumerate(betas):
            W = TipAdapterWeightsStrategy(beta).transform(Q, Kc)   # un-normalised inputs, like train_loop
            if vi == 0 and bi == 2:
                att["weights_b5.5"] = W.numpy()
            O = W @ V.to(W.dtype)                                   # documented dtype deviation
            att[f"cache_logits_v{vi}_b{bi}"] = O.numpy()
            accs = []
            for alpha in alphas:
                out = Z + O * alpha                                 # image_attention.py:111
                accs.append(compute_accuracy(out, labels))
            att[f"acc_v{vi}_b{bi}"] = np.array(accs)
    # all-logits cache (Nk = N)
    W = TipAdapterWeightsStrategy(5.5).transform(Q, K)
    att["cache_logits_all_hard_b5.5"] = (W @ HardCacheStrategy().transform(L).to(W.dtype)).numpy()
    att["acc_zero_shot"] = np.array(compute_accuracy(Z, labels))
    np.savez_compressed(HERE / "image_attention.npz", **att)

    # ------------------------------------------------------------------ Tip-Adapter goldens
    banks = synthetic_banks(400, 16 * 11, 64, 11, seed=13, sigma=0.5, sigma_text=0.8, shared=3.0)
    feats = torch.nn.functional.normalize(banks["test_image_features"].t(), dim=1).contiguous()   # [Nq, D] rows
    keys = torch.nn.functional.normalize(banks["cache_image_features"].t(), dim=1).t()           # [D, Nk] VIEW (utils.py:61)
    vals = torch.nn.functional.one_hot(banks["cache_labels"].long(), 11).half()                   # utils.py:62
    clip_w = banks["text_features"]
    tl = banks["test_labels"].long()
    cfg = {"search_hp": True, "search_scale": [7, 3], "search_step": [20, 5]}
    # the reference multiplies an fp32 weight matrix by the fp16 one-hot: cast like the oracle does
    with contextlib.redirect_stdout(io.StringIO()):
        best_beta, best_alpha = tip_utils.search_hp(cfg, keys, vals.float(), feats, tl, clip_w)
    tip = dict(features=feats.numpy(), cache_keys=keys.contiguous().numpy(), cache_labels=banks["cache_labels"].numpy(),
               clip_weights=clip_w.numpy(), test_labels=tl.numpy(), search_scale=np.array([7, 3]),
               search_step=np.array([20, 5]), best_beta=np.array(best_beta), best_alpha=np.array(best_alpha))
    beta, alpha = 5.5, 1.0                                           # conf/tip_adapter_imagenet.yaml:25-26
    clip_logits = 100.0 * feats @ clip_w
    affinity = feats @ keys
    cache_logits = ((-1) * (beta - beta * affinity)).exp() @ vals.float()
    tip_logits = clip_logits + cache_logits * alpha
    tip.update(tip_logits=tip_logits.numpy(), acc_zero_shot=np.array(tip_utils.cls_acc(clip_logits, tl)),
               acc_tip=np.array(tip_utils.cls_acc(tip_logits, tl)))
    np.savez_compressed(HERE / "tip_adapter.npz", **tip)
    for f in ("selection.npz", "image_attention.npz", "tip_adapter.npz"):
        print(f, (HERE / f).stat().st_size, "bytes")


if __name__ == "__main__":
    main(*sys.argv[1:])
