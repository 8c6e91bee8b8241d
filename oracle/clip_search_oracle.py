"""CPU oracle for the CLIP-search hot path — TEST INFRASTRUCTURE ONLY.

A torch-CPU fp32 / numpy restatement of what myrachins/summer-clip computes on the path
BASELINE.json names.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module, and only as the checker or the timed CPU
baseline — never the product path (summer_clip_b200/ has no import of it and fails loudly
without its CUDA library).

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md §8c).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: in the
dev container tests/golden/make_golden.py imports the reference's own modules from
/root/reference (stubbing the absent clip/hydra/omegaconf imports), runs them on seeded
synthetic banks and commits inputs + outputs as tests/golden/*.npz; tests/test_oracle_golden.py
checks every function below against those files.

Deviations from the reference, all documented in DESIGN.md:
  * `W @ V.to(W.dtype)`: HardCacheStrategy returns fp16 whatever the bank dtype and the reference's
    matmul raises on fp32 banks (cache_value_strategy.py:16 vs image_attention.py:109); one-hot
    values are exact in fp16 so the cast is lossless.
  * torch.topk breaks ties arbitrarily; the oracle fixes the order (confidence descending, then row
    index ascending).  On tie-free inputs it equals the reference (golden-checked).
  * image_attention() chunks the queries so that the [Nq, Nk] matrix is bounded; numerically
    identical per row.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

CLIP_SCALE = 100.00000762939453  # conf/cache_strategy/topk_prob.yaml:3, conf/cache_value_strategy/softmax_cache.yaml:2


# --------------------------------------------------------------------------- weights
def normalize_columns(x: torch.Tensor) -> torch.Tensor:
    """cache_weights_strategy.py:19-20 — x / x.norm(dim=0, keepdim=True) for a [D, N] bank."""
    return x / x.norm(dim=0, keepdim=True)


def tip_weights(q_norm: torch.Tensor, k_norm: torch.Tensor, beta: float) -> torch.Tensor:
    """cache_weights_strategy.py:33-36 — A = Q^T K ; W = exp(-1 * beta * (1 - A)).  Inputs [D, N]."""
    a = q_norm.t() @ k_norm
    return (-1 * beta * (1 - a)).exp()


def tip_weights_tipadapter(q_rows: torch.Tensor, keys: torch.Tensor, beta: float) -> torch.Tensor:
    """tip_adapter/utils.py:114-116 — affinity = features @ cache_keys ; exp(-(beta - beta*affinity)).
    q_rows [Nq, D] row-normalised, keys [D, Nk]."""
    affinity = q_rows @ keys
    return ((-1) * (beta - beta * affinity)).exp()


# --------------------------------------------------------------------------- values
def hard_values(cache_outs: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """cache_value_strategy.py:14-17 — one_hot(argmax_c L); `.half()` replaced by `dtype` (lossless)."""
    _, labels_ids = cache_outs.max(dim=1)
    return torch.nn.functional.one_hot(labels_ids, num_classes=cache_outs.shape[1]).to(dtype)


def softmax_values(cache_outs: torch.Tensor, clip_scale: float, scale: float) -> torch.Tensor:
    """cache_value_strategy.py:26-28 — softmax(clip_scale * scale * L, dim=1)."""
    return torch.softmax(clip_scale * scale * cache_outs, dim=1)


def onehot_values(labels: torch.Tensor, n_classes: int, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """image_attention.py:65-66 / tip_adapter/utils.py:62 — one_hot(gold labels)."""
    return torch.nn.functional.one_hot(labels.long(), num_classes=n_classes).to(dtype)


# --------------------------------------------------------------------------- selection
def row_confidence(image_outs: torch.Tensor, prob: bool, scale: float = CLIP_SCALE) -> Tuple[torch.Tensor, torch.Tensor]:
    """cache_strategy.py:68 (raw) and :80 + :68 (softmax(L * scale) first): (confidence, label)."""
    x = image_outs.float()
    if prob:
        x = torch.softmax(x * scale, dim=1)
    conf, label = x.max(dim=1)
    return conf, label


def select_topk_per_label(image_labels: np.ndarray, image_logits: np.ndarray, topk: int) -> np.ndarray:
    """cache_strategy.py:48-59 — for label in unique(labels) ascending: the min(topk, n) most confident
    members, most confident first.  Ties: smaller row index first (the reference's torch.topk is
    unspecified there)."""
    image_labels = np.asarray(image_labels)
    image_logits = np.asarray(image_logits, dtype=np.float32)
    picked: List[np.ndarray] = []
    for label in np.unique(image_labels):
        members = np.nonzero(image_labels == label)[0]
        order = np.argsort(-image_logits[members], kind="stable")[: min(topk, members.shape[0])]
        picked.append(members[order])
    return np.concatenate(picked).astype(np.int64) if picked else np.zeros(0, np.int64)


def topk_select(image_outs: torch.Tensor, topk: int) -> np.ndarray:
    """TopKStrategy.select, cache_strategy.py:67-70."""
    conf, label = row_confidence(image_outs, prob=False)
    return select_topk_per_label(label.numpy(), conf.numpy(), topk)


def topk_prob_select(image_outs: torch.Tensor, topk: int, scale: float = CLIP_SCALE) -> np.ndarray:
    """TopKProbStrategy.select, cache_strategy.py:79-81."""
    conf, label = row_confidence(image_outs, prob=True, scale=scale)
    return select_topk_per_label(label.numpy(), conf.numpy(), topk)


def all_logits_select(image_outs: torch.Tensor) -> np.ndarray:
    """AllLogitsStrategy.select, cache_strategy.py:30-32."""
    return np.arange(image_outs.shape[0], dtype=np.int64)


# --------------------------------------------------------------------------- logits / accuracy
def zero_shot_logits(test_image_features: torch.Tensor, text_features: torch.Tensor) -> torch.Tensor:
    """image_attention.py:80-83 — 100 * normalise_cols(Q)^T @ T."""
    norm = test_image_features / test_image_features.norm(dim=0, keepdim=True)
    return 100.0 * norm.t() @ text_features


def image_outs(image_features: torch.Tensor, text_features: torch.Tensor) -> torch.Tensor:
    """clip_searcher/save_image_outs.py:21-25 — the logits bank normalise_cols(X)^T @ T (no x100), [N, C]."""
    norm = image_features / image_features.norm(dim=0, keepdim=True)
    return norm.t() @ text_features


def accuracy_counts(output: torch.Tensor, target: torch.Tensor, topk: Sequence[int] = (1, 5)) -> List[float]:
    """clip_adapter/train_adapter.py:156-159 — number of rows whose target is within the top-k."""
    pred = output.topk(max(topk), 1, True, True)[1].t()
    correct = pred.eq(target.view(1, -1).expand_as(pred))
    return [float(correct[:k].reshape(-1).float().sum().item()) for k in topk]


def compute_accuracy(outputs: torch.Tensor, target: torch.Tensor, topk: Sequence[int] = (1, 5)) -> List[float]:
    """clip_searcher/utils.py:15-21 — 100 * correct / N per k."""
    return [100.0 * acc / target.shape[0] for acc in accuracy_counts(outputs, target, topk)]


def cls_acc(output: torch.Tensor, target: torch.Tensor, topk: int = 1) -> float:
    """tip_adapter/utils.py:10-15."""
    return 100.0 * accuracy_counts(output, target, (topk,))[0] / target.shape[0]


# --------------------------------------------------------------------------- the fused path
def image_attention(test_image_features: torch.Tensor, cache_image_features: torch.Tensor,
                    cache_values: torch.Tensor, beta: float, chunk: int = 2048) -> torch.Tensor:
    """cache_weights_strategy.py:18-21,33-36 + image_attention.py:109 — O = exp(-beta(1 - Q^T K)) @ V with
    un-normalised [D, N] banks; queries processed `chunk` at a time (bounded memory)."""
    qn = normalize_columns(test_image_features.float())
    kn = normalize_columns(cache_image_features.float())
    v = cache_values.float()
    out = torch.empty((qn.shape[1], v.shape[1]), dtype=torch.float32)
    for s in range(0, qn.shape[1], chunk):
        w = tip_weights(qn[:, s:s + chunk], kn, beta)
        out[s:s + chunk] = w @ v.to(w.dtype)
    return out


def searcher_logits(clip_logits: torch.Tensor, cache_logits: torch.Tensor, alpha: float) -> torch.Tensor:
    """image_attention.py:111 — clip_logits + cache_logits * alpha."""
    return clip_logits + cache_logits * alpha


def softmax_attention(test_image_features: torch.Tensor, cache_image_features: torch.Tensor,
                      cache_values: torch.Tensor, tau: float, chunk: int = 2048) -> torch.Tensor:
    """North-star extension with NO reference implementation (parity unpinned by the reference):
    softmax(tau * Q^T K, dim=1) @ V."""
    qn = normalize_columns(test_image_features.float())
    kn = normalize_columns(cache_image_features.float())
    v = cache_values.float()
    out = torch.empty((qn.shape[1], v.shape[1]), dtype=torch.float32)
    for s in range(0, qn.shape[1], chunk):
        out[s:s + chunk] = torch.softmax(tau * (qn[:, s:s + chunk].t() @ kn), dim=1) @ v
    return out


# --------------------------------------------------------------------------- Tip-Adapter
def tip_head(features: torch.Tensor, cache_keys: torch.Tensor, cache_values: torch.Tensor,
             clip_weights: torch.Tensor, beta: float, alpha: float) -> torch.Tensor:
    """tip_adapter/tip_adapter.py:58-68 — Z = 100 Q T ; A = Q K ; out = Z + alpha * exp(-(beta - beta A)) V."""
    clip_logits = 100.0 * features @ clip_weights
    w = tip_weights_tipadapter(features, cache_keys, beta)
    cache_logits = w @ cache_values.to(w.dtype)
    return clip_logits + cache_logits * alpha


def tip_cache_keys(train_features: torch.Tensor) -> torch.Tensor:
    """tip_adapter/utils.py:59-61 — mean over the augment epochs [E, Nk, D], row-normalise, permute -> [D, Nk]
    (in the dtype given: the reference runs these lines on fp16 CUDA tensors)."""
    cache_keys = train_features.mean(dim=0)
    cache_keys = cache_keys / cache_keys.norm(dim=-1, keepdim=True)
    return cache_keys.permute(1, 0)


def tip_normalize_rows(features: torch.Tensor) -> torch.Tensor:
    """tip_adapter/utils.py:84 — image_features /= image_features.norm(dim=-1, keepdim=True)."""
    return features / features.norm(dim=-1, keepdim=True)


def golds_as_outs(cache_labels: torch.Tensor, n_classes: int) -> torch.Tensor:
    """image_attention.py:65-66 — cache.replace_outs_with_golds: the selected logits are replaced by
    one_hot(gold).half(); the value strategy is then applied to THAT matrix."""
    return torch.nn.functional.one_hot(cache_labels.long(), num_classes=n_classes).half()


def search_grid(search_scale: Sequence[float], search_step: Sequence[int]) -> Tuple[List[float], List[float]]:
    """tip_adapter/utils.py:103-104 — the beta / alpha grids of search_hp."""
    beta_list = [i * (search_scale[0] - 0.1) / search_step[0] + 0.1 for i in range(search_step[0])]
    alpha_list = [i * (search_scale[1] - 0.1) / search_step[1] + 0.1 for i in range(search_step[1])]
    return beta_list, alpha_list


def search_hp(search_scale, search_step, cache_keys, cache_values, features, labels, clip_weights):
    """tip_adapter/utils.py:99-129 — exhaustive (beta, alpha) search; strict `>` keeps the first best.
    Returns (best_beta, best_alpha, best_acc).  GEMM-1 is hoisted out of the loops (identical values)."""
    beta_list, alpha_list = search_grid(search_scale, search_step)
    best_acc, best_beta, best_alpha = 0.0, 0, 0
    affinity = features @ cache_keys
    clip_logits = 100.0 * features @ clip_weights
    for beta in beta_list:
        cache_logits = ((-1) * (beta - beta * affinity)).exp() @ cache_values.to(affinity.dtype)
        for alpha in alpha_list:
            acc = cls_acc(clip_logits + cache_logits * alpha, labels)
            if acc > best_acc:
                best_acc, best_beta, best_alpha = acc, beta, alpha
    return best_beta, best_alpha, best_acc


# --------------------------------------------------------------------------- synthetic banks
def synthetic_banks(n_query: int, n_key: int, dim: int, n_classes: int, seed: int, sigma: float = 1.0,
                    sigma_text: float = 3.0, shared: float = 1.0, dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """clustered banks stored feature-major ([D, N]) like save_features.py:36.
    Class prototypes mu_c = normalise(shared * u0 + g_c / sqrt(D)) share a common direction u0 (CLIP
    embeddings of different classes are far from orthogonal; with shared = 1 prototypes have cosine
    ~0.5).  Samples x = mu_y + sigma / sqrt(D) * eps are left UN-normalised (random positive scale);
    the text classifier T[:, c] = normalise(mu_c + sigma_text / sqrt(D) * eps) is a noisy copy of the
    prototypes so zero-shot accuracy is well below 100 %; L = K_norm^T T (save_image_outs.py:25)."""
    g = torch.Generator().manual_seed(seed)
    u0 = torch.nn.functional.normalize(torch.randn(dim, generator=g), dim=0)
    protos = torch.nn.functional.normalize(shared * u0 + torch.randn(n_classes, dim, generator=g) / dim ** 0.5, dim=1)
    yq = torch.randint(0, n_classes, (n_query,), generator=g)
    yk = torch.randint(0, n_classes, (n_key,), generator=g)
    s = sigma / dim ** 0.5
    q = (protos[yq] + s * torch.randn(n_query, dim, generator=g)) * (0.5 + torch.rand(n_query, 1, generator=g))
    k = (protos[yk] + s * torch.randn(n_key, dim, generator=g)) * (0.5 + torch.rand(n_key, 1, generator=g))
    t = torch.nn.functional.normalize(protos + sigma_text / dim ** 0.5 * torch.randn(n_classes, dim, generator=g), dim=1)
    q_bank = q.t().contiguous().to(dtype)          # [D, Nq]
    k_bank = k.t().contiguous().to(dtype)          # [D, Nk]
    text = t.t().contiguous().to(dtype)            # [D, C]
    kn = k_bank.float() / k_bank.float().norm(dim=0, keepdim=True)
    outs = (kn.t() @ text.float()).to(dtype)       # save_image_outs.py:25 (no x100)
    return {"test_image_features": q_bank, "cache_image_features": k_bank, "text_features": text,
            "cache_image_outs": outs, "test_labels": yq.to(torch.int32), "cache_labels": yk.to(torch.int32)}
