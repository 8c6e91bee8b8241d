"""B200 mirror of summer_clip/clip_searcher/save_image_outs.py — the producer of the `[N, C]` zero-shot logits bank
`image_outs = normalise(X)^T @ T` (save_image_outs.py:21-27; no x100) that `cache.image_outs_path` points at.

Same trainer shape and config keys (conf/save_image_outs.yaml); the text classifier, which the reference builds
with the CLIP text tower (`eval_clip.zeroshot_classifier`, outside this path), is read from
`data.text_features_path` (T [D, C]).  The product is the split-fp16 tensor-core GEMM (`ops.zero_shot_logits`,
fp32-accurate), cast to the feature bank's dtype like the reference's result, rows computed in chunks so that the
fp32 tile never exceeds `data.rows_per_chunk`.  For pseudo-label selection the bank need not exist at all
(`cache.image_outs_path: null` -> `LazyLogitsBank`); this entry is for the users that read it: soft cache values,
the notebooks, other tools.

    python -m summer_clip_b200.clip_searcher.save_image_outs [CONFIG.yaml] [key=value ...]
"""
from __future__ import annotations

import typing as tp
from pathlib import Path

import torch

from .. import ops
from ..utils.config import Config
from .image_attention import _load_tensor, compose_from_argv, run_trainer


class SaveImageOuts:
    def __init__(self, cfg: tp.Mapping, run_dir: tp.Union[str, Path] = ".") -> None:
        self.cfg = cfg if isinstance(cfg, Config) else Config(cfg)
        self.run_dir = Path(run_dir)

    def setup(self) -> None:
        self.device = ops.require_cuda_device((self.cfg.get("meta") or {}).get("device"), "save_image_outs")
        data = self.cfg["data"]
        if not data.get("text_features_path"):
            raise ops._lib.SummerClipError("save_image_outs needs data.text_features_path (the text classifier T [D, C]; "
                                           "the CLIP text tower is outside this path)")
        self.image_features = _load_tensor(data["image_features_path"], self.device)          # [D, N]
        self.text_features = _load_tensor(data["text_features_path"], self.device)            # [D, C]

    @torch.no_grad()
    def train_loop(self) -> torch.Tensor:
        X, T = self.image_features, self.text_features
        n, c = X.shape[1], T.shape[1]
        chunk = int(self.cfg["data"].get("rows_per_chunk") or (1 << 20))
        t_split = ops.text_split(T.float().contiguous())
        out = torch.empty((n, c), dtype=X.dtype, device=self.device)
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            out[lo:hi] = ops.zero_shot_logits(X[:, lo:hi], True, T, scale=1.0, normalize=True, t_split=t_split,
                                                two_pass=True).to(X.dtype)          # fp16 banks: raw rows + 1/norm
        path = Path(self.cfg["data"]["output_image_outs"])
        if not path.is_absolute():
            path = self.run_dir / path
        path.parent.mkdir(parents=True, exist_ok=True)
        torch.save(out.cpu(), path)
        self.output_path, self.image_outs = path, out
        return out


def run(argv: tp.Optional[tp.Sequence[str]] = None) -> SaveImageOuts:
    """The reference's `save_image_outs.py` command line (save_image_outs.py:30-32)."""
    cfg = compose_from_argv(argv, "save_image_outs")
    return run_trainer(SaveImageOuts, cfg, run_dir=(cfg.get("run_dir") or "."))


if __name__ == "__main__":
    run()
