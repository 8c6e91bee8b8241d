"""BASELINE.json config 1 as the reference runs it: the default image_attention sweep (conf/image_attention.yaml)
on a SUN397-shaped synthetic problem — 19 850 test x 19 850 unlabeled-train features, 1024-d, 397 classes; 25
caches or more (TopK / TopKProb / per-class random / global random / per-gold for k in 1..32, AllLogits) x 8 beta x
7 alpha accuracy records — through `ImageAttention` (one gather+normalise per cache, one attention launch per beta, one
epilogue launch for all alphas).  Prints one JSON line: wall time of setup and of the sweep.

    python tools/bench_sweep.py
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from bench import make_banks  # noqa: E402
from summer_clip_b200 import build as _build  # noqa: E402
from summer_clip_b200.clip_searcher.image_attention import ImageAttention  # noqa: E402
from summer_clip_b200.utils.config import load_config  # noqa: E402


def main():
    _build.build_library()
    dev = torch.device("cuda")
    nq, nk, dim, c = 19850, 19850, 1024, 397
    q_bank, k_bank, outs, text, labels = make_banks(torch, nq, 0, nk, dim, c, seed=1, device=dev)
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        paths = {"q": tmp / "test_features.pt", "k": tmp / "train_features.pt", "l": tmp / "train_outs.pt",
                 "t": tmp / "text.pt", "y": tmp / "labels.pt", "ky": tmp / "train_labels.pt"}
        torch.save(q_bank.cpu(), paths["q"])
        torch.save(k_bank.cpu(), paths["k"])
        torch.save(outs.cpu(), paths["l"])
        torch.save(text.cpu(), paths["t"])
        torch.save(labels.cpu(), paths["y"])
        # gold labels of the train bank (the per-gold strategies and the cache-quality records need them): the
        # zero-shot prediction, a quarter of them reassigned at random
        g = torch.Generator(device=dev).manual_seed(7)
        gold = outs.float().argmax(dim=1)
        flip = torch.rand(nk, generator=g, device=dev) < 0.25
        gold = torch.where(flip, torch.randint(0, c, (nk,), generator=g, device=dev), gold)
        torch.save(gold.cpu(), paths["ky"])
        del q_bank, k_bank, outs
        conf = Path(__file__).resolve().parent.parent / "summer_clip_b200" / "conf" / "image_attention.yaml"
        cfg = load_config(conf, {"data": {"image_features_path": str(paths["q"]), "text_features_path": str(paths["t"]),
                                          "labels_path": str(paths["y"])},
                                 "cache": {"image_features_path": str(paths["k"]), "image_outs_path": str(paths["l"]),
                                           "labels_path": str(paths["ky"])}})
        times = []
        for rep in range(2):                                       # second repetition: warm allocator / page cache
            trainer = ImageAttention(cfg, tmp / f"run{rep}")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            trainer.setup()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            trainer.train_loop()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            times.append((t1 - t0, t2 - t1))
        records = [json.loads(l) for l in (tmp / "run1" / "image_attention.log").read_text().splitlines()]
        results = [r for r in records if r.get("type") == "searcher_result"]
        caches = [r for r in records if r.get("type") == "cache_info"]
        best = max(results, key=lambda r: r["acc1"])
        zs = next(r for r in records if r.get("type") == "zero_shot")
    print(json.dumps({"config": "cfg1_sun397_default_sweep", "n_queries": nq, "n_train": nk, "dim": dim, "n_classes": c,
                      "caches": len(caches), "records": len(results), "setup_s": times[1][0], "sweep_s": times[1][1],
                      "first_run_sweep_s": times[0][1], "ms_per_cache_beta": 1e3 * times[1][1] / (len(caches) * 8),
                      "zero_shot_acc1": zs["acc1"], "best_acc1": best["acc1"],
                      "best": {k: best[k] for k in ("cache_strategy", "cache_weights_strategy", "alpha")}}))


if __name__ == "__main__":
    main()
