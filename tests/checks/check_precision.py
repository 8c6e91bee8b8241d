"""Acceptance criteria of the north star for every operand type, on clustered synthetic banks: the CLIP-search logits
`Z + alpha * O` of the GPU path against the fp32 oracle on the same banks — max-abs difference after softmax over the
classes (bar: 2e-3) and argmax agreement (bar: 99.9 %).  One JSON line per (shape, operand type).

    python tests/checks/check_precision.py            # fp16, bf16 and the opt-in e4m3
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch  # noqa: E402

from oracle import clip_search_oracle as orc  # noqa: E402  (checker only)
from summer_clip_b200 import build as _build, ops  # noqa: E402
from summer_clip_b200.searcher import ClipSearcher  # noqa: E402


def main():
    _build.build_library()
    torch.set_num_threads(os.cpu_count() or 1)
    shapes = [(2048, 65536, 1024, 1000, 3), (2048, 19850, 1024, 397, 1), (1024, 32768, 768, 1000, 4), (1024, 16000, 512, 1000, 2)]
    betas, alphas = [1.0, 5.5, 11.5], [0.5, 1.0, 4.0]
    for nq, nk, dim, c, seed in shapes:
        banks = orc.synthetic_banks(nq, nk, dim, c, seed=seed)
        Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
        Z = orc.zero_shot_logits(Q, T)
        V = orc.hard_values(L)
        want = {b: orc.image_attention(Q, K, V, b, chunk=256) for b in betas}
        for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16), ("e4m3", ops.E4M3)):
            s = ClipSearcher("cuda", op_dtype=dt)
            s.set_text(T.cuda())
            s.set_cache(K.cuda(), L.cuda())
            res = s.search(Q.cuda(), betas, alphas, want_logits=True)
            worst_p, worst_rel, agree, total = 0.0, 0.0, 0, 0
            for r, b in zip(res, betas):
                o = r["cache_logits"].float().cpu()
                worst_rel = max(worst_rel, float((o - want[b]).abs().max() / want[b].abs().max()))
                for ai, a in enumerate(alphas):
                    ref = orc.searcher_logits(Z, want[b], a)
                    got = r["logits"][ai].float().cpu()
                    worst_p = max(worst_p, float((torch.softmax(got, 1) - torch.softmax(ref, 1)).abs().max()))
                    agree += int((got.argmax(1) == ref.argmax(1)).sum())
                    total += nq
            print(json.dumps({"shape": {"n_queries": nq, "n_keys": nk, "dim": dim, "n_classes": c}, "operands": name,
                              "softmax_max_abs": worst_p, "cache_logits_rel_max": worst_rel,
                              "argmax_agreement": agree / total, "points": len(betas) * len(alphas),
                              "pass": worst_p <= 2e-3 and agree / total >= 0.999}), flush=True)
            del s
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
