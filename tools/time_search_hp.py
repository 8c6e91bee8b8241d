"""Where search_hp's time goes at the ImageNet 16-shot shape (50 000 queries x 16 000 keys x 1024-d x 1000 classes):
multi-beta attention launches, the alpha epilogue, and the whole 200 x 20 grid.  One JSON line."""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from summer_clip_b200 import build as _build, ops  # noqa: E402
from summer_clip_b200.tip_adapter import utils as tip_utils  # noqa: E402


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    _build.build_library()
    dev = torch.device("cuda")
    nq, nk, dim, c = 50000, 16000, 1024, 1000
    g = torch.Generator(device=dev).manual_seed(2)
    protos = torch.nn.functional.normalize(torch.randn(c, dim, generator=g, device=dev), dim=1)
    yk = torch.arange(nk, device=dev) % c
    keys = torch.nn.functional.normalize(protos[yk] + torch.randn(nk, dim, generator=g, device=dev) / dim ** 0.5, dim=1).half().t()
    vals = torch.nn.functional.one_hot(yk, c).half()
    yq = torch.randint(0, c, (nq,), generator=g, device=dev)
    feats = torch.nn.functional.normalize(protos[yq] + torch.randn(nq, dim, generator=g, device=dev) / dim ** 0.5, dim=1).half()
    clip_w = torch.nn.functional.normalize(protos + 2.0 / dim ** 0.5 * torch.randn(c, dim, generator=g, device=dev), dim=1).t().contiguous().half()
    head = tip_utils.TipAdapterHead(keys, vals, feats, clip_w)
    labels = yq.int()
    alphas = [0.1 + 0.15 * i for i in range(20)]
    out = {"attn_1beta_ms": timed(lambda: head.cache_logits(5.5))}
    for nb in (2, 4, 16):
        out[f"attn_{nb}beta_ms"] = timed(lambda: head.cache_logits_many([0.5 + 0.3 * i for i in range(nb)]))
    o = head.cache_logits(5.5)
    out["epilogue_20alpha_ms"] = timed(lambda: ops.epilogue(head.clip_logits, o, alphas, labels=labels, want_pred=False))
    out["epilogue_1alpha_ms"] = timed(lambda: ops.epilogue(head.clip_logits, o, alphas[:1], labels=labels, want_pred=False))
    out["top1_counts_many_16beta_ms"] = timed(lambda: head.top1_counts_many([0.5 + 0.3 * i for i in range(16)], alphas, labels), iters=3)
    cfg = {"search_hp": True, "search_scale": [7, 3], "search_step": [200, 20]}
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            tip_utils.search_hp(cfg, keys, vals, feats, labels, clip_w)
        torch.cuda.synchronize()
        out[f"search_hp_s_rep{rep}"] = time.perf_counter() - t0
    print(json.dumps(out))


if __name__ == "__main__":
    main()
