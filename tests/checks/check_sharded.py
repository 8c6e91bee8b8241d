"""N-rank NCCL check of the sharded ClipSearcher.search (key shards with the query-blocked overlap pipeline, query
shards, per-rank query slices, the temperature-softmax (m, l, O) merge) against the single-rank result and the
oracle.  Run under torchrun:  python -m torch.distributed.run --nproc-per-node 2 tests/checks/check_sharded.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist

from oracle import clip_search_oracle as orc
from summer_clip_b200.searcher import ClipSearcher, query_slice

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ok = True
BETAS = [5.5, 0.1, 1.0, 3.5, 11.5]                 # one-hot banks: a 4-beta launch + a 1-beta launch
ALPHAS = [0.5, 2.0]


def compare(gots, refs, what):
    """every rank: same pred / counts as the single-rank search; logits of its own pieces equal."""
    e, same_pred, same_cnt, rows = 0.0, True, True, 0
    for got, r in zip(gots, refs):
        for lo, hi, logits, _, _ in got["pieces"]:
            e = max(e, (logits - r["logits"][:, lo:hi]).abs().max().item() / r["logits"].abs().max().item())
            rows += hi - lo
        same_pred &= got["pred"].shape == r["pred"].shape and bool((got["pred"] == r["pred"]).float().mean() > 0.999)
        same_cnt &= bool((got["top1"] - r["top1"]).abs().max() <= 1) and bool((got["top5"] - r["top5"]).abs().max() <= 1)
    good = e < 1e-4 and same_pred and same_cnt
    print(f"rank {rank} {what}: rows={rows // len(gots)} rel_err={e:.2e} pred={same_pred} counts={same_cnt} {'OK' if good else 'FAIL'}", flush=True)
    return good


for nq, hard in ((1001, True), (512, True), (777, False)):
    banks = orc.synthetic_banks(nq, 5000, 256, 300, seed=5, sigma=0.5, sigma_text=0.8, shared=3.0)
    Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    labels = banks["test_labels"]
    kw = {} if hard else {"softmax_scale": orc.CLIP_SCALE * 0.1}
    single = ClipSearcher(dev)
    single.set_text(T)
    single.set_cache(K, L, **kw)
    refs = single.search(Q, BETAS, ALPHAS, labels=labels, want_logits=True)
    V = orc.hard_values(L.float()) if hard else orc.softmax_values(L.float(), orc.CLIP_SCALE, 0.1)
    want = orc.searcher_logits(orc.zero_shot_logits(Q.float(), T.float()), orc.image_attention(Q, K, V, 5.5), 0.5)
    e0 = (refs[0]["logits"][0].cpu() - want).abs().max().item() / want.abs().max().item()
    print(f"rank {rank} nq={nq} hard={hard}: single-rank result vs oracle rel_err={e0:.2e}", flush=True)
    ok &= e0 < 2e-3
    lo, hi = query_slice(nq, rank, world)
    for shard in ("keys", "queries"):
        sharded = ClipSearcher(dev, group=dist.group.WORLD, shard=shard)
        sharded.set_text(T)
        sharded.set_cache(K, L, **kw)
        for blocks in ((1, 3) if shard == "keys" else (None,)):
            gots = sharded.search(Q, BETAS, ALPHAS, labels=labels, want_logits=True, blocks=blocks)
            ok &= compare(gots, refs, f"nq={nq} hard={hard} shard={shard} blocks={blocks}")
        # every rank brings only ITS query slice (sharded host->device copy)
        gots = sharded.search(Q[:, lo:hi].contiguous(), BETAS, ALPHAS, labels=labels[lo:hi].contiguous(), want_logits=True,
                              query_shard=True, blocks=2 if shard == "keys" else None)
        ok &= compare(gots, refs, f"nq={nq} hard={hard} shard={shard} query_shard")

# temperature-softmax mode: (m, l, O) partials of the key shards merged by log-sum-exp
for hard in (True, False):
    banks = orc.synthetic_banks(600, 4000, 256, 50, seed=6, sigma=2.0, sigma_text=0.8, shared=3.0, dtype=torch.float16)
    Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    sm = ClipSearcher(dev, group=dist.group.WORLD, shard="keys")
    sm.set_text(T.float())
    sm.set_cache(K, L, softmax_normalize=True, softmax_scale=None if hard else 2.0)
    got = sm.search(Q, [100.0], [1.0], labels=banks["test_labels"], want_logits=True)[0]
    V = orc.hard_values(L.float()) if hard else torch.softmax(2.0 * L.float(), dim=1)
    ref = orc.softmax_attention(Q.float(), K.float(), V, 100.0)
    e = max((o.cpu() - ref[lo_:hi_]).abs().max().item() for lo_, hi_, _, o, _ in got["pieces"])
    good = e <= 2e-3
    ok &= good
    print(f"rank {rank} softmax mode tau=100 hard={hard}: max-abs vs oracle {e:.2e} {'OK' if good else 'FAIL'}", flush=True)

# row-sharded pseudo-label selection: every rank scans ITS rows of the train bank (stored logits, and
# features + text classifier without a logits bank), per-class top-k candidates all-gathered and merged
from summer_clip_b200 import selection
from summer_clip_b200.clip_searcher.cache_strategy import LazyLogitsBank
banks = orc.synthetic_banks(8, 6001, 256, 100, seed=8, sigma=0.5, sigma_text=0.8, shared=3.0, dtype=torch.float16)
K, L, T = (banks[n] for n in ("cache_image_features", "cache_image_outs", "text_features"))
n = K.shape[1]
r_lo, r_hi = n * rank // world, n * (rank + 1) // world
L_exact = (orc.normalize_columns(K.double()).t() @ T.double()).float()     # what the lazy bank computes (L itself is fp16-rounded)
for prob_scale in (None, orc.CLIP_SCALE):
    want = orc.topk_select(L.float(), 8) if prob_scale is None else orc.topk_prob_select(L.float(), 8, prob_scale)
    got = selection.topk_select_sharded(L[r_lo:r_hi].to(dev), 8, prob_scale, dist.group.WORLD)
    good = np.array_equal(got.cpu().numpy(), want)
    lazy = LazyLogitsBank(K[:, r_lo:r_hi].to(dev), T.to(dev))
    got_lazy = selection.topk_select_sharded(lazy, 8, prob_scale, dist.group.WORLD).cpu().numpy()
    want_lazy = orc.topk_select(L_exact, 8) if prob_scale is None else orc.topk_prob_select(L_exact, 8, prob_scale)
    same = got_lazy.shape == want_lazy.shape and float((got_lazy == want_lazy).mean()) >= 0.99    # a flip needs a gap < 1e-6
    ok &= good and same
    print(f"rank {rank} row-sharded selection prob_scale={prob_scale}: {got.numel()} picked, equals oracle={good}, "
          f"without the logits bank equals oracle={same} {'OK' if good and same else 'FAIL'}", flush=True)

dist.barrier()
print(f"rank {rank} SHARDED {'OK' if ok else 'FAIL'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
