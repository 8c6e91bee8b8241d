"""2-rank NCCL check of the key-sharded ClipSearcher.search against the single-rank result (run under torchrun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

from oracle import clip_search_oracle as orc
from summer_clip_b200.searcher import ClipSearcher

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ok = True
for nq, hard in ((1001, True), (512, True), (777, False)):
    banks = orc.synthetic_banks(nq, 5000, 256, 300, seed=5, sigma=0.5, sigma_text=0.8, shared=3.0)
    Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    labels = banks["test_labels"]
    kw = {} if hard else {"softmax_scale": orc.CLIP_SCALE * 0.1}
    single = ClipSearcher(dev)
    single.set_text(T)
    single.set_cache(K, L, **kw)
    BETAS = [5.5, 0.1, 1.0, 3.5, 11.5]                 # one-hot banks: a 4-beta launch + a 1-beta launch
    refs = single.search(Q, BETAS, [0.5, 2.0], labels=labels, want_logits=True)
    ref = refs[0]
    V = orc.hard_values(L.float()) if hard else orc.softmax_values(L.float(), orc.CLIP_SCALE, 0.1)
    want = orc.searcher_logits(orc.zero_shot_logits(Q.float(), T.float()), orc.image_attention(Q, K, V, 5.5), 0.5)
    e0 = (ref["logits"][0].cpu() - want).abs().max().item() / want.abs().max().item()
    print(f"rank {rank} nq={nq} hard={hard}: single-rank result vs oracle rel_err={e0:.2e}", flush=True)
    ok &= e0 < 2e-3
    for shard in ("keys", "queries"):
        sharded = ClipSearcher(dev, group=dist.group.WORLD, shard=shard)
        sharded.set_text(T)
        sharded.set_cache(K, L, **kw)
        gots = sharded.search(Q, BETAS, [0.5, 2.0], labels=labels, want_logits=True)
        e, same_pred, same_cnt = 0.0, True, True
        for got, r in zip(gots, refs):
            lo, hi = got["lo"], got["hi"]
            e = max(e, (got["logits"] - r["logits"][:, lo:hi]).abs().max().item() / r["logits"].abs().max().item())
            same_pred &= bool((got["pred"] == r["pred"]).float().mean() > 0.999) and got["pred"].shape == r["pred"].shape
            same_cnt &= bool((got["top1"] - r["top1"]).abs().max() <= 1) and bool((got["top5"] - r["top5"]).abs().max() <= 1)
        got = gots[0]
        good = e < 1e-4 and same_pred and same_cnt and got["pred"].shape == ref["pred"].shape
        ok &= good
        print(f"rank {rank} nq={nq} hard={hard} shard={shard}: slice=[{lo},{hi}) rel_err={e:.2e} pred={same_pred} "
              f"counts={same_cnt} {'OK' if good else 'FAIL'}", flush=True)
dist.barrier()
print(f"rank {rank} SHARDED {'OK' if ok else 'FAIL'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
