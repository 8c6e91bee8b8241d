"""BASELINE.json config 5: online classification latency, query batch 1..4096 against the 1.28M-key RN50 bank,
on 1 / 2 / 4 / 8 GPUs (key-sharded ClipSearcher; run under torchrun for N > 1).  One JSON object per batch size
on rank 0: p50 / p99 of the host-observed latency of `search` (queries already on the device, predictions copied
back), queries/s and the bank bytes streamed per second.

    python tools/bench_latency.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_latency.py
"""
from __future__ import annotations

import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import make_banks  # noqa: E402
from summer_clip_b200 import build as _build  # noqa: E402
from summer_clip_b200.searcher import ClipSearcher, shard_range  # noqa: E402


def main():
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    if rank == 0:
        _build.build_library()
    if world > 1:
        dist.barrier()
    n, dim, c = 1281167, 1024, 1000
    lo, hi = shard_range(n, rank, world)
    q_bank, k_bank, outs, text, labels = make_banks(torch, 4096, lo, hi, dim, c, seed=5, device=dev)
    # the searcher shards a GLOBAL bank by rank; here every rank generated only its shard, so build the local cache
    # with a single-rank searcher and attach the group afterwards (same layout ClipSearcher.set_cache produces)
    s = ClipSearcher(dev)
    s.set_text(text)
    s.set_cache(k_bank, outs)
    del k_bank, outs
    if world > 1:
        s.group, s.world, s.rank = group, world, rank
    bank_bytes = 2.0 * 1024 * (s.hard_bank.n_sorted if s.hard_bank is not None else (hi - lo))
    for b in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
        qb = q_bank[:, :b].contiguous()
        lab = labels[:b].contiguous()
        lat = []
        for it in range(14):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = s.search(qb, [5.5], [1.0], labels=lab)[0]
            pred = res["pred"].cpu()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = sorted(lat[3:])
        p50 = lat[len(lat) // 2]
        # the same search replayed from a CUDA graph (fixed batch shape): launch latency out of the way
        g50 = None
        try:
            graph, gres = s.capture_search(qb, [5.5], [1.0], labels=lab)
            glat = []
            for it in range(14):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                graph.replay()
                gpred = gres[0]["pred"].cpu()
                glat.append((time.perf_counter() - t0) * 1e3)
            glat = sorted(glat[3:])
            g50 = glat[len(glat) // 2]
            same = bool((gpred == pred).all())
            del graph, gres
        except Exception as exc:  # noqa: BLE001
            same, g50 = str(exc)[:80], None
        if world > 1:
            t = torch.tensor([p50, lat[-1], g50 if g50 is not None else -1.0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            p50, p99, g50 = float(t[0]), float(t[1]), (float(t[2]) if g50 is not None else None)
        else:
            p99 = lat[-1]
        if rank == 0:
            print(json.dumps({"config": "cfg5_latency", "n_gpus": world, "batch": b, "p50_ms": p50, "p99_ms": p99,
                              "graph_p50_ms": g50, "graph_same_pred": same,
                              "queries_per_s": b / ((g50 or p50) * 1e-3), "bank_gbs_per_gpu": bank_bytes / ((g50 or p50) * 1e-3) / 1e9,
                              "top1": int(res["top1"][0])}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
