"""Host logic of the key-sharded N > 1 path on CPU, world_size 2, gloo: shard ranges, per-shard sorted-bank
layout, the partial-tile exchange (`searcher.exchange_partials`) and the query slices each rank finishes.
The CUDA kernel is replaced by the oracle here (the product path itself has no CPU fallback)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import clip_search_oracle as orc


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, nq: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from summer_clip_b200.searcher import exchange_partials, query_slice, shard_range
        banks = orc.synthetic_banks(nq, 700, 64, 23, seed=91, sigma=0.5, sigma_text=0.8, shared=3.0)
        Q, K, L = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs"))
        lo, hi = shard_range(K.shape[1], rank, world)
        labels = L[lo:hi].argmax(1).int()
        from bank_layout_spec import hard_bank_layout_spec
        bank = hard_bank_layout_spec(labels, 23)                       # the shard's own label-sorted layout
        assert sorted(bank.perm[bank.perm >= 0].tolist()) == list(range(hi - lo))
        part = orc.image_attention(Q, K[:, lo:hi], orc.hard_values(L[lo:hi]), 5.5)   # stand-in for the kernel
        rows, qlo, qhi = exchange_partials(part, dist.group.WORLD)
        assert (qlo, qhi) == query_slice(nq, rank, world) and rows.shape == (qhi - qlo, 23)
        whole = orc.image_attention(Q, K, orc.hard_values(L), 5.5)
        np.save(os.path.join(out_dir, f"err_{rank}.npy"), (rows - whole[qlo:qhi]).abs().max().numpy() if qhi > qlo else np.float32(0))
        np.save(os.path.join(out_dir, f"slice_{rank}.npy"), np.array([qlo, qhi, lo, hi]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nq", [50, 51, 1])
def test_key_sharded_exchange_two_ranks_gloo(tmp_path, nq):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, nq, str(tmp_path)), nprocs=2, join=True)
    slices = [np.load(tmp_path / f"slice_{r}.npy") for r in range(2)]
    assert slices[0][0] == 0 and slices[0][1] == slices[1][0] and slices[1][1] == nq          # query slices tile [0, nq)
    assert slices[0][2] == 0 and slices[0][3] == slices[1][2] and slices[1][3] == 700        # key shards tile the bank
    assert slices[0][3] % 128 == 0                                                          # 128-aligned shard boundary
    for r in range(2):
        assert float(np.load(tmp_path / f"err_{r}.npy")) < 1e-4


def _oracle_local_topk(conf: torch.Tensor, label: torch.Tensor, n_classes: int, k: int, row_offset: int):
    """Stand-in for sc_topk_per_class + selection.local_candidates on CPU: the oracle's select_topk_per_label per
    class, as (confidence [C, k], global row [C, k]; -1 = no candidate)."""
    cc = torch.zeros((n_classes, k), dtype=torch.float32)
    cr = torch.full((n_classes, k), -1, dtype=torch.int64)
    picked = orc.select_topk_per_label(label.numpy(), conf.numpy(), k)
    for c in range(n_classes):
        mine = [int(i) for i in picked if int(label[i]) == c]
        for j, i in enumerate(mine):
            cc[c, j], cr[c, j] = conf[i], i + row_offset
    return cc, cr


def _select_worker(rank: int, world: int, port: int, cut: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from summer_clip_b200 import selection
        banks = orc.synthetic_banks(4, 900, 32, 17, seed=92, sigma=0.5, sigma_text=0.8, shared=3.0)
        conf, label = orc.row_confidence(banks["cache_image_outs"], prob=True, scale=orc.CLIP_SCALE)
        conf = (conf * 64).round() / 64                                  # many equal confidences: ties -> smaller row
        lo, hi = (0, cut) if rank == 0 else (cut, 900)                   # rank shards are consecutive row ranges
        offs = selection.row_offsets(hi - lo, torch.device("cpu"), dist.group.WORLD)
        assert offs == [0, cut, 900]
        cc, cr = _oracle_local_topk(conf[lo:hi], label[lo:hi], 17, 5, offs[rank])
        got = selection.exchange_and_merge(cc, cr, 5, dist.group.WORLD)
        flat = got.reshape(-1)
        np.save(os.path.join(out_dir, f"sel_{rank}.npy"), flat[flat >= 0].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cut", [450, 13, 0])
def test_sharded_selection_two_ranks_gloo(tmp_path, cut):
    """SURVEY.md 8e "Selection kernel": per-rank per-class top-k candidates, all-gathered and merged, equal the
    reference's select_topk_per_label over the whole bank (uneven and empty shards; tied confidences)."""
    port = _free_port()
    mp.spawn(_select_worker, args=(2, port, cut, str(tmp_path)), nprocs=2, join=True)
    banks = orc.synthetic_banks(4, 900, 32, 17, seed=92, sigma=0.5, sigma_text=0.8, shared=3.0)
    conf, label = orc.row_confidence(banks["cache_image_outs"], prob=True, scale=orc.CLIP_SCALE)
    conf = (conf * 64).round() / 64
    want = orc.select_topk_per_label(label.numpy(), conf.numpy(), 5)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"sel_{r}.npy"), want)


def test_merge_candidates_total_order():
    """NaN above +inf, equal confidences by row, -inf real candidates before missing ones, fewer than k candidates."""
    from summer_clip_b200.selection import merge_candidates
    nan, inf = float("nan"), float("inf")
    conf = torch.tensor([[[0.5, 0.5, 0.0]], [[nan, inf, -inf]], [[0.5, 0.0, 0.0]]])          # [W=3, C=1, k=3]
    row = torch.tensor([[[7, 9, -1]], [[20, 21, 22]], [[3, -1, -1]]])
    assert merge_candidates(conf, row, 3).tolist() == [[20, 21, 3]]
    assert merge_candidates(conf, row, 8).tolist() == [[20, 21, 3, 7, 9, 22, -1, -1]]


def test_shard_and_slice_arithmetic():
    from summer_clip_b200.searcher import query_slice, shard_range
    for n in (1, 127, 128, 1000, 1281167):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert all(lo % 128 == 0 for lo, _ in edges if lo < n)
            q = [query_slice(n, r, world) for r in range(world)]
            assert q[0][0] == 0 and q[-1][1] == n and all(a[1] == b[0] for a, b in zip(q, q[1:]))
