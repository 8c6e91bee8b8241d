"""Parity of the CUDA path (called through the C ABI via summer_clip_b200.ops / the strategy classes)
against (a) the golden outputs of the reference's own code and (b) the CPU oracle on seeded banks.

Acceptance (BASELINE.json north_star): pseudo-label indices and top-k sets bit-exact; output
logits within 2e-3 max-abs after softmax; argmax agreement >= 99.9 %.
"""
import numpy as np
import pytest
import torch

import bank_layout_spec
from oracle import clip_search_oracle as orc

pytestmark = pytest.mark.gpu

SOFTMAX_TOL = 2e-3       # north_star: max-abs after softmax
ARGMAX_AGREE = 0.999     # north_star: argmax agreement


@pytest.fixture(scope="module")
def ops(cuda_lib):
    from summer_clip_b200 import ops as _ops
    assert torch.cuda.is_available()
    return _ops


def assert_logits_match(out: torch.Tensor, ref: torch.Tensor, what: str = ""):
    out, ref = out.detach().float().cpu(), ref.detach().float().cpu()
    diff = (torch.softmax(out, dim=1) - torch.softmax(ref, dim=1)).abs().max().item()
    assert diff <= SOFTMAX_TOL, f"{what}: softmax max-abs {diff:.3e} > {SOFTMAX_TOL}"
    agree = (out.argmax(1) == ref.argmax(1)).float().mean().item()
    n = out.shape[0]
    assert agree >= min(ARGMAX_AGREE, 1.0 - 1.0 / n), f"{what}: argmax agreement {agree:.5f}"
    return diff, agree


def cuda(x, dtype=None):
    t = torch.from_numpy(np.asarray(x)) if not isinstance(x, torch.Tensor) else x
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ----------------------------------------------------------------------------- selection (bit-exact)
@pytest.fixture(scope="module")
def sel(golden_dir):
    return np.load(golden_dir / "selection.npz")


def test_rowconf_matches_reference(ops, sel):
    outs = cuda(sel["image_outs"])
    conf, label = ops.rowconf(outs, prob=False)
    assert np.array_equal(label.cpu().numpy(), sel["label"])
    assert np.array_equal(conf.cpu().numpy(), sel["conf_raw"])            # max is exact
    conf_p, label_p = ops.rowconf(outs, scale=orc.CLIP_SCALE, prob=True)
    assert np.array_equal(label_p.cpu().numpy(), sel["label"])
    np.testing.assert_allclose(conf_p.cpu().numpy(), sel["conf_prob"], rtol=2e-6)


@pytest.mark.parametrize("k", [1, 4, 16, 64])
def test_topk_strategies_bit_exact_vs_reference_golden(ops, sel, k):
    from summer_clip_b200.clip_searcher.cache_strategy import TopKProbStrategy, TopKStrategy
    outs = cuda(sel["image_outs"])
    feats = torch.empty(1, outs.shape[0], device="cuda")
    idx = TopKStrategy(k).select(feats, outs)
    assert idx.dtype == torch.int64
    assert np.array_equal(idx.cpu().numpy(), sel[f"topk_{k}"])
    idx_p = TopKProbStrategy(k, orc.CLIP_SCALE).select(feats, outs)
    assert np.array_equal(idx_p.cpu().numpy(), sel[f"topk_prob_{k}"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_selection_vs_oracle_sun397_shape(ops, dtype):
    """19 850 x 397 logits bank (config 1 shape), k = 16, raw and prob ranking, incl. fp16 storage (ties!)."""
    from summer_clip_b200.clip_searcher.cache_strategy import TopKProbStrategy, TopKStrategy
    g = torch.Generator().manual_seed(21)
    outs = (0.25 + 0.02 * torch.randn(19850, 397, generator=g)).to(dtype)
    feats = torch.empty(1, 1, device="cuda")
    got = TopKStrategy(16).select(feats, outs.cuda()).cpu().numpy()
    assert np.array_equal(got, orc.topk_select(outs, 16))                  # fp16 ties -> index-ascending policy
    got_p = TopKProbStrategy(16, orc.CLIP_SCALE).select(feats, outs.cuda()).cpu().numpy()
    want_p = orc.topk_prob_select(outs, 16)
    if not np.array_equal(got_p, want_p):
        # the two softmax sums may differ in the last ulp: then the SETS per class must still agree
        # wherever the oracle's k/(k+1) boundary is not a near-tie
        conf, label = orc.row_confidence(outs, prob=True)
        conf, label = conf.numpy(), label.numpy()
        for c in np.unique(label):
            a, b = set(got_p[label[got_p] == c].tolist()), set(want_p[label[want_p] == c].tolist())
            if a != b:
                vals = np.sort(conf[label == c])[::-1]
                k = min(16, vals.size)
                assert k < vals.size and abs(vals[k - 1] - vals[k]) <= 4e-7 * vals[k - 1], f"class {c} differs off a tie"


def test_remaining_strategies_vs_reference_golden(ops, golden_dir):
    """Threshold, per-gold top-k (raw / softmax) and the per-predicted-class random sampler against outputs of the
    reference's own classes (tests/golden/strategies.npz): indices bit-exact, same order."""
    from summer_clip_b200.clip_searcher import cache_strategy as cs
    st = np.load(golden_dir / "strategies.npz")
    outs = cuda(st["image_outs"])
    gold = cuda(st["gold_labels"])
    feats = torch.empty(1, outs.shape[0], device="cuda")
    for thr in (0.06, 0.07):
        assert np.array_equal(cs.ThresholdStrategy(thr, True).select(feats, outs).cpu().numpy(), st[f"threshold_softmax_{thr}"])
    for thr in (0.3, 0.36):
        assert np.array_equal(cs.ThresholdStrategy(thr, False).select(feats, outs).cpu().numpy(), st[f"threshold_raw_{thr}"])
    for k in (1, 4, 80):
        got = cs.TopKPerGoldStrategy(k, cache_labels=gold).select(feats, outs).cpu().numpy()
        assert np.array_equal(got, st[f"topk_per_gold_{k}"])
        got = cs.TopKPerGoldProbStrategy(k, scale=orc.CLIP_SCALE, cache_labels=gold).select(feats, outs).cpu().numpy()
        assert np.array_equal(got, st[f"topk_per_gold_prob_{k}"])
    for k in (1, 3, 60):
        np.random.seed(42)
        got = cs.PerPredClassRandomSampleStrategy(k).select(feats, outs).cpu().numpy()
        assert np.array_equal(got, st[f"per_pred_random_{k}"])


def test_selection_edge_cases(ops):
    # empty classes, classes with fewer than k members, k larger than N, a single row, all rows one class
    conf = cuda(np.array([0.9, 0.1, 0.5, 0.5, 0.7], dtype=np.float32))
    label = cuda(np.array([2, 2, 0, 0, 2], dtype=np.int32))
    idx, cnt = ops.topk_per_class(conf, label, 4, 2)
    assert cnt.cpu().tolist() == [2, 0, 2, 0]
    assert idx.cpu().tolist() == [[2, 3], [-1, -1], [0, 4], [-1, -1]]      # tie 0.5/0.5 -> smaller index first
    assert ops.select_topk_per_label(conf, label, 4, 64).cpu().tolist() == [2, 3, 0, 4, 1]
    one = ops.select_topk_per_label(cuda(np.array([0.3], dtype=np.float32)), cuda(np.array([1], dtype=np.int32)), 3, 5)
    assert one.cpu().tolist() == [1 - 1]
    n = 70000
    g = torch.Generator().manual_seed(5)
    c = torch.rand(n, generator=g)
    big = ops.select_topk_per_label(c.cuda(), torch.zeros(n, dtype=torch.int32).cuda(), 1, 1000).cpu().numpy()
    assert np.array_equal(big, np.argsort(-c.numpy(), kind="stable")[:1000])   # radix-select path, one huge bucket
    neg = cuda(np.array([-1.0, -0.0, 0.0, -3.5, float("-inf")], dtype=np.float32))
    order = ops.select_topk_per_label(neg, torch.zeros(5, dtype=torch.int32).cuda(), 1, 5).cpu().tolist()
    assert order == [1, 2, 0, 3, 4]                                        # -0.0 == 0.0, then index ascending


# ----------------------------------------------------------------------------- attention vs reference golden
@pytest.fixture(scope="module")
def att(golden_dir):
    return np.load(golden_dir / "image_attention.npz")


@pytest.mark.parametrize("op_dtype", [torch.float16, torch.bfloat16])
def test_image_attention_sweep_vs_reference_golden(ops, att, op_dtype, monkeypatch):
    from summer_clip_b200.clip_searcher.cache_value_strategy import HardCacheStrategy, SoftmaxCacheStrategy
    from summer_clip_b200.clip_searcher.cache_weights_strategy import TipAdapterWeightsStrategy, _BANKS
    monkeypatch.setattr(ops, "OP_DTYPE", op_dtype)
    _BANKS.clear()
    Q, K, L = cuda(att["test_image_features"]), cuda(att["cache_image_features"]), cuda(att["cache_image_outs"])
    Z = cuda(att["clip_logits"])
    labels = cuda(att["test_labels"])
    idx = cuda(att["cache_idx"])
    Kc, Lc = K[:, idx], L[idx]                                             # reference loop: image_attention.py:54-55
    strategies = [HardCacheStrategy(), SoftmaxCacheStrategy(orc.CLIP_SCALE, 0.1), SoftmaxCacheStrategy(orc.CLIP_SCALE, 10.0)]
    for vi, vstrat in enumerate(strategies):
        values = vstrat.transform(Lc)
        np.testing.assert_allclose(values.dense().cpu().numpy(), att[f"values_{vi}"], atol=4e-3 if op_dtype == torch.bfloat16 else 5e-4)
        for bi, beta in enumerate(att["betas"]):
            weights = TipAdapterWeightsStrategy(float(beta)).transform(Q, Kc)      # un-normalised [D, N] banks in
            assert weights.shape == (Q.shape[1], idx.numel())
            cache_logits = weights @ values                                         # the reference's own expression
            ref = torch.from_numpy(att[f"cache_logits_v{vi}_b{bi}"])
            rel = (cache_logits.cpu() - ref).abs().max().item() / ref.abs().max().item()
            # bf16 operands (opt-in) carry 3 fewer mantissa bits: ~8x the fp16 error, largest at beta = 11.5
            assert rel < (2e-2 if op_dtype == torch.bfloat16 else 1e-3), (vi, bi, rel)
            res = ops.epilogue(Z, cache_logits, att["alphas"].tolist(), labels=labels, want_logits=True)
            for ai, alpha in enumerate(att["alphas"]):
                want = torch.from_numpy(att["clip_logits"]) + ref * float(alpha)
                if op_dtype == torch.float16:
                    assert_logits_match(res["logits"][ai], want, f"v{vi} b{bi} a{ai}")
                acc = att[f"acc_v{vi}_b{bi}"][ai]
                n = labels.numel()
                assert abs(100.0 * int(res["top1"][ai]) / n - acc[0]) <= 100.0 / n + 1e-9
                assert abs(100.0 * int(res["top5"][ai]) / n - acc[1]) <= 100.0 / n + 1e-9


def test_materialized_weights_vs_reference_golden(ops, att):
    from summer_clip_b200.clip_searcher.cache_weights_strategy import TipAdapterWeightsStrategy
    Q, K = cuda(att["test_image_features"]), cuda(att["cache_image_features"])
    idx = cuda(att["cache_idx"])
    W = TipAdapterWeightsStrategy(5.5).transform(Q, K[:, idx]).materialize()
    np.testing.assert_allclose(W.cpu().numpy(), att["weights_b5.5"], rtol=0, atol=3e-3)


def test_all_logits_cache_and_zero_shot_vs_golden(ops, att):
    Q, K, L, T = (cuda(att[n]) for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    Z = ops.zero_shot_logits(Q, True, T)
    np.testing.assert_allclose(Z.cpu().numpy(), att["clip_logits"], rtol=0, atol=2e-4)
    Qn, Kn = ops.normalize_cast(Q, True), ops.normalize_cast(K, True)
    Vt = ops.values_prepare(L, L.shape[1])
    O = ops.attn_fwd(Qn, Kn, Vt, K.shape[1], L.shape[1], 5.5)
    ref = torch.from_numpy(att["cache_logits_all_hard_b5.5"])
    assert (O.cpu() - ref).abs().max().item() / ref.abs().max().item() < 1e-3
    from summer_clip_b200.clip_searcher.utils import compute_accuracy
    assert np.allclose(compute_accuracy(Z, cuda(att["test_labels"])), att["acc_zero_shot"])


def test_tip_adapter_head_and_search_vs_golden(ops, golden_dir, capsys):
    from summer_clip_b200.tip_adapter import utils as tip_utils
    tip = np.load(golden_dir / "tip_adapter.npz")
    feats = cuda(tip["features"], torch.float16)                           # Tip-Adapter caches are fp16
    keys = cuda(tip["cache_keys"].T.copy(), torch.float16).t()             # [D, Nk] permuted VIEW (utils.py:61)
    assert not keys.is_contiguous()
    vals = torch.nn.functional.one_hot(cuda(tip["cache_labels"]).long(), 11).half()
    clip_w = cuda(tip["clip_weights"], torch.float16)
    labels = cuda(tip["test_labels"])
    # fp32 inputs: against the reference's golden logits
    head32 = tip_utils.TipAdapterHead(cuda(tip["cache_keys"]), vals, cuda(tip["features"]), cuda(tip["clip_weights"]))
    out32 = head32.logits(5.5, 1.0)
    assert_logits_match(out32, torch.from_numpy(tip["tip_logits"]), "tip head fp32 inputs")
    assert abs(tip_utils.cls_acc(out32, labels) - float(tip["acc_tip"])) <= 100.0 / labels.numel() + 1e-9
    # fp16 caches (what tip_adapter/utils.py saves on a GPU): against the oracle on the SAME fp16-rounded inputs
    head = tip_utils.TipAdapterHead(keys, vals, feats, clip_w)
    out = head.logits(5.5, 1.0)
    want16 = orc.tip_head(feats.float().cpu(), keys.float().cpu(), orc.onehot_values(torch.from_numpy(tip["cache_labels"]), 11),
                          clip_w.float().cpu(), 5.5, 1.0)
    assert_logits_match(out, want16, "tip head fp16 inputs")
    cfg = {"search_hp": True, "search_scale": tip["search_scale"].tolist(), "search_step": tip["search_step"].tolist()}
    best_beta, best_alpha = tip_utils.search_hp(cfg, keys, vals, feats, labels, clip_w)
    printed = capsys.readouterr().out
    assert "After searching, the best accuarcy" in printed
    # fp16 inputs can move a near-tie in accuracy: the found optimum must be as good as the reference's
    _, _, ref_best = orc.search_hp(cfg["search_scale"], cfg["search_step"], torch.from_numpy(tip["cache_keys"]),
                                   orc.onehot_values(torch.from_numpy(tip["cache_labels"]), 11),
                                   torch.from_numpy(tip["features"]), torch.from_numpy(tip["test_labels"]),
                                   torch.from_numpy(tip["clip_weights"]))
    got_acc = tip_utils.cls_acc(head.logits(best_beta, best_alpha), labels)
    assert got_acc >= ref_best - 100.0 / labels.numel() - 1e-9
    assert (best_beta, best_alpha) == (pytest.approx(float(tip["best_beta"])), pytest.approx(float(tip["best_alpha"]))) \
        or got_acc >= ref_best - 1e-9


# ----------------------------------------------------------------------------- vs the oracle on seeded banks
@pytest.mark.parametrize("shape", [(1000, 6000, 512, 397, torch.float16), (777, 4097, 1024, 1000, torch.float32),
                                   (130, 129, 768, 100, torch.float32)])
def test_fused_path_vs_oracle(ops, shape):
    """Ragged Nq / Nk (not multiples of 128), D in {512, 768, 1024}, C in {100, 397, 1000}, fp16 and fp32 banks."""
    from summer_clip_b200.searcher import ClipSearcher
    nq, nk, dim, c, dtype = shape
    banks = orc.synthetic_banks(nq, nk, dim, c, seed=31, sigma=0.5, sigma_text=0.8, shared=2.0, dtype=dtype)
    Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    labels = banks["test_labels"]
    betas, alphas = [1.0, 5.5, 11.5], [0.0, 0.5, 1.0, 4.0]
    s = ClipSearcher("cuda")
    s.set_text(T.float())
    Zref = orc.zero_shot_logits(Q.float(), T.float())
    for hard in (True, False):
        s.set_cache(K, L, softmax_scale=None if hard else orc.CLIP_SCALE * 0.1)
        V = orc.hard_values(L.float()) if hard else orc.softmax_values(L.float(), orc.CLIP_SCALE, 0.1)
        res = s.search(Q, betas, alphas, labels=labels, want_logits=True)
        for r, beta in zip(res, betas):
            Oref = orc.image_attention(Q, K, V, beta)
            for ai, alpha in enumerate(alphas):
                want = orc.searcher_logits(Zref, Oref, alpha)
                assert_logits_match(r["logits"][ai], want, f"{shape} hard={hard} beta={beta} alpha={alpha}")
                c1, c5 = orc.accuracy_counts(want, labels.long())
                assert abs(int(r["top1"][ai]) - c1) <= max(1, nq // 1000)
                assert abs(int(r["top5"][ai]) - c5) <= max(1, nq // 1000)
                assert torch.equal(r["pred"][ai].long().cpu(), r["logits"][ai].argmax(1).cpu())


def test_pseudo_label_cache_pipeline_vs_oracle(ops):
    """select (TopKProb k=8) -> gather -> attention, the default CLIP-search configuration."""
    from summer_clip_b200.clip_searcher.cache_strategy import TopKProbStrategy
    from summer_clip_b200.searcher import ClipSearcher
    banks = orc.synthetic_banks(600, 9000, 512, 100, seed=32, sigma=0.5, sigma_text=0.8, shared=3.0)
    Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    idx = TopKProbStrategy(8, orc.CLIP_SCALE).select(K.cuda(), L.cuda())
    want_idx = orc.topk_prob_select(L, 8)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    s = ClipSearcher("cuda")
    s.set_text(T)
    s.set_cache(K, L, idx=idx)
    res = s.search(Q, [5.5], [1.0, 2.0], labels=banks["test_labels"], want_logits=True)[0]
    widx = torch.from_numpy(want_idx)
    Oref = orc.image_attention(Q, K[:, widx], orc.hard_values(L[widx]), 5.5)
    Zref = orc.zero_shot_logits(Q, T)
    for ai, alpha in enumerate([1.0, 2.0]):
        assert_logits_match(res["logits"][ai], orc.searcher_logits(Zref, Oref, alpha), f"alpha={alpha}")


# ----------------------------------------------------------------------------- structural properties
def test_key_splits_and_shards_sum_to_the_whole(ops):
    g = torch.Generator().manual_seed(41)
    nq, nk, dim, c = 300, 5000, 256, 300
    Qn = ops.normalize_cast(torch.randn(nq, dim, generator=g).cuda(), False)
    Kn = ops.normalize_cast(torch.randn(nk, dim, generator=g).cuda(), False)
    L = torch.randn(nk, c, generator=g).cuda()
    Vt = ops.values_prepare(L, c, softmax_scale=1.0)
    whole = ops.attn_fwd(Qn, Kn, Vt, nk, c, 3.0, splits=1)
    for splits in (2, 7, 40):
        parts = ops.attn_fwd(Qn, Kn, Vt, nk, c, 3.0, splits=splits, merge=False)
        assert parts.shape == (splits, nq, c)
        torch.testing.assert_close(ops.merge_partials(parts), whole, rtol=2e-4, atol=1e-4)   # fp32 summation order differs
    # key shards as the multi-GPU path cuts them (128-aligned contiguous ranges) merge to the same result
    from summer_clip_b200.searcher import shard_range
    shards = []
    for r in range(3):
        lo, hi = shard_range(nk, r, 3)
        idx = torch.arange(lo, hi, device="cuda")
        vt = ops.values_prepare(L, c, idx=idx, softmax_scale=1.0)
        shards.append(ops.attn_fwd(Qn, Kn[lo:hi].contiguous(), vt, hi - lo, c, 3.0))
    torch.testing.assert_close(ops.merge_partials(torch.stack(shards)), whole, rtol=2e-4, atol=1e-4)


def test_linearity_and_rowsum_properties(ops):
    g = torch.Generator().manual_seed(42)
    nq, nk, dim, c = 257, 1111, 128, 50
    Qn = ops.normalize_cast(torch.randn(nq, dim, generator=g).cuda(), False)
    Kn = ops.normalize_cast(torch.randn(nk, dim, generator=g).cuda(), False)
    L = torch.randn(nk, c, generator=g).cuda()
    hard = ops.values_prepare(L, c, ones_row=True)
    O = ops.attn_fwd(Qn, Kn, hard, nk, c + 1, 2.0)
    # one-hot values: the class columns partition the keys, so they add up to the all-ones column
    torch.testing.assert_close(O[:, :c].sum(1), O[:, c], rtol=2e-5, atol=1e-5)
    # beta = 0 -> every weight is 1 -> O counts the keys per class exactly
    O0 = ops.attn_fwd(Qn, Kn, hard, nk, c + 1, 0.0)
    counts = torch.bincount(L.argmax(1), minlength=c).float()
    assert torch.equal(O0[:, :c], counts.expand(nq, c))
    assert torch.all(O0[:, c] == nk)
    # weights are bounded by exp(beta * (1 - 1)) = 1 up to rounding and positive
    assert O.min().item() >= 0.0


def test_tiny_and_degenerate_shapes(ops):
    g = torch.Generator().manual_seed(43)
    for nq, nk, dim, c in [(1, 1, 64, 1), (3, 2, 100, 7), (128, 128, 64, 16), (129, 257, 192, 257)]:
        Q, K = torch.randn(nq, dim, generator=g), torch.randn(nk, dim, generator=g)
        L = torch.randn(nk, c, generator=g)
        Qn, Kn = ops.normalize_cast(Q.cuda(), False), ops.normalize_cast(K.cuda(), False)
        Vt = ops.values_prepare(L.cuda(), c)
        O = ops.attn_fwd(Qn, Kn, Vt, nk, c, 4.0)
        ref = orc.image_attention(Q.t(), K.t(), orc.hard_values(L), 4.0)
        torch.testing.assert_close(O.cpu(), ref, rtol=0, atol=2e-3 * max(1.0, ref.abs().max().item()))


# ----------------------------------------------------------------------------- one-hot values: segmented kernel
@pytest.mark.parametrize("shape", [(1, 1, 64, 1), (3, 2, 100, 7), (128, 256, 64, 16), (129, 257, 192, 257),
                                   (300, 5000, 512, 1000), (513, 4099, 320, 100), (130, 700, 192, 37)])
def test_hard_label_segmented_kernel(ops, shape):
    """sc_attn_fwd_hard on a label-sorted bank == W @ one_hot(labels): against fp32 on the same rounded operands,
    against the dense-values kernel, with invalid labels (dropped), absent classes (zero columns), every
    split count, and beta = 0 counting keys exactly."""
    nq, nk, dim, c = shape
    g = torch.Generator().manual_seed(81)
    Qn = ops.normalize_cast(torch.randn(nq, dim, generator=g).cuda(), False)
    Kn = ops.normalize_cast(torch.randn(nk, dim, generator=g).cuda(), False)
    labels = torch.randint(0, max(1, c - c // 4), (nk,), generator=g).int()       # top quarter of the classes is absent
    if nk > 10:
        labels[torch.randperm(nk, generator=g)[: nk // 10]] = -1                  # invalid labels select no class
    labels = labels.cuda()
    bank = ops.hard_bank_layout(labels, c).gather(Kn)
    valid = labels >= 0
    assert bank.n_sorted % 16 == 0 and int((bank.perm >= 0).sum()) == int(valid.sum())
    W = torch.exp(3.0 * (Qn.float() @ Kn.float().t() - 1.0))
    ref = torch.zeros(nq, c, device="cuda").index_add_(1, labels[valid].long(), W[:, valid])
    steps = -(-bank.n_sorted // 256)
    for splits in sorted({1, min(3, steps), min(7, steps), 0}):
        O = ops.attn_fwd_hard(Qn, bank, 3.0, splits=splits)
        assert O.shape == (nq, c)
        torch.testing.assert_close(O, ref, rtol=2e-5, atol=1e-6 * max(1.0, ref.max().item()))
    parts = ops.attn_fwd_hard(Qn, bank, 3.0, splits=min(3, steps), merge=False)
    assert parts.shape[0] == min(3, steps) and float(parts[:, :, c - c // 4:].abs().sum()) == 0.0 if c >= 4 else True
    Vt = ops.values_prepare(None, c, labels=labels)
    dense = ops.attn_fwd(Qn, Kn, Vt, nk, c, 3.0)
    torch.testing.assert_close(O, dense, rtol=0, atol=2e-3 * max(1.0, ref.max().item()))
    O0 = ops.attn_fwd_hard(Qn, bank, 0.0)
    assert torch.equal(O0, torch.bincount(labels[valid].long(), minlength=c).float().expand(nq, c))


@pytest.mark.parametrize("n_keys,n_classes", [(1, 1), (17, 3), (1000, 37), (5000, 1000), (70000, 397), (2049, 5)])
def test_sorted_bank_layout_kernel_matches_the_specification(ops, n_keys, n_classes):
    """sc_hard_bank_layout (stable counting sort on the GPU) == the torch index arithmetic, element for element;
    sc_gather_rows == index_select with zero padding rows."""
    g = torch.Generator().manual_seed(84)
    labels = torch.randint(-1, n_classes + 1, (n_keys,), generator=g).int()
    want = bank_layout_spec.hard_bank_layout_spec(labels, n_classes)
    for lab in (labels.cuda(), ops.hard_labels(None, n_classes, labels=labels.cuda())[:n_keys]):
        got = ops.hard_bank_layout(lab, n_classes)
        assert got.n_sorted == want.n_sorted and got.n_keys == n_keys
        assert torch.equal(got.perm.cpu(), want.perm) and torch.equal(got.gcls.cpu(), want.gcls)
        assert torch.equal(got.kbits.cpu(), want.kbits)
    rows = torch.randn(n_keys, 64, generator=g).half().cuda()
    got.gather(rows)
    ref = rows[got.perm.clamp_min(0)]
    ref[got.perm < 0] = 0
    assert torch.equal(got.rows, ref)


def test_hard_label_segmented_kernel_random_shapes(ops):
    """Seeded sweep of ragged shapes through sc_attn_fwd_hard (both operand types): queries / keys off every tile
    boundary, D not a multiple of 64, more classes than keys, one huge class, splits forced high."""
    rng = np.random.default_rng(87)
    for case in range(14):
        nq = int(rng.integers(1, 700))
        nk = int(rng.integers(1, 3000))
        dim = int(rng.choice([24, 64, 100, 192, 257, 512]))
        c = int(rng.choice([1, 2, 17, 256, 257, 1000, 4000]))
        beta = float(rng.choice([0.1, 1.0, 5.5, 11.5]))
        op_dtype = torch.float16 if case % 2 == 0 else torch.bfloat16
        g = torch.Generator().manual_seed(1000 + case)
        Qn = ops.normalize_cast(torch.randn(nq, dim, generator=g).cuda(), False, op_dtype=op_dtype)
        Kn = ops.normalize_cast(torch.randn(nk, dim, generator=g).cuda(), False, op_dtype=op_dtype)
        if case % 3 == 0:
            labels = torch.zeros(nk, dtype=torch.int32)                       # every key in one class
        else:
            labels = torch.randint(0, c, (nk,), generator=g).int()
        labels = labels.cuda()
        bank = ops.hard_bank_layout(labels, c).gather(Kn)
        W = torch.exp(beta * (Qn.float() @ Kn.float().t() - 1.0))
        ref = torch.zeros(nq, c, device="cuda").index_add_(1, labels.long(), W)
        steps = max(1, -(-bank.n_sorted // 256))
        for splits in (0, steps):
            O = ops.attn_fwd_hard(Qn, bank, beta, splits=splits)
            torch.testing.assert_close(O, ref, rtol=3e-5, atol=1e-6 * max(1.0, ref.max().item()),
                                       msg=lambda m: f"case {case} nq={nq} nk={nk} D={dim} C={c} splits={splits}: {m}")


@pytest.mark.parametrize("nq,c", [(1, 1), (37, 5), (300, 128), (257, 129), (100, 397), (64, 1000), (50, 1024), (40, 1025), (33, 3000)])
def test_alpha_epilogue_kernel(ops, nq, c):
    """sc_epilogue / sc_epilogue_parts == torch `Z + O * alpha` -> argmax / top-1 / top-5 (image_attention.py:111-117,
    clip_searcher/utils.py:15-21) bit for bit, on the register path (C <= 1024) and the generic one: ties with the
    label's value and with the maximum (first index wins), -inf rows, labels out of range (not counted), missing Z,
    row-sum normalisation, and unmerged partial tiles (== merge_partials first)."""
    g = torch.Generator().manual_seed(93 + c)
    Z = (torch.randn(nq, c, generator=g) * 3).cuda()
    O = torch.rand(3, nq, c, generator=g).cuda()
    O[:, :, ::3] = (O[:, :, ::3] * 4).round() / 4                  # many exact ties
    Z[:, ::3] = Z[:, ::3].round()
    if nq > 4:
        Z[1], O[:, 1] = 0.0, 1.0                                   # a whole row of equal values
        Z[2] = float("-inf")
        Z[3, c // 2] = float("inf")
    labels = torch.randint(0, c, (nq,), generator=g).int().cuda()
    if nq > 8:
        labels[5], labels[6] = -1, c + 3
    alphas = [0.0, 0.25, 1.0, 4.0, 17.5]
    merged = ops.merge_partials(O)
    for z, rowsum in ((Z, None), (None, None), (Z, merged.sum(1) + 1.0)):
        res = ops.epilogue(z, merged, alphas, labels=labels, rowsum=rowsum, want_logits=True)
        from_parts = ops.epilogue(z, O, alphas, labels=labels, rowsum=rowsum, want_logits=True)
        counts_only = ops.epilogue(z, O, alphas, labels=labels, rowsum=rowsum, want_pred=False)
        o = merged if rowsum is None else merged * (1.0 / rowsum)[:, None]
        valid = (labels >= 0) & (labels < c)
        lab = labels.long().clamp(0, c - 1)
        for ai, alpha in enumerate(alphas):
            want = o * alpha if z is None else z + o * alpha
            assert torch.equal(res["logits"][ai], want)
            assert torch.equal(res["pred"][ai].long(), want.argmax(dim=1)), (nq, c, alpha)
            vlab = want.gather(1, lab[:, None])
            ahead = ((want > vlab) | ((want == vlab) & (torch.arange(c, device="cuda")[None, :] < lab[:, None]))).sum(1)
            assert int(res["top1"][ai]) == int(((ahead == 0) & valid).sum())
            assert int(res["top5"][ai]) == int(((ahead < 5) & valid).sum())
        for k in ("logits", "pred", "top1", "top5"):
            assert torch.equal(res[k], from_parts[k]), k
        assert counts_only["pred"] is None
        assert torch.equal(counts_only["top1"], res["top1"]) and torch.equal(counts_only["top5"], res["top5"])


@pytest.mark.parametrize("n_parts,rows,cols", [(1, 5, 8), (2, 300, 1000), (3, 513, 1000), (5, 100, 400), (8, 64, 1024),
                                               (3, 77, 397), (4, 10, 3)])
def test_merge_partials_is_the_ordered_sum(ops, n_parts, rows, cols):
    """sc_merge_partials (16-byte vector path when the row length allows it, scalar otherwise) == p0 + p1 + ... in that
    order, bit for bit — the order sc_epilogue_parts and the key-sharded exchange rely on."""
    g = torch.Generator().manual_seed(97)
    parts = (torch.randn(n_parts, rows, cols, generator=g) * 10.0 ** torch.randint(-3, 4, (n_parts, 1, 1), generator=g)).cuda()
    want = parts[0].clone()
    for p in range(1, n_parts):
        want += parts[p]
    assert torch.equal(ops.merge_partials(parts), want)


def test_e4m3_feature_banks(ops):
    """Opt-in 8-bit operands (north star part 1: "bf16/fp8 cast"): sc_normalize_cast writes 256 * x as e4m3
    (round-to-nearest, within one e4m3 step of torch's own cast of the fp32 normalised value), the segmented
    kernel (tcgen05 kind::f8f6f4, fp32 accumulate) reproduces fp32 arithmetic on THOSE operands to the same
    tolerance as the 16-bit types, and against unquantised fp32 the class sums move by a few 1e-3 relative."""
    e4m3 = torch.float8_e4m3fn
    g = torch.Generator().manual_seed(101)
    for nq, nk, dim, c, feature_major in [(300, 3000, 1024, 37, True), (129, 700, 200, 300, False), (64, 5000, 512, 1000, True)]:
        q_src = torch.randn((dim, nq) if feature_major else (nq, dim), generator=g).cuda()
        k_src = torch.randn((dim, nk) if feature_major else (nk, dim), generator=g).half().cuda()
        Qn = ops.normalize_cast(q_src, feature_major, op_dtype=e4m3)
        Kn = ops.normalize_cast(k_src, feature_major, op_dtype=e4m3)
        d_pad = -(-dim // 128) * 128
        assert Qn.dtype == e4m3 and Qn.shape == (nq, d_pad) and Kn.shape == (nk, d_pad)
        for src, got in ((q_src, Qn), (k_src, Kn)):
            x = src.float().t() if feature_major else src.float()
            want = torch.nn.functional.normalize(x, dim=1) * 256.0
            ref = want.to(e4m3).float()
            gf = got.float()
            assert float(gf[:, dim:].abs().sum()) == 0.0
            step = torch.maximum(torch.full_like(ref, 2.0 ** -9), 2.0 ** (torch.floor(torch.log2(ref.abs().clamp_min(2.0 ** -6))) - 3))
            assert bool(((gf[:, :dim] - ref).abs() <= step).all())
            assert float((gf[:, :dim] != ref).float().mean()) < 0.01          # only values on a rounding boundary move
        labels = torch.randint(0, c, (nk,), generator=g).int().cuda()
        bank = ops.hard_bank_layout(labels, c).gather(Kn)
        qf, kf = Qn.float() / 256.0, Kn.float() / 256.0
        for betas in ([5.5], [0.1, 1.0, 3.5, 11.5]):
            W = [torch.exp(b * (qf @ kf.t() - 1.0)) for b in betas]
            outs = ops.attn_fwd_hard_multi(Qn, bank, betas)
            for b, w, O in zip(betas, W, outs):
                ref = torch.zeros(nq, c, device="cuda").index_add_(1, labels.long(), w)
                torch.testing.assert_close(O, ref, rtol=3e-5, atol=1e-6 * max(1.0, ref.max().item()))
                assert torch.equal(O, ops.attn_fwd_hard(Qn, bank, b))
        # against unquantised fp32 features: the quantisation error itself (reported, loosely bounded)
        x_q = torch.nn.functional.normalize(q_src.float().t() if feature_major else q_src.float(), dim=1)
        x_k = torch.nn.functional.normalize(k_src.float().t() if feature_major else k_src.float(), dim=1)
        exact = torch.zeros(nq, c, device="cuda").index_add_(1, labels.long(), torch.exp(5.5 * (x_q @ x_k.t() - 1.0)))
        got = ops.attn_fwd_hard(Qn, bank, 5.5)
        rel = (got - exact).abs() / exact.abs().clamp_min(1e-6 * exact.max())
        assert float(rel.mean()) < 1e-2 and float(rel.max()) < 0.2
    with pytest.raises(TypeError):
        ops._op(e4m3)                                  # dense-values operands stay 16-bit


def test_search_over_several_betas_equals_one_search_per_beta(ops):
    """ClipSearcher.search groups the betas of a one-hot bank four to a launch; every result equals the search for
    that beta alone bit for bit."""
    from summer_clip_b200.searcher import ClipSearcher
    banks = orc.synthetic_banks(500, 4000, 256, 120, seed=11, sigma=0.5, sigma_text=0.8, shared=3.0)
    Q, K, L, T = (banks[n].cuda() for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    labels = banks["test_labels"].cuda()
    s = ClipSearcher("cuda")
    s.set_text(T)
    s.set_cache(K, L)
    betas, alphas = [0.1, 1.0, 1.5, 3.5, 5.5, 7.5], [0.0, 1.0, 4.0]
    many = s.search(Q, betas, alphas, labels=labels, want_logits=True)
    assert [r["beta"] for r in many] == betas
    for r in many:
        one = s.search(Q, [r["beta"]], alphas, labels=labels, want_logits=True)[0]
        for k in ("cache_logits", "logits", "pred", "top1", "top5"):
            assert torch.equal(r[k], one[k]), (r["beta"], k)


def test_beta_sweep_shares_the_tensor_core_pass(ops):
    """sc_attn_fwd_hard_multi: up to 4 betas per launch off one S = Q.K^T; every beta's slab is bit-identical to
    its own single-beta launch (same arithmetic, same summation order), for any group size and ragged shapes; the
    strategy-level sweep (FusedWeights.matmul_many, TipAdapterHead.cache_logits_many) goes through it."""
    from summer_clip_b200.clip_searcher.cache_value_strategy import GoldCacheValues
    from summer_clip_b200.clip_searcher.cache_weights_strategy import FusedWeights, TipAdapterWeightsStrategy, _BANKS
    g = torch.Generator().manual_seed(91)
    for nq, nk, dim, c in [(300, 3000, 192, 37), (129, 257, 64, 300), (513, 9000, 512, 1000)]:
        Qn = ops.normalize_cast(torch.randn(nq, dim, generator=g).cuda(), False)
        Kn = ops.normalize_cast(torch.randn(nk, dim, generator=g).cuda(), False)
        labels = torch.randint(0, c, (nk,), generator=g).int().cuda()
        bank = ops.hard_bank_layout(labels, c).gather(Kn)
        betas = [0.1, 1.0, 1.5, 3.5, 5.5, 7.5, 9.5, 11.5, 0.0]          # image_attention.yaml's sweep + a 1-beta tail group
        for n in (1, 2, 3, 4, 9):
            outs = ops.attn_fwd_hard_multi(Qn, bank, betas[:n])
            assert len(outs) == n
            for beta, O in zip(betas[:n], outs):
                assert torch.equal(O, ops.attn_fwd_hard(Qn, bank, beta)), (nq, nk, dim, c, n, beta)
        outs = ops.attn_fwd_hard_multi(Qn, bank, betas[:4], splits=3)
        for beta, O in zip(betas[:4], outs):
            assert torch.equal(O, ops.attn_fwd_hard(Qn, bank, beta, splits=3))
    _BANKS.clear()
    Q, K = torch.randn(64, 200, generator=g).cuda(), torch.randn(64, 1500, generator=g).cuda()
    gold = GoldCacheValues(20).transform(torch.randint(0, 20, (1500,), generator=g).cuda())
    ws = [TipAdapterWeightsStrategy(b).transform(Q, K) for b in (0.1, 1.0, 5.5, 7.5, 11.5)]
    for w, many in zip(ws, FusedWeights.matmul_many(ws, gold)):
        assert torch.equal(many, w @ gold)


def test_hard_values_route_through_the_segmented_kernel(ops, monkeypatch):
    """HardCacheStrategy / gold labels / one-hot Tip-Adapter cache values all become a label-sorted bank; the
    dense-values route (SUMMER_CLIP_B200_DENSE_VALUES=1) gives the same logits."""
    from summer_clip_b200.clip_searcher.cache_value_strategy import CacheValues, GoldCacheValues, HardCacheStrategy
    from summer_clip_b200.clip_searcher.cache_weights_strategy import TipAdapterWeightsStrategy, _BANKS
    banks = orc.synthetic_banks(300, 2000, 256, 50, seed=82, sigma=0.5, sigma_text=0.8, shared=3.0)
    Q, K, L = (banks[n].cuda() for n in ("test_image_features", "cache_image_features", "cache_image_outs"))
    _BANKS.clear()
    w = TipAdapterWeightsStrategy(5.5).transform(Q, K)
    hard = HardCacheStrategy().transform(L)
    assert hard.is_hard and hard.shape == (2000, 50)
    got = w @ hard
    ref = orc.image_attention(banks["test_image_features"], banks["cache_image_features"], orc.hard_values(banks["cache_image_outs"]), 5.5)
    assert (got.cpu() - ref).abs().max().item() / ref.abs().max().item() < 1e-3
    gold = GoldCacheValues(50).transform(banks["cache_labels"].cuda())
    one_hot = torch.nn.functional.one_hot(banks["cache_labels"].long(), 50).half().cuda()
    dense_in = CacheValues.from_dense(one_hot)
    assert gold.is_hard and dense_in.is_hard
    torch.testing.assert_close(w @ gold, w @ dense_in, rtol=0, atol=0)
    torch.testing.assert_close(w @ one_hot, w @ gold, rtol=0, atol=0)              # plain tensor operand, reference style
    monkeypatch.setenv("SUMMER_CLIP_B200_DENSE_VALUES", "1")
    dense = HardCacheStrategy().transform(L)
    assert not dense.is_hard
    torch.testing.assert_close(w @ dense, got, rtol=0, atol=2e-3 * got.max().item())


@pytest.mark.parametrize("shape", [(512, 333, 101, True, torch.float32), (1024, 1000, 1000, True, torch.float16),
                                   (768, 257, 397, False, torch.float32), (100, 5, 3, True, torch.float32),
                                   (64, 700, 513, False, torch.float16)])
def test_zero_shot_logits_tensor_core_route(ops, shape):
    """Z = 100 * normalise(X)^T T through split-fp16 operands on the tcgen05 pipeline (sc_normalize_split +
    sc_gemm_split_nt) against float64 torch and against the fp32 SIMT kernel: fp32-accurate."""
    D, N, C, fm, dtype = shape
    g = torch.Generator().manual_seed(83)
    X = torch.randn((D, N) if fm else (N, D), generator=g).to(dtype).cuda()
    T = torch.nn.functional.normalize(torch.randn(D, C, generator=g), dim=0).cuda()
    Z = ops.zero_shot_logits(X, fm, T)
    Zs = ops.zero_shot_logits(X, fm, T, tensor_cores=False)
    Xd = X.double() if not fm else X.double().t()
    ref = 100.0 * torch.nn.functional.normalize(Xd, dim=1) @ T.double()
    assert Z.shape == (N, C) and Z.dtype == torch.float32
    # |Z| <= 100; fp32 accumulation over D terms is what limits both routes (the golden check uses atol 2e-4)
    assert (Z.double() - ref).abs().max().item() < 3e-4
    assert (Zs.double() - ref).abs().max().item() < 3e-4
    assert (Z - Zs).abs().max().item() < 3e-4
    hi, lo = ops.normalize_split(X, fm)
    assert hi.shape == (N, ops.pad_dim(D)) and float(hi[:, D:].abs().sum() + lo[:, D:].abs().sum()) == 0.0
    rec = hi.double() + lo.double()
    assert (rec[:, :D] - torch.nn.functional.normalize(Xd, dim=1)).abs().max().item() < 3e-7      # 22 bits of each value
    raw = ops.zero_shot_logits(X, fm, T, scale=1.0, normalize=False)                              # the L producer
    assert (raw.double() - Xd @ T.double()).abs().max().item() < 1e-4 * max(1.0, float(Xd.abs().max()) * D ** 0.5)
    if dtype == torch.float16:      # the two-pass route of bank-sized fp16 inputs (sc_transpose_norms + sc_gemm_rows_nt)
        Z2 = ops.zero_shot_logits(X, fm, T, two_pass=True)
        assert (Z2.double() - ref).abs().max().item() < 3e-4
        raw2 = ops.zero_shot_logits(X, fm, T, scale=1.0, normalize=False, two_pass=True)
        assert (raw2.double() - Xd @ T.double()).abs().max().item() < 1e-4 * max(1.0, float(Xd.abs().max()) * D ** 0.5)


def test_sidecar_bank_gives_the_same_search(ops, tmp_path):
    """ClipSearcher.save_bank / load_bank (bank_io sidecar): identical predictions and logits."""
    from summer_clip_b200.searcher import ClipSearcher
    banks = orc.synthetic_banks(300, 3000, 256, 40, seed=85, sigma=0.5, sigma_text=0.8, shared=3.0, dtype=torch.float16)
    Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    a = ClipSearcher("cuda")
    a.set_text(T.float())
    a.set_cache(K, L)
    ra = a.search(Q, [5.5], [1.0], labels=banks["test_labels"], want_logits=True)[0]
    a.save_bank(tmp_path / "bank", key="k1")
    b = ClipSearcher("cuda")
    b.set_text(T.float())
    assert not b.load_bank(tmp_path / "bank", key="k2") and b.load_bank(tmp_path / "bank", key="k1")
    rb = b.search(Q, [5.5], [1.0], labels=banks["test_labels"], want_logits=True)[0]
    assert torch.equal(ra["logits"], rb["logits"]) and torch.equal(ra["pred"], rb["pred"])


def test_cuda_graph_replay_of_a_search(ops):
    """ClipSearcher.capture_search: new queries copied into the static inputs + graph.replay() == eager search."""
    from summer_clip_b200.searcher import ClipSearcher
    banks = orc.synthetic_banks(64, 3000, 256, 40, seed=86, sigma=0.5, sigma_text=0.8, shared=3.0, dtype=torch.float16)
    Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    s = ClipSearcher("cuda")
    s.set_text(T.float())
    s.set_cache(K, L)
    q_static = Q[:, :32].contiguous().cuda()
    lab_static = banks["test_labels"][:32].int().cuda()
    graph, res = s.capture_search(q_static, [5.5], [1.0, 2.0], labels=lab_static)
    for lo in (0, 32):
        q_static.copy_(Q[:, lo:lo + 32])
        lab_static.copy_(banks["test_labels"][lo:lo + 32])
        graph.replay()
        torch.cuda.synchronize()
        eager = s.search(Q[:, lo:lo + 32].contiguous(), [5.5], [1.0, 2.0], labels=banks["test_labels"][lo:lo + 32])[0]
        assert torch.equal(res[0]["pred"], eager["pred"]) and torch.equal(res[0]["top1"], eager["top1"])


def test_operand_caches_are_keyed_by_identity_not_address(ops):
    """Three problems of the SAME shapes back to back, each freeing its tensors: the caching allocator reuses the
    addresses, so any operand cache keyed by data_ptr would serve the previous problem's text split / normalised
    bank / sorted bank (a bug the 2-rank check caught).  Every result must match ITS oracle."""
    from summer_clip_b200.clip_searcher.cache_value_strategy import HardCacheStrategy
    from summer_clip_b200.clip_searcher.cache_weights_strategy import TipAdapterWeightsStrategy
    from summer_clip_b200.searcher import ClipSearcher
    for seed in (11, 12, 13):
        banks = orc.synthetic_banks(200, 1500, 128, 30, seed=seed, sigma=0.5, sigma_text=0.8, shared=3.0)
        Q, K, L, T = (banks[n] for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
        Zref = orc.zero_shot_logits(Q, T)
        Oref = orc.image_attention(Q, K, orc.hard_values(L), 5.5)
        s = ClipSearcher("cuda")
        s.set_text(T)
        s.set_cache(K, L)
        got = s.search(Q, [5.5], [1.0], want_logits=True)[0]
        assert_logits_match(got["logits"][0], orc.searcher_logits(Zref, Oref, 1.0), f"searcher seed={seed}")
        Qc, Kc, Lc = Q.cuda(), K.cuda(), L.cuda()
        o = TipAdapterWeightsStrategy(5.5).transform(Qc, Kc) @ HardCacheStrategy().transform(Lc)
        assert (o.cpu() - Oref).abs().max().item() / Oref.abs().max().item() < 1e-3, f"strategies seed={seed}"
        z = ops.zero_shot_logits(Qc, True, T.cuda())
        assert (z.cpu() - Zref).abs().max().item() < 3e-4, f"zero-shot seed={seed}"
        del s, got, Qc, Kc, Lc, o, z


def test_cpu_tensors_are_rejected(ops):
    from summer_clip_b200._lib import SummerClipError
    with pytest.raises(SummerClipError):
        ops.normalize_cast(torch.randn(8, 8), True)
    with pytest.raises(SummerClipError):
        ops.rowconf(torch.randn(8, 8))


def test_image_attention_runner_matches_oracle_sweep(ops, tmp_path):
    """The hydra-free ImageAttention trainer on .pt banks: same records, same order, accuracies equal to the
    oracle's sweep (image_attention.py:89-120)."""
    import json
    from summer_clip_b200.clip_searcher.image_attention import ImageAttention, run_trainer
    from summer_clip_b200.utils.config import load_config
    from pathlib import Path
    banks = orc.synthetic_banks(500, 3000, 256, 40, seed=51, sigma=0.5, sigma_text=0.8, shared=3.0, dtype=torch.float16)
    paths = {}
    for name, t in banks.items():
        paths[name] = tmp_path / f"{name}.pt"
        torch.save(t, paths[name])
    conf = Path(__file__).resolve().parent.parent / "summer_clip_b200" / "conf" / "image_attention.yaml"
    cfg = load_config(conf, {
        "data": {"image_features_path": str(paths["test_image_features"]), "text_features_path": str(paths["text_features"]),
                 "labels_path": str(paths["test_labels"])},
        "cache": {"image_features_path": str(paths["cache_image_features"]), "image_outs_path": str(paths["cache_image_outs"]),
                  "labels_path": str(paths["cache_labels"]), "alpha": [0.0, 1.0, 4.0]},
        "cache_weights_strategy": {"beta": [1.0, 5.5]},
        "run_saves": {"save_cache_inds": True},
        "cache_strategies": {"topk": {"topk": [2, 16]}, "topk_prob": {"topk": [4]},
                             "per_pred_class_random": {"topk": [2]}, "global_random": {"topk": [1]}},
    })
    for group in ("topk_per_gold", "topk_prob_per_gold", "per_gold_class_random"):      # covered by test_gpu_round2.py
        cfg.cache_strategies.pop(group)
    trainer = run_trainer(ImageAttention, cfg, run_dir=tmp_path / "run")
    records = [json.loads(l) for l in (tmp_path / "run" / "image_attention.log").read_text().splitlines()]
    kinds = [r.get("type") for r in records]
    assert kinds[0] is None and kinds[1] == "zero_shot"
    n_caches = 2 + 1 + 1 + 1 + 1
    assert kinds.count("cache_info") == n_caches and kinds.count("searcher_result") == n_caches * 2 * 3
    # oracle sweep over the deterministic strategies
    Q, K, L, T = (banks[n].float() for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    labels = banks["test_labels"].long()
    Z = orc.zero_shot_logits(Q, T)
    zs = next(r for r in records if r.get("type") == "zero_shot")
    assert np.allclose([zs["acc1"], zs["acc5"]], orc.compute_accuracy(Z, labels), atol=0.2 + 1e-9)
    results = [r for r in records if r.get("type") == "searcher_result"]
    assert set(results[0]) >= {"cache_strategy", "cache_value_strategy", "cache_weights_strategy", "alpha", "acc1", "acc5", "message"}
    assert results[0]["cache_strategy"]["_target_"] == "summer_clip.clip_searcher.cache_strategy.TopKStrategy"
    for strat, sel in (("TopKStrategy", lambda k: orc.topk_select(banks["cache_image_outs"], k)),
                       ("TopKProbStrategy", lambda k: orc.topk_prob_select(banks["cache_image_outs"], k)),
                       ("AllLogitsStrategy", lambda k: orc.all_logits_select(L))):
        for r in results:
            if not r["cache_strategy"]["_target_"].endswith(strat):
                continue
            idx = torch.from_numpy(sel(r["cache_strategy"].get("topk", 0)))
            O = orc.image_attention(Q, K[:, idx], orc.hard_values(L[idx]), r["cache_weights_strategy"]["beta"])
            acc = orc.compute_accuracy(orc.searcher_logits(Z, O, r["alpha"]), labels)
            assert abs(r["acc1"] - acc[0]) <= 0.2 + 1e-9 and abs(r["acc5"] - acc[1]) <= 0.2 + 1e-9, r
    info = [r for r in records if r.get("type") == "cache_info"]
    assert info[0]["cache_size"] == len(orc.topk_select(banks["cache_image_outs"], 2)) and "acc1" in info[0]
    saved = np.load(info[0]["cache_inds_path"])
    assert np.array_equal(saved, orc.topk_select(banks["cache_image_outs"], 2))


# ----------------------------------------------------------------------------- full BASELINE size
@pytest.mark.parametrize("nq", [4096])
def test_full_size_key_bank_properties(ops, nq):
    """Config-3 key bank (1 281 167 keys x 1024-d, 1000 classes) against a query block: the oracle cannot
    run this size, so check size-independent properties — key shards sum to the whole, one-hot class
    columns sum to the ones column, beta = 0 counts keys — plus a sampled exact check of a few rows."""
    nk, dim, c = 1281167, 1024, 1000
    g = torch.Generator(device="cuda").manual_seed(61)
    protos = torch.nn.functional.normalize(torch.randn(c, dim, generator=g, device="cuda"), dim=1)
    yk = torch.randint(0, c, (nk,), generator=g, device="cuda")
    Kn = torch.empty((nk, dim), dtype=ops.OP_DTYPE, device="cuda")
    step = 1 << 17
    for s in range(0, nk, step):
        e = min(nk, s + step)
        x = protos[yk[s:e]] + torch.randn(e - s, dim, generator=g, device="cuda") / dim ** 0.5
        ops.normalize_cast(x, False, out=Kn[s:e])
    yq = torch.randint(0, c, (nq,), generator=g, device="cuda")
    Qn = ops.normalize_cast(protos[yq] + torch.randn(nq, dim, generator=g, device="cuda") / dim ** 0.5, False)
    Vt = ops.values_prepare(None, c, labels=yk.int(), ones_row=True)
    O = ops.attn_fwd(Qn, Kn, Vt, nk, c + 1, 5.5)
    torch.testing.assert_close(O[:, :c].sum(1), O[:, c], rtol=1e-3, atol=1e-3)   # 1.28M-term fp32 sums, different order
    from summer_clip_b200.searcher import shard_range
    parts = []
    for r in range(8):
        lo, hi = shard_range(nk, r, 8)
        vt = ops.values_prepare(None, c, labels=yk[lo:hi].int(), ones_row=True)
        parts.append(ops.attn_fwd(Qn, Kn[lo:hi], vt, hi - lo, c + 1, 5.5))
    torch.testing.assert_close(ops.merge_partials(torch.stack(parts)), O, rtol=1e-3, atol=1e-3)
    # sampled rows against fp32 torch on the same device (chunked; 8 rows x 1.28M keys)
    rows = torch.tensor([0, 1, 127, 128, 1000, 2047, 4000, nq - 1], device="cuda")
    A = Qn[rows].float() @ Kn.float().t()
    W = torch.exp(5.5 * (A - 1.0))
    ref = torch.zeros(rows.numel(), c, device="cuda").index_add_(1, yk, W)
    torch.testing.assert_close(O[rows][:, :c], ref, rtol=2e-3, atol=1e-2)
    O0 = ops.attn_fwd(Qn[:128].contiguous(), Kn, Vt, nk, c + 1, 0.0)
    assert torch.all(O0[:, c] == nk)
    # the same bank through the label-sorted segmented kernel (the bench's path for one-hot values)
    bank = ops.hard_bank_layout(yk.int(), c).gather(Kn)
    Oh = ops.attn_fwd_hard(Qn, bank, 5.5)
    torch.testing.assert_close(Oh[rows], ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(Oh, O[:, :c], rtol=2e-3, atol=1e-2)
    torch.testing.assert_close(Oh.sum(1), O[:, c], rtol=1e-3, atol=1e-3)
    hparts = []
    for r in range(8):
        lo, hi = shard_range(nk, r, 8)
        hparts.append(ops.attn_fwd_hard(Qn, ops.hard_bank_layout(yk[lo:hi].int(), c).gather(Kn[lo:hi]), 5.5))
    torch.testing.assert_close(ops.merge_partials(torch.stack(hparts)), Oh, rtol=1e-4, atol=1e-4)
    Oh0 = ops.attn_fwd_hard(Qn[:128].contiguous(), bank, 0.0)
    assert torch.equal(Oh0, torch.bincount(yk, minlength=c).float().expand(128, c))
