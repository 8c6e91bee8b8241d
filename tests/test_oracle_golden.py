"""The oracle restatement vs outputs of the reference's own code (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import clip_search_oracle as orc


@pytest.fixture(scope="module")
def sel(golden_dir):
    return np.load(golden_dir / "selection.npz")


@pytest.fixture(scope="module")
def att(golden_dir):
    return np.load(golden_dir / "image_attention.npz")


@pytest.fixture(scope="module")
def tip(golden_dir):
    return np.load(golden_dir / "tip_adapter.npz")


def test_row_confidence_matches_reference(sel):
    outs = torch.from_numpy(sel["image_outs"])
    conf, label = orc.row_confidence(outs, prob=False)
    assert np.array_equal(label.numpy(), sel["label"])
    assert np.array_equal(conf.numpy(), sel["conf_raw"])
    conf_p, label_p = orc.row_confidence(outs, prob=True)
    assert np.array_equal(label_p.numpy(), sel["label"])
    assert np.array_equal(conf_p.numpy(), sel["conf_prob"])


@pytest.mark.parametrize("k", [1, 4, 16, 64])
def test_topk_selection_bit_exact(sel, k):
    outs = torch.from_numpy(sel["image_outs"])
    assert np.array_equal(orc.topk_select(outs, k), sel[f"topk_{k}"])
    assert np.array_equal(orc.topk_prob_select(outs, k), sel[f"topk_prob_{k}"])


def test_selection_edge_cases(sel):
    outs = torch.from_numpy(sel["image_outs"])
    idx = orc.topk_select(outs, 16)
    labels = sel["label"][idx]
    # classes ascending, empty predicted classes absent, rare classes contribute what they have
    assert np.all(np.diff(labels) >= 0)
    assert not set(labels.tolist()) & {4, 9, 17}
    counts = np.bincount(sel["label"], minlength=23)
    for c in range(23):
        assert (labels == c).sum() == min(16, counts[c])
    assert np.array_equal(orc.all_logits_select(outs), sel["all_logits"])


def test_tie_policy_is_value_then_index():
    labels = np.array([0, 0, 0, 0, 1, 1])
    conf = np.array([0.5, 0.7, 0.7, 0.1, 0.2, 0.2], dtype=np.float32)
    assert orc.select_topk_per_label(labels, conf, 2).tolist() == [1, 2, 4, 5]


def test_zero_shot_logits_and_accuracy(att):
    Q = torch.from_numpy(att["test_image_features"])
    T = torch.from_numpy(att["text_features"])
    Z = orc.zero_shot_logits(Q, T)
    assert np.array_equal(Z.numpy(), att["clip_logits"])
    labels = torch.from_numpy(att["test_labels"])
    assert np.allclose(orc.compute_accuracy(Z, labels), att["acc_zero_shot"])


def test_weights_and_values_match_reference(att):
    Q = torch.from_numpy(att["test_image_features"])
    K = torch.from_numpy(att["cache_image_features"])
    L = torch.from_numpy(att["cache_image_outs"])
    idx = torch.from_numpy(att["cache_idx"])
    W = orc.tip_weights(orc.normalize_columns(Q), orc.normalize_columns(K[:, idx]), 5.5)
    assert np.array_equal(W.numpy(), att["weights_b5.5"])
    Lc = L[idx]
    assert np.array_equal(orc.hard_values(Lc).numpy(), att["values_0"])
    assert np.array_equal(orc.softmax_values(Lc, orc.CLIP_SCALE, 0.1).numpy(), att["values_1"])
    assert np.array_equal(orc.softmax_values(Lc, orc.CLIP_SCALE, 10.0).numpy(), att["values_2"])


def test_image_attention_sweep_matches_reference(att):
    Q = torch.from_numpy(att["test_image_features"])
    K = torch.from_numpy(att["cache_image_features"])
    L = torch.from_numpy(att["cache_image_outs"])
    Z = torch.from_numpy(att["clip_logits"])
    labels = torch.from_numpy(att["test_labels"])
    idx = torch.from_numpy(att["cache_idx"])
    Kc, Lc = K[:, idx], L[idx]
    values = [orc.hard_values(Lc), orc.softmax_values(Lc, orc.CLIP_SCALE, 0.1), orc.softmax_values(Lc, orc.CLIP_SCALE, 10.0)]
    for vi, V in enumerate(values):
        for bi, beta in enumerate(att["betas"]):
            # chunk smaller than Nq: the chunked driver must not change a single bit per row
            O = orc.image_attention(Q, Kc, V, float(beta), chunk=64)
            ref = att[f"cache_logits_v{vi}_b{bi}"]
            assert np.allclose(O.numpy(), ref, rtol=1e-6, atol=1e-7), (vi, bi)
            accs = [orc.compute_accuracy(orc.searcher_logits(Z, torch.from_numpy(ref), float(a)), labels) for a in att["alphas"]]
            assert np.allclose(np.array(accs), att[f"acc_v{vi}_b{bi}"])
    O_all = orc.image_attention(Q, K, orc.hard_values(L), 5.5)
    assert np.allclose(O_all.numpy(), att["cache_logits_all_hard_b5.5"], rtol=1e-6, atol=1e-7)


def test_tip_adapter_head_and_search(tip):
    feats = torch.from_numpy(tip["features"])
    keys = torch.from_numpy(tip["cache_keys"])
    vals = orc.onehot_values(torch.from_numpy(tip["cache_labels"]), 11)
    clip_w = torch.from_numpy(tip["clip_weights"])
    labels = torch.from_numpy(tip["test_labels"])
    out = orc.tip_head(feats, keys, vals, clip_w, 5.5, 1.0)
    assert np.allclose(out.numpy(), tip["tip_logits"], rtol=1e-6, atol=1e-6)
    assert orc.cls_acc(out, labels) == pytest.approx(float(tip["acc_tip"]))
    bb, ba, _ = orc.search_hp(tip["search_scale"].tolist(), tip["search_step"].tolist(), keys, vals, feats, labels, clip_w)
    assert bb == pytest.approx(float(tip["best_beta"])) and ba == pytest.approx(float(tip["best_alpha"]))


def test_synthetic_banks_are_seeded_and_feature_major():
    a = orc.synthetic_banks(10, 20, 16, 5, seed=3)
    b = orc.synthetic_banks(10, 20, 16, 5, seed=3)
    assert a["test_image_features"].shape == (16, 10) and a["cache_image_outs"].shape == (20, 5)
    for k in a:
        assert torch.equal(a[k], b[k])


def test_round2_goldens_gold_replace_and_tip_tail(golden_dir):
    """tests/golden/round2.npz (reference's build_cache with replace_outs_with_golds; build_cache_model tail)."""
    r2 = np.load(golden_dir / "round2.npz")
    Q, K, L, T = (torch.from_numpy(r2[f"gr_{n}"]) for n in ("test_image_features", "cache_image_features", "cache_image_outs", "text_features"))
    gold, labels = torch.from_numpy(r2["gr_cache_labels"]), torch.from_numpy(r2["gr_test_labels"]).long()
    idx = torch.from_numpy(orc.topk_select(L, 4))
    assert idx.numel() == int(r2["gr_info"][0])
    outs = orc.golds_as_outs(gold[idx], L.shape[1])
    assert np.array_equal(outs.float().numpy(), r2["gr_outs_replaced"])
    acc = orc.compute_accuracy(L[idx], gold[idx].long())
    assert np.allclose(acc, r2["gr_info"][1:3]) and orc.compute_accuracy(outs.float(), gold[idx].long()) == [100.0, 100.0]
    Z = orc.zero_shot_logits(Q, T)
    for vi, values in enumerate((orc.hard_values(outs), orc.softmax_values(outs, orc.CLIP_SCALE, 0.1), orc.softmax_values(outs, orc.CLIP_SCALE, 10.0))):
        np.testing.assert_allclose(values.float().numpy(), r2[f"gr_values_{vi}"], atol=1e-6)
        O = orc.image_attention(Q, K[:, idx], values.float(), 5.5)
        np.testing.assert_allclose(O.numpy(), r2[f"gr_cache_logits_{vi}"], rtol=2e-5, atol=2e-5)
        accs = np.array([orc.compute_accuracy(orc.searcher_logits(Z, O, a), labels) for a in (0.5, 1.0, 4.0)])
        np.testing.assert_allclose(accs, r2[f"gr_acc_{vi}"], atol=1e-9)
    keys = orc.tip_cache_keys(torch.from_numpy(r2["tip_train_features"]))
    assert np.array_equal(keys.contiguous().numpy(), r2["tip_cache_keys"])
    assert np.array_equal(orc.tip_normalize_rows(torch.from_numpy(r2["tip_test_features"])).numpy(), r2["tip_test_f"])
    f32, k32 = torch.from_numpy(r2["tip_test_f"]).float(), torch.from_numpy(r2["tip_cache_keys"]).float()
    v32, w32 = torch.from_numpy(r2["tip_cache_values"]).float(), torch.from_numpy(r2["tip_clip_weights"]).float()
    tl = torch.from_numpy(r2["tip_test_labels"])
    np.testing.assert_allclose(orc.tip_head(f32, k32, v32, w32, 5.5, 1.0).numpy(), r2["tip_logits"], rtol=1e-5, atol=1e-4)
    assert orc.cls_acc(orc.tip_head(f32, k32, v32, w32, 5.5, 1.0), tl) == float(r2["tip_acc"])
    bb, ba, _ = orc.search_hp([7, 3], [20, 5], k32, v32, f32, tl, w32)
    assert np.allclose([bb, ba], r2["tip_best"])
