"""Host-side mirror of the reference's plug-in plumbing (no GPU): the YAML + `${...}` config loader, the
`_target_` resolution / Cartesian expansion of `instantiate_all` (utils/hydra_utils.py:38-50), the JSON-lines
record schema, and the host-RNG cache strategies against outputs of the reference's own code
(tests/golden/strategies.npz, made by tests/golden/make_golden_strategies.py)."""
import json

import numpy as np
import pytest
import torch


def test_config_interpolation_and_overrides(tmp_path):
    from summer_clip_b200.utils.config import load_config
    (tmp_path / "c.yaml").write_text(
        "dataset_cfg:\n  name: sun397\n  root: /data/${dataset_cfg.name}\n"
        "data:\n  image_features_path: ${dataset_cfg.root}/test.pt\n  alpha: ${cache.alpha}\n"
        "cache:\n  alpha: [0.0, 1.0]\n  k: 4\n")
    cfg = load_config(tmp_path / "c.yaml", {"cache": {"k": 8}, "extra": {"x": "${cache.k}"}})
    assert cfg.data.image_features_path == "/data/sun397/test.pt"
    assert cfg.data.alpha == [0.0, 1.0] and cfg.cache.k == 8 and cfg.extra.x == 8        # `${x}` alone keeps the type
    assert cfg.get("missing") is None and cfg["cache"]["alpha"][1] == 1.0
    (tmp_path / "cyc.yaml").write_text("a: ${b}\nb: ${a}\n")
    with pytest.raises(ValueError):
        load_config(tmp_path / "cyc.yaml")


def test_instantiate_all_grid_order_and_reference_targets():
    from summer_clip_b200.clip_searcher.cache_strategy import TopKProbStrategy
    from summer_clip_b200.utils import hydra_utils
    cfg = {"_target_": "summer_clip.clip_searcher.cache_strategy.TopKProbStrategy", "topk": [1, 4], "scale": [100.0, 50.0]}
    got = list(hydra_utils.instantiate_all(cfg))
    assert [(p["topk"], p["scale"]) for _, p in got] == [(1, 100.0), (1, 50.0), (4, 100.0), (4, 50.0)]   # key order, last fastest
    assert all(isinstance(obj, TopKProbStrategy) for obj, _ in got)                   # reference target -> this package
    assert all(p["_target_"] == cfg["_target_"] for _, p in got)                      # records keep the reference's name
    assert got[2][0].topk == 4 and got[1][0].scale == 50.0
    assert hydra_utils.resolve_target("summer_clip.tip_adapter.utils.cls_acc") == "summer_clip_b200.tip_adapter.utils.cls_acc"
    assert hydra_utils.resolve_target("summer_clip_b200.x.Y") == "summer_clip_b200.x.Y"
    assert hydra_utils.type_full_name(TopKProbStrategy).endswith("cache_strategy.TopKProbStrategy")
    with pytest.raises(AttributeError):
        hydra_utils.load_obj("summer_clip_b200.clip_searcher.cache_strategy.NoSuchStrategy")


def test_json_lines_records(tmp_path):
    from summer_clip_b200.utils.log_utils import JsonLinesLogger
    log = JsonLinesLogger("image_attention", tmp_path / "run" / "image_attention.log")
    log.log_info("original-data-size: 19850")
    log.log_info(dict(acc1=61.5, acc5=88.0, type="zero_shot"))
    log.log_info_wandb(dict(alpha=1.0, acc1=70.0, type="searcher_result"))
    log.close()
    recs = [json.loads(l) for l in (tmp_path / "run" / "image_attention.log").read_text().splitlines()]
    assert [r.get("type") for r in recs] == [None, "zero_shot", "searcher_result"]
    assert recs[0]["message"] == "original-data-size: 19850" and recs[1]["message"] is None
    assert all({"asctime", "name", "levelname"} <= set(r) for r in recs) and recs[2]["alpha"] == 1.0


@pytest.fixture(scope="module")
def strat(golden_dir):
    return np.load(golden_dir / "strategies.npz")


@pytest.mark.parametrize("k", [1, 3, 60])
def test_host_rng_strategies_reproduce_the_reference_stream(strat, k):
    """GlobalRandomSampleStrategy / PerGoldClassRandomSampleStrategy draw from numpy's global RNG exactly like the
    reference (cache_strategy.py:103-141), so a seeded run selects the same cache."""
    from summer_clip_b200.clip_searcher.cache_strategy import GlobalRandomSampleStrategy, PerGoldClassRandomSampleStrategy
    outs = torch.from_numpy(strat["image_outs"])
    gold = torch.from_numpy(strat["gold_labels"])
    feats = torch.empty(1, outs.shape[0])
    np.random.seed(42)
    assert np.array_equal(GlobalRandomSampleStrategy(k).select(feats, outs).numpy(), strat[f"global_random_{k}"])
    np.random.seed(42)
    got = PerGoldClassRandomSampleStrategy(k, cache_labels=gold).select(feats, outs).numpy()
    assert np.array_equal(got, strat[f"per_gold_random_{k}"])


def test_hydra_defaults_composition_of_the_hot_path_config_tree():
    """utils/config.compose: the defaults list of conf/image_attention.yaml (sibling configs, group: option,
    group@package: option, nested /group, _self_), ${...} interpolation, key=value and group=option overrides."""
    from pathlib import Path
    from summer_clip_b200.utils.config import compose
    conf = Path(__file__).resolve().parent.parent / "summer_clip_b200" / "conf"
    cfg = compose(conf, "image_attention")
    assert list(cfg.cache_strategies) == ["topk", "topk_prob", "topk_per_gold", "topk_prob_per_gold", "per_pred_class_random",
                                          "per_gold_class_random", "global_random", "all_logits"]
    assert cfg.cache_value_strategy["_target_"].endswith("HardCacheStrategy")
    assert cfg.cache_weights_strategy.beta == [0.1, 1.0, 1.5, 3.5, 5.5, 7.5, 9.5, 11.5]
    assert cfg.dataset_name == "sun397" and cfg.prompting == cfg.dataset_cfg.prompting          # nested /prompting lands in dataset_cfg
    assert cfg.data.image_features_path == cfg.saved_paths.image_features["SUN397_tip_test-RN50"]
    assert cfg.cache.dataset.split == "train" and cfg.dataset.split == "test" and cfg.cache.dataset.dataset == "sun397"
    assert cfg.cache_strategies.topk_per_gold.cache_dataset == [cfg.cache.dataset]              # ${cache.dataset} keeps its type
    assert cfg.meta.random_state == 42 and cfg.cache.alpha[-1] == 4.0
    over = compose(conf, "image_attention", ["cache_value_strategy=softmax_cache", "img_attn_dataset@dataset_cfg=imagenet",
                                             "cache.alpha=[1.0]", "cache_strategies.topk.topk=[16]", "run_dir=/tmp/x"])
    assert over.cache_value_strategy.scale == [0.1, 1.0, 10.0, 20.0] and over.dataset_name == "imagenet"
    assert over.cache.image_outs_path == over.saved_paths.logits["ImageNet_train-RN50-tip_adapter"]
    assert over.cache.alpha == [1.0] and over.cache_strategies.topk.topk == [16] and over.run_dir == "/tmp/x"
    tip = compose(conf, "tip_adapter_imagenet", ["search_step=[20,5]"])
    assert tip.search_scale == [7, 3] and tip.search_step == [20, 5] and tip.init_beta == 5.5 and tip.shots == 16
    assert compose(conf, "tip_adapter").search_scale == [20, 10]


def test_reference_yaml_tree_composes_unmodified():
    """The reference's own conf/ directory (dev container only) through the same composer: same key structure as the
    package's tree apart from the four file inputs this path adds."""
    from pathlib import Path
    from summer_clip_b200.utils.config import compose
    ref = Path("/root/reference/summer_clip/conf")
    if not ref.exists():
        pytest.skip("reference tree not present (GPU box)")
    ours = compose(Path(__file__).resolve().parent.parent / "summer_clip_b200" / "conf", "image_attention")
    theirs = compose(ref, "image_attention")

    def keys(d, pre=""):
        out = set()
        for k, v in d.items():
            out.add(pre + k)
            if isinstance(v, dict) and not pre.startswith(("saved_paths", "hydra")):
                out |= keys(v, pre + k + ".")
        return out
    skip = ("saved_paths.", "hydra.")
    a = {k for k in keys(ours) if not k.startswith(skip)}
    b = {k for k in keys(theirs) if not k.startswith(skip)}
    assert b <= a and a - b == {"cache.labels_path", "data.clip_logits_path", "data.labels_path", "data.text_features_path"}
    assert theirs.cache_strategies.topk_prob == ours.cache_strategies.topk_prob
    assert theirs.cache.alpha == ours.cache.alpha and theirs.cache_weights_strategy == ours.cache_weights_strategy
    # the logits-bank producer's config (save_image_outs.py): the reference's keys plus the text classifier file
    ours = compose(Path(__file__).resolve().parent.parent / "summer_clip_b200" / "conf", "save_image_outs")
    theirs = compose(ref, "save_image_outs")
    a = {k for k in keys(ours) if not k.startswith(skip)}
    b = {k for k in keys(theirs) if not k.startswith(skip)}
    assert b <= a and a - b == {"data.text_features_path", "data.rows_per_chunk"}
    assert theirs.data.output_image_outs == ours.data.output_image_outs == "image_outs.pt" and theirs.exp == ours.exp
    for name in ("tip_adapter", "tip_adapter_imagenet"):
        t, o = compose(ref, name), compose(Path(__file__).resolve().parent.parent / "summer_clip_b200" / "conf", name)
        for k in ("search_hp", "search_scale", "search_step", "init_beta", "init_alpha", "dataset", "shots", "backbone", "augment_epoch"):
            assert t[k] == o[k], (name, k)


def test_entry_points_fail_loudly_without_a_cuda_device(tmp_path):
    """No CPU fallback anywhere on the product path: the three entry points refuse `meta.device=cpu`, and a machine
    without a GPU (this container) gets the library's own error, not a silent slow path."""
    import torch
    from summer_clip_b200._lib import SummerClipError
    from summer_clip_b200.clip_searcher import image_attention, save_image_outs
    from summer_clip_b200.tip_adapter import tip_adapter, tip_adapter_imagenet
    for mod in (image_attention, save_image_outs, tip_adapter, tip_adapter_imagenet):
        with pytest.raises(SummerClipError, match="no CPU fallback"):
            mod.run(["meta.device=cpu", f"run_dir={tmp_path}", "data.text_features_path=/nonexistent.pt"])
        if not torch.cuda.is_available():
            with pytest.raises(SummerClipError, match="no CPU fallback"):
                mod.run([f"run_dir={tmp_path}", "data.text_features_path=/nonexistent.pt"])


def test_command_line_of_the_entry_points(tmp_path):
    """`[CONFIG.yaml | --config-dir DIR --config-name NAME] [key=value ...]` like the reference's hydra entry points."""
    from pathlib import Path
    from summer_clip_b200.clip_searcher.image_attention import compose_from_argv
    conf = Path(__file__).resolve().parent.parent / "summer_clip_b200" / "conf"
    a = compose_from_argv(["cache.alpha=[2.0]"], "image_attention")
    b = compose_from_argv(["--config-dir", str(conf), "--config-name", "image_attention", "cache.alpha=[2.0]"], "ignored")
    c = compose_from_argv([str(conf / "image_attention.yaml"), "cache.alpha=[2.0]"], "ignored")
    assert a == b == c and a.cache.alpha == [2.0]
    assert compose_from_argv([], "save_image_outs").data.output_image_outs == "image_outs.pt"
    assert compose_from_argv(["search_step=[5,5]"], "tip_adapter").search_step == [5, 5]
