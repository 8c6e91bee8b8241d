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
