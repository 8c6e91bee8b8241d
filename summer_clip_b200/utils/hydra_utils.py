"""Hydra-free mirror of summer_clip/utils/hydra_utils.py.

The reference selects its strategy plug-ins by `_target_` strings in Hydra YAML and expands list
valued parameters into a Cartesian grid (`instantiate_all`, hydra_utils.py:38-50).  hydra/omegaconf
are not installable in this environment, so this module implements the same contract on plain
dict configs.  `_target_` strings that name the reference package (`summer_clip.…`) resolve to the
same class name inside this package, which is what makes the CUDA path a drop-in for the YAML.
"""
from __future__ import annotations

import copy
import importlib
import itertools
import typing as tp

_PREFIX_MAP = (("summer_clip.", "summer_clip_b200."),)


def load_obj(obj_path: str, default_obj_path: str = "") -> tp.Any:
    """The object a dotted path names (`pkg.mod.Name`; a bare `Name` is looked up in `default_obj_path`) — the
    reference's helper of the same name (hydra_utils.py:9-27)."""
    module_name, dot, attr = obj_path.rpartition(".")
    if not dot:
        module_name = default_obj_path
    module = importlib.import_module(module_name)
    try:
        return getattr(module, attr)
    except AttributeError:
        raise AttributeError(f"Object `{attr}` cannot be loaded from `{module_name}`.") from None


def type_full_name(type_: tp.Optional[type]) -> tp.Optional[str]:
    """`module.QualifiedName` of a class, bare for builtins (hydra_utils.py:30-36)."""
    if type_ is None:
        return None
    owner = getattr(type_, "__module__", None)
    return type_.__name__ if owner in (None, "builtins") else owner + "." + type_.__name__


def resolve_target(target: str) -> str:
    for old, new in _PREFIX_MAP:
        if target.startswith(old) and not target.startswith(new):
            return new + target[len(old):]
    return target


def instantiate(cfg: tp.Mapping[str, tp.Any], **overrides: tp.Any) -> tp.Any:
    """hydra.utils.instantiate for the flat configs this path uses (nested `_target_` dicts recurse)."""
    params = {k: v for k, v in cfg.items() if k != "_target_"}
    params.update(overrides)
    for k, v in list(params.items()):
        if isinstance(v, dict) and "_target_" in v:
            params[k] = instantiate(v)
    return load_obj(resolve_target(cfg["_target_"]))(**params)


def instantiate_all(cfg: tp.Mapping[str, tp.Any], inject: tp.Optional[tp.Mapping[str, tp.Tuple[str, tp.Any]]] = None) \
        -> tp.Iterator[tp.Tuple[tp.Any, tp.Dict[str, tp.Any]]]:
    """hydra_utils.py:38-50 — every non-`_target_` key holds a LIST; yield (instance, params) for the
    Cartesian product in key order.  The yielded params dict keeps the ORIGINAL `_target_` string and the
    ORIGINAL parameter values, so log records are identical to the reference's (notebooks map by class name).
    `inject` maps a config key to (constructor argument, object): the key's configured value is logged as is but
    the constructor receives the object under the other name — e.g. `cache_dataset: [${cache.dataset}]`
    (conf/cache_strategy/topk_per_gold.yaml:3-4) is a dataset description in the record and the loaded label
    tensor (`cache_labels`) for the strategy, without the tensor ever entering the JSON record or a deepcopy."""
    cfg_dict = copy.deepcopy(dict(cfg))
    target = cfg_dict.pop("_target_")
    inject = dict(inject or {})
    for param_values in itertools.product(*cfg_dict.values()):
        param_to_value = dict(zip(cfg_dict.keys(), param_values))
        ctor_args = {k: v for k, v in param_to_value.items() if k not in inject}
        ctor_args.update({new: obj for key, (new, obj) in inject.items() if key in param_to_value})
        yield _construct(target, ctor_args), {"_target_": target, **copy.deepcopy(param_to_value)}


def _construct(target: str, params: tp.Mapping[str, tp.Any]) -> tp.Any:
    """`instantiate` for arguments that are already objects (no recursion into dict values: an injected-for
    parameter such as a dataset description must not be instantiated)."""
    return load_obj(resolve_target(target))(**params)
