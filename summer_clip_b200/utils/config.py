"""YAML config loading for the reference's Hydra config tree without hydra/omegaconf (neither is installable
here): OmegaConf-style `${a.b.c}` interpolation plus the part of Hydra's defaults-list composition the hot-path
configs use (conf/image_attention.yaml:1-19, conf/tip_adapter*.yaml:1-4, conf/img_attn_dataset/*.yaml:1-2):

  - name                      a config of the same directory, merged at the parent's package
  - group: option             conf/<group>/<option>.yaml (or the key <option> of conf/<group>.yaml) under the key `group`
  - group@pkg.path: option    ... placed under `pkg.path` (relative to the parent's package)
  - /group: option            group path taken from the config root instead of the parent's directory
  - _self_                    where the file's own keys merge (last when absent, Hydra >= 1.1)

so that the reference's own YAML files load unmodified (`compose("/…/summer_clip/conf", "image_attention")`)."""
from __future__ import annotations

import re
import typing as tp
from pathlib import Path

import yaml

_INTERP = re.compile(r"\$\{([^${}]+)\}")


class Config(dict):
    """dict with attribute access (cfg.cache.alpha) like DictConfig; `.get` works as usual."""

    def __getattr__(self, name: str) -> tp.Any:
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc

    def __setattr__(self, name: str, value: tp.Any) -> None:
        self[name] = value


def _wrap(node: tp.Any) -> tp.Any:
    if isinstance(node, dict):
        return Config({k: _wrap(v) for k, v in node.items()})
    if isinstance(node, list):
        return [_wrap(v) for v in node]
    return node


def _lookup(root: tp.Mapping, dotted: str) -> tp.Any:
    cur: tp.Any = root
    for part in dotted.strip().split("."):
        cur = cur[int(part)] if isinstance(cur, list) else cur[part]
    return cur


def _resolve(node: tp.Any, root: tp.Mapping, depth: int = 0) -> tp.Any:
    if depth > 32:
        raise ValueError("interpolation cycle")
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    if isinstance(node, str):
        whole = _INTERP.fullmatch(node)
        if whole:                                   # `${x}` alone keeps the referenced node's type
            return _resolve(_lookup(root, whole.group(1)), root, depth + 1)
        if _INTERP.search(node):
            return _INTERP.sub(lambda m: str(_resolve(_lookup(root, m.group(1)), root, depth + 1)), node)
    return node


def resolve(cfg: tp.Mapping) -> Config:
    return _wrap(_resolve(dict(cfg), cfg))


def merge(base: tp.MutableMapping, override: tp.Mapping) -> tp.MutableMapping:
    for k, v in override.items():
        if isinstance(v, dict) and isinstance(base.get(k), dict):
            merge(base[k], v)
        else:
            base[k] = v
    return base


def _set_at(root: dict, package: tp.Sequence[str], content: tp.Mapping) -> None:
    cur = root
    for key in package:
        nxt = cur.get(key)
        if not isinstance(nxt, dict):
            nxt = cur[key] = {}
        cur = nxt
    merge(cur, content)


def _load_node(conf_root: Path, rel: str) -> dict:
    """The mapping of config `rel`: conf_root/<rel>.yaml (Hydra's layout: one file per group option), or — this
    package's own conf/ keeps the options of a group together — the key <option> of conf_root/<group>.yaml."""
    path = conf_root / (rel + ".yaml")
    if path.exists():
        with open(path) as f:
            return yaml.safe_load(f) or {}
    group, _, option = rel.rpartition("/")
    pack = conf_root / (group + ".yaml")
    if group and pack.exists():
        with open(pack) as f:
            options = yaml.safe_load(f) or {}
        if option in options:
            return dict(options[option] or {})
    raise FileNotFoundError(f"config {rel!r} not found under {conf_root}")


def _compose_file(conf_root: Path, rel: str, package: tp.List[str], out: dict, choices: tp.Mapping[str, str]) -> None:
    """Merge conf_root/<rel>.yaml (and, recursively, its defaults list) into `out` at `package`."""
    content = _load_node(conf_root, rel)
    defaults = content.pop("defaults", None) or []
    if "_self_" not in defaults:
        defaults = list(defaults) + ["_self_"]
    parent_dir = rel.rsplit("/", 1)[0] if "/" in rel else ""
    for item in defaults:
        if item == "_self_":
            _set_at(out, package, content)
            continue
        if isinstance(item, str):                                   # a sibling config, same package
            _compose_file(conf_root, f"{parent_dir}/{item}" if parent_dir else item, package, out, choices)
            continue
        (key, option), = item.items()
        key = str(key)
        if key.startswith("override "):
            key = key[len("override "):]
        option = choices.get(key, option)
        if option is None:
            continue
        group, _, pkg = key.partition("@")
        if group.startswith("/"):
            group_dir = group.lstrip("/")
        else:
            group_dir = f"{parent_dir}/{group}" if parent_dir else group
        if pkg == "_global_":
            sub_package: tp.List[str] = []
        elif pkg:
            sub_package = package + pkg.split(".")
        else:
            sub_package = package + group.lstrip("/").split("/")
        _compose_file(conf_root, f"{group_dir}/{option}", sub_package, out, choices)


def compose(conf_root: tp.Union[str, Path], config_name: str, overrides: tp.Optional[tp.Sequence[str]] = None) -> Config:
    """hydra.compose for the subset above.  `overrides`: `a.b=value` sets a key after composition (value parsed as
    YAML); `group=option` / `group@pkg=option` picks another option of a defaults-list entry of the primary config
    (e.g. `cache_value_strategy=softmax_cache`, `img_attn_dataset@dataset_cfg=imagenet`)."""
    conf_root = Path(conf_root)
    with open(conf_root / (config_name + ".yaml")) as f:
        primary = yaml.safe_load(f) or {}
    group_keys = {str(next(iter(d))) for d in (primary.get("defaults") or []) if isinstance(d, dict)}
    choices: tp.Dict[str, str] = {}
    sets: tp.List[tp.Tuple[str, tp.Any]] = []
    for item in overrides or []:
        key, _, value = item.partition("=")
        key = key.lstrip("+")
        if key in group_keys:
            choices[key] = value
        else:
            sets.append((key, yaml.safe_load(value)))
    out: dict = {}
    _compose_file(conf_root, config_name, [], out, choices)
    for key, value in sets:
        parts = key.split(".")
        _set_at(out, parts[:-1], {parts[-1]: value})
    return resolve(out)


def split_overrides(items: tp.Sequence[str]) -> tp.List[str]:
    """Command-line `key=value` items as given to hydra (kept as strings for `compose`)."""
    return [it for it in items if "=" in it]


def load_config(path: tp.Union[str, Path], overrides: tp.Optional[tp.Mapping] = None) -> Config:
    path = Path(path)
    with open(path) as f:
        raw = yaml.safe_load(f) or {}
    if "defaults" in raw:                      # a Hydra primary config: compose it from its own directory
        raw = {}
        _compose_file(path.parent, path.stem, [], raw, {})
    if overrides:
        merge(raw, overrides)
    return resolve(raw)
