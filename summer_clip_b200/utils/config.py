"""YAML config loading with OmegaConf-style `${a.b.c}` interpolation (the subset the hot-path
configs use: conf/image_attention.yaml:21-44, conf/tip_adapter*.yaml)."""
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


def load_config(path: tp.Union[str, Path], overrides: tp.Optional[tp.Mapping] = None) -> Config:
    with open(path) as f:
        raw = yaml.safe_load(f) or {}
    if overrides:
        merge(raw, overrides)
    return resolve(raw)
