"""Global model registry: name -> (model class, config).

Behavioural contract (reference tfimm/models/registry.py:27-159):
  * ``register_model(fn)`` calls ``fn()`` immediately to get ``(cls, cfg)``, requires
    ``fn.__name__ == cfg.name`` (``ValueError`` otherwise), stores a deep copy of ``cfg``,
    records the defining module's stem, and appends the name to that module's ``__all__``.
  * ``list_models`` filters with fnmatch include / exclude patterns, optionally by module or by
    "has pretrained url", and returns names in natural sort order.
"""
import fnmatch
import re
import sys
from copy import deepcopy
from typing import Dict, List, Set, Union

__all__ = [
    "list_models",
    "is_model",
    "is_model_in_modules",
    "is_model_pretrained",
    "list_modules",
    "model_class",
    "model_config",
    "register_model",
]

_classes: Dict[str, type] = {}
_configs: Dict[str, object] = {}
_by_module: Dict[str, Set[str]] = {}
_with_url: Set[str] = set()


def register_model(fn):
    cls, cfg = fn()
    if fn.__name__ != cfg.name:
        raise ValueError(f"Model name({cfg.name}) != function name ({fn.__name__}).")
    _register(cfg.name, cls, cfg, fn.__module__)
    return fn


def _register(name, cls, cfg, module_path):
    """Registration core, also used by the table-driven zoos in ``tfimm.architectures``."""
    module = sys.modules.get(module_path)
    stem = module_path.rsplit(".", 1)[-1]
    if module is not None:
        exported = getattr(module, "__all__", None)
        if exported is None:
            module.__all__ = [name]
        elif name not in exported:
            exported.append(name)
    _classes[name] = cls
    _configs[name] = deepcopy(cfg)
    _by_module.setdefault(stem, set()).add(name)
    if cfg.url:
        _with_url.add(name)


def _natural_key(s: str):
    return [int(tok) if tok.isdigit() else tok for tok in re.split(r"(\d+)", s.lower())]


def _as_list(x) -> List[str]:
    if not x:
        return []
    return list(x) if isinstance(x, (tuple, list)) else [x]


def list_models(
    name_filter: Union[str, List[str]] = "",
    module: str = "",
    pretrained: Union[bool, str] = False,
    exclude_filters: Union[str, List[str]] = "",
) -> List[str]:
    """Names of registered models, naturally sorted.

    Args:
        name_filter: fnmatch pattern(s); a model is kept if it matches any of them.
        module: restrict to models defined in this architecture module (e.g. ``"resnet"``).
        pretrained: ``True`` keeps only models whose config has a non-empty ``url``;
            ``"timm"`` intersects with ``timm.list_models(pretrained=True)`` (needs timm).
        exclude_filters: fnmatch pattern(s) removed after inclusion.
    """
    pool = list(_by_module.get(module, ())) if module else list(_classes)
    includes = _as_list(name_filter)
    if includes:
        selected = set()
        for pat in includes:
            selected.update(fnmatch.filter(pool, pat))
    else:
        selected = set(pool)
    for pat in _as_list(exclude_filters):
        selected.difference_update(fnmatch.filter(selected, pat))
    if pretrained is True:
        selected &= _with_url
    elif pretrained == "timm":
        import timm  # optional dependency, imported lazily exactly like the reference

        selected &= set(timm.list_models(pretrained=True))
    return sorted(selected, key=_natural_key)


def is_model(model_name: str) -> bool:
    return model_name in _classes


def model_class(model_name: str):
    return _classes[model_name]


def model_config(model_name: str):
    return _configs[model_name]


def list_modules() -> List[str]:
    return sorted(_by_module)


def is_model_in_modules(model_name: str, module_names) -> bool:
    assert isinstance(module_names, (tuple, list, set))
    return any(model_name in _by_module.get(m, ()) for m in module_names)


def is_model_pretrained(model_name: str) -> bool:
    return model_name in _with_url
