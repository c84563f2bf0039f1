"""Table-driven registration: each family's registrations are data (``zoo/<family>.json``,
generated from the reference's own registry by ``tools/extract_zoo.py``).  For every entry we
register ``(model class, config)`` and expose an entry-point function of the same name in the
family module, exactly what ``@register_model`` produces in the reference
(tfimm/models/registry.py:34-59)."""
import json
import sys
from pathlib import Path

from ..models.registry import _register

_ZOO_DIR = Path(__file__).resolve().parent / "zoo"


def _tuplify(v):
    if isinstance(v, list):
        return tuple(_tuplify(x) for x in v)
    return v


def load_table(family: str):
    with open(_ZOO_DIR / f"{family}.json") as f:
        raw = json.load(f)
    return {
        name: {k: _tuplify(v) for k, v in fields.items() if not k.startswith("__")}
        for name, fields in raw.items()
    }


def register_zoo(module_name: str, family: str, model_cls, cfg_cls):
    module = sys.modules[module_name]
    for name, fields in load_table(family).items():
        cfg = cfg_cls(**fields)

        def entry(_cls=model_cls, _cfg=cfg):
            from copy import deepcopy

            return _cls, deepcopy(_cfg)

        entry.__name__ = entry.__qualname__ = name
        entry.__module__ = module_name
        entry.__doc__ = f"Entry point for ``{name}``; returns ``(model class, config)``."
        setattr(module, name, entry)
        _register(name, model_cls, cfg, module_name)
