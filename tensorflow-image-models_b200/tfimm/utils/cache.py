"""Model cache locations (API parity with the reference's tfimm/utils/cache.py:1-94).

Resolution order for the cache root: ``set_dir`` > ``$TFIMM_HOME`` > ``$XDG_CACHE_HOME/tfimm``
> ``~/.cache/tfimm``.  Per-model overrides registered with ``set_model_cache`` win over the
directory lookup.
"""
import os
from typing import Dict, List, Optional

_root_override: Optional[str] = None
_per_model: Dict[str, str] = {}


def get_dir() -> str:
    if _root_override is not None:
        return _root_override
    xdg = os.getenv("XDG_CACHE_HOME", "~/.cache")
    home = os.getenv("TFIMM_HOME", os.path.join(xdg, "tfimm"))
    return os.path.expanduser(home)


def set_dir(d: str):
    global _root_override
    _root_override = d


def set_model_cache(model_name: str, model_path: str):
    _per_model[model_name] = model_path


def clear_model_cache(model_name: str):
    _per_model.pop(model_name, None)


def list_cached_models() -> List[str]:
    return sorted(_per_model)


def cached_model_path(model_name: str) -> Optional[str]:
    if model_name in _per_model:
        return _per_model[model_name]
    candidate = os.path.join(get_dir(), model_name)
    return candidate if os.path.exists(candidate) else None
