"""PyTorch / timm ``state_dict`` -> engine weights.

Same conversion contract as the reference's ``tfimm/utils/timm.py:39-229``:

* name: drop ``remove/`` path segments, ``/`` -> ``.``, then ``kernel`` / ``depthwise_kernel`` / ``gamma`` ->
  ``weight``, ``beta`` -> ``bias``, ``moving_mean`` -> ``running_mean``, ``moving_variance`` -> ``running_var``;
* layout: rank-4 kernels are ``weight.permute(2, 3, 1, 0)`` (depthwise ``(C,1,k,k)`` then reshapes to
  ``(k,k,C,1)``), rank-2 kernels are ``weight.T``; a missing / extra leading axis is squeezed / expanded and a
  final reshape reconciles same-size shapes (e.g. ``cls_token``, ECA's Conv1d kernel);
* non-trainable buffers rebuilt at construction (``keys_to_ignore_on_load_missing``) are skipped;
* a weight with no counterpart raises ``AttributeError`` unless ``allow_missing_keys``.

Downloading from timm / the HF hub needs a network and timm; only the local conversion is provided.
"""
import re
from typing import Dict, List, Tuple

import numpy as np
import torch

_LEAF = {"kernel": "weight", "depthwise_kernel": "weight", "gamma": "weight", "embeddings": "weight",
         "beta": "bias", "moving_mean": "running_mean", "moving_variance": "running_var"}


def pytorch_key(weight_key: str) -> str:
    """``"remove/fc/kernel"`` -> ``"fc.weight"``, ``"blocks.0.1/bn1/moving_mean"`` -> ``"blocks.0.1.bn1.running_mean"``."""
    parts = [p for p in weight_key.split("/") if p and p != "remove"]
    parts[-1] = _LEAF.get(parts[-1], parts[-1])
    return ".".join(parts)


def _to_engine_layout(array: np.ndarray, key: str, target_shape: Tuple[int, ...]) -> np.ndarray:
    """Layout rule of the reference (tfimm/utils/timm.py:76-90, 164-188): the TRANSPOSE is chosen from the rank of the
    TF-side variable -- rank-4 ``kernel`` / ``depthwise_kernel`` take ``(2, 3, 1, 0)`` of the PyTorch tensor (CONV2D),
    every other kernel takes the full axis reversal (SIMPLE, so a conv-as-linear ``(out, in, 1, 1)`` checkpoint
    becomes ``(1, 1, in, out)`` and squeezes to ``(in, out)``) -- then a missing / extra leading axis is expanded /
    squeezed and a final same-size reshape reconciles e.g. ``cls_token`` or depthwise ``(1, k, k, C) -> (k, k, C, 1)``."""
    leaf = key.rsplit("/", 1)[-1]
    if leaf in ("kernel", "depthwise_kernel"):
        if len(target_shape) == 4:
            array = np.transpose(array, (2, 3, 1, 0))
        else:
            array = np.transpose(array)
    if array.ndim > len(target_shape):
        array = np.squeeze(array)
    elif array.ndim < len(target_shape):
        array = np.expand_dims(array, axis=0)
    if tuple(array.shape) != tuple(target_shape):
        if array.size != int(np.prod(target_shape)):
            raise ValueError(f"{key}: PyTorch tensor {array.shape} cannot become {tuple(target_shape)}")
        array = np.reshape(array, target_shape)
    return np.ascontiguousarray(array)


def convert_state_dict(model, state_dict: Dict[str, object], allow_missing_keys: bool = False):
    """Returns ``(weights, missing, unexpected)``: a flat dict in the model's own names/layouts plus bookkeeping."""
    sd = {}
    for k, v in state_dict.items():
        if k.endswith(".gamma") and k[:-6] + ".weight" not in state_dict:
            k = k[:-6] + ".weight"  # old-style LayerNorm parameter names (reference timm.py:119-135)
        elif k.endswith(".beta") and k[:-5] + ".bias" not in state_dict:
            k = k[:-5] + ".bias"
        sd[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    ignore = list(getattr(model, "keys_to_ignore_on_load_missing", []))
    weights, missing, used = {}, [], set()
    for key, cur in model.params.items():
        if any(re.search(pat, key) is not None for pat in ignore):
            continue
        name = pytorch_key(key)
        if name not in sd:
            if allow_missing_keys:
                missing.append(name)
                continue
            raise AttributeError(f"{name} not found in PyTorch model")
        weights[key] = _to_engine_layout(sd[name], key, tuple(cur.shape))
        used.add(name)
    unexpected: List[str] = [k for k in sd if k not in used and "num_batches_tracked" not in k]
    return weights, missing, unexpected


def load_pytorch_weights_in_model(model, state_dict, allow_missing_keys: bool = False):
    """In-place load of a PyTorch ``state_dict`` (timm naming) into an engine model."""
    weights, missing, unexpected = convert_state_dict(model, state_dict, allow_missing_keys)
    model.load_weights_dict(weights, strict=False)
    return missing, unexpected
