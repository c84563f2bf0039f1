"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference (``/root/reference/tfimm``) on CPU.

TensorFlow is not installable in this image, so ``oracle/tf_shim/tensorflow`` (a torch-CPU restatement of the
TF/Keras calls the reference makes) is put first on ``sys.path`` and the reference's own modules are imported as
they are: ``tfimm.models.factory.create_model`` builds the reference's Keras model classes
(``tfimm/architectures/{vit,swin,convnext,efficientnet,resnet}.py``) and their ``call()`` code executes op for op.
This is what pins ``oracle/*.py`` (and the committed ``tests/golden`` fixtures) to the reference itself.

``/root/reference`` exists only in the build container: callers check ``available()`` first (the GPU box runs
against the committed fixtures instead).  The reference package is called ``tfimm`` like the engine package, so the
two cannot be imported side by side: ``_reference_modules()`` swaps ``sys.modules`` / ``sys.path`` for the duration
of a reference call (``tfimm`` = reference, ``tensorflow`` = shim inside the block; the engine outside of it).
Only ``tests/`` and ``tools/`` import this module.
"""
import importlib
import importlib.abc
import importlib.util
import sys
from contextlib import contextmanager
from pathlib import Path

import numpy as np
import torch

REFERENCE = Path("/root/reference")
SHIM = Path(__file__).resolve().parent / "tf_shim"
FAMILIES = ("vit", "swin", "convnext", "efficientnet", "resnet")

_loaded = None


def available() -> bool:
    return (REFERENCE / "tfimm" / "architectures" / "vit.py").exists()


@contextmanager
def _reference_modules():
    """Swap ``sys.modules`` so that ``tfimm`` / ``tensorflow`` mean the reference and the shim inside the block,
    and the engine package (also called ``tfimm``) outside of it."""
    global _loaded
    saved = {k: v for k, v in sys.modules.items()
             if k == "tfimm" or k.startswith("tfimm.") or k == "tensorflow" or k.startswith("tensorflow.")}
    for k in saved:
        del sys.modules[k]
    saved_path = list(sys.path)
    sys.path[:] = [str(SHIM)] + [p for p in sys.path if "tensorflow-image-models_b200" not in p]
    if _loaded is not None:
        sys.modules.update(_loaded)
    try:
        yield
    finally:
        _loaded = {k: v for k, v in sys.modules.items()
                   if k == "tfimm" or k.startswith("tfimm.") or k == "tensorflow" or k.startswith("tensorflow.")}
        for k in _loaded:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = saved_path


def _import_reference():
    """Imports the reference's registry, factory and the five in-scope architecture modules.  ``tfimm/__init__``
    and ``tfimm/architectures/__init__`` are bypassed (they import the ten out-of-scope families, the training
    framework and ``timm``); every module that IS imported is the reference's file, unmodified."""
    import types

    if "tfimm" not in sys.modules:
        pkg = types.ModuleType("tfimm")
        pkg.__path__ = [str(REFERENCE / "tfimm")]
        sys.modules["tfimm"] = pkg
        arch = types.ModuleType("tfimm.architectures")
        arch.__path__ = [str(REFERENCE / "tfimm" / "architectures")]
        sys.modules["tfimm.architectures"] = arch
        pkg.architectures = arch
    mods = {"factory": importlib.import_module("tfimm.models.factory"),
            "registry": importlib.import_module("tfimm.models.registry"),
            "timm": importlib.import_module("tfimm.utils.timm")}
    for fam in FAMILIES:
        mods[fam] = importlib.import_module(f"tfimm.architectures.{fam}")
    return mods


class ReferenceModel:
    """A reference Keras model instance built by the reference's ``create_model`` on the shim."""

    def __init__(self, model, mods):
        self.model, self._mods = model, mods
        self.cfg = model.cfg

    def weight_names(self):
        """Variable names as the reference's loaders see them: model-name prefix and ``:0`` stripped
        (tfimm/models/factory.py:294-305)."""
        strip = self._mods["factory"]._get_weight_name
        return [strip(w.name) for w in self.model.weights]

    def weight_shapes(self):
        strip = self._mods["factory"]._get_weight_name
        return {strip(w.name): tuple(w.shape) for w in self.model.weights}

    def weights_dict(self):
        strip = self._mods["factory"]._get_weight_name
        return {strip(w.name): w.numpy() for w in self.model.weights}

    def assign(self, weights, ignore_missing=()):
        """weights: {name: array} in reference (TF) layouts; every model variable must be given unless its name
        matches ``ignore_missing`` (the reference's keys_to_ignore_on_load_missing)."""
        import re

        strip = self._mods["factory"]._get_weight_name
        for w in self.model.weights:
            key = strip(w.name)
            if key not in weights:
                if any(re.search(p, key) for p in ignore_missing):
                    continue
                raise KeyError(f"no value for reference variable {key}")
            v = weights[key]
            w.assign(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))

    def __call__(self, x, return_features=False):
        with _reference_modules(), torch.no_grad():
            xin = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
            out = self.model(xin, training=False, return_features=return_features)
        conv = lambda t: t.as_subclass(torch.Tensor).detach().clone()  # noqa: E731
        if return_features:
            y, feats = out
            return conv(y), {k: conv(v) for k, v in feats.items()}
        return conv(out)


def set_floatx(name: str):
    with _reference_modules():
        import tensorflow as tf

        tf.keras.backend.set_floatx(name)


def create_model(model_name: str, **kwargs) -> ReferenceModel:
    """``tfimm.create_model(model_name, **kwargs)`` of the reference (tfimm/models/factory.py:18-125)."""
    with _reference_modules():
        mods = _import_reference()
        model = mods["factory"].create_model(model_name, **kwargs)
    return ReferenceModel(model, mods)


def list_models(module: str = ""):
    with _reference_modules():
        mods = _import_reference()
        return mods["registry"].list_models(module=module)


def create_preprocessing(model_name: str, **kwargs):
    with _reference_modules():
        mods = _import_reference()
        return mods["factory"].create_preprocessing(model_name, **kwargs)


def load_pytorch_weights(ref_model: ReferenceModel, state_dict):
    """The reference's own PyTorch -> TF conversion (tfimm/utils/timm.py:109-229) applied to ``ref_model``."""
    with _reference_modules():
        ref_model._mods["timm"].load_pytorch_weights_in_tf2_model(ref_model.model, state_dict)


def transfer_weights(src: ReferenceModel, dst: ReferenceModel):
    with _reference_modules():
        src._mods["factory"].transfer_weights(src.model, dst.model)
