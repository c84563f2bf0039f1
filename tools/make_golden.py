"""Writes the committed golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Each fixture = one configuration of an in-scope family: seeded inputs (the reference's own generator,
tests/test_timm.py:56-59), seeded fully-random parameters (oracle/params.py) and the float32 logits computed by the
UNMODIFIED reference code (``/root/reference/tfimm``: ``create_model`` + the model's ``call()``) executed on the
torch-CPU TensorFlow shim (``oracle/ref_runner.py`` / ``oracle/tf_shim``; TensorFlow itself cannot be installed in
this image).  The oracle's own logits for the same case are compared on the spot and the agreement is stored in the
fixture's meta (``oracle_vs_reference``).

Two groups:
  * small configurations: images are stored too (the GPU box has no /root/reference and no need to regenerate);
  * the BASELINE.json configurations at full size (vit_base_patch16_224, convnext_base,
    swin_base_patch4_window7_224, efficientnet_b4 @ 380, resnet50): 4 images, only the logits are stored -- images
    and parameters are regenerated from their seeds (``images=None`` in the file).

The CPU suite checks that the oracle reproduces every fixture; the GPU suite checks the engine against them.
Parameters are regenerated from the seed, never stored.  Run from the repo root (needs /root/reference):

    python tools/make_golden.py
"""
import importlib
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tensorflow-image-models_b200"))

import tfimm  # noqa: E402
from oracle import params  # noqa: E402
from oracle import ref_runner  # noqa: E402

IGNORE = ("attn_mask", "relative_position_index", "blur_kernel")

SMALL = [
    ("vit", "vit_tiny_patch16_224", {"input_size": (64, 64), "nb_blocks": 3}, 2, 11),
    ("vit", "deit_tiny_distilled_patch16_224", {"input_size": (48, 48), "nb_blocks": 2}, 2, 12),
    ("swin", "swin_tiny_patch4_window7_224", {"input_size": (112, 112), "nb_blocks": (2, 2), "nb_heads": (3, 6)}, 1, 13),
    ("convnext", "convnext_tiny", {"input_size": (64, 64), "nb_blocks": (1, 1, 2, 1)}, 2, 14),
    ("efficientnet", "efficientnet_b0", {"input_size": (64, 64)}, 2, 15),
    ("efficientnet", "pt_efficientnet_b0", {"input_size": (64, 80)}, 1, 16),
    ("resnet", "resnet18", {"input_size": (64, 64)}, 2, 17),
    ("resnet", "seresnext26d_32x4d", {"input_size": (64, 64)}, 1, 18),
]
# BASELINE.json configs at their own input size; file name carries a "full_" prefix
FULL = [
    ("vit", "vit_base_patch16_224", {}, 4, 21),
    ("convnext", "convnext_base", {}, 4, 22),
    ("swin", "swin_base_patch4_window7_224", {}, 4, 23),
    ("efficientnet", "efficientnet_b4", {}, 4, 24),
    ("resnet", "resnet50", {}, 4, 25),
]


def _case(family, model, overrides, batch, seed, store_images):
    mod = importlib.import_module(f"oracle.{family}")
    base = tfimm.models.model_config(model)
    cfg = type(base)(**{**base.__dict__, **overrides})
    w = params.random_params(mod.param_shapes(cfg), seed=seed)
    x = params.test_images(batch, *cfg.input_size, cfg.in_channels)
    ref = ref_runner.create_model(model, **overrides)
    ref.assign(w, ignore_missing=IGNORE)
    y_ref = ref(x)
    with torch.no_grad():
        y_or = mod.forward(cfg, w, x)
    agree = (y_or - y_ref).abs().max().item() / (y_ref.abs().max().item() + 1e-6)
    assert agree < 5e-6, (model, agree)
    meta = {"family": family, "model": model, "overrides": overrides, "batch": batch, "seed": seed,
            "images_seed": 2021, "source": "reference tfimm code executed on oracle/tf_shim (float32)",
            "oracle_vs_reference": agree}
    return (x.numpy() if store_images else None), y_ref.numpy().astype(np.float32), meta


def main():
    if not ref_runner.available():
        raise SystemExit("/root/reference is needed to regenerate the fixtures")
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    for group, prefix, store in ((SMALL, "", True), (FULL, "full_", False)):
        for family, model, overrides, batch, seed in group:
            images, logits, meta = _case(family, model, overrides, batch, seed, store)
            path = out_dir / f"{prefix}{model}.npz"
            arrays = {"logits": logits, "meta": np.array(meta, dtype=object)}
            if images is not None:
                arrays["images"] = images
            np.savez_compressed(path, **arrays)
            print(path.name, tuple(logits.shape), f"max|logit|={float(np.abs(logits).max()):.3f}",
                  f"oracle vs reference {meta['oracle_vs_reference']:.2e}", f"{path.stat().st_size / 1024:.0f} KB")


if __name__ == "__main__":
    main()
