"""Writes the committed golden fixtures under tests/golden/.

Each fixture = one small configuration of an in-scope family: seeded inputs (the reference's own
generator, tests/test_timm.py:56-59), seeded fully-random parameters (oracle/params.py) and the fp32 logits
of the CPU oracle.  The oracle itself is pinned against torchvision in tests/test_oracle_cpu.py; these files
freeze its output so that (a) an accidental change to the oracle is caught on CPU and (b) the GPU parity
tests can run against stored numbers.  Parameters are regenerated from the seed, only inputs and logits are
stored.  Run from the repo root:  python tools/make_golden.py
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

CASES = [
    ("vit", "vit_tiny_patch16_224", {"input_size": (64, 64), "nb_blocks": 3}, 2, 11),
    ("vit", "deit_tiny_distilled_patch16_224", {"input_size": (48, 48), "nb_blocks": 2}, 2, 12),
    ("swin", "swin_tiny_patch4_window7_224", {"input_size": (112, 112), "nb_blocks": (2, 2), "nb_heads": (3, 6)}, 1, 13),
    ("convnext", "convnext_tiny", {"input_size": (64, 64), "nb_blocks": (1, 1, 2, 1)}, 2, 14),
    ("efficientnet", "efficientnet_b0", {"input_size": (64, 64)}, 2, 15),
    ("efficientnet", "pt_efficientnet_b0", {"input_size": (64, 80)}, 1, 16),
    ("resnet", "resnet18", {"input_size": (64, 64)}, 2, 17),
    ("resnet", "seresnext26d_32x4d", {"input_size": (64, 64)}, 1, 18),
]


def main():
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    for family, model, overrides, batch, seed in CASES:
        mod = importlib.import_module(f"oracle.{family}")
        base = tfimm.models.model_config(model)
        cfg = type(base)(**{**base.__dict__, **overrides})
        w = params.random_params(mod.param_shapes(cfg), seed=seed)
        x = params.test_images(batch, *cfg.input_size, cfg.in_channels)
        with torch.no_grad():
            y = mod.forward(cfg, w, x)
        meta = {"family": family, "model": model, "overrides": overrides, "batch": batch, "seed": seed}
        path = out_dir / f"{model}.npz"
        np.savez_compressed(path, images=x.numpy(), logits=y.numpy(), meta=np.array(meta, dtype=object))
        print(path.name, tuple(y.shape), f"max|logit|={float(y.abs().max()):.3f}", f"{path.stat().st_size / 1024:.0f} KB")


if __name__ == "__main__":
    main()
