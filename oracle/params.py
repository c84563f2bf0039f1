"""Seeded "randomise everything" parameter generator for parity fixtures.

The reference's initialisers make whole branches inert (zeros cls/pos tokens, ConvNeXt gamma 1e-6,
ResNet last-BN gamma 0), and its conversion-test script warns that default-
initialised norm layers hide mismatches (scripts/test_conversion.py:33-35).  So fixtures draw
EVERY tensor at random, scaled by role so activations stay O(1) through deep stacks.
"""
import math

import numpy as np
import torch


def random_params(shapes, seed=0, dtype=torch.float32):
    """shapes: ordered {name: shape}.  Returns {name: tensor}."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        leaf = name.rsplit("/", 1)[-1]
        if leaf in ("kernel", "depthwise_kernel"):
            fan_in = int(np.prod(shape[:-1])) if leaf == "kernel" else shape[0] * shape[1]
            v = rng.standard_normal(shape) / math.sqrt(fan_in)
        elif leaf == "moving_variance":
            v = rng.uniform(0.5, 1.5, shape)
        elif leaf == "gamma":
            parts = name.split("/")
            layer_scale = len(parts) >= 3 and parts[-2].isdigit() and parts[-3] == "blocks"
            if layer_scale:  # ConvNeXt layer scale (convnext.py:211-217): make the branch matter
                v = 0.5 + 0.2 * rng.standard_normal(shape)
            else:  # LayerNorm / BatchNorm scale
                v = 1.0 + 0.2 * rng.standard_normal(shape)
        elif leaf in ("beta", "bias", "moving_mean"):
            v = 0.2 * rng.standard_normal(shape)
        elif leaf == "relative_position_bias_table":
            v = 0.5 * rng.standard_normal(shape)
        else:  # cls_token, dist_token, pos_embed, ...
            v = 0.3 * rng.standard_normal(shape)
        out[name] = torch.from_numpy(np.asarray(v)).to(dtype)
    return out


def test_images(batch, height, width, channels=3, seed=2021):
    """Inputs exactly as the reference's parity test draws them (tests/test_timm.py:56-59)."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.random((batch, height, width, channels), dtype=np.float32))
