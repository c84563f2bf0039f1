"""TensorFlow/Keras op semantics the reference relies on, restated on torch CPU tensors.

All activations are channels-last (NHWC / (B, T, C)); kernels keep their TF layouts:
Dense ``(in, out)``, Conv2D ``(kh, kw, in/groups, out)``, DepthwiseConv2D ``(kh, kw, C, 1)``.
These are the eight third-party behaviours listed in SURVEY.md 8(c); tests/test_oracle_ops.py
checks each against closed-form cases.
"""
import math

import torch
import torch.nn.functional as F


def dense(x, kernel, bias=None):
    """tf.keras.layers.Dense: contracts the last axis; kernel is (in, out)."""
    y = x @ kernel
    return y if bias is None else y + bias


def act(x, name):
    """tfimm/layers/factory.py:6-13 -> Keras activations.  "gelu" is the exact erf form."""
    if name in ("linear", "", None):
        return x
    if name == "gelu":
        return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))
    if name == "swish":
        return x * torch.sigmoid(x)
    if name == "relu":
        return torch.relu(x)
    if name == "relu6":
        return torch.clamp(x, 0.0, 6.0)
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "tanh":
        return torch.tanh(x)
    raise ValueError(f"Unknown activation: {name}.")


def layer_norm(x, gamma, beta, eps):
    """tf.keras.layers.LayerNormalization(axis=-1): biased variance, eps inside the rsqrt."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def batch_norm(x, gamma, beta, moving_mean, moving_variance, eps):
    """tf.keras.layers.BatchNormalization(training=False) over the channel (last) axis."""
    return (x - moving_mean) * torch.rsqrt(moving_variance + eps) * gamma + beta


def norm(x, weights, prefix, kind):
    """norm_layer_factory (tfimm/layers/factory.py:16-60) restricted to the in-scope kinds."""
    if kind == "":
        return x
    if kind in ("layer_norm", "layer_norm_eps_1e-6"):
        eps = 1e-5 if kind == "layer_norm" else 1e-6
        return layer_norm(x, weights[f"{prefix}/gamma"], weights[f"{prefix}/beta"], eps)
    if kind in ("batch_norm", "batch_norm_tf"):
        eps = 1e-5 if kind == "batch_norm" else 1e-3
        return batch_norm(x, weights[f"{prefix}/gamma"], weights[f"{prefix}/beta"],
                          weights[f"{prefix}/moving_mean"], weights[f"{prefix}/moving_variance"], eps)
    if kind in ("group_norm", "group_norm_1grp"):
        return group_norm(x, weights[f"{prefix}/gamma"], weights[f"{prefix}/beta"],
                          32 if kind == "group_norm" else 1, 1e-5)
    raise ValueError(f"Unknown normalization layer: {kind}")


def group_norm(x, gamma, beta, nb_groups, eps):
    """group_normalize (tfimm/layers/norm.py:22-101): NHWC -> N,H,W,G,S; moments over every axis except N and G
    (biased variance); per-channel gamma / beta."""
    shape = x.shape
    c = shape[-1]
    xg = x.reshape(shape[0], -1, nb_groups, c // nb_groups)
    mean = xg.mean(dim=(1, 3), keepdim=True)
    var = ((xg - mean) ** 2).mean(dim=(1, 3), keepdim=True)
    xg = (xg - mean) * torch.rsqrt(var + eps)
    return xg.reshape(shape) * gamma + beta


def blur_pool2d(x, stride=2):
    """BlurPool2D.call (tfimm/layers/blurpool.py:54-62), kernel_size 3: REFLECT pad by (3 + stride) // 2 - 1,
    then a VALID depthwise convolution with [[1,2,1],[2,4,2],[1,2,1]] / 16 and the given stride."""
    p = (3 + stride) // 2 - 1
    xc = x.permute(0, 3, 1, 2)
    xc = torch.nn.functional.pad(xc, (p, p, p, p), mode="reflect")
    k1 = torch.tensor([1.0, 2.0, 1.0], dtype=x.dtype)
    k = (k1[:, None] * k1[None, :] / 16.0)[None, None].repeat(x.shape[-1], 1, 1, 1)
    y = torch.nn.functional.conv2d(xc, k, stride=stride, groups=x.shape[-1])
    return y.permute(0, 2, 3, 1)


def same_padding(size, k, s, d=1):
    """TF "SAME": out = ceil(in/s); total = max((out-1)*s + (k-1)*d + 1 - in, 0); the extra
    pixel goes AFTER (bottom / right)."""
    out = -(-size // s)
    total = max((out - 1) * s + (k - 1) * d + 1 - size, 0)
    return total // 2, total - total // 2


def symmetric_padding(k, s, d=1):
    """get_padding, tfimm/layers/conv.py:15-28 (PyTorch-style)."""
    return ((s - 1) + d * (k - 1)) // 2


def _pad_nhwc(x, pt, pb, pl, pr):
    if pt or pb or pl or pr:
        x = F.pad(x, (0, 0, pl, pr, pt, pb))
    return x


def conv2d(x, kernel, bias=None, stride=1, padding="valid", groups=1, dilation=1):
    """tf.keras.layers.Conv2D on NHWC with kernel (kh, kw, in/groups, out).
    padding: "valid" | "same" (TF asymmetric) | "symmetric" (PadConv2D, tfimm/layers/conv.py:31-88)
    | int (explicit ZeroPadding2D then VALID)."""
    kh, kw = kernel.shape[0], kernel.shape[1]
    H, W = x.shape[1], x.shape[2]
    if padding == "same":
        pt, pb = same_padding(H, kh, stride, dilation)
        pl, pr = same_padding(W, kw, stride, dilation)
    elif padding == "symmetric":
        pt = pb = symmetric_padding(kh, stride, dilation)
        pl = pr = symmetric_padding(kw, stride, dilation)
    elif padding == "valid":
        pt = pb = pl = pr = 0
    else:
        pt = pb = pl = pr = int(padding)
    x = _pad_nhwc(x, pt, pb, pl, pr)
    w = kernel.permute(3, 2, 0, 1)  # (out, in/groups, kh, kw)
    y = F.conv2d(x.permute(0, 3, 1, 2), w, bias, stride=stride, dilation=dilation, groups=groups)
    return y.permute(0, 2, 3, 1)


def depthwise_conv2d(x, kernel, bias=None, stride=1, padding="valid", dilation=1):
    """tf.keras.layers.DepthwiseConv2D (depth_multiplier 1), kernel (kh, kw, C, 1)."""
    C = kernel.shape[2]
    return conv2d(x, kernel.permute(0, 1, 3, 2), bias, stride, padding, groups=C, dilation=dilation)


def max_pool2d(x, k, s, padding="valid"):
    if padding == "same":
        pt, pb = same_padding(x.shape[1], k, s)
        pl, pr = same_padding(x.shape[2], k, s)
        x = F.pad(x, (0, 0, pl, pr, pt, pb), value=float("-inf"))
    return F.max_pool2d(x.permute(0, 3, 1, 2), k, s).permute(0, 2, 3, 1)


def avg_pool2d_same(x, k, s):
    """tf.keras.layers.AveragePooling2D(k, s, padding="same"): padded cells are EXCLUDED from the
    average (TF semantics)."""
    pt, pb = same_padding(x.shape[1], k, s)
    pl, pr = same_padding(x.shape[2], k, s)
    xp = F.pad(x, (0, 0, pl, pr, pt, pb)).permute(0, 3, 1, 2)
    ones = F.pad(torch.ones_like(x[..., :1]), (0, 0, pl, pr, pt, pb)).permute(0, 3, 1, 2)
    num = F.avg_pool2d(xp, k, s) * (k * k)
    den = F.avg_pool2d(ones, k, s) * (k * k)
    return (num / den).permute(0, 2, 3, 1)


def softmax(x):
    """tf.nn.softmax(axis=-1) (subtracts the row max)."""
    return torch.softmax(x, dim=-1)


def roll(x, shift, axes):
    """tf.roll: y[i] = x[(i - shift) mod n]."""
    return torch.roll(x, shifts=shift, dims=axes)


def _keys_cubic(x):
    a = -0.5
    x = x.abs()
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    far = ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a
    return torch.where(x <= 1.0, near, torch.where(x < 2.0, far, torch.zeros_like(x)))


def resize_bicubic(images, size):
    """tf.image.resize(method="bicubic", antialias=False) = ResizeBicubic(half_pixel_centers=True)
    (resize_bicubic_op.cc): Keys cubic A=-0.5 tabulated at 1024 steps, float32 source coordinate, out-of-image taps
    get weight 0 and the rest are renormalised."""
    import numpy as np

    def matrix(n_in, n_out):
        table, a = 1024, -0.5
        scale = np.float32(n_in) / np.float32(n_out)
        w = np.zeros((n_out, n_in), dtype=np.float32)
        for o in range(n_out):
            src = np.float32(np.float32(o + 0.5) * scale) - np.float32(0.5)
            base = int(np.floor(src))
            off = int(np.rint(np.float32(src - np.float32(base)) * np.float32(table)))
            taps = []
            for k, x in ((-1, off / table + 1.0), (0, off / table), (1, (table - off) / table),
                         (2, (table - off) / table + 1.0)):
                val = ((a + 2) * x - (a + 3)) * x * x + 1 if x <= 1.0 else ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
                if 0 <= base + k < n_in:
                    taps.append((base + k, np.float32(val)))
            tot = np.float32(sum(t[1] for t in taps))
            for idx, val in taps:
                w[o, idx] += np.float32(val / tot)
        return torch.from_numpy(w).to(images.dtype)

    out = torch.einsum("oh,bhwc->bowc", matrix(images.shape[1], size[0]), images)
    return torch.einsum("pw,bowc->bopc", matrix(images.shape[2], size[1]), out)