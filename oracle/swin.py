"""Oracle restatement of the reference Swin Transformer forward (tfimm/architectures/swin.py).

Written to mirror the reference op for op (explicit roll / partition / reverse, gathered bias, additive
mask) -- deliberately NOT the index-folded formulation the engine uses, so the two are independent.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import tf_ops as tf


def _stage_dims(cfg):
    res = (cfg.input_size[0] // cfg.patch_size, cfg.input_size[1] // cfg.patch_size)
    for i, depth in enumerate(cfg.nb_blocks):
        yield i, depth, (res[0] // 2 ** i, res[1] // 2 ** i), int(cfg.embed_dim * 2 ** i), cfg.nb_heads[i]


def _block_window(cfg, input_size, j):
    """SwinTransformerBlock.__init__ clamp, swin.py:219-223; shift schedule swin.py:387."""
    ws, shift = cfg.window_size, (0 if j % 2 == 0 else cfg.window_size // 2)
    if min(input_size) <= ws:
        shift, ws = 0, min(input_size)
    return ws, shift


def param_shapes(cfg):
    """Trainable variables only (the non-trainable attn_mask / relative_position_index are rebuilt at
    construction and never loaded: swin.py:479-486)."""
    s = OrderedDict()
    D = cfg.embed_dim
    s["patch_embed/proj/kernel"] = (cfg.patch_size, cfg.patch_size, cfg.in_channels, D)
    s["patch_embed/proj/bias"] = (D,)
    if cfg.patch_norm:
        s["patch_embed/norm/gamma"] = (D,)
        s["patch_embed/norm/beta"] = (D,)
    nb_stages = len(cfg.nb_blocks)
    for i, depth, size, dim, heads in _stage_dims(cfg):
        for j in range(depth):
            p = f"layers/{i}/blocks/{j}"
            s[f"{p}/norm1/gamma"] = (dim,)
            s[f"{p}/norm1/beta"] = (dim,)
            s[f"{p}/attn/qkv/kernel"] = (dim, 3 * dim)
            if cfg.qkv_bias:
                s[f"{p}/attn/qkv/bias"] = (3 * dim,)
            s[f"{p}/attn/proj/kernel"] = (dim, dim)
            s[f"{p}/attn/proj/bias"] = (dim,)
            s[f"{p}/attn/relative_position_bias_table"] = ((2 * cfg.window_size - 1) ** 2, heads)
            s[f"{p}/norm2/gamma"] = (dim,)
            s[f"{p}/norm2/beta"] = (dim,)
            hid = int(dim * cfg.mlp_ratio)
            s[f"{p}/mlp/fc1/kernel"] = (dim, hid)
            s[f"{p}/mlp/fc1/bias"] = (hid,)
            s[f"{p}/mlp/fc2/kernel"] = (hid, dim)
            s[f"{p}/mlp/fc2/bias"] = (dim,)
        if i < nb_stages - 1:
            s[f"layers/{i}/downsample/norm/gamma"] = (4 * dim,)
            s[f"layers/{i}/downsample/norm/beta"] = (4 * dim,)
            s[f"layers/{i}/downsample/reduction/kernel"] = (4 * dim, 2 * dim)
    last = int(D * 2 ** (nb_stages - 1))
    s["norm/gamma"] = (last,)
    s["norm/beta"] = (last,)
    if cfg.nb_classes > 0:
        s["head/kernel"] = (last, cfg.nb_classes)
        s["head/bias"] = (cfg.nb_classes,)
    return s


def window_partition(x, ws):
    """swin.py:72-87."""
    b, h, w, c = x.shape
    x = x.reshape(b, h // ws, ws, w // ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws, ws, c)


def window_reverse(windows, ws, h, w, c):
    """swin.py:90-108."""
    x = windows.reshape(-1, h // ws, w // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, h, w, c)


def relative_position_index(ws):
    """WindowAttention.build, swin.py:143-152."""
    coords = np.stack(np.meshgrid(np.arange(ws), np.arange(ws), indexing="ij"))
    flat = coords.reshape(2, -1)
    rel = (flat[:, :, None] - flat[:, None, :]).transpose((1, 2, 0)).copy()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return torch.from_numpy(rel.sum(-1).astype(np.int64))


def attention_mask(h, w, ws, shift, dtype):
    """SwinTransformerBlock.build, swin.py:249-285."""
    if shift == 0:
        return torch.zeros((1,), dtype=dtype)
    img_mask = np.zeros([1, h, w, 1])
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img_mask[:, hs, wsl, :] = cnt
            cnt += 1
    mw = window_partition(torch.from_numpy(img_mask), ws).reshape(-1, ws * ws)
    diff = mw[:, None, :] - mw[:, :, None]
    return torch.where(diff != 0, torch.tensor(-100.0, dtype=torch.float64), torch.tensor(0.0, dtype=torch.float64)).to(dtype)


def window_attention(x, mask, w, prefix, cfg, dim, heads):
    """WindowAttention.call, swin.py:159-198 (uses cfg.window_size for the bias, as the reference does)."""
    _, n, c = x.shape
    qkv = tf.dense(x, w[f"{prefix}/qkv/kernel"], w.get(f"{prefix}/qkv/bias") if cfg.qkv_bias else None)
    qkv = qkv.reshape(-1, n, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (dim // heads) ** -0.5
    attn = q @ k.transpose(-1, -2)
    ws = cfg.window_size
    bias = w[f"{prefix}/relative_position_bias_table"][relative_position_index(ws).reshape(-1)]
    bias = bias.reshape(ws * ws, ws * ws, -1).permute(2, 0, 1)
    attn = attn + bias[None]
    nw = mask.shape[0]
    m = mask[None, :, None] if mask.dim() == 3 else mask  # (1, nW, 1, n, n) or the broadcastable zeros((1,))
    attn = attn.reshape(-1, nw, heads, n, n) + m
    attn = tf.softmax(attn.reshape(-1, heads, n, n))
    x = (attn @ v).permute(0, 2, 1, 3).reshape(-1, n, c)
    return tf.dense(x, w[f"{prefix}/proj/kernel"], w[f"{prefix}/proj/bias"])


def block(x, w, prefix, cfg, input_size, dim, heads, j):
    """SwinTransformerBlock.call, swin.py:287-327."""
    h, wd = input_size
    ws, shift = _block_window(cfg, input_size, j)
    b, l, c = x.shape
    shortcut = x
    x = tf.norm(x, w, f"{prefix}/norm1", cfg.norm_layer).reshape(-1, h, wd, c)
    shifted = tf.roll(x, (-shift, -shift), (1, 2))
    xw = window_partition(shifted, ws).reshape(-1, ws * ws, c)
    aw = window_attention(xw, attention_mask(h, wd, ws, shift, x.dtype), w, f"{prefix}/attn", cfg, dim, heads)
    shifted = window_reverse(aw.reshape(-1, ws, ws, c), ws, h, wd, c)
    x = tf.roll(shifted, (shift, shift), (1, 2)).reshape(-1, h * wd, c)
    x = x + shortcut
    shortcut = x
    y = tf.norm(x, w, f"{prefix}/norm2", cfg.norm_layer)
    y = tf.dense(y, w[f"{prefix}/mlp/fc1/kernel"], w[f"{prefix}/mlp/fc1/bias"])
    y = tf.act(y, cfg.act_layer)
    y = tf.dense(y, w[f"{prefix}/mlp/fc2/kernel"], w[f"{prefix}/mlp/fc2/bias"])
    return y + shortcut


def patch_merging(x, w, prefix, cfg, input_size):
    """PatchMerging.call, swin.py:348-362."""
    h, wd = input_size
    c = x.shape[-1]
    x = x.reshape(-1, h, wd, c)
    x = torch.cat((x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]), dim=-1)
    x = x.reshape(-1, (h // 2) * (wd // 2), 4 * c)
    x = tf.norm(x, w, f"{prefix}/norm", cfg.norm_layer)
    return tf.dense(x, w[f"{prefix}/reduction/kernel"])


def forward_features(cfg, w, x, return_features=False):
    """SwinTransformer.forward_features, swin.py:488-508; PatchEmbeddings, layers/transformers.py:142-173."""
    features = OrderedDict()
    x = tf.conv2d(x, w["patch_embed/proj/kernel"], w["patch_embed/proj/bias"], stride=cfg.patch_size)
    x = x.reshape(x.shape[0], -1, x.shape[-1])
    if cfg.patch_norm:
        x = tf.norm(x, w, "patch_embed/norm", cfg.norm_layer)
    features["patch_embedding"] = x
    block_idx = 0
    nb_stages = len(cfg.nb_blocks)
    for i, depth, size, dim, heads in _stage_dims(cfg):
        for j in range(depth):
            x = block(x, w, f"layers/{i}/blocks/{j}", cfg, size, dim, heads, j)
            if return_features:
                features[f"block_{block_idx}"] = x
            block_idx += 1
        if i < nb_stages - 1:
            x = patch_merging(x, w, f"layers/{i}/downsample", cfg, size)
        if return_features:
            features[f"stage_{i}"] = x
    x = tf.norm(x, w, "norm", cfg.norm_layer)
    features["features_all"] = x
    x = x.mean(dim=1)  # GlobalAveragePooling1D
    features["features"] = x
    return (x, features) if return_features else x


def forward(cfg, w, x, return_features=False):
    """SwinTransformer.call, swin.py:510-517."""
    features = {}
    x = forward_features(cfg, w, x, return_features)
    if return_features:
        x, features = x
    if cfg.nb_classes > 0:
        x = tf.dense(x, w["head/kernel"], w["head/bias"])
    features["logits"] = x
    return (x, features) if return_features else x
