"""Oracle restatement of the reference ViT / DeiT forward (tfimm/architectures/vit.py)."""
from collections import OrderedDict

import torch

from . import tf_ops as tf


def _get(cfg, name, default=None):
    return getattr(cfg, name, default)


def param_shapes(cfg):
    """Variable names (without the "<model>/" prefix and ":0") and shapes, in creation order.
    Follows vit.py:142-146,200-216,314,344-371,378-400 and layers/transformers.py:131-140,192-205."""
    D = cfg.embed_dim
    hid = int(D * cfg.mlp_ratio)
    nb_tokens = 2 if cfg.distilled else 1
    gh, gw = cfg.input_size[0] // cfg.patch_size, cfg.input_size[1] // cfg.patch_size
    s = OrderedDict()
    s["patch_embed/proj/kernel"] = (cfg.patch_size, cfg.patch_size, cfg.in_channels, D)
    s["patch_embed/proj/bias"] = (D,)
    s["cls_token"] = (1, 1, D)
    if cfg.distilled:
        s["dist_token"] = (1, 1, D)
    s["pos_embed"] = (1, gh * gw + nb_tokens, D)
    for j in range(cfg.nb_blocks):
        p = f"blocks/{j}"
        s[f"{p}/norm1/gamma"] = (D,)
        s[f"{p}/norm1/beta"] = (D,)
        s[f"{p}/attn/qkv/kernel"] = (D, 3 * D)
        if cfg.qkv_bias:
            s[f"{p}/attn/qkv/bias"] = (3 * D,)
        s[f"{p}/attn/proj/kernel"] = (D, D)
        s[f"{p}/attn/proj/bias"] = (D,)
        s[f"{p}/norm2/gamma"] = (D,)
        s[f"{p}/norm2/beta"] = (D,)
        s[f"{p}/mlp/fc1/kernel"] = (D, hid)
        s[f"{p}/mlp/fc1/bias"] = (hid,)
        s[f"{p}/mlp/fc2/kernel"] = (hid, D)
        s[f"{p}/mlp/fc2/bias"] = (D,)
    s["norm/gamma"] = (D,)
    s["norm/beta"] = (D,)
    feat = D
    if cfg.representation_size:
        s["pre_logits/fc/kernel"] = (D, cfg.representation_size)
        s["pre_logits/fc/bias"] = (cfg.representation_size,)
        feat = cfg.representation_size
    if cfg.nb_classes > 0:
        s["head/kernel"] = (feat, cfg.nb_classes)
        s["head/bias"] = (cfg.nb_classes,)
        if cfg.distilled:
            s["head_dist/kernel"] = (feat, cfg.nb_classes)
            s["head_dist/bias"] = (cfg.nb_classes,)
    return s


def attention(x, w, prefix, nb_heads, qkv_bias):
    """ViTMultiHeadAttention.call, vit.py:149-171."""
    B, N, D = x.shape
    qkv = tf.dense(x, w[f"{prefix}/qkv/kernel"], w.get(f"{prefix}/qkv/bias") if qkv_bias else None)
    qkv = qkv.reshape(B, N, 3, nb_heads, D // nb_heads).permute(2, 0, 3, 1, 4)  # (3, B, H, N, dh)
    q, k, v = qkv[0], qkv[1], qkv[2]
    scale = (D // nb_heads) ** -0.5
    attn = tf.softmax(scale * (q @ k.transpose(-1, -2)))
    y = (attn @ v).permute(0, 2, 1, 3).reshape(B, N, D)
    return tf.dense(y, w[f"{prefix}/proj/kernel"], w[f"{prefix}/proj/bias"]), attn


def block(x, w, prefix, cfg):
    """ViTBlock.call, vit.py:219-235 (DropPath is the identity at inference, layers/drop.py:27-28)."""
    shortcut = x
    y = tf.norm(x, w, f"{prefix}/norm1", cfg.norm_layer)
    y, attn = attention(y, w, f"{prefix}/attn", cfg.nb_heads, cfg.qkv_bias)
    x = y + shortcut
    shortcut = x
    y = tf.norm(x, w, f"{prefix}/norm2", cfg.norm_layer)
    y = tf.dense(y, w[f"{prefix}/mlp/fc1/kernel"], w[f"{prefix}/mlp/fc1/bias"])  # MLP.call, transformers.py:208-214
    y = tf.act(y, cfg.act_layer)
    y = tf.dense(y, w[f"{prefix}/mlp/fc2/kernel"], w[f"{prefix}/mlp/fc2/bias"])
    return y + shortcut, attn


def interpolate_pos_embeddings(pos_embed, src_grid, tgt_grid, nb_tokens):
    """layers/transformers.py:13-47."""
    if tuple(src_grid) == tuple(tgt_grid):
        return pos_embed
    grid = pos_embed[:, nb_tokens:].reshape(1, *src_grid, -1)
    grid = tf.resize_bicubic(grid, tgt_grid).reshape(1, tgt_grid[0] * tgt_grid[1], -1)
    return torch.cat((pos_embed[:, :nb_tokens], grid), dim=1)


def forward_features(cfg, w, x, return_features=False):
    """ViT.forward_features, vit.py:422-464; PatchEmbeddings.call, layers/transformers.py:142-173."""
    features = OrderedDict()
    B = x.shape[0]
    nb_tokens = 2 if cfg.distilled else 1
    x = tf.conv2d(x, w["patch_embed/proj/kernel"], w["patch_embed/proj/bias"], stride=cfg.patch_size)
    gh, gw = x.shape[1], x.shape[2]
    x = x.reshape(B, gh * gw, -1)
    toks = [w["cls_token"].expand(B, -1, -1)]
    if cfg.distilled:
        toks.append(w["dist_token"].expand(B, -1, -1))
    x = torch.cat(toks + [x], dim=1)
    src_grid = (cfg.input_size[0] // cfg.patch_size, cfg.input_size[1] // cfg.patch_size)
    pos = w["pos_embed"]
    if _get(cfg, "interpolate_input", False):
        pos = interpolate_pos_embeddings(pos, src_grid, (gh, gw), nb_tokens)
    x = x + pos
    features["patch_embedding"] = x
    for j in range(cfg.nb_blocks):
        x, attn = block(x, w, f"blocks/{j}", cfg)
        if return_features:
            features[f"block_{j}/attn"] = attn
        features[f"block_{j}"] = x
    x = tf.norm(x, w, "norm", cfg.norm_layer)
    features["features_all"] = x
    if cfg.distilled:
        x = x[:, :2]
    elif cfg.representation_size:
        x = torch.tanh(tf.dense(x[:, 0], w["pre_logits/fc/kernel"], w["pre_logits/fc/bias"]))
    else:
        x = x[:, 0]
    features["features"] = x
    return (x, features) if return_features else x


def forward(cfg, w, x, return_features=False):
    """ViT.call, vit.py:466-478."""
    features = {}
    x = forward_features(cfg, w, x, return_features)
    if return_features:
        x, features = x
    if cfg.nb_classes > 0:
        if not cfg.distilled:
            x = tf.dense(x, w["head/kernel"], w["head/bias"])
        else:
            y = tf.dense(x[:, 0], w["head/kernel"], w["head/bias"])
            y_dist = tf.dense(x[:, 1], w["head_dist/kernel"], w["head_dist/bias"])
            x = torch.stack((y, y_dist), dim=1)
    features["logits"] = x
    return (x, features) if return_features else x
