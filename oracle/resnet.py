"""Oracle restatement of the reference ResNet family forward (tfimm/architectures/resnet.py,
tfimm/layers/attention.py, tfimm/layers/classifier.py).  BatchNorm, padding, pooling and the attention
gates are separate ops in the reference's order.  BlurPool / GroupNorm variants are not restated."""
import math
from collections import OrderedDict

import torch

from . import tf_ops as tf


def make_divisible(value, divisor, min_value=None, round_limit=0.9):
    min_value = min_value or divisor
    new_value = max(min_value, int(value + divisor / 2) // divisor * divisor)
    if new_value < round_limit * value:
        new_value += divisor
    return new_value


def _bn(s, prefix, ch, kind="batch_norm"):
    # GroupNormalization has gamma / beta only, no moving statistics (tfimm/layers/norm.py:143-160)
    leaves = ("gamma", "beta") if kind.startswith("group_norm") else ("gamma", "beta", "moving_mean", "moving_variance")
    for leaf in leaves:
        s[f"{prefix}/{leaf}"] = (ch,)


def _plan(cfg):
    """Block list following make_stage (resnet.py:333-382): yields dicts with the channel bookkeeping."""
    expansion = 1 if cfg.block == "basic_block" else 4
    in_channels = cfg.stem_width * 2 if cfg.stem_type in {"deep", "deep_tiered"} else 64
    actual_in = in_channels
    blocks = []
    for idx in range(4):
        nb_channels = cfg.nb_channels[idx]
        out_channels = nb_channels * expansion
        for block_idx in range(cfg.nb_blocks[idx]):
            stride = 1 if idx == 0 or block_idx > 0 else 2
            down = (block_idx == 0) and (stride != 1 or in_channels != out_channels)
            blocks.append(dict(name=f"layer{idx + 1}/{block_idx}", nb_channels=nb_channels, stride=stride,
                               down=down, cin=actual_in, cout=out_channels))
            in_channels = nb_channels
            actual_in = out_channels
    return blocks


def _eca_k(channels, gamma=2, beta=1):
    t = int(abs(math.log(channels, 2) + beta) / gamma)
    return max(t if t % 2 else t + 1, 3)


def param_shapes(cfg):
    s = OrderedDict()
    if cfg.stem_type in {"deep", "deep_tiered"}:
        c0 = 3 * (cfg.stem_width // 4) if cfg.stem_type == "deep_tiered" else cfg.stem_width
        s["conv1/0/kernel"] = (3, 3, cfg.in_channels, c0)
        _bn(s, "conv1/1", c0, cfg.norm_layer)
        s["conv1/3/kernel"] = (3, 3, c0, cfg.stem_width)
        _bn(s, "conv1/4", cfg.stem_width, cfg.norm_layer)
        s["conv1/6/kernel"] = (3, 3, cfg.stem_width, cfg.stem_width * 2)
        stem_out = cfg.stem_width * 2
    else:
        s["conv1/kernel"] = (7, 7, cfg.in_channels, 64)
        stem_out = 64
    _bn(s, "bn1", stem_out, cfg.norm_layer)
    if cfg.replace_stem_pool:
        s["maxpool/0/kernel"] = (3, 3, stem_out, stem_out)
        _bn(s, "maxpool/1", stem_out, cfg.norm_layer)
    for b in _plan(cfg):
        p, ch, cin, cout = b["name"], b["nb_channels"], b["cin"], b["cout"]
        if cfg.block == "basic_block":
            first = ch // cfg.block_reduce_first
            s[f"{p}/conv1/kernel"] = (3, 3, cin, first)
            _bn(s, f"{p}/bn1", first, cfg.norm_layer)
            s[f"{p}/conv2/kernel"] = (3, 3, first, cout)
            _bn(s, f"{p}/bn2", cout, cfg.norm_layer)
        else:
            width = int(math.floor(ch * (cfg.base_width / 64)) * cfg.cardinality)
            first = width // cfg.block_reduce_first
            s[f"{p}/conv1/kernel"] = (1, 1, cin, first)
            _bn(s, f"{p}/bn1", first, cfg.norm_layer)
            s[f"{p}/conv2/kernel"] = (3, 3, first // cfg.cardinality, width)
            _bn(s, f"{p}/bn2", width, cfg.norm_layer)
            s[f"{p}/conv3/kernel"] = (1, 1, width, cout)
            _bn(s, f"{p}/bn3", cout, cfg.norm_layer)
        if cfg.attn_layer == "se":
            rd = make_divisible(cout * cfg.se_ratio, 8, round_limit=0.0)
            s[f"{p}/se/fc1/kernel"] = (1, 1, cout, rd)
            s[f"{p}/se/fc1/bias"] = (rd,)
            s[f"{p}/se/fc2/kernel"] = (1, 1, rd, cout)
            s[f"{p}/se/fc2/bias"] = (cout,)
        elif cfg.attn_layer == "eca":
            s[f"{p}/se/conv/kernel"] = (_eca_k(cout), 1, 1)
        if b["down"]:
            if cfg.downsample_mode == "conv":
                s[f"{p}/downsample/0/kernel"] = (cfg.down_kernel_size, cfg.down_kernel_size, cin, cout)
                _bn(s, f"{p}/downsample/1", cout, cfg.norm_layer)
            else:
                s[f"{p}/downsample/1/kernel"] = (1, 1, cin, cout)
                _bn(s, f"{p}/downsample/2", cout, cfg.norm_layer)
    if cfg.nb_classes > 0:
        s["remove/fc/kernel"] = (_plan(cfg)[-1]["cout"], cfg.nb_classes)
        s["remove/fc/bias"] = (cfg.nb_classes,)
    return s


def _attn(x, w, prefix, cfg):
    if cfg.attn_layer == "se":  # SEModule.call, layers/attention.py:67-75 (norm_layer "" -> identity)
        x_se = x.mean(dim=(1, 2), keepdim=True)
        x_se = tf.conv2d(x_se, w[f"{prefix}/fc1/kernel"], w[f"{prefix}/fc1/bias"])
        x_se = tf.act(x_se, "relu")
        x_se = tf.conv2d(x_se, w[f"{prefix}/fc2/kernel"], w[f"{prefix}/fc2/bias"])
        return x * tf.act(x_se, "sigmoid")
    if cfg.attn_layer == "eca":  # EcaModule.call, layers/attention.py:120-130
        y = x.mean(dim=(1, 2))                      # (N, C)
        k = w[f"{prefix}/conv/kernel"].reshape(-1)  # Conv1D(filters=1) kernel (k, 1, 1)
        pad = (k.numel() - 1) // 2
        y = torch.nn.functional.conv1d(torch.nn.functional.pad(y, (pad, pad))[:, None, :], k[None, None, :])[:, 0]
        return x * torch.sigmoid(y)[:, None, None, :]
    return x


def _downsample(x, w, prefix, cfg, stride):
    if cfg.downsample_mode == "avg":  # downsample_avg, resnet.py:295-312
        if stride != 1:
            x = tf.avg_pool2d_same(x, 2, stride)
        x = tf.conv2d(x, w[f"{prefix}/downsample/1/kernel"])
        return tf.norm(x, w, f"{prefix}/downsample/2", cfg.norm_layer)
    p = (stride + cfg.down_kernel_size) // 2 - 1  # downsample_conv, resnet.py:315-330
    x = tf.conv2d(x, w[f"{prefix}/downsample/0/kernel"], stride=stride, padding=p)
    return tf.norm(x, w, f"{prefix}/downsample/1", cfg.norm_layer)


def block(x, w, b, cfg):
    p, stride, act = b["name"], b["stride"], cfg.act_layer
    shortcut = x
    use_aa = bool(cfg.aa_layer) and stride == 2  # the blur layer takes over the stride (resnet.py:127-140, 218-241)
    if cfg.block == "basic_block":  # BasicBlock.call, resnet.py:166-189
        x = tf.conv2d(x, w[f"{p}/conv1/kernel"], stride=1 if use_aa else stride, padding=1)
        x = tf.act(tf.norm(x, w, f"{p}/bn1", cfg.norm_layer), act)
        if use_aa:
            x = tf.blur_pool2d(x, stride)
        x = tf.conv2d(x, w[f"{p}/conv2/kernel"], padding=1)
        x = tf.norm(x, w, f"{p}/bn2", cfg.norm_layer)
    else:  # Bottleneck.call, resnet.py:266-292
        x = tf.conv2d(x, w[f"{p}/conv1/kernel"])
        x = tf.act(tf.norm(x, w, f"{p}/bn1", cfg.norm_layer), act)
        x = tf.conv2d(x, w[f"{p}/conv2/kernel"], stride=1 if use_aa else stride, padding=1, groups=cfg.cardinality)
        x = tf.act(tf.norm(x, w, f"{p}/bn2", cfg.norm_layer), act)
        if use_aa:
            x = tf.blur_pool2d(x, stride)
        x = tf.conv2d(x, w[f"{p}/conv3/kernel"])
        x = tf.norm(x, w, f"{p}/bn3", cfg.norm_layer)
    x = _attn(x, w, f"{p}/se", cfg)
    if b["down"]:
        shortcut = _downsample(shortcut, w, p, cfg, stride)
    return tf.act(x + shortcut, act)


def forward_features(cfg, w, x, return_features=False):
    """ResNet.forward_features, resnet.py:570-584; stem built at :466-540."""
    features = OrderedDict()
    act = cfg.act_layer
    if cfg.stem_type in {"deep", "deep_tiered"}:
        x = tf.conv2d(x, w["conv1/0/kernel"], stride=2, padding=1)
        x = tf.act(tf.norm(x, w, "conv1/1", cfg.norm_layer), act)
        x = tf.conv2d(x, w["conv1/3/kernel"], padding="same")
        x = tf.act(tf.norm(x, w, "conv1/4", cfg.norm_layer), act)
        x = tf.conv2d(x, w["conv1/6/kernel"], padding="same")
    else:
        x = tf.conv2d(x, w["conv1/kernel"], stride=2, padding=3)
    x = tf.act(tf.norm(x, w, "bn1", cfg.norm_layer), act)
    if cfg.replace_stem_pool:
        x = tf.conv2d(x, w["maxpool/0/kernel"], stride=2, padding=1)
        x = tf.act(tf.norm(x, w, "maxpool/1", cfg.norm_layer), act)
    else:
        x = torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))  # ZeroPadding2D(1): zeros take part in the max
        if cfg.aa_layer:  # MaxPool2D(3, strides=1) + BlurPool2D(stride=2), resnet.py:532-536
            x = tf.blur_pool2d(tf.max_pool2d(x, 3, 1), 2)
        else:
            x = tf.max_pool2d(x, 3, 2)
    features["stem"] = x
    for j, b in enumerate(_plan(cfg)):
        x = block(x, w, b, cfg)
        features[f"block_{j}"] = x
    features["features"] = x
    return (x, features) if return_features else x


def forward(cfg, w, x, return_features=False):
    """ResNet.call (resnet.py:586-593) + ClassifierHead.call (layers/classifier.py:65-74)."""
    features = {}
    x = forward_features(cfg, w, x, return_features)
    if return_features:
        x, features = x
    x = x.mean(dim=(1, 2))
    if cfg.nb_classes > 0:
        x = tf.dense(x, w["remove/fc/kernel"], w["remove/fc/bias"])
    features["logits"] = x
    return (x, features) if return_features else x
