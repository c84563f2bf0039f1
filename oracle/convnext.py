"""Oracle restatement of the reference ConvNeXt forward (tfimm/architectures/convnext.py)."""
from collections import OrderedDict

from . import tf_ops as tf


def param_shapes(cfg):
    """Names/shapes per convnext.py:191-217,258-266,319-327,353-360 and layers/transformers.py:192-205,
    240-253 (ConvMLP uses 1x1 Conv2D kernels)."""
    s = OrderedDict()
    d0 = cfg.embed_dim[0]
    s["stem/0/kernel"] = (cfg.patch_size, cfg.patch_size, cfg.in_channels, d0)
    s["stem/0/bias"] = (d0,)
    s["stem/1/gamma"] = (d0,)
    s["stem/1/beta"] = (d0,)
    for j, (dim, depth) in enumerate(zip(cfg.embed_dim, cfg.nb_blocks)):
        if j > 0:
            prev = cfg.embed_dim[j - 1]
            s[f"stages/{j}/downsample/0/gamma"] = (prev,)
            s[f"stages/{j}/downsample/0/beta"] = (prev,)
            s[f"stages/{j}/downsample/1/kernel"] = (2, 2, prev, dim)
            s[f"stages/{j}/downsample/1/bias"] = (dim,)
        hid = int(cfg.mlp_ratio * dim)
        lead = (1, 1) if cfg.conv_mlp_block else ()
        for k in range(depth):
            p = f"stages/{j}/blocks/{k}"
            s[f"{p}/conv_dw/depthwise_kernel"] = (7, 7, dim, 1)
            s[f"{p}/conv_dw/bias"] = (dim,)
            s[f"{p}/norm/gamma"] = (dim,)
            s[f"{p}/norm/beta"] = (dim,)
            s[f"{p}/mlp/fc1/kernel"] = (*lead, dim, hid)
            s[f"{p}/mlp/fc1/bias"] = (hid,)
            s[f"{p}/mlp/fc2/kernel"] = (*lead, hid, dim)
            s[f"{p}/mlp/fc2/bias"] = (dim,)
            s[f"{p}/gamma"] = (dim,)
    last = cfg.embed_dim[-1]
    s["head/norm/gamma"] = (last,)
    s["head/norm/beta"] = (last,)
    if cfg.nb_classes > 0:
        s["head/fc/kernel"] = (last, cfg.nb_classes)
        s["head/fc/bias"] = (cfg.nb_classes,)
    return s


def _mlp(x, w, prefix, cfg):
    """MLP.call / ConvMLP.call, layers/transformers.py:208-214, 256-262."""
    if cfg.conv_mlp_block:
        x = tf.conv2d(x, w[f"{prefix}/fc1/kernel"], w[f"{prefix}/fc1/bias"])
        x = tf.act(x, cfg.act_layer)
        return tf.conv2d(x, w[f"{prefix}/fc2/kernel"], w[f"{prefix}/fc2/bias"])
    x = tf.dense(x, w[f"{prefix}/fc1/kernel"], w[f"{prefix}/fc1/bias"])
    x = tf.act(x, cfg.act_layer)
    return tf.dense(x, w[f"{prefix}/fc2/kernel"], w[f"{prefix}/fc2/bias"])


def block(x, w, prefix, cfg):
    """ConvNeXtBlock.call, convnext.py:219-228."""
    shortcut = x
    x = tf.depthwise_conv2d(x, w[f"{prefix}/conv_dw/depthwise_kernel"], w[f"{prefix}/conv_dw/bias"], padding=3)
    x = tf.norm(x, w, f"{prefix}/norm", cfg.norm_layer)
    x = _mlp(x, w, f"{prefix}/mlp", cfg)
    x = x * w[f"{prefix}/gamma"]
    return x + shortcut


def forward_features(cfg, w, x, return_features=False):
    """ConvNeXt.forward_features, convnext.py:375-409; ConvNeXtStage.call, :286-295."""
    features = OrderedDict()
    x = tf.conv2d(x, w["stem/0/kernel"], w["stem/0/bias"], stride=cfg.patch_size)
    x = tf.norm(x, w, "stem/1", cfg.norm_layer)
    features["stem"] = x
    for j, depth in enumerate(cfg.nb_blocks):
        if j > 0:
            x = tf.norm(x, w, f"stages/{j}/downsample/0", cfg.norm_layer)
            x = tf.conv2d(x, w[f"stages/{j}/downsample/1/kernel"], w[f"stages/{j}/downsample/1/bias"], stride=2)
            if return_features:
                features[f"stage_{j}/downsample"] = x
        for k in range(depth):
            x = block(x, w, f"stages/{j}/blocks/{k}", cfg)
            if return_features:
                features[f"stage_{j}/block_{k}"] = x
    features["conv_features"] = x
    return (x, features) if return_features else x


def forward(cfg, w, x, return_features=False):
    """ConvNeXt.call, convnext.py:411-440."""
    features = OrderedDict()
    x = forward_features(cfg, w, x, return_features)
    if return_features:
        x, features = x
    x = x.mean(dim=(1, 2))  # GlobalAveragePooling2D
    x = tf.norm(x, w, "head/norm", cfg.norm_layer)
    features["features"] = x
    if cfg.nb_classes > 0:
        x = tf.dense(x, w["head/fc/kernel"], w["head/fc/bias"])
    features["logits"] = x
    return (x, features) if return_features else x
