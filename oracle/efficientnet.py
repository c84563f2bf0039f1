"""Oracle restatement of the reference EfficientNet / MobileNet-V2 forward
(tfimm/architectures/efficientnet.py, efficientnet_blocks.py, efficientnet_builder.py).

BatchNorm is applied as a separate op (no folding) and SqueezeExcite as mean -> conv -> act -> conv ->
sigmoid -> multiply, exactly in the reference's order.
"""
import math
import re
from collections import OrderedDict
from copy import deepcopy

from . import tf_ops as tf


def make_divisible(value, divisor, min_value=None, round_limit=0.9):
    """utils/etc.py:14-26."""
    min_value = min_value or divisor
    new_value = max(min_value, int(value + divisor / 2) // divisor * divisor)
    if new_value < round_limit * value:
        new_value += divisor
    return new_value


def round_channels(channels, multiplier=1.0):
    """efficientnet_builder.py:31-44."""
    return make_divisible(channels * multiplier, 8)


def decode_block(block_string):
    """BlockArgs.decode, efficientnet_blocks.py:114-169 (fields the in-scope models use)."""
    ops = block_string.split("_")
    options = {"block_type": ops[0]}
    for op in ops[1:]:
        if op == "noskip":
            options["skip"] = False
        elif op == "skip":
            options["skip"] = True
        elif op.startswith("n"):
            options["n"] = {"re": "relu", "r6": "relu6", "hs": "hard_swish", "sw": "swish", "mi": "mish"}[op[1:]]
        else:
            splits = re.split(r"(\d.*)", op)
            if len(splits) >= 2:
                options[splits[0]] = splits[1]
    bt = options["block_type"]

    def ksize(ss):
        return int(ss) if ss.isdigit() else int(ss.split(".")[0])

    return dict(
        block_type=bt,
        nb_repeats=int(options.get("r")),
        filters=int(options.get("c")),
        force_in_channels=int(options.get("fc", 0)) or None,
        exp_kernel_size=ksize(options.get("a", "1")) if bt != "er" else ksize(options.get("k")),
        dw_kernel_size=ksize(options.get("k")) if bt != "er" else 1,
        stride=int(options.get("s")),
        exp_ratio=float(options.get("e", 1.0)),
        pw_act=bt == "dsa",
        se_ratio=float(options.get("se", 0.0)),
        act_layer=options.get("n", None),
        skip_connection=False if bt == "dsa" else options.get("skip", True),
    )


def scale_stage_depth(stack_args, depth_multiplier):
    """_scale_stage_depth with depth_trunc="ceil", efficientnet_builder.py:47-93."""
    repeats = [ba["nb_repeats"] for ba in stack_args]
    nb_repeats = sum(repeats)
    nb_repeats_scaled = int(math.ceil(nb_repeats * depth_multiplier))
    repeats_scaled = []
    for r in repeats[::-1]:
        rs = max(1, round((r / nb_repeats * nb_repeats_scaled)))
        repeats_scaled.append(rs)
        nb_repeats -= r
        nb_repeats_scaled -= rs
    repeats_scaled = repeats_scaled[::-1]
    out = []
    for ba, rep in zip(stack_args, repeats_scaled):
        out.extend([deepcopy(ba) for _ in range(rep)])
    return out


def build_blocks(cfg):
    """decode_architecture (efficientnet_builder.py:96-144) + EfficientNetBuilder.__call__/_make_block
    (:173-283) + the channel bookkeeping the Keras layers do in build().  Returns a list of dicts."""
    n = len(cfg.architecture)
    blocks = []
    in_ch = cfg.stem_size
    for stack_idx, block_strings in enumerate(cfg.architecture):
        stack = [decode_block(s) for s in block_strings]
        fix = cfg.fix_first_last and stack_idx in {0, n - 1}
        stack = scale_stage_depth(stack, 1.0 if fix else cfg.depth_multiplier)
        for block_idx, ba in enumerate(stack):
            if block_idx >= 1:
                ba["stride"] = 1
            ba["filters"] = round_channels(ba["filters"], cfg.channel_multiplier)
            if ba["force_in_channels"] is not None:
                ba["force_in_channels"] = round_channels(ba["force_in_channels"], cfg.channel_multiplier)
            ba["act_layer"] = ba["act_layer"] or cfg.act_layer
            if ba["block_type"] != "cn":
                ba["se_ratio"] /= ba["exp_ratio"]
            ba["name"] = f"blocks.{stack_idx}.{block_idx}"
            ba["key"] = f"stage_{stack_idx}/block_{block_idx}"
            ba["in_channels"] = in_ch
            bt = ba["block_type"]
            if bt == "ir":
                ba["mid"] = make_divisible(in_ch * ba["exp_ratio"], 8)
            elif bt == "er":
                ba["mid"] = make_divisible((ba["force_in_channels"] or in_ch) * ba["exp_ratio"], 8)
            ba["has_skip"] = ba["stride"] == 1 and ba["filters"] == in_ch and ba["skip_connection"]
            blocks.append(ba)
            in_ch = ba["filters"]
    return blocks


def _bn_shapes(s, prefix, ch):
    for leaf in ("gamma", "beta", "moving_mean", "moving_variance"):
        s[f"{prefix}/{leaf}"] = (ch,)


def _se_shapes(s, prefix, ch, ratio):
    rd = round(ch * ratio)  # SqueezeExcite.build, efficientnet_blocks.py:225
    s[f"{prefix}/conv_reduce/kernel"] = (1, 1, ch, rd)
    s[f"{prefix}/conv_reduce/bias"] = (rd,)
    s[f"{prefix}/conv_expand/kernel"] = (1, 1, rd, ch)
    s[f"{prefix}/conv_expand/bias"] = (ch,)


def param_shapes(cfg):
    s = OrderedDict()
    s["conv_stem/kernel"] = (3, 3, cfg.in_channels, cfg.stem_size)
    _bn_shapes(s, "bn1", cfg.stem_size)
    last = cfg.stem_size
    for ba in build_blocks(cfg):
        p, bt, cin, cout = ba["name"], ba["block_type"], ba["in_channels"], ba["filters"]
        use_se = ba["se_ratio"] > 0.0
        if bt == "ir":
            mid = ba["mid"]
            s[f"{p}/conv_pw/kernel"] = (ba["exp_kernel_size"], ba["exp_kernel_size"], cin, mid)
            _bn_shapes(s, f"{p}/bn1", mid)
            s[f"{p}/conv_dw/depthwise_kernel"] = (ba["dw_kernel_size"], ba["dw_kernel_size"], mid, 1)
            _bn_shapes(s, f"{p}/bn2", mid)
            if use_se:
                _se_shapes(s, f"{p}/se", mid, ba["se_ratio"])
            s[f"{p}/conv_pwl/kernel"] = (1, 1, mid, cout)
            _bn_shapes(s, f"{p}/bn3", cout)
        elif bt in ("ds", "dsa"):
            s[f"{p}/conv_dw/depthwise_kernel"] = (ba["dw_kernel_size"], ba["dw_kernel_size"], cin, 1)
            _bn_shapes(s, f"{p}/bn1", cin)
            if use_se:
                _se_shapes(s, f"{p}/se", cin, ba["se_ratio"])
            s[f"{p}/conv_pw/kernel"] = (1, 1, cin, cout)
            _bn_shapes(s, f"{p}/bn2", cout)
        elif bt == "er":
            mid = ba["mid"]
            s[f"{p}/conv_exp/kernel"] = (ba["exp_kernel_size"], ba["exp_kernel_size"], cin, mid)
            _bn_shapes(s, f"{p}/bn1", mid)
            if use_se:
                _se_shapes(s, f"{p}/se", mid, ba["se_ratio"])
            s[f"{p}/conv_pwl/kernel"] = (1, 1, mid, cout)
            _bn_shapes(s, f"{p}/bn2", cout)
        elif bt == "cn":
            s[f"{p}/conv/kernel"] = (ba["dw_kernel_size"], ba["dw_kernel_size"], cin, cout)
            _bn_shapes(s, f"{p}/bn1", cout)
        else:
            raise ValueError(bt)
        last = cout
    s["conv_head/kernel"] = (1, 1, last, cfg.nb_features)
    _bn_shapes(s, "bn2", cfg.nb_features)
    if cfg.nb_classes > 0:
        s["classifier/kernel"] = (cfg.nb_features, cfg.nb_classes)
        s["classifier/bias"] = (cfg.nb_classes,)
    return s


def squeeze_excite(x, w, prefix, act):
    """SqueezeExcite.call, efficientnet_blocks.py:241-248."""
    x_se = x.mean(dim=(1, 2), keepdim=True)
    x_se = tf.conv2d(x_se, w[f"{prefix}/conv_reduce/kernel"], w[f"{prefix}/conv_reduce/bias"])
    x_se = tf.act(x_se, act)
    x_se = tf.conv2d(x_se, w[f"{prefix}/conv_expand/kernel"], w[f"{prefix}/conv_expand/bias"])
    return x * tf.act(x_se, "sigmoid")


def run_block(x, w, ba, cfg):
    p, bt, act, pad = ba["name"], ba["block_type"], ba["act_layer"], cfg.padding
    use_se = ba["se_ratio"] > 0.0
    shortcut = x
    if bt == "ir":  # InvertedResidual.call, efficientnet_blocks.py:438-453
        x = tf.conv2d(x, w[f"{p}/conv_pw/kernel"], padding=pad)
        x = tf.act(tf.norm(x, w, f"{p}/bn1", cfg.norm_layer), act)
        x = tf.depthwise_conv2d(x, w[f"{p}/conv_dw/depthwise_kernel"], stride=ba["stride"], padding=pad)
        x = tf.act(tf.norm(x, w, f"{p}/bn2", cfg.norm_layer), act)
        if use_se:
            x = squeeze_excite(x, w, f"{p}/se", act)
        x = tf.conv2d(x, w[f"{p}/conv_pwl/kernel"], padding=pad)
        x = tf.norm(x, w, f"{p}/bn3", cfg.norm_layer)
    elif bt in ("ds", "dsa"):  # DepthwiseSeparableConv.call, :348-362
        x = tf.depthwise_conv2d(x, w[f"{p}/conv_dw/depthwise_kernel"], stride=ba["stride"], padding=pad)
        x = tf.act(tf.norm(x, w, f"{p}/bn1", cfg.norm_layer), act)
        if use_se:
            x = squeeze_excite(x, w, f"{p}/se", act)
        x = tf.conv2d(x, w[f"{p}/conv_pw/kernel"], padding=pad)
        x = tf.norm(x, w, f"{p}/bn2", cfg.norm_layer)
        if ba["pw_act"]:
            x = tf.act(x, act)
    elif bt == "er":  # EdgeResidual.call, :520-535
        x = tf.conv2d(x, w[f"{p}/conv_exp/kernel"], stride=ba["stride"], padding=pad)
        x = tf.act(tf.norm(x, w, f"{p}/bn1", cfg.norm_layer), act)
        if use_se:
            x = squeeze_excite(x, w, f"{p}/se", act)
        x = tf.conv2d(x, w[f"{p}/conv_pwl/kernel"], padding=pad)
        x = tf.norm(x, w, f"{p}/bn2", cfg.norm_layer)
    else:  # ConvBnAct.call, :283-293
        x = tf.conv2d(x, w[f"{p}/conv/kernel"], stride=ba["stride"], padding=pad)
        x = tf.act(tf.norm(x, w, f"{p}/bn1", cfg.norm_layer), act)
    if ba["has_skip"]:
        x = x + shortcut
    return x


def forward_features(cfg, w, x, return_features=False):
    """EfficientNet.forward_features, efficientnet.py:278-314."""
    features = OrderedDict()
    x = tf.conv2d(x, w["conv_stem/kernel"], stride=2, padding=cfg.padding)
    x = tf.act(tf.norm(x, w, "bn1", cfg.norm_layer), cfg.act_layer)
    features["stem"] = x
    for ba in build_blocks(cfg):
        x = run_block(x, w, ba, cfg)
        features[ba["key"]] = x
    x = tf.conv2d(x, w["conv_head/kernel"], padding=cfg.padding)
    x = tf.act(tf.norm(x, w, "bn2", cfg.norm_layer), cfg.act_layer)
    features["conv_features"] = x
    return (x, features) if return_features else x


def forward(cfg, w, x, return_features=False):
    """EfficientNet.call, efficientnet.py:316-345."""
    features = OrderedDict()
    x = forward_features(cfg, w, x, return_features)
    if return_features:
        x, features = x
    x = x.mean(dim=(1, 2))
    features["features"] = x
    if cfg.nb_classes > 0:
        x = tf.dense(x, w["classifier/kernel"], w["classifier/bias"])
    features["logits"] = x
    return (x, features) if return_features else x
