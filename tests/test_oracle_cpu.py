"""The oracle is only worth something if it is pinned.  TensorFlow is not installed, so:

1. each TF/Keras op semantic the restatement assumes is checked against a closed form;
2. the restated GRAPHS are checked against torchvision's independent implementations of the same
   timm-parameterised networks (ViT, ConvNeXt, Swin, ResNet, EfficientNet-B0 with symmetric padding),
   weights mapped by the reference's own conversion rules (tfimm/utils/timm.py:39-106: Dense kernels are
   weight.T, conv kernels weight.permute(2,3,1,0), depthwise (C,1,k,k)->(k,k,C,1), BN weight/bias/
   running_mean/running_var -> gamma/beta/moving_mean/moving_variance).  Error metric and threshold are the
   reference's (tests/test_timm.py:71, scripts/test_conversion.py:84-86);
3. parameter counts equal the reference's published flops.csv values;
4. committed golden logits (tests/golden/*.npz, written by tools/make_golden.py) pin the oracle itself.
"""
import math
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torchvision

import tfimm
from oracle import convnext as oconvnext
from oracle import efficientnet as oeff
from oracle import params
from oracle import resnet as oresnet
from oracle import swin as oswin
from oracle import tf_ops as tf
from oracle import vit as ovit

from pathlib import Path

GOLDEN = Path(__file__).resolve().parent / "golden"


def nerr(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-6)).item()


# ------------------------------------------------------------------------------------------ op semantics
def test_gelu_is_exact_erf_form():
    x = torch.linspace(-4, 4, 101, dtype=torch.float64)
    ref = torch.tensor([0.5 * v * (1 + math.erf(v / math.sqrt(2))) for v in x.tolist()], dtype=torch.float64)
    assert (tf.act(x, "gelu") - ref).abs().max() < 1e-12
    tanh_form = 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * x ** 3)))
    assert (tf.act(x, "gelu") - tanh_form).abs().max() > 1e-4  # and is NOT the tanh approximation


def test_layer_norm_biased_variance_eps_inside():
    x = torch.tensor([[1.0, 2.0, 3.0, 6.0]], dtype=torch.float64)
    g, b = torch.tensor([1.0, 2.0, 1.0, 0.5], dtype=torch.float64), torch.tensor([0.0, 1.0, 0.0, -1.0], dtype=torch.float64)
    mean, var = 3.0, (4 + 1 + 0 + 9) / 4  # biased
    ref = (x - mean) / math.sqrt(var + 1e-3) * g + b
    assert (tf.layer_norm(x, g, b, 1e-3) - ref).abs().max() < 1e-12


def test_batch_norm_inference():
    x = torch.tensor([[[[2.0, -1.0]]]], dtype=torch.float64)
    out = tf.batch_norm(x, torch.tensor([2.0, 1.0]), torch.tensor([0.5, 0.0]), torch.tensor([1.0, 1.0]),
                        torch.tensor([3.0, 0.0]), 1.0)
    assert torch.allclose(out, torch.tensor([[[[2.0 * 1 / 2 + 0.5, -2.0]]]], dtype=torch.float64))


@pytest.mark.parametrize("size,k,s,expect", [(224, 3, 2, (0, 1)), (225, 3, 2, (1, 1)), (380, 5, 2, (1, 2)), (7, 3, 1, (1, 1)),
                                             (8, 1, 2, (0, 0)), (5, 7, 1, (3, 3))])
def test_tf_same_padding_is_asymmetric(size, k, s, expect):
    assert tf.same_padding(size, k, s) == expect
    out = -(-size // s)
    assert (size + sum(expect) - k) // s + 1 == out


def test_conv_same_stride2_matches_manual_padding():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 6, 6, 2, generator=g, dtype=torch.float64)
    k = torch.randn(3, 3, 2, 4, generator=g, dtype=torch.float64)
    y = tf.conv2d(x, k, stride=2, padding="same")
    xp = torch.nn.functional.pad(x, (0, 0, 0, 1, 0, 1))  # extra row/col at bottom/right only
    ref = torch.stack([(xp[0, 2 * i:2 * i + 3, 2 * j:2 * j + 3, :, None] * k).sum(dim=(0, 1, 2))
                       for i in range(3) for j in range(3)]).reshape(1, 3, 3, 4)
    assert (y - ref).abs().max() < 1e-12
    sym = tf.conv2d(x, k, stride=2, padding="symmetric")
    assert sym.shape == y.shape and (sym - y).abs().max() > 1e-3  # PyTorch-style padding differs


def test_depthwise_kernel_layout():
    x = torch.zeros(1, 3, 3, 2, dtype=torch.float64)
    x[0, 1, 1, 0], x[0, 1, 1, 1] = 1.0, 10.0
    k = torch.zeros(3, 3, 2, 1, dtype=torch.float64)
    k[0, 2, 0, 0], k[2, 0, 1, 0] = 5.0, 7.0
    y = tf.depthwise_conv2d(x, k, padding="same")
    assert y[0, 2, 0, 0] == 5.0 and y[0, 0, 2, 1] == 70.0 and y.abs().sum() == 75.0


def test_roll_and_dense_and_softmax():
    x = torch.arange(5.0)
    assert tf.roll(x, -2, 0).tolist() == [2, 3, 4, 0, 1]  # y[i] = x[(i + 2) mod n]
    g = torch.Generator().manual_seed(0)
    a, w = torch.randn(2, 3, 4, generator=g), torch.randn(4, 5, generator=g)
    assert torch.allclose(tf.dense(a, w), torch.einsum("bnk,kj->bnj", a, w), atol=1e-5)
    s = tf.softmax(torch.tensor([[1000.0, 1000.0]]))
    assert torch.allclose(s, torch.tensor([[0.5, 0.5]]))


def test_group_norm_matches_torch_and_blur_pool_matches_direct_sum():
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 5, 6, 64, generator=g, dtype=torch.float64)
    gamma, beta = torch.randn(64, generator=g, dtype=torch.float64), torch.randn(64, generator=g, dtype=torch.float64)
    ref = torch.nn.functional.group_norm(x.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5).permute(0, 2, 3, 1)
    assert (tf.group_norm(x, gamma, beta, 32, 1e-5) - ref).abs().max() < 1e-12
    # BlurPool2D: REFLECT pad 1 (index -1 -> 1, H -> H-2), [1 2 1] x [1 2 1] / 16, stride 2, VALID
    y = tf.blur_pool2d(x, 2)
    assert y.shape == (2, 3, 3, 64)
    k = [1.0, 2.0, 1.0]
    refl = lambda i, n: -i if i < 0 else (2 * n - 2 - i if i >= n else i)
    for (oy, ox) in [(0, 0), (2, 2), (1, 0)]:
        acc = sum(k[a] * k[b] / 16.0 * x[:, refl(2 * oy + a - 1, 5), refl(2 * ox + b - 1, 6)]
                  for a in range(3) for b in range(3))
        assert (y[:, oy, ox] - acc).abs().max() < 1e-12


def test_avg_pool_same_excludes_padding():
    x = torch.ones(1, 3, 3, 1)
    assert torch.allclose(tf.avg_pool2d_same(x, 2, 2), torch.ones(1, 2, 2, 1))


def test_tf_bicubic_resize_properties():
    g = torch.Generator().manual_seed(1)
    img = torch.randn(1, 5, 7, 3, generator=g, dtype=torch.float64)
    assert (tf.resize_bicubic(img, (5, 7)) - img).abs().max() < 1e-12          # identity at native size
    const = torch.full((1, 4, 4, 1), 2.5, dtype=torch.float64)
    assert (tf.resize_bicubic(const, (9, 6)) - 2.5).abs().max() < 1e-6          # weights renormalised (in float32)
    up = tf.resize_bicubic(img, (10, 14))
    from tfimm.layers import tf_bicubic_resize
    assert (tf_bicubic_resize(img, (10, 14)) - up).abs().max() < 1e-6           # engine and oracle agree (fp32 weights)


# ------------------------------------------------------------------------------------------ parameter counts
@pytest.mark.parametrize("name,count", [("vit_base_patch16_224", 86567656), ("vit_tiny_patch16_224", 5717416),
                                        ("swin_base_patch4_window7_224", 88104377), ("resnet50", 25610152)])
def test_parameter_counts_match_reference_flops_csv(name, count):
    """results/profiling/flops.csv:142,175,123,81."""
    assert tfimm.create_model(name, device="meta").count_params() == count


@pytest.mark.parametrize("fam,mod", [("vit", ovit), ("swin", oswin), ("convnext", oconvnext), ("efficientnet", oeff), ("resnet", oresnet)])
def test_engine_and_oracle_agree_on_every_variable_name_and_shape(fam, mod):
    """Two independently written enumerations of the reference's variables (engine: _param_specs)."""
    for name in tfimm.list_models(module=fam):
        try:
            m = tfimm.create_model(name, device="meta")
        except NotImplementedError:
            continue
        ignore = set(getattr(m, "keys_to_ignore_on_load_missing", [])) if fam == "swin" else set()
        a = {k: tuple(v.shape) for k, v in m.params.items() if k not in ignore}
        b = {k: tuple(v) for k, v in mod.param_shapes(m.cfg).items()}
        assert a == b, name


# ------------------------------------------------------------------------------------------ torchvision pins
def _randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in module.named_parameters():
            if p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(p[0].numel()))
            elif "weight" in n or n.endswith("layer_scale"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
        for n, b in module.named_buffers():
            if n.endswith("running_mean"):
                b.copy_(0.2 * torch.randn(b.shape, generator=g))
            elif n.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=g))
    return module.eval().double()


def _lin(sd, name):
    return sd[name + ".weight"].t().contiguous()


def _conv(sd, name):
    return sd[name + ".weight"].permute(2, 3, 1, 0).contiguous()


def _bn(w, dst, sd, src):
    w[dst + "/gamma"], w[dst + "/beta"] = sd[src + ".weight"], sd[src + ".bias"]
    w[dst + "/moving_mean"], w[dst + "/moving_variance"] = sd[src + ".running_mean"], sd[src + ".running_var"]


def _images(b, h, w, seed=2021):
    return torch.from_numpy(np.random.default_rng(seed).random((b, h, w, 3))).double()


def test_vit_oracle_matches_torchvision():
    tv = _randomize(torchvision.models.VisionTransformer(image_size=64, patch_size=16, num_layers=3, num_heads=3,
                                                         hidden_dim=48, mlp_dim=96, num_classes=11), 1)
    with torch.no_grad():
        tv.class_token.copy_(torch.randn(tv.class_token.shape, dtype=torch.float64) * 0.3)
        tv.encoder.pos_embedding.copy_(torch.randn(tv.encoder.pos_embedding.shape, dtype=torch.float64) * 0.3)
    sd = tv.state_dict()
    cfg = SimpleNamespace(nb_classes=11, in_channels=3, input_size=(64, 64), patch_size=16, embed_dim=48, nb_blocks=3,
                          nb_heads=3, mlp_ratio=2.0, qkv_bias=True, representation_size=None, distilled=False,
                          norm_layer="layer_norm_eps_1e-6", act_layer="gelu", interpolate_input=False)
    w = {"patch_embed/proj/kernel": _conv(sd, "conv_proj"), "patch_embed/proj/bias": sd["conv_proj.bias"],
         "cls_token": sd["class_token"], "pos_embed": sd["encoder.pos_embedding"],
         "norm/gamma": sd["encoder.ln.weight"], "norm/beta": sd["encoder.ln.bias"],
         "head/kernel": _lin(sd, "heads.head"), "head/bias": sd["heads.head.bias"]}
    for j in range(3):
        s, d = f"encoder.layers.encoder_layer_{j}", f"blocks/{j}"
        w[f"{d}/norm1/gamma"], w[f"{d}/norm1/beta"] = sd[f"{s}.ln_1.weight"], sd[f"{s}.ln_1.bias"]
        w[f"{d}/attn/qkv/kernel"], w[f"{d}/attn/qkv/bias"] = sd[f"{s}.self_attention.in_proj_weight"].t().contiguous(), sd[f"{s}.self_attention.in_proj_bias"]
        w[f"{d}/attn/proj/kernel"], w[f"{d}/attn/proj/bias"] = _lin(sd, f"{s}.self_attention.out_proj"), sd[f"{s}.self_attention.out_proj.bias"]
        w[f"{d}/norm2/gamma"], w[f"{d}/norm2/beta"] = sd[f"{s}.ln_2.weight"], sd[f"{s}.ln_2.bias"]
        w[f"{d}/mlp/fc1/kernel"], w[f"{d}/mlp/fc1/bias"] = _lin(sd, f"{s}.mlp.0"), sd[f"{s}.mlp.0.bias"]
        w[f"{d}/mlp/fc2/kernel"], w[f"{d}/mlp/fc2/bias"] = _lin(sd, f"{s}.mlp.3"), sd[f"{s}.mlp.3.bias"]
    assert set(w) == set(ovit.param_shapes(cfg))
    x = _images(2, 64, 64)
    with torch.no_grad():
        ref = tv(x.permute(0, 3, 1, 2))
    assert nerr(ovit.forward(cfg, w, x), ref) < 1e-9


def test_convnext_oracle_matches_torchvision():
    from torchvision.models.convnext import CNBlockConfig, ConvNeXt

    tv = _randomize(ConvNeXt([CNBlockConfig(16, 32, 2), CNBlockConfig(32, 64, 2), CNBlockConfig(64, None, 1)],
                             stochastic_depth_prob=0.0, layer_scale=0.7, num_classes=9), 2)
    sd = tv.state_dict()
    cfg = SimpleNamespace(nb_classes=9, in_channels=3, input_size=(64, 64), patch_size=4, embed_dim=(16, 32, 64),
                          nb_blocks=(2, 2, 1), mlp_ratio=4.0, conv_mlp_block=False, norm_layer="layer_norm_eps_1e-6",
                          act_layer="gelu")
    w = {"stem/0/kernel": _conv(sd, "features.0.0"), "stem/0/bias": sd["features.0.0.bias"],
         "stem/1/gamma": sd["features.0.1.weight"], "stem/1/beta": sd["features.0.1.bias"],
         "head/norm/gamma": sd["classifier.0.weight"], "head/norm/beta": sd["classifier.0.bias"],
         "head/fc/kernel": _lin(sd, "classifier.2"), "head/fc/bias": sd["classifier.2.bias"]}
    for j, depth in enumerate(cfg.nb_blocks):
        if j > 0:
            s = f"features.{2 * j}"
            w[f"stages/{j}/downsample/0/gamma"], w[f"stages/{j}/downsample/0/beta"] = sd[f"{s}.0.weight"], sd[f"{s}.0.bias"]
            w[f"stages/{j}/downsample/1/kernel"], w[f"stages/{j}/downsample/1/bias"] = _conv(sd, f"{s}.1"), sd[f"{s}.1.bias"]
        for k in range(depth):
            s, d = f"features.{2 * j + 1}.{k}", f"stages/{j}/blocks/{k}"
            w[f"{d}/conv_dw/depthwise_kernel"] = sd[f"{s}.block.0.weight"].permute(2, 3, 0, 1).contiguous()
            w[f"{d}/conv_dw/bias"] = sd[f"{s}.block.0.bias"]
            w[f"{d}/norm/gamma"], w[f"{d}/norm/beta"] = sd[f"{s}.block.2.weight"], sd[f"{s}.block.2.bias"]
            w[f"{d}/mlp/fc1/kernel"], w[f"{d}/mlp/fc1/bias"] = _lin(sd, f"{s}.block.3"), sd[f"{s}.block.3.bias"]
            w[f"{d}/mlp/fc2/kernel"], w[f"{d}/mlp/fc2/bias"] = _lin(sd, f"{s}.block.5"), sd[f"{s}.block.5.bias"]
            w[f"{d}/gamma"] = sd[f"{s}.layer_scale"].reshape(-1)
    assert set(w) == set(oconvnext.param_shapes(cfg))
    x = _images(2, 64, 64)
    with torch.no_grad():
        ref = tv(x.permute(0, 3, 1, 2))
    assert nerr(oconvnext.forward(cfg, w, x), ref) < 1e-9


def test_swin_oracle_matches_torchvision():
    """The reference never pinned Swin (tests/test_timm.py:29-30 is commented out); this does, including
    the shifted-window mask, the relative-position gather and PatchMerging's neighbour order."""
    tv = _randomize(torchvision.models.SwinTransformer(patch_size=[4, 4], embed_dim=32, depths=[2, 2], num_heads=[1, 2],
                                                       window_size=[7, 7], stochastic_depth_prob=0.0, num_classes=13), 3)
    sd = tv.state_dict()
    cfg = SimpleNamespace(nb_classes=13, in_channels=3, input_size=(112, 112), patch_size=4, embed_dim=32, nb_blocks=(2, 2),
                          nb_heads=(1, 2), window_size=7, mlp_ratio=4.0, qkv_bias=True, norm_layer="layer_norm",
                          act_layer="gelu", patch_norm=True)
    w = {"patch_embed/proj/kernel": _conv(sd, "features.0.0"), "patch_embed/proj/bias": sd["features.0.0.bias"],
         "patch_embed/norm/gamma": sd["features.0.2.weight"], "patch_embed/norm/beta": sd["features.0.2.bias"],
         "norm/gamma": sd["norm.weight"], "norm/beta": sd["norm.bias"],
         "head/kernel": _lin(sd, "head"), "head/bias": sd["head.bias"]}
    for i, depth in enumerate(cfg.nb_blocks):
        for j in range(depth):
            s, d = f"features.{2 * i + 1}.{j}", f"layers/{i}/blocks/{j}"
            w[f"{d}/norm1/gamma"], w[f"{d}/norm1/beta"] = sd[f"{s}.norm1.weight"], sd[f"{s}.norm1.bias"]
            w[f"{d}/attn/qkv/kernel"], w[f"{d}/attn/qkv/bias"] = _lin(sd, f"{s}.attn.qkv"), sd[f"{s}.attn.qkv.bias"]
            w[f"{d}/attn/proj/kernel"], w[f"{d}/attn/proj/bias"] = _lin(sd, f"{s}.attn.proj"), sd[f"{s}.attn.proj.bias"]
            w[f"{d}/attn/relative_position_bias_table"] = sd[f"{s}.attn.relative_position_bias_table"]
            w[f"{d}/norm2/gamma"], w[f"{d}/norm2/beta"] = sd[f"{s}.norm2.weight"], sd[f"{s}.norm2.bias"]
            w[f"{d}/mlp/fc1/kernel"], w[f"{d}/mlp/fc1/bias"] = _lin(sd, f"{s}.mlp.0"), sd[f"{s}.mlp.0.bias"]
            w[f"{d}/mlp/fc2/kernel"], w[f"{d}/mlp/fc2/bias"] = _lin(sd, f"{s}.mlp.3"), sd[f"{s}.mlp.3.bias"]
        if i < len(cfg.nb_blocks) - 1:
            s = f"features.{2 * i + 2}"
            w[f"layers/{i}/downsample/norm/gamma"], w[f"layers/{i}/downsample/norm/beta"] = sd[f"{s}.norm.weight"], sd[f"{s}.norm.bias"]
            w[f"layers/{i}/downsample/reduction/kernel"] = _lin(sd, f"{s}.reduction")
    assert set(w) == set(oswin.param_shapes(cfg))
    x = _images(2, 112, 112)
    with torch.no_grad():
        ref = tv(x.permute(0, 3, 1, 2))
    assert nerr(oswin.forward(cfg, w, x), ref) < 1e-9


@pytest.mark.parametrize("arch,block,depths", [("resnet18", "basic_block", (2, 2, 2, 2)), ("resnet50", "bottleneck", (3, 4, 6, 3)),
                                               ("resnext50_32x4d", "bottleneck", (3, 4, 6, 3))])
def test_resnet_oracle_matches_torchvision(arch, block, depths):
    tv = _randomize(getattr(torchvision.models, arch)(num_classes=7), 4)
    sd = tv.state_dict()
    cfg = tfimm.models.model_config(arch)
    cfg = SimpleNamespace(**{**cfg.__dict__, "nb_classes": 7, "input_size": (96, 96)})
    w = {"conv1/kernel": _conv(sd, "conv1"), "remove/fc/kernel": _lin(sd, "fc"), "remove/fc/bias": sd["fc.bias"]}
    _bn(w, "bn1", sd, "bn1")
    for i, depth in enumerate(depths):
        for b in range(depth):
            s, d = f"layer{i + 1}.{b}", f"layer{i + 1}/{b}"
            for c in range(1, 3 if block == "basic_block" else 4):
                w[f"{d}/conv{c}/kernel"] = _conv(sd, f"{s}.conv{c}")
                _bn(w, f"{d}/bn{c}", sd, f"{s}.bn{c}")
            if f"{s}.downsample.0.weight" in sd:
                w[f"{d}/downsample/0/kernel"] = _conv(sd, f"{s}.downsample.0")
                _bn(w, f"{d}/downsample/1", sd, f"{s}.downsample.1")
    assert set(w) == set(oresnet.param_shapes(cfg))
    x = _images(2, 96, 96)
    with torch.no_grad():
        ref = tv(x.permute(0, 3, 1, 2))
    assert nerr(oresnet.forward(cfg, w, x), ref) < 1e-9


def test_efficientnet_oracle_matches_torchvision():
    """torchvision's efficientnet_b0 == the reference's pt_efficientnet_b0 (symmetric padding, BN eps 1e-5)."""
    tv = _randomize(torchvision.models.efficientnet_b0(num_classes=5), 5)
    sd = tv.state_dict()
    base = tfimm.models.model_config("pt_efficientnet_b0")
    cfg = SimpleNamespace(**{**base.__dict__, "nb_classes": 5, "input_size": (96, 96)})
    assert tv.features[1][0].block[0][1].eps == 1e-5
    w = {"conv_stem/kernel": _conv(sd, "features.0.0"), "conv_head/kernel": _conv(sd, "features.8.0"),
         "classifier/kernel": _lin(sd, "classifier.1"), "classifier/bias": sd["classifier.1.bias"]}
    _bn(w, "bn1", sd, "features.0.1")
    _bn(w, "bn2", sd, "features.8.1")

    def se(dst, src):
        w[f"{dst}/conv_reduce/kernel"], w[f"{dst}/conv_reduce/bias"] = _conv(sd, f"{src}.fc1"), sd[f"{src}.fc1.bias"]
        w[f"{dst}/conv_expand/kernel"], w[f"{dst}/conv_expand/bias"] = _conv(sd, f"{src}.fc2"), sd[f"{src}.fc2.bias"]

    def dw(name):
        return sd[name + ".weight"].permute(2, 3, 0, 1).contiguous()

    for ba in oeff.build_blocks(cfg):
        si, bi = (int(v) for v in ba["name"].split(".")[1:])
        s, d = f"features.{si + 1}.{bi}.block", ba["name"]
        if ba["block_type"] == "ds":
            w[f"{d}/conv_dw/depthwise_kernel"] = dw(f"{s}.0.0")
            _bn(w, f"{d}/bn1", sd, f"{s}.0.1")
            se(f"{d}/se", f"{s}.1")
            w[f"{d}/conv_pw/kernel"] = _conv(sd, f"{s}.2.0")
            _bn(w, f"{d}/bn2", sd, f"{s}.2.1")
        else:
            w[f"{d}/conv_pw/kernel"] = _conv(sd, f"{s}.0.0")
            _bn(w, f"{d}/bn1", sd, f"{s}.0.1")
            w[f"{d}/conv_dw/depthwise_kernel"] = dw(f"{s}.1.0")
            _bn(w, f"{d}/bn2", sd, f"{s}.1.1")
            se(f"{d}/se", f"{s}.2")
            w[f"{d}/conv_pwl/kernel"] = _conv(sd, f"{s}.3.0")
            _bn(w, f"{d}/bn3", sd, f"{s}.3.1")
    shapes = oeff.param_shapes(cfg)
    assert set(w) == set(shapes)
    assert all(tuple(w[k].shape) == tuple(shapes[k]) for k in w)
    x = _images(2, 96, 96)
    with torch.no_grad():
        ref = tv(x.permute(0, 3, 1, 2))
    assert nerr(oeff.forward(cfg, w, x), ref) < 1e-9


# ------------------------------------------------------------------------------------------ invariants / golden
def test_return_features_does_not_change_logits():
    cfg = tfimm.models.model_config("vit_tiny_patch16_224")
    cfg = SimpleNamespace(**{**cfg.__dict__, "nb_blocks": 2, "input_size": (32, 32), "distilled": False,
                             "representation_size": None})
    w = params.random_params(ovit.param_shapes(cfg), seed=1)
    x = params.test_images(2, 32, 32)
    y, feats = ovit.forward(cfg, w, x, return_features=True)
    assert torch.equal(y, ovit.forward(cfg, w, x)) and list(feats)[-1] == "logits"


@pytest.mark.parametrize("fixture", sorted(p.name for p in GOLDEN.glob("*.npz")))
def test_oracle_reproduces_committed_golden_logits(fixture):
    import importlib

    data = np.load(GOLDEN / fixture, allow_pickle=True)
    meta = data["meta"].item()
    mod = importlib.import_module(f"oracle.{meta['family']}")
    cfg = tfimm.models.model_config(meta["model"])
    cfg = type(cfg)(**{**cfg.__dict__, **meta["overrides"]})
    w = params.random_params(mod.param_shapes(cfg), seed=meta["seed"])
    x = params.test_images(meta["batch"], *cfg.input_size, cfg.in_channels, seed=meta.get("images_seed", 2021))
    if "images" in data.files:
        assert np.array_equal(x.numpy(), data["images"])
    if fixture.startswith("full_"):
        x, want = x[:1], data["logits"][:1]  # full-size BASELINE configs: one image keeps the CPU suite short
    else:
        want = data["logits"]
    with torch.no_grad():
        y = mod.forward(cfg, w, x)
    # the stored logits were computed by the reference's own code on the TF shim (tools/make_golden.py)
    assert "reference" in meta.get("source", "")
    assert nerr(y, torch.from_numpy(want)) < 5e-6


def test_pytorch_state_dict_ingestion_reproduces_torchvision_resnet():
    """N1: a PyTorch state_dict in timm naming (torchvision's ResNet uses the same names) converted by
    tfimm.utils.timm (rules of reference utils/timm.py:39-106) and run through the oracle gives torchvision's
    own logits."""
    from tfimm.utils.timm import convert_state_dict, pytorch_key

    assert pytorch_key("remove/fc/kernel") == "fc.weight"
    assert pytorch_key("layer1/0/downsample/1/moving_variance") == "layer1.0.downsample.1.running_var"
    tv = _randomize(torchvision.models.resnet50(), 6).float()
    model = tfimm.create_model("resnet50", device="cpu")
    weights, missing, unexpected = convert_state_dict(model, tv.state_dict())
    assert not missing and not unexpected
    assert set(weights) == set(model.params)
    model.load_weights_dict(weights)
    x = _images(1, 96, 96).float()
    with torch.no_grad():
        ref = tv(x.permute(0, 3, 1, 2))
    got = oresnet.forward(model.cfg, {k: v for k, v in model.params.items()}, x)
    assert nerr(got, ref) < 1e-4
    with pytest.raises(AttributeError):
        convert_state_dict(tfimm.create_model("resnet18", device="meta"), {"conv1.weight": np.zeros((64, 3, 7, 7))})
