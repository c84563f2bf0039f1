"""SURVEY.md 8(f) rows on the GPU, through the engine:

N1  a PyTorch ``state_dict`` converted by ``tfimm.utils.timm`` (the reference's rules, tfimm/utils/timm.py:109-229),
    loaded into the ENGINE, reproduces the PyTorch model's own fp32 output (what tests/test_timm.py:38-71 of the
    reference checks for the TF model);
N3  ``return_features`` for EfficientNet and ResNet (tests/models/test_factory.py:205-222);
N4  ``in_channels`` 1 / 6 and ``nb_classes`` changes through ``create_model`` + ``transfer_weights`` preserve outputs /
    features (tests/models/test_factory.py:37-90).
"""
import importlib
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nerr(out, ref):
    out, ref = out.float().cpu(), ref.float().cpu()
    return (out - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)


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
    return module.eval().float()


def _images(b, h, w, c=3, seed=2021):
    return torch.from_numpy(np.random.default_rng(seed).random((b, h, w, c), dtype=np.float32))


def _vit_timm_names(sd):
    out = {}
    for k, v in sd.items():
        k = (k.replace("encoder.layers.encoder_layer_", "blocks.").replace("ln_1", "norm1").replace("ln_2", "norm2")
             .replace("self_attention.in_proj_", "attn.qkv.").replace("self_attention.out_proj", "attn.proj")
             .replace("mlp.0", "mlp.fc1").replace("mlp.3", "mlp.fc2").replace("encoder.ln", "norm")
             .replace("conv_proj", "patch_embed.proj").replace("heads.head", "head")
             .replace("class_token", "cls_token").replace("encoder.pos_embedding", "pos_embed"))
        out[k] = v
    return out


def _swin_timm_names(sd, depths):
    out = {}
    for k, v in sd.items():
        if "relative_position_index" in k:
            continue
        k = k.replace("features.0.0.", "patch_embed.proj.").replace("features.0.2.", "patch_embed.norm.")
        for i in range(len(depths)):
            k = k.replace(f"features.{2 * i + 1}.", f"layers.{i}.blocks.")
            k = k.replace(f"features.{2 * i + 2}.", f"layers.{i}.downsample.")
        k = k.replace(".mlp.0.", ".mlp.fc1.").replace(".mlp.3.", ".mlp.fc2.")
        out[k] = v
    return out


@pytest.mark.parametrize("arch", ["resnet50", "vit", "swin"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 1.2e-2)])
def test_pytorch_state_dict_loaded_into_engine_reproduces_pytorch_logits(arch, precision, tol):
    import torchvision

    import tfimm
    from tfimm.utils.timm import load_pytorch_weights_in_model

    if arch == "resnet50":
        tv = _randomize(torchvision.models.resnet50(num_classes=17), 4)
        model = tfimm.create_model("resnet50", precision=precision, device="cuda", nb_classes=17, input_size=(96, 96))
        sd, size = tv.state_dict(), (96, 96)
    elif arch == "vit":
        tv = _randomize(torchvision.models.VisionTransformer(image_size=64, patch_size=16, num_layers=3, num_heads=3,
                                                             hidden_dim=192, mlp_dim=768, num_classes=11), 1)
        with torch.no_grad():
            tv.class_token.copy_(torch.randn(tv.class_token.shape) * 0.3)
            tv.encoder.pos_embedding.copy_(torch.randn(tv.encoder.pos_embedding.shape) * 0.3)
        model = tfimm.create_model("vit_tiny_patch16_224", precision=precision, device="cuda", nb_classes=11,
                                   input_size=(64, 64), nb_blocks=3)
        sd, size = _vit_timm_names(tv.state_dict()), (64, 64)
    else:
        tv = _randomize(torchvision.models.SwinTransformer(patch_size=[4, 4], embed_dim=96, depths=[2, 2],
                                                           num_heads=[3, 6], window_size=[7, 7],
                                                           stochastic_depth_prob=0.0, num_classes=13), 3)
        model = tfimm.create_model("swin_tiny_patch4_window7_224", precision=precision, device="cuda", nb_classes=13,
                                   input_size=(112, 112), nb_blocks=(2, 2), nb_heads=(3, 6))
        sd, size = _swin_timm_names(tv.state_dict(), (2, 2)), (112, 112)
    missing, unexpected = load_pytorch_weights_in_model(model, sd)
    assert not missing and not unexpected, (missing, unexpected)
    x = _images(2, *size)
    with torch.no_grad():
        ref = tv(x.permute(0, 3, 1, 2))
    out = model(x.cuda())
    err = _nerr(out, ref)
    print(f"{arch} {precision}: engine(converted state_dict) vs PyTorch fp32 {err:.3e}")
    assert out.shape == ref.shape
    assert err < tol


@pytest.mark.parametrize("name,family,overrides", [
    ("efficientnet_b0", "efficientnet", {"input_size": (96, 96)}),
    ("efficientnet_v2_b0", "efficientnet", {"input_size": (96, 96)}),
    ("resnet50", "resnet", {"input_size": (96, 96)}),
    ("seresnext26d_32x4d", "resnet", {"input_size": (96, 96)}),
])
def test_return_features_efficientnet_and_resnet(name, family, overrides):
    import tfimm
    from oracle import params

    omod = importlib.import_module(f"oracle.{family}")
    model = tfimm.create_model(name, precision="fp32", device="cuda", **overrides)
    w = params.random_params(omod.param_shapes(model.cfg), seed=41)
    model.load_weights_dict(w)
    x = params.test_images(2, *model.cfg.input_size)
    y = model(x.cuda())
    y2, feats = model(x.cuda(), return_features=True)
    assert (y - y2).abs().max().item() < 1e-5
    with torch.no_grad():
        _, ofeats = omod.forward(model.cfg, w, x, return_features=True)
    assert list(feats.keys()) == list(ofeats.keys()) == model.feature_names
    for k in feats:
        assert tuple(feats[k].shape) == tuple(ofeats[k].shape), k
        assert _nerr(feats[k], ofeats[k]) < 5e-5, k
    # bf16 engine: same keys and shapes, values within the bf16 budget
    mb = tfimm.create_model(name, precision="bf16", device="cuda", **overrides)
    mb.load_weights_dict(w)
    yb, fb = mb(x.cuda(), return_features=True)
    assert list(fb.keys()) == list(ofeats.keys())
    assert _nerr(fb["logits"], ofeats["logits"]) < 8e-3


@pytest.mark.parametrize("name,family,overrides", [
    ("vit_tiny_patch16_224", "vit", {"nb_blocks": 2}),
    ("convnext_tiny", "convnext", {"input_size": (64, 64), "nb_blocks": (1, 1, 1, 1)}),
    ("resnet18", "resnet", {"input_size": (64, 64)}),
    ("efficientnet_b0", "efficientnet", {"input_size": (64, 64)}),
    ("swin_tiny_patch4_window7_224", "swin", {"input_size": (112, 112), "nb_blocks": (2, 2), "nb_heads": (3, 6)}),
])
def test_nb_classes_and_in_channels_changes_preserve_outputs(name, family, overrides):
    """reference tests/models/test_factory.py:37-52 (nb_classes keeps features) and :55-90 (in_channels 1 / 6: a
    grey image fed as 1 channel equals the same image tiled to 3 channels, up to the channel rescaling rule)."""
    import tfimm
    from oracle import params

    omod = importlib.import_module(f"oracle.{family}")
    src = tfimm.create_model(name, precision="fp32", device="cuda", **overrides)
    w = params.random_params(omod.param_shapes(src.cfg), seed=43)
    src.load_weights_dict(w)
    h, wd = src.cfg.input_size
    x3 = params.test_images(2, h, wd, 3).cuda()
    _, f_src = src(x3, return_features=True)

    # nb_classes: new classifier, identical features
    dst = tfimm.create_model(name, precision="fp32", device="cuda", nb_classes=7, **overrides)
    tfimm.models.transfer_weights(src, dst)
    y, f_dst = dst(x3, return_features=True)
    assert y.shape[-1] == 7
    assert _nerr(f_dst["features"], f_src["features"]) < 1e-5
    # nb_classes = 0: the classifier is removed and the model returns the features
    nocls = tfimm.create_model(name, precision="fp32", device="cuda", nb_classes=0, **overrides)
    tfimm.models.transfer_weights(src, nocls)
    feats = f_src["features"]
    if feats.dim() == 4:   # ResNet reports the un-pooled feature map (resnet.py:570-584); the head pools it
        feats = feats.float().mean(dim=(1, 2))
    assert _nerr(nocls(x3), feats) < 1e-5

    # in_channels = 1: first conv summed over the input channels -> grey image == grey image repeated 3x
    grey = params.test_images(2, h, wd, 1).cuda()
    one = tfimm.create_model(name, precision="fp32", device="cuda", in_channels=1, **overrides)
    tfimm.models.transfer_weights(src, one)
    assert _nerr(one(grey), src(grey.repeat(1, 1, 1, 3))) < 2e-5
    # in_channels = 6: first conv tiled and rescaled by 3/6 -> image stacked twice == original image
    six = tfimm.create_model(name, precision="fp32", device="cuda", in_channels=6, **overrides)
    tfimm.models.transfer_weights(src, six)
    assert _nerr(six(torch.cat([x3, x3], dim=-1)), src(x3)) < 2e-5
