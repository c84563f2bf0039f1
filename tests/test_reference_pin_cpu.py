"""Pins the oracle (and the engine's host-side API) to the REFERENCE ITSELF.

TensorFlow cannot be installed here, so ``oracle/ref_runner.py`` executes the unmodified reference sources
(``/root/reference/tfimm/architectures/{vit,swin,convnext,efficientnet,resnet}.py`` + ``tfimm/layers`` +
``tfimm/models/{factory,registry}.py`` + ``tfimm/utils/timm.py``) on a torch-CPU restatement of the TF/Keras calls
they make (``oracle/tf_shim``).  Everything below compares against what the reference's own ``call()`` code
computes, with its own variable names, on identical seeded inputs (tests/test_timm.py:56-71 of the reference).

``/root/reference`` exists only in the build container; on the GPU box these tests skip and the committed
fixtures in ``tests/golden`` (generated from the same reference run, ``tools/make_golden.py``) stand in.
"""
import importlib
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tensorflow-image-models_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import params  # noqa: E402
from oracle import ref_runner as rr  # noqa: E402

pytestmark = pytest.mark.skipif(not rr.available(), reason="/root/reference is not present on this machine")

IGNORE = ("attn_mask", "relative_position_index", "blur_kernel")

# (family, registered name, create_model overrides) -- the small configurations of the reference's own test-suite
# (tests/models/architectures.py: *_test_model) expressed as overrides of registered models, plus real registrations.
CASES = [
    ("vit", "vit_tiny_patch16_224", {"input_size": (32, 32), "patch_size": 8, "embed_dim": 4, "nb_blocks": 2, "nb_heads": 2, "nb_classes": 12}),
    ("vit", "deit_tiny_distilled_patch16_224", {"input_size": (32, 32), "patch_size": 8, "embed_dim": 4, "nb_blocks": 2, "nb_heads": 2, "nb_classes": 12}),
    ("vit", "vit_base_patch32_224_in21k", {"input_size": (64, 64), "embed_dim": 24, "nb_blocks": 2, "nb_heads": 3, "representation_size": 16, "nb_classes": 7}),
    ("vit", "vit_tiny_patch16_224", {"input_size": (96, 64), "nb_blocks": 3}),
    ("swin", "swin_tiny_patch4_window7_224", {"input_size": (32, 32), "patch_size": 2, "embed_dim": 4, "nb_blocks": (2, 2), "nb_heads": (1, 2), "window_size": 4, "nb_classes": 12}),
    ("swin", "swin_tiny_patch4_window7_224", {"input_size": (112, 112), "nb_blocks": (2, 2, 2), "nb_heads": (3, 6, 12)}),
    ("convnext", "convnext_tiny", {"input_size": (32, 32), "embed_dim": (3, 4, 5, 6), "nb_blocks": (1, 1, 1, 1), "nb_classes": 12}),
    ("convnext", "convnext_tiny", {"input_size": (64, 96), "nb_blocks": (1, 1, 2, 1)}),
    ("convnext", "convnext_tiny_in22k", {"input_size": (64, 64), "nb_blocks": (1, 1, 1, 1), "conv_mlp_block": True}),
    ("efficientnet", "efficientnet_b0", {"input_size": (64, 64)}),
    ("efficientnet", "efficientnet_b4", {"input_size": (76, 76)}),
    ("efficientnet", "pt_efficientnet_b0", {"input_size": (64, 80)}),
    ("efficientnet", "mobilenet_v2_100", {"input_size": (64, 64)}),
    ("efficientnet", "efficientnet_es", {"input_size": (64, 64)}),
    ("efficientnet", "efficientnet_lite0", {"input_size": (64, 64)}),
    ("efficientnet", "efficientnet_v2_b0", {"input_size": (64, 64)}),
    ("resnet", "resnet18", {"input_size": (64, 64)}),
    ("resnet", "resnet50", {"input_size": (64, 64)}),
    ("resnet", "resnet50d", {"input_size": (64, 64)}),
    ("resnet", "resnext50_32x4d", {"input_size": (64, 64)}),
    ("resnet", "seresnext26d_32x4d", {"input_size": (64, 64)}),
    ("resnet", "ecaresnet26t", {"input_size": (64, 64)}),
    ("resnet", "resnetblur50", {"input_size": (64, 64)}),
    ("resnet", "resnet50_gn", {"input_size": (64, 64)}),
    ("resnet", "resnetrs50", {"input_size": (64, 64)}),
]


def _nerr(a, b):
    a, b = a.double(), b.double()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-6)


def _engine_cfg(name, overrides):
    import tfimm

    base = tfimm.models.model_config(name)
    return type(base)(**{**base.__dict__, **overrides})


@pytest.fixture()
def float64_reference():
    rr.set_floatx("float64")
    yield
    rr.set_floatx("float32")


@pytest.mark.parametrize("family,name,overrides", CASES, ids=[f"{c[1]}-{i}" for i, c in enumerate(CASES)])
def test_oracle_equals_reference_code_run_on_the_shim(family, name, overrides, float64_reference):
    """Same variables (names + shapes) and, in float64, the same logits to 1e-12: the oracle restates the
    reference's graph exactly.  (float32 run of the same comparison: ~3e-7, see tools/make_golden.py.)"""
    omod = importlib.import_module(f"oracle.{family}")
    ref = rr.create_model(name, **overrides)
    cfg = _engine_cfg(name, overrides)
    shapes = omod.param_shapes(cfg)
    ref_shapes = ref.weight_shapes()
    loadable = {k: v for k, v in ref_shapes.items() if not any(p in k for p in IGNORE)}
    assert set(loadable) == set(shapes), (sorted(set(loadable) ^ set(shapes))[:6])
    for k, shp in shapes.items():
        assert tuple(shp) == loadable[k], (k, shp, loadable[k])
    w = params.random_params(shapes, seed=31, dtype=torch.float64)
    ref.assign(w, ignore_missing=IGNORE)
    x = params.test_images(2, *cfg.input_size, cfg.in_channels).double()
    y_ref, f_ref = ref(x, return_features=True)
    with torch.no_grad():
        y_or, f_or = omod.forward(cfg, w, x, return_features=True)
    assert y_ref.dtype == torch.float64 and y_ref.shape == y_or.shape
    assert _nerr(y_or, y_ref) < 1e-12
    # intermediate features: same keys in the same order, same values (tests/models/test_factory.py:205-222)
    assert list(f_ref.keys()) == list(f_or.keys())
    for k in f_ref:
        assert f_ref[k].shape == f_or[k].shape, k
        assert _nerr(f_or[k], f_ref[k]) < 1e-11, k


def test_oracle_equals_reference_in_float32_at_full_size():
    """The reference's default dtype and a real registration at its native 224 px."""
    from oracle import vit as ovit

    ref = rr.create_model("vit_tiny_patch16_224")
    cfg = _engine_cfg("vit_tiny_patch16_224", {})
    w = params.random_params(ovit.param_shapes(cfg), seed=3)
    ref.assign(w)
    x = params.test_images(1, 224, 224)
    assert _nerr(ovit.forward(cfg, w, x), ref(x)) < 2e-6


def test_vit_interpolate_input_equals_reference(float64_reference):
    """interpolate_input=True resamples pos_embed with tf.image.resize(bicubic) (layers/transformers.py:13-47)."""
    from oracle import vit as ovit

    ov = {"input_size": (64, 64), "nb_blocks": 1, "interpolate_input": True}
    ref = rr.create_model("vit_tiny_patch16_224", **ov)
    cfg = _engine_cfg("vit_tiny_patch16_224", ov)
    w = params.random_params(ovit.param_shapes(cfg), seed=4, dtype=torch.float64)
    ref.assign(w)
    x = params.test_images(1, 96, 128).double()
    # tf.image.resize returns float32, so agreement is at float32 rounding of the position table
    assert _nerr(ovit.forward(cfg, w, x), ref(x)) < 1e-6


def test_reference_initial_values_match_engine_initialisers():
    """Variables created by build(): the engine's ParamSpec initialisers name the same constants
    (zeros cls/pos tokens vit.py:378-400, ConvNeXt layer scale 1e-6 convnext.py:211-217, zero-init last BN gamma
    with moving_variance = zeros only where the reference passes it, resnet.py:147-155)."""
    import tfimm

    for name, ov in (("convnext_tiny", {"input_size": (32, 32), "nb_blocks": (1, 1, 1, 1)}),
                     ("vit_tiny_patch16_224", {"input_size": (32, 32), "nb_blocks": 1}),
                     ("resnet18", {"input_size": (32, 32)}), ("resnet50_gn", {"input_size": (32, 32)})):
        ref = rr.create_model(name, **ov).weights_dict()
        eng = tfimm.create_model(name, device="cpu", **ov)
        for key, spec in eng.param_specs().items():
            kind, _, arg = spec.init.partition(":")
            if kind in ("zeros", "ones", "const"):
                want = {"zeros": 0.0, "ones": 1.0}.get(kind, float(arg) if arg else 0.0)
                assert np.allclose(ref[key], want), (name, key, spec.init, float(np.ravel(ref[key])[0]))


def test_list_models_and_configs_equal_the_reference_registry():
    import dataclasses

    import tfimm

    for fam in rr.FAMILIES:
        ref_names = rr.list_models(module=fam)
        assert tfimm.list_models(module=fam) == ref_names
    with rr._reference_modules():
        mods = rr._import_reference()
        ref_cfgs = {n: dataclasses.asdict(mods["registry"].model_config(n)) for n in
                    ("vit_base_patch16_224", "swin_base_patch4_window7_224", "convnext_base", "efficientnet_b4",
                     "resnet50")}
    for n, rc in ref_cfgs.items():
        ec = dataclasses.asdict(tfimm.models.model_config(n))
        for k, v in rc.items():
            assert ec[k] == v or list(ec[k]) == list(v), (n, k, ec[k], v)


@pytest.mark.parametrize("name", ["vit_base_patch16_224", "convnext_base", "efficientnet_b4", "resnet50"])
def test_create_preprocessing_equals_reference(name):
    import tfimm

    img = np.random.default_rng(0).integers(0, 256, (2, 16, 16, 3)).astype(np.uint8)
    ref = rr.create_preprocessing(name, dtype="float32")
    with rr._reference_modules():
        a = ref(img)
    a = a.numpy() if hasattr(a, "numpy") else np.asarray(a)
    b = np.asarray(tfimm.create_preprocessing(name, dtype="float32")(img))
    assert np.abs(a - b).max() < 1e-6
    with pytest.raises(ValueError):
        rr.create_preprocessing("not_a_model")
    with pytest.raises(ValueError):
        tfimm.create_preprocessing("not_a_model")


@pytest.mark.parametrize("name,ov", [("resnet18", {"input_size": (32, 32)}),
                                     ("vit_tiny_patch16_224", {"input_size": (32, 32), "nb_blocks": 1}),
                                     ("convnext_tiny", {"input_size": (32, 32), "nb_blocks": (1, 1, 1, 1)})])
@pytest.mark.parametrize("change", [{"in_channels": 1}, {"in_channels": 5}, {"nb_classes": 7}])
def test_transfer_weights_equals_reference(name, ov, change):
    """in_channels / nb_classes adaptation (tfimm/models/factory.py:174-305; tests/models/test_factory.py:37-90):
    the engine's transfer_weights writes the same values into the same variables as the reference's."""
    import tfimm

    fam = {"resnet18": "resnet", "vit_tiny_patch16_224": "vit", "convnext_tiny": "convnext"}[name]
    omod = importlib.import_module(f"oracle.{fam}")
    cfg = _engine_cfg(name, ov)
    w = params.random_params(omod.param_shapes(cfg), seed=17)
    src_ref = rr.create_model(name, **ov)
    src_ref.assign(w, ignore_missing=IGNORE)
    dst_ref = rr.create_model(name, **ov, **change)
    before = dst_ref.weights_dict()
    rr.transfer_weights(src_ref, dst_ref)
    after = dst_ref.weights_dict()

    src = tfimm.create_model(name, device="cpu", **ov)
    src.load_weights_dict(w)
    dst = tfimm.create_model(name, device="cpu", **ov, **change)
    init = dst.weights_dict()
    tfimm.models.transfer_weights(src, dst)
    got = dst.weights_dict()
    for k, v in after.items():
        if any(p in k for p in IGNORE):
            continue
        if np.array_equal(v, before[k]) and not np.array_equal(got[k], init[k]):
            raise AssertionError(f"{k}: the reference left it at its initial value, the engine overwrote it")
        if not np.array_equal(v, before[k]):
            assert np.abs(got[k] - v).max() < 1e-6, k


@pytest.mark.parametrize("arch", ["resnet50", "vit_b_16"])
def test_pytorch_state_dict_conversion_equals_reference(arch):
    """N1: tfimm.utils.timm.convert_state_dict produces exactly what the reference's
    load_pytorch_weights_in_tf2_model (tfimm/utils/timm.py:109-229) writes into its variables."""
    import torchvision

    import tfimm
    from tfimm.utils import timm as etimm

    if arch == "resnet50":
        tv = torchvision.models.resnet50(weights=None)
        name, ov = "resnet50", {"input_size": (32, 32)}
        sd = {k: v for k, v in tv.state_dict().items()}
    else:
        tv = torchvision.models.VisionTransformer(image_size=32, patch_size=8, num_layers=2, num_heads=2, hidden_dim=16,
                                                  mlp_dim=64, num_classes=10)
        name, ov = "vit_tiny_patch16_224", {"input_size": (32, 32), "patch_size": 8, "embed_dim": 16, "nb_blocks": 2,
                                             "nb_heads": 2, "nb_classes": 10}
        # torchvision -> timm key names (the reference converts timm checkpoints)
        sd = {}
        for k, v in tv.state_dict().items():
            k = (k.replace("encoder.layers.encoder_layer_", "blocks.").replace("ln_1", "norm1").replace("ln_2", "norm2")
                 .replace("self_attention.in_proj_", "attn.qkv.").replace("self_attention.out_proj", "attn.proj")
                 .replace("mlp.0", "mlp.fc1").replace("mlp.3", "mlp.fc2").replace("encoder.ln", "norm")
                 .replace("conv_proj", "patch_embed.proj").replace("heads.head", "head")
                 .replace("class_token", "cls_token").replace("encoder.pos_embedding", "pos_embed"))
            sd[k] = v
    g = torch.Generator().manual_seed(0)
    sd = {k: (torch.randn(v.shape, generator=g) if v.is_floating_point() else v) for k, v in sd.items()}
    ref = rr.create_model(name, **ov)
    rr.load_pytorch_weights(ref, {k: v.clone() for k, v in sd.items()})
    want = ref.weights_dict()
    eng = tfimm.create_model(name, device="cpu", **ov)
    got, missing, unexpected = etimm.convert_state_dict(eng, sd)
    assert not missing
    for k, v in got.items():
        assert np.array_equal(np.asarray(v, dtype=np.float32), want[k]), k
    assert set(got) == {k for k in want if not any(p in k for p in IGNORE)}
