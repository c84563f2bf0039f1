"""bf16 parity as a MEASURED BUDGET.

north_star: logits within 1e-3 (bf16) / 1e-5 (fp32) of the reference.  Three separate statements are tested:

A. ``engine(fp32 mode)`` vs oracle <= 1e-5                                   -- tests/test_models_gpu.py
B. ``engine(bf16)`` vs ``oracle/emulate_bf16`` (the SAME graph, the SAME bf16 storage points, exact per-op arithmetic
   in torch)  <= 1e-3: everything the kernels add on top of the inherent rounding of bf16-stored tensors
   (tanh-form GELU, tanh.approx swish, ex2.approx softmax, fp16 staging in the depthwise kernel, accumulation order).
C. ``engine(bf16)`` vs fp32 oracle: the total, dominated by bf16 operand rounding.  Asserted at 1.3x the value
   measured on B200 (DESIGN.md section 5 lists them) so that a regression shows.

and D: the emulation graph with no bf16 storage anywhere (precision="fp32" models) reproduces the oracle -- which is
pinned to the reference's own code (tests/test_reference_pin_cpu.py) -- i.e. B compares against the reference's graph.

Also here: the BASELINE.json configurations at their OWN batch size (256 per GPU) through ``model.cuda_graph`` -- the
CTA-pair GEMM, the persistent multi-wave attention schedule, the pruned last ViT block and graph replay, which the
batch-2 tests never reach.
"""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu

# metric of the reference's own parity test (tests/test_timm.py:71)


def _nerr(out, ref):
    out, ref = out.float().cpu(), ref.float().cpu()
    d = (out - ref).abs().max().item()
    return d / (ref.abs().max().item() + 1e-6), d


def _model(name, family, precision, overrides=None, seed=3):
    import tfimm
    from oracle import params

    omod = importlib.import_module(f"oracle.{family}")
    model = tfimm.create_model(name, precision=precision, device="cuda", **(overrides or {}))
    w = params.random_params(omod.param_shapes(model.cfg), seed=seed)
    model.load_weights_dict(w)
    return model, omod, w


SMALL = [
    ("vit", "vit_tiny_patch16_224", {"nb_blocks": 4}),
    ("vit", "deit_tiny_distilled_patch16_224", {"nb_blocks": 2}),
    ("swin", "swin_tiny_patch4_window7_224", {"input_size": (112, 112), "nb_blocks": (2, 2), "nb_heads": (3, 6)}),
    ("convnext", "convnext_tiny", {"input_size": (64, 96), "nb_blocks": (1, 1, 2, 1)}),
    ("efficientnet", "efficientnet_b0", {"input_size": (96, 96)}),
    ("efficientnet", "efficientnet_v2_b0", {"input_size": (96, 96)}),
    ("resnet", "resnet50", {"input_size": (96, 96)}),
    ("resnet", "seresnext26d_32x4d", {"input_size": (96, 96)}),
    ("resnet", "ecaresnet26t", {"input_size": (96, 96)}),
    ("resnet", "resnetblur50", {"input_size": (96, 96)}),
    ("resnet", "resnet50_gn", {"input_size": (96, 96)}),
]


@pytest.mark.parametrize("family,name,overrides", SMALL, ids=[c[1] for c in SMALL])
def test_emulated_graph_equals_oracle_in_fp32(family, name, overrides):
    """D: engine orchestration + exact torch ops, no bf16 storage == the (reference-pinned) oracle graph."""
    from oracle import emulate_bf16, params

    model, omod, w = _model(name, family, "fp32", overrides)
    x = params.test_images(2, *model.cfg.input_size, model.cfg.in_channels)
    with emulate_bf16.emulated_ops():
        y = model(x.cuda()).float().cpu()
    with torch.no_grad():
        ref = omod.forward(model.cfg, w, x)
    rel, ab = _nerr(y, ref)
    print(f"{name} emulated fp32 graph vs oracle: normalised {rel:.3e}")
    assert rel < 1e-5


# (family, model, overrides, batch, bound on B, bound on C = 1.3 x measured on B200)
BUDGET = [
    ("vit", "vit_tiny_patch16_224", {}, 2, 1e-3, 1.0e-2),
    ("vit", "vit_base_patch16_224", {}, 2, 1e-3, 8e-3),
    ("swin", "swin_tiny_patch4_window7_224", {}, 2, 1e-3, 1.0e-2),
    ("swin", "swin_base_patch4_window7_224", {}, 2, 1e-3, 1.0e-2),
    ("convnext", "convnext_tiny", {}, 2, 1e-3, 8e-3),
    ("convnext", "convnext_base", {}, 2, 1e-3, 8e-3),
    ("efficientnet", "efficientnet_b0", {}, 2, 1e-3, 3e-2),
    ("efficientnet", "efficientnet_b4", {}, 2, 1e-3, 3e-2),
    ("resnet", "resnet50", {}, 2, 1e-3, 3e-2),
    ("resnet", "seresnet50", {}, 2, 1e-3, 3e-2),
]


@pytest.mark.parametrize("family,name,overrides,batch,tol_kernels,tol_total", BUDGET, ids=[c[1] for c in BUDGET])
def test_bf16_error_budget(family, name, overrides, batch, tol_kernels, tol_total):
    """B and C on the same inputs; prints the split so DESIGN.md can quote it."""
    from oracle import emulate_bf16, params

    model, omod, w = _model(name, family, "bf16", overrides)
    x = params.test_images(batch, *model.cfg.input_size, model.cfg.in_channels)
    y = model(x.cuda()).float().cpu()
    with emulate_bf16.emulated_ops():
        y_ideal = model(x.cuda()).float().cpu()
    with torch.no_grad():
        ref = omod.forward(model.cfg, w, x)
    kern, _ = _nerr(y, y_ideal)
    inherent, _ = _nerr(y_ideal, ref)
    total, _ = _nerr(y, ref)
    print(f"BUDGET {name}: engine-vs-emulated {kern:.3e} | emulated-vs-fp32-oracle {inherent:.3e} | "
          f"engine-vs-fp32-oracle {total:.3e}")
    assert kern < tol_kernels, f"kernel-internal error {kern:.3e}"
    assert total < tol_total, f"total error {total:.3e}"


FULL = [
    ("vit", "vit_base_patch16_224", 256, 16, 8e-3),
    ("convnext", "convnext_base", 256, 8, 8e-3),
    ("swin", "swin_base_patch4_window7_224", 256, 8, 1.0e-2),
    ("efficientnet", "efficientnet_b4", 256, 4, 3e-2),      # native 380 px; per-GPU share of the 2048 batch
]


@pytest.mark.parametrize("family,name,batch,nref,tol", FULL, ids=[c[1] for c in FULL])
def test_baseline_config_at_its_own_batch_through_cuda_graph(family, name, batch, nref, tol):
    """BASELINE.json configs[1..4] exactly as bench.py runs them: per-GPU batch 256, captured CUDA graph, bf16.
    The oracle (CPU) is evaluated on images spread over the batch; the kernel-internal part is bounded against the
    emulation on the same slice."""
    from oracle import emulate_bf16, params

    model, omod, w = _model(name, family, "bf16", seed=29)
    h, wd = model.cfg.input_size
    x = params.test_images(batch, h, wd, model.cfg.in_channels, seed=77)
    fwd = model.cuda_graph(batch)
    y = fwd(x.cuda()).float().cpu().clone()
    y_again = fwd(x.cuda()).float().cpu()
    if family == "efficientnet":   # the fused squeeze (global pool) uses fp32 atomics: sums differ in the last bits
        assert _nerr(y_again, y)[0] < 1e-3
    else:
        assert torch.equal(y, y_again)                               # replay is deterministic
    idx = torch.linspace(0, batch - 1, nref).round().long()          # first, last and in between: every CTA wave
    with torch.no_grad():
        ref = omod.forward(model.cfg, w, x[idx])
    total, _ = _nerr(y[idx], ref)
    with emulate_bf16.emulated_ops():
        y_ideal = model(x[idx].cuda()).float().cpu()
    kern, _ = _nerr(y[idx], y_ideal)
    print(f"FULL {name} batch {batch}: engine-vs-fp32-oracle {total:.3e} | engine-vs-emulated {kern:.3e} "
          f"({fwd.launches} launches per replay)")
    assert total < tol
    assert kern < 1e-3
    # batch invariance: image i of the 256-batch equals the same image run in a small eager batch
    small = model(x[idx].cuda()).float().cpu()
    inv, _ = _nerr(y[idx], small)
    print(f"FULL {name}: batch-{batch} graph vs batch-{nref} eager {inv:.3e}")
    assert inv < 3e-3


@pytest.mark.parametrize("fixture", ["full_vit_base_patch16_224.npz", "full_convnext_base.npz",
                                     "full_swin_base_patch4_window7_224.npz", "full_efficientnet_b4.npz",
                                     "full_resnet50.npz"])
def test_fp32_mode_matches_reference_generated_logits_at_full_size(fixture):
    """precision="fp32" against logits the REFERENCE'S OWN CODE produced (tools/make_golden.py): 1e-5."""
    from pathlib import Path

    import numpy as np

    import tfimm
    from oracle import params

    data = np.load(Path(__file__).resolve().parent / "golden" / fixture, allow_pickle=True)
    meta = data["meta"].item()
    mod = importlib.import_module(f"oracle.{meta['family']}")
    model = tfimm.create_model(meta["model"], precision="fp32", device="cuda")
    model.load_weights_dict(params.random_params(mod.param_shapes(model.cfg), seed=meta["seed"]))
    x = params.test_images(meta["batch"], *model.cfg.input_size, model.cfg.in_channels, seed=meta["images_seed"])
    rel, ab = _nerr(model(x.cuda()), torch.from_numpy(data["logits"]))
    print(f"{fixture} fp32 vs reference-generated logits: normalised {rel:.3e} abs {ab:.3e}")
    assert rel < 1e-5
