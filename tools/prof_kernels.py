"""Launches ONE hot kernel a few times at its benchmark shape so that `ncu --set full` can capture it quickly.

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 2 -c 1 -o gpurun_out/<name> \
        python tools/prof_kernels.py <case>
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tensorflow-image-models_b200"))

from tfimm.backend import ops  # noqa: E402


def main(case):
    g = torch.Generator(device="cuda").manual_seed(0)
    B = 256
    if case == "attn":          # ViT-B/16: N=197, H=12
        qkv = torch.randn(B * 197, 2304, device="cuda", generator=g).to(torch.bfloat16)
        fn = lambda: ops.attention(qkv, B, 197, 12, 64, 0.125)
    elif case == "gemm_fc1":    # ViT-B fc1 + GELU
        a = torch.randn(B * 197, 768, device="cuda", generator=g).to(torch.bfloat16)
        w = (torch.randn(3072, 768, device="cuda", generator=g) / 28).to(torch.bfloat16)
        bias = torch.randn(3072, device="cuda", generator=g)
        fn = lambda: ops.gemm(a, w, bias=bias, act="gelu")
    elif case == "gemm_proj":   # ViT-B proj + fp32 residual in place
        a = torch.randn(B * 197, 768, device="cuda", generator=g).to(torch.bfloat16)
        w = (torch.randn(768, 768, device="cuda", generator=g) / 28).to(torch.bfloat16)
        bias = torch.randn(768, device="cuda", generator=g)
        x = torch.randn(B * 197, 768, device="cuda", generator=g)
        fn = lambda: ops.gemm(a, w, bias=bias, residual=x, out=x)
    elif case == "dwconv_ln":   # ConvNeXt-B stage 2
        x = torch.randn(B, 14, 14, 512, device="cuda", generator=g)
        wt = torch.randn(49, 512, device="cuda", generator=g) / 7
        v = torch.randn(512, device="cuda", generator=g)
        fn = lambda: ops.dwconv_ln(x, wt, v, v, v, 1e-6, torch.bfloat16)
    elif case == "dwconv_ln0":  # ConvNeXt-B stage 0
        x = torch.randn(B, 56, 56, 128, device="cuda", generator=g)
        wt = torch.randn(49, 128, device="cuda", generator=g) / 7
        v = torch.randn(128, device="cuda", generator=g)
        fn = lambda: ops.dwconv_ln(x, wt, v, v, v, 1e-6, torch.bfloat16)
    elif case == "window_attn":  # Swin-B stage 2, shifted
        from tfimm.architectures.swin import window_tables

        qkv = torch.randn(B * 196, 1536, device="cuda", generator=g).to(torch.bfloat16)
        bias = torch.randn(16, 49, 49, device="cuda", generator=g)
        rm, lab = window_tables(14, 14, 7, 3)
        lb = torch.from_numpy(lab).view(4, 49)
        diff = (lb[:, :, None] != lb[:, None, :]).to(torch.int64)
        bits = torch.zeros(4, 64, dtype=torch.int64)
        bits[:, :49] = (diff << torch.arange(49, dtype=torch.int64)[None, None, :]).sum(dim=-1)
        rm, bits = torch.from_numpy(rm).cuda(), bits.cuda()
        bias_pad = torch.zeros(16, 64, 64, device="cuda")
        bias_pad[:, :49, :49] = bias
        fn = lambda: ops.window_attention_tc(qkv, bias_pad, rm, bits, B, 4, 49, 16, 32, 32 ** -0.5)
    elif case == "window_attn0":  # Swin-B stage 0, shifted: 64 windows x 4 heads per image
        from tfimm.architectures.swin import window_tables

        qkv = torch.randn(B * 3136, 384, device="cuda", generator=g).to(torch.bfloat16)
        bias = torch.randn(4, 49, 49, device="cuda", generator=g)
        rm, lab = window_tables(56, 56, 7, 3)
        lb = torch.from_numpy(lab).view(64, 49)
        diff = (lb[:, :, None] != lb[:, None, :]).to(torch.int64)
        bits = torch.zeros(64, 64, dtype=torch.int64)
        bits[:, :49] = (diff << torch.arange(49, dtype=torch.int64)[None, None, :]).sum(dim=-1)
        rm, bits = torch.from_numpy(rm).cuda(), bits.cuda()
        bias_pad = torch.zeros(4, 64, 64, device="cuda")
        bias_pad[:, :49, :49] = bias
        fn = lambda: ops.window_attention_tc(qkv, bias_pad, rm, bits, B, 64, 49, 4, 32, 32 ** -0.5)
    elif case == "dwconv_act":   # EfficientNet-B4 stage 1 (380 px): 95x95x192, k3 s1
        x = torch.randn(B, 95, 95, 192, device="cuda", generator=g).to(torch.bfloat16)
        wt = torch.randn(9, 192, device="cuda", generator=g) / 3
        bias = torch.randn(192, device="cuda", generator=g)
        pool = torch.zeros(B, 192, device="cuda")
        fn = lambda: ops.dwconv_bias_act(x, wt, bias, 3, 1, "same", act="swish", pool_sum=pool)
    elif case == "layernorm":
        x = torch.randn(B * 197, 768, device="cuda", generator=g)
        v = torch.randn(768, device="cuda", generator=g)
        fn = lambda: ops.layernorm(x, v, v, 1e-6, torch.bfloat16)
    else:
        raise SystemExit(f"unknown case {case}")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(5):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    print(f"{case}: {ev[0].elapsed_time(ev[1]) / 5 * 1000:.1f} us per launch")


if __name__ == "__main__":
    main(sys.argv[1])
