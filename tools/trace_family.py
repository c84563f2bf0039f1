"""Per-launch device time of one kernel family in one forward pass (CUDA events around every launch).
    python tools/trace_family.py MODEL FAMILY [batch]      e.g. efficientnet_b4 dwconv_bias_act 256"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tensorflow-image-models_b200"))
import tfimm  # noqa: E402
from tfimm.backend import ops  # noqa: E402


def main():
    name, fam = sys.argv[1], sys.argv[2]
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    model = tfimm.create_model(name, precision="bf16", device="cuda", seed=0)
    x = torch.randn(B, *model.cfg.input_size, model.cfg.in_channels, device="cuda")
    for _ in range(2):
        model(x)
    torch.cuda.synchronize()
    ops.trace = []
    model(x)
    torch.cuda.synchronize()
    trace, ops.trace = ops.trace, None
    total = 0.0
    for i, (n, e0, e1, flops, nbytes) in enumerate(trace):
        if n != fam:
            continue
        ms = e0.elapsed_time(e1)
        total += ms
        print(f"{i:4d} {n}: {ms * 1e3:8.1f} us  {nbytes / 1e6:9.1f} MB  {nbytes / ms * 1e-6:7.0f} GB/s  "
              f"{flops / ms * 1e-9:8.1f} TFLOP/s")
    print(f"total {fam}: {total:.3f} ms")


if __name__ == "__main__":
    main()
