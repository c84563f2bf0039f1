"""fc1 -> act -> fc2 (+ gamma, + fp32 residual): the fused kernel against the two-GEMM form.
    python tools/bench_mlp.py M C HIDDEN"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tensorflow-image-models_b200"))
from tfimm.backend import ops  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    M, C, Hd = (int(v) for v in sys.argv[1:4])
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(M, C, device="cuda", generator=g).to(torch.bfloat16)
    w1 = (torch.randn(Hd, C, device="cuda", generator=g) / C ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, Hd, device="cuda", generator=g) / Hd ** 0.5).to(torch.bfloat16)
    b1 = torch.randn(Hd, device="cuda", generator=g)
    b2 = torch.randn(C, device="cuda", generator=g)
    gamma = torch.randn(C, device="cuda", generator=g)
    x = torch.randn(M, C, device="cuda", generator=g)
    hid = torch.empty(M, Hd, device="cuda", dtype=torch.bfloat16)

    def two():
        ops.gemm(a, w1, bias=b1, act="gelu", out=hid)
        ops.gemm(hid, w2, bias=b2, gamma=gamma, residual=x, out=x)

    def fused():
        ops.mlp_fused(a, w1, b1, w2, b2, "gelu", gamma=gamma, residual=x, out=x)

    t2, t1 = timed(two), timed(fused)
    alg = M * C * (2 + 4 + 4) + 2 * (w1.numel() + w2.numel())
    print(f"mlp M={M} C={C} hidden={Hd}: two GEMMs {t2:.1f} us, fused {t1:.1f} us "
          f"({alg / t1 * 1e-6:.2f} TB/s of algorithmic bytes, {4.0 * M * C * Hd / t1 * 1e-6:.0f} TFLOP/s)")


if __name__ == "__main__":
    main()
