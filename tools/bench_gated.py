"""Gated GEMM (squeeze-excite gate applied in shared memory) against scale_channels_ + GEMM.
    python tools/bench_gated.py B HW K N"""
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
    B, HW, K, N = (int(v) for v in sys.argv[1:5])
    M = B * HW
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    gate = torch.sigmoid(torch.randn(B, K, device="cuda", generator=g))
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    a3 = a.view(B, HW, K)
    t_gated = timed(lambda: ops.gemm_gated(a, gate, HW, w, bias=bias, residual=res))
    t_scale = timed(lambda: ops.scale_channels_(a3, gate))
    t_gemm = timed(lambda: ops.gemm(a, w, bias=bias, residual=res))
    alg = M * K * 2 + 2 * M * N * 2 + M * N * 2
    print(f"gated B={B} HW={HW} K={K} N={N}: gated {t_gated:.1f} us ({alg / t_gated * 1e-6:.2f} TB/s), "
          f"scale {t_scale:.1f} + gemm {t_gemm:.1f} = {t_scale + t_gemm:.1f} us")


if __name__ == "__main__":
    main()
