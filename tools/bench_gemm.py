"""GEMM micro-benchmark: python tools/bench_gemm.py M N K [act] [out=bf16|f32] [res=0|1] [block_n]
Prints the device time per launch (CUDA events, 20 launches after warm-up) and the achieved TFLOP/s."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tensorflow-image-models_b200"))

from tfimm.backend import ops  # noqa: E402


def main():
    M, N, K = (int(v) for v in sys.argv[1:4])
    act = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "none" else None
    out_dtype = torch.float32 if len(sys.argv) > 5 and sys.argv[5] == "f32" else torch.bfloat16
    res = len(sys.argv) > 6 and sys.argv[6] == "1"
    block_n = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g).to(out_dtype)
    fn = (lambda: ops.gemm(a, w, bias=bias, act=act, residual=x, out=x, block_n=block_n)) if res else \
         (lambda: ops.gemm(a, w, bias=bias, act=act, out=x, block_n=block_n))
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"gemm M={M} N={N} K={K} act={act} out={str(out_dtype)[6:]} res={int(res)} block_n={block_n}: "
          f"{us:.1f} us  {2.0 * M * N * K / us * 1e-6:.0f} TFLOP/s")


if __name__ == "__main__":
    main()
