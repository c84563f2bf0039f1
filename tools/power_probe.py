"""Steady-state probe of one GEMM shape: runs it back to back for a few seconds while sampling SM clock and board power
(nvidia-smi), so that a power-capped clock shows up next to the time per launch.
    python tools/power_probe.py M N K [act] [out=bf16|f32] [res=0|1] [seconds] [cublas]"""
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tensorflow-image-models_b200"))


def main():
    M, N, K = (int(v) for v in sys.argv[1:4])
    act = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "none" else None
    out_dtype = torch.float32 if len(sys.argv) > 5 and sys.argv[5] == "f32" else torch.bfloat16
    res = len(sys.argv) > 6 and sys.argv[6] == "1"
    seconds = float(sys.argv[7]) if len(sys.argv) > 7 else 3.0
    cublas = len(sys.argv) > 8 and sys.argv[8] == "cublas"
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g).to(out_dtype)
    if cublas:
        wt = w.t()
        fn = lambda: torch.matmul(a, wt)
    else:
        from tfimm.backend import ops
        fn = (lambda: ops.gemm(a, w, bias=bias, act=act, residual=x, out=x)) if res else \
             (lambda: ops.gemm(a, w, bias=bias, act=act, out=x))
    samples = []
    stop = threading.Event()

    def sampler():
        while not stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active",
                                    "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True,
                                   timeout=5).stdout.strip().split(",")
                samples.append((time.time(), float(o[0]), float(o[1]), o[2].strip()))
            except Exception:
                pass
            time.sleep(0.05)

    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    th = threading.Thread(target=sampler, daemon=True)
    th.start()
    t_end = time.time() + seconds
    windows = []
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        torch.cuda.synchronize()
        windows.append(e0.elapsed_time(e1) * 1e3 / 200)
    stop.set()
    th.join()
    clk = sorted(s[1] for s in samples)
    pw = sorted(s[2] for s in samples)
    med = lambda v: v[len(v) // 2] if v else float("nan")
    name = "cublas" if cublas else "tfimm"
    print(f"{name} M={M} N={N} K={K} act={act} out={str(out_dtype)[6:]} res={int(res)}: first window {windows[0]:.1f} us, "
          f"last {windows[-1]:.1f} us ({2.0 * M * N * K / windows[-1] * 1e-6:.0f} TFLOP/s), sm clock median {med(clk):.0f} MHz "
          f"(min {clk[0] if clk else 0:.0f}), power median {med(pw):.0f} W (max {pw[-1] if pw else 0:.0f}), "
          f"reasons {sorted(set(s[3] for s in samples))}")


if __name__ == "__main__":
    main()
