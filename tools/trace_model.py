"""Per-launch device times of one eager forward (CUDA events around every kernel launch), grouped by
(kernel family, algorithmic flops, algorithmic bytes): python tools/trace_model.py <model> [batch]"""
import sys
from collections import OrderedDict
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tensorflow-image-models_b200"))
import tfimm  # noqa: E402
from tfimm.backend import ops  # noqa: E402

name = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
model = tfimm.create_model(name, device="cuda")
h, w = model.cfg.input_size
x = torch.randn(batch, h, w, 3, device="cuda")
for _ in range(3):
    model(x)
torch.cuda.synchronize()
ops.trace = []
model(x)
torch.cuda.synchronize()
agg = OrderedDict()
for fam, e0, e1, flops, nbytes in ops.trace:
    k = (fam, round(flops / 1e9, 1), round(nbytes / 1e6, 1))
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1) * 1e3
ops.trace = None
total = sum(v[1] for v in agg.values())
print(f"# {name} batch {batch}: {total / 1e3:.2f} ms summed over {sum(v[0] for v in agg.values())} launches")
print("# family | GFLOP | MB | launches | avg us | TFLOP/s | GB/s | share")
for (fam, gf, mb), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    avg = us / n
    print(f"{fam:24s} {gf:9.1f} {mb:9.1f} {n:4d} {avg:9.1f} {gf / avg * 1e3 / 1e3:8.0f} {mb / avg * 1e3:8.0f} {100 * us / total:6.1f}%")
