"""In-process A/B of an environment switch that the forward pass reads at call time (same GPU, alternating):
    python tools/ab_env.py <model> <ENV_VAR> <value_a> <value_b> [batch]
Each setting is captured into its own CUDA graph and replayed 30 times, three rounds."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tensorflow-image-models_b200"))
import tfimm  # noqa: E402

name, var, va, vb = sys.argv[1:5]
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 256
model = tfimm.create_model(name, device="cuda")
h, w = model.cfg.input_size
x = torch.randn(batch, h, w, 3, device="cuda")
for rep in range(3):
    for val in (va, vb):
        os.environ[var] = val
        run = model.cuda_graph(batch)
        for _ in range(5):
            run(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            run(x)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name} {var}={val}: {e0.elapsed_time(e1) / 30:.3f} ms/step ({run.launches} launches)")
