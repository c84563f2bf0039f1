"""A/B of the ViT last-block pruning inside one process (same GPU, alternating), graph replay timing."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tensorflow-image-models_b200"))
import tfimm  # noqa: E402
from tfimm.backend import ops  # noqa: E402

model = tfimm.create_model("vit_base_patch16_224", device="cuda")
x = torch.randn(256, 224, 224, 3, device="cuda")
for rep in range(3):
    for prune in ("1", "0"):
        os.environ["TFIMM_B200_VIT_PRUNE"] = prune
        before = ops.launch_count
        run = model.cuda_graph(256)
        for _ in range(5):
            run(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            run(x)
        e1.record()
        torch.cuda.synchronize()
        print(f"prune={prune}: {e0.elapsed_time(e1) / 30:.3f} ms/step, {run.launches} launches per forward")
