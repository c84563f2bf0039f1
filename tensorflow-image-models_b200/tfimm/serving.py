"""Host-to-logits inference pipeline: pinned host batches in, logits out, with the H2D copy of batch i+1
overlapping the forward pass of batch i.

``depth`` CUDA graphs of the same model (one static input / output buffer each) are replayed round-robin on
the compute stream; uploads go through a separate copy stream; logits come back into pinned host memory.
This is the end-to-end path ``bench.py`` times as ``e2e`` (every step pays its own H2D and D2H).
"""
from typing import List, Optional

import torch


class InferencePipeline:
    def __init__(self, model, batch_size: int, depth: int = 2, input_dtype=torch.float32, input_size=None,
                 gather=None):
        """gather: optional ``f(logits) -> tensor`` run after each forward (e.g. the NCCL all-gather)."""
        self.model = model
        self.depth = depth
        self.gather = gather
        self.compute = torch.cuda.current_stream(model.device)
        self.copy = torch.cuda.Stream(device=model.device)
        self.slots = [model.cuda_graph(batch_size, input_size=input_size, dtype=input_dtype) for _ in range(depth)]
        self.h2d_done = [torch.cuda.Event() for _ in range(depth)]
        self.slot_free = [torch.cuda.Event() for _ in range(depth)]
        self.out_host: List[Optional[torch.Tensor]] = [None] * depth
        self.step = 0
        for ev in self.slot_free:
            ev.record(self.compute)

    def submit(self, host_batch: torch.Tensor) -> torch.Tensor:
        """Enqueues one batch (pinned host tensor) and returns the pinned host tensor its logits will land in
        (valid after ``synchronize()`` or after ``depth`` further submits)."""
        from .backend import ops

        i = self.step % self.depth
        slot = self.slots[i]
        self.copy.wait_event(self.slot_free[i])          # the previous forward on this slot has consumed its input
        with torch.cuda.stream(self.copy):
            slot.static_input.copy_(host_batch, non_blocking=True)
            self.h2d_done[i].record(self.copy)
        self.compute.wait_event(self.h2d_done[i])
        slot.graph.replay()
        ops.launch_count += slot.launches
        out = slot.static_output if self.gather is None else self.gather(slot.static_output)
        if self.out_host[i] is None:
            self.out_host[i] = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        self.out_host[i].copy_(out, non_blocking=True)   # D2H of this step's result
        self.slot_free[i].record(self.compute)
        self.step += 1
        return self.out_host[i]

    def synchronize(self):
        self.compute.synchronize()
        self.copy.synchronize()
