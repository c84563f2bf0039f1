"""Host-side logic of the multi-GPU path on CPU: world_size 2 over gloo (the -m gpu bench uses NCCL)."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tensorflow-image-models_b200"))


def _worker(rank, world, port, batch, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tfimm import parallel

    x = torch.arange(batch * 3, dtype=torch.float32).reshape(batch, 3)

    class Fake:  # stands in for a model: logits = 2 * x (rows stay identifiable)
        def __call__(self, t):
            return t * 2.0

    out = parallel.data_parallel_forward(Fake(), x)
    ok = torch.equal(out, x * 2.0)
    lo, hi = parallel.shard_bounds(batch, rank, world)
    ok = ok and torch.equal(parallel.shard_batch(x), x[lo:hi])
    results[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_shard_and_gather_world2(batch):
    world = 2
    port = 29500 + (os.getpid() % 2000) + batch
    with mp.Manager() as manager:
        results = manager.dict()
        mp.spawn(_worker, args=(world, port, batch, results), nprocs=world, join=True)
        assert dict(results) == {0: True, 1: True}


def test_shard_bounds_cover_batch():
    from tfimm.parallel import shard_bounds

    for batch in (1, 7, 8, 256, 2048):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
