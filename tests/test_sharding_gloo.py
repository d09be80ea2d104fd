"""World-size-2 (and 3) CPU tests of the multi-GPU host logic over gloo: image sharding + output all-gather
(SURVEY 8e).  The "model" is a per-image function, so the sharded result must equal the single-process one exactly."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _per_image_model(x):                 # any function without cross-image coupling
    return torch.sigmoid(x * 3.0 - 1.0) + x.flip(-1) * 0.25 + x.mean(dim=(1, 2, 3), keepdim=True)


def _worker(rank, world, port, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import xrd_b200
        g = torch.Generator().manual_seed(11)
        x = torch.rand(batch, 1, 16, 24, generator=g)
        lo, hi = xrd_b200.shard_bounds(batch, rank, world)
        assert xrd_b200.shard_batch(x, rank, world).shape[0] == hi - lo
        y = xrd_b200.run_sharded(_per_image_model, x, rank, world)
        torch.save({"y": y, "lo": lo, "hi": hi}, os.path.join(out_dir, f"r{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 8), (2, 5), (3, 4)])
def test_sharded_equals_single_process(tmp_path, world, batch):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, batch, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(batch, 1, 16, 24, generator=g)
    ref = _per_image_model(x)
    covered = []
    for r in range(world):
        d = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert torch.equal(d["y"], ref), f"rank {r}: gathered output differs from the single-process result"
        covered += list(range(d["lo"], d["hi"]))
    assert covered == list(range(batch))       # every image exactly once, in order


def test_shard_bounds_properties():
    import xrd_b200
    for batch in (0, 1, 7, 16, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [xrd_b200.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        xrd_b200.shard_bounds(4, 2, 2)
