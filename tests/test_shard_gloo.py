"""CPU, world_size 2, gloo: the N>1 host logic (shard ranges, detection gather, timing reduction)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_images, chunks, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from centernet_b200 import shard
    lo, hi = shard.shard_range(n_images, rank, world, chunks)
    # fake per-image detections: value encodes the global image index
    local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1).expand(hi - lo, 3, 6).contiguous()
    full = shard.gather_detections(local, n_images, chunks)
    mx = shard.max_over_ranks(10.0 + rank)
    if rank == 0:
        torch.save({"full": full, "max": mx, "range": (lo, hi)}, out)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images,chunks", [(8, None), (7, None), (8, [3, 5])])
def test_two_rank_shard_gather(tmp_path, n_images, chunks):
    out = str(tmp_path / "r0.pt")
    port = 29500 + os.getpid() % 1000 + (n_images if chunks is None else 17)
    mp.spawn(_worker, args=(2, port, n_images, chunks, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["full"].shape == (n_images, 3, 6)
    assert torch.equal(r["full"][:, 0, 0], torch.arange(n_images, dtype=torch.float32))
    assert r["max"] == 11.0
    assert r["range"][0] == 0


def test_shard_range_covers_batch():
    from centernet_b200.shard import shard_range
    for n in (1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(ValueError):
        shard_range(8, 0, 2, [3, 4])
