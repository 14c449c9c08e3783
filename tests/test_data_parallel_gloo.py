"""CPU, world_size 2, gloo: E2 -- the one-process-per-GPU trainer shim behind the reference's DataParallel call
shape (trains/base_trainer.py:31-35, 70-76).  The bucketed, hook-driven gradient all-reduce must reproduce what the
reference's single-process DataParallel computes: grad = mean over replicas of d(loss_replica)/dw, each replica
with ITS OWN normaliser, for even and uneven (`chunk_sizes`) batch splits."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


class ModelWithLoss(nn.Module):
    """Shape of trains/base_trainer.py:12-21: loss computed inside the replicated module, per-replica normaliser."""

    def __init__(self):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 4, 3, padding=1))
        self.unused_head = nn.Linear(4, 4)      # gets no gradient: its bucket must still be reduced on every rank

    def forward(self, batch):
        out = self.net(batch["input"])
        num_pos = batch["mask"].sum().clamp(min=1.0)          # per-replica normaliser (like _neg_loss's num_pos)
        loss = ((out - batch["target"]) ** 2 * batch["mask"]).sum() / num_pos
        return out, loss, {"loss": loss}


def _batch(n):
    g = torch.Generator().manual_seed(3)
    return {"input": torch.randn(n, 3, 8, 8, generator=g), "target": torch.randn(n, 4, 8, 8, generator=g),
            "mask": (torch.rand(n, 4, 8, 8, generator=g) > 0.6).float(), "meta": {"ids": list(range(n))}}


def _reference_grads(chunks, n):
    """What the reference's DataParallel + `loss.mean()` yields, computed in one process."""
    torch.manual_seed(0)
    m = ModelWithLoss()
    batch = _batch(n)
    losses, lo = [], 0
    for c in chunks:
        sl = {k: (v if k == "meta" else v[lo:lo + c]) for k, v in batch.items()}
        losses.append(m(sl)[1])
        lo += c
    torch.stack(losses).mean().backward()
    return [None if p.grad is None else p.grad.clone() for p in m.parameters()], [p.detach().clone() for p in m.parameters()]


def _worker(rank, world, port, chunks, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from centernet_b200.data_parallel import DataParallel, ShardedDataParallel
    torch.manual_seed(0 if rank == 0 else 123)       # rank 1 starts with DIFFERENT weights: the wrapper must sync them
    m = ModelWithLoss()
    dp = DataParallel(m, device_ids=[0, 1], chunk_sizes=chunks)
    assert isinstance(dp, ShardedDataParallel)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    assert len(dp.reducer.buckets) >= 1
    grads = None
    for step in range(2):                            # second step goes through optimizer.zero_grad(set_to_none=True)
        opt.zero_grad()
        output, loss, stats = dp(_batch(n))
        loss = loss.mean()
        loss.backward()
        if step == 0:
            grads = [None if p.grad is None else p.grad.clone() for p in m.parameters()]
        opt.step()
    if rank == 0:
        torch.save({"grads": grads, "out_rows": output.shape[0], "params": [p.detach().clone() for p in m.parameters()]}, out)
    else:
        torch.save({"params": [p.detach().clone() for p in m.parameters()]}, out + ".r1")
    dist.destroy_process_group()


@pytest.mark.parametrize("chunks,n", [(None, 8), ([3, 5], 8)])
def test_two_rank_gradient_allreduce_matches_dataparallel(tmp_path, chunks, n):
    out = str(tmp_path / "r0.pt")
    port = 29700 + os.getpid() % 1000 + (0 if chunks is None else 31)
    mp.spawn(_worker, args=(2, port, chunks, n, out), nprocs=2, join=True)
    r0, r1 = torch.load(out), torch.load(out + ".r1")
    want, _ = _reference_grads(chunks or [4, 4], n)
    assert r0["out_rows"] == (chunks[0] if chunks else 4)
    for g, w in zip(r0["grads"], want):
        if w is None:
            assert g is None or float(g.abs().max()) == 0.0
        else:
            assert torch.allclose(g, w, rtol=1e-5, atol=1e-6)
    for a, b in zip(r0["params"], r1["params"]):      # replicas stay bit-identical after the optimizer steps
        assert torch.equal(a, b)


def test_bucketing_and_fallback():
    from centernet_b200.data_parallel import DataParallel, GradientAllReducer
    m = nn.Sequential(nn.Linear(256, 256), nn.Linear(256, 256), nn.Linear(256, 10))
    red = GradientAllReducer(m.parameters(), bucket_mb=0.3)          # 263 KB per big layer -> several buckets
    assert len(red.buckets) >= 2 and red.nbytes == sum(p.numel() * 4 for p in m.parameters())
    # reverse order: the LAST layer's parameters sit in the FIRST bucket
    assert red.buckets[0]["params"][0] is list(m.parameters())[-1]
    m(torch.randn(4, 256)).sum().backward()                           # world size 1: hooks run, nothing to reduce
    assert all(p.grad is not None and p.grad.data_ptr() >= red._of[p]["flat"].data_ptr() for p in m.parameters())
    red.remove()
    assert isinstance(DataParallel(nn.Linear(2, 2), device_ids=[0]), torch.nn.DataParallel)   # no process group
