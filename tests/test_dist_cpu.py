"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: contiguous batch sharding, the flat
gradient bucket all-reduce (== big-batch gradient), output gathering.  The transform itself needs no
collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 3, 3, padding=1))


def _worker(rank, world, port, n):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from rpst.dist import GradBucket, gather_outputs, shard_batch, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(n, 3, 8, 8, generator=g)
        y = torch.randn(n, 3, 8, 8, generator=g)
        lo, hi = shard_range(n, rank, world)
        xs, ys = shard_batch([x, y], rank, world)
        assert xs.shape[0] == hi - lo
        # reference: single-process big-batch gradient of a per-sample-mean loss
        ref = _model()
        ((ref(x) - y) ** 2).mean(dim=(1, 2, 3)).sum().div(n).backward()
        m = _model()
        # each rank weights its shard by its share of the batch so that the mean over ranks is exact
        local = ((m(xs) - ys) ** 2).mean(dim=(1, 2, 3)).sum().div(n) * world
        local.backward()
        bucket = GradBucket(m.parameters())
        extras = bucket.allreduce_mean({"loss": local.detach() / world})
        for p, q in zip(m.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, q.grad, atol=1e-6), (p.grad - q.grad).abs().max()
        full_loss = ((ref(x) - y) ** 2).mean(dim=(1, 2, 3)).sum().div(n)
        assert abs(float(extras["loss"]) * world - float(full_loss)) < 1e-5
        out = gather_outputs(m(xs).detach(), n)
        assert torch.allclose(out, ref(x).detach(), atol=1e-6)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 5])
def test_sharded_gradients_equal_big_batch(n):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)


def test_shard_range_partitions_exactly():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from rpst.dist import shard_range
    for n in (0, 1, 7, 16, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
