"""Data-parallel plumbing on CPU: 2 processes over gloo.  The bucketed all-reduce (senas_b200.dp.GradBuckets, hooks
fired during backward) together with the global-batch dice_ce loss must give every rank exactly the gradient of the
un-sharded batch (SUM all-reduce, BatchNorm-free model so that shard statistics do not enter)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(7)
    return nn.Sequential(nn.Conv2d(1, 6, 3, padding=1), nn.ReLU(), nn.Conv2d(6, 6, 3, padding=1), nn.ReLU(),
                         nn.Conv2d(6, 2, 1))


def _worker(rank, world, port, ret):
    from senas_b200.dp import GradBuckets, broadcast_parameters
    from senas_b200.loss import SegmentationLosses
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    x = torch.randn(4 * world, 1, 12, 12)
    y = (torch.rand(4 * world, 12, 12) > 0.7).long()
    model = _model()
    if rank != 0:  # replicas start different; the broadcast must fix that
        for p in model.parameters():
            p.data.add_(1.0)
    broadcast_parameters(model)
    arch = [nn.Parameter(torch.zeros(3))]
    buckets = GradBuckets(list(model.parameters()) + arch, arch, bucket_floats=100)
    assert len(buckets.buckets) >= 3
    crit = SegmentationLosses('dice_ce', group=dist.group.WORLD)
    sl = slice(4 * rank, 4 * (rank + 1))
    for _ in range(2):  # twice: the bucket state must reset between passes
        model.zero_grad()
        arch[0].grad = None
        loss = crit([model(x[sl]) * (1 + arch[0].sum())], y[sl])
        loss.backward()
        buckets.finish()
    ref = _model()
    a0 = nn.Parameter(torch.zeros(3))
    SegmentationLosses('dice_ce')([ref(x) * (1 + a0.sum())], y).backward()
    err = max((p.grad - q.grad).abs().max().item() for p, q in zip(model.parameters(), ref.parameters()))
    err = max(err, (arch[0].grad - a0.grad).abs().max().item())
    ret[rank] = err
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_global_batch_gradient():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == world
        assert max(ret.values()) < 1e-5, dict(ret)
