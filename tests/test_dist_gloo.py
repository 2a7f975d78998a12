"""World-size-2 test of the multi-GPU host logic on CPU (gloo): image sharding and the single flat-bucket gradient
all-reduce that replaces DataParallel."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out, bucket_views=False):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, '3d-sdn_b200'))
    from sdn_b200.dist import FlatGradSync, shard_range
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    data = torch.randn(10, 6, generator=torch.Generator().manual_seed(1))
    lo, hi = shard_range(10, rank, world)
    sync = FlatGradSync(model, bucket_views=bucket_views)
    if bucket_views:
        for _ in range(2):          # a second round checks that zero() really clears the shared bucket
            sync.zero()
            model(data[lo:hi]).pow(2).mean().backward()
            sync()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(sync.params, sync.views))
        flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        assert torch.equal(flat, sync.flat)
        out[rank] = (lo, hi, flat.clone(), sync.numel)
        dist.destroy_process_group()
        return
    # per-rank loss is the MEAN over its shard; equal shard sizes -> mean of means == global mean
    model(data[lo:hi]).pow(2).mean().backward()
    sync()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    out[rank] = (lo, hi, flat.clone(), sync.numel)
    dist.destroy_process_group()


@pytest.mark.parametrize('bucket_views', [False, True], ids=['copy', 'bucket_views'])
def test_shard_and_flat_allreduce_two_ranks(bucket_views):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, bucket_views), nprocs=world, join=True)
    assert (out[0][0], out[0][1], out[1][0], out[1][1]) == (0, 5, 5, 10)
    assert torch.equal(out[0][2], out[1][2])                         # identical gradients on every rank
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
    data = torch.randn(10, 6, generator=torch.Generator().manual_seed(1))
    model(data).pow(2).mean().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(out[0][2], ref, atol=1e-6)                 # == the single-process gradient
    assert out[0][3] == ref.numel()


def test_shard_range_covers_everything():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, '3d-sdn_b200'))
    from sdn_b200.dist import shard_range
    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_flat_bucket_views_follow_channels_last_parameters():
    """Gradient views of the flat bucket take the strides of channels-last parameters (autograd's layout contract), so
    backward accumulates in place: after backward every p.grad still IS its view, and the bucket holds the gradients."""
    import torch
    from sdn_b200.dist import FlatGradSync
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 4, 3, padding=1))
    net = net.to(memory_format=torch.channels_last)
    ref = [p.detach().clone() for p in net.parameters()]
    x = torch.randn(2, 3, 8, 8)
    net(x).square().sum().backward()
    want = [p.grad.detach().clone() for p in net.parameters()]
    for p in net.parameters():
        p.grad = None
    sync = FlatGradSync(net, bucket_views=True)
    for p, v in zip(sync.params, sync.views):
        assert v.stride() == p.stride() and v.shape == p.shape
    sync.zero()
    net(x).square().sum().backward()
    assert sync._attached()
    sync()                                  # world size 1: no collective, gradients stay where they are
    for p, w, r in zip(net.parameters(), want, ref):
        assert torch.allclose(p.grad, w, rtol=1e-5, atol=1e-6) and torch.equal(p.detach(), r)
    assert abs(float(sync.flat.abs().sum()) - float(sum(w.abs().sum() for w in want))) < 1e-2
