"""world_size-2 gloo test of the data-parallel host logic (flat buffers, bucketed overlapped all-reduce)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fissure_segmentation_b200.ddp import FlatDataParallel
    torch.manual_seed(100 + rank)          # deliberately different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                              torch.nn.Linear(16, 3))
    dp = FlatDataParallel(net, n_buckets=3)
    assert 2 <= len(dp.buckets) <= 3 and dp.buckets[0][0] == 0 and dp.buckets[-1][1] == dp.flat_param.numel()
    gen = torch.Generator().manual_seed(7)
    full_x = torch.randn(8, 6, generator=gen)
    full_y = torch.randn(8, 3, generator=gen)
    shard = slice(rank * 4, rank * 4 + 4)   # batch-sharded clouds
    for step in range(2):
        dp.zero_grad()
        loss = ((dp(full_x[shard]) - full_y[shard]) ** 2).sum()
        loss.backward()
        dp.finish_backward()
    # reference: single-process gradient over the whole batch with rank 0's weights
    params0 = [p.detach().clone() for p in net.parameters()]
    out[rank] = (dp.flat_grad.clone(), [p.clone() for p in params0], full_x, full_y)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_data_parallel_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    g0, p0, x, y = out[0]
    g1, p1, _, _ = out[1]
    assert torch.equal(g0, g1)                                   # all ranks hold the same reduced gradient
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)                                 # broadcast from rank 0
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                              torch.nn.Linear(16, 3))
    with torch.no_grad():
        for p, v in zip(net.parameters(), p0):
            p.copy_(v)
    ((net(x) - y) ** 2).sum().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in reversed(list(net.parameters()))])
    assert torch.allclose(g0, ref, rtol=1e-5, atol=1e-6)         # SUM over ranks == full-batch gradient
