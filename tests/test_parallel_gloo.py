"""Host-side multi-GPU logic on CPU: world_size-2 gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from smilecode_b200.parallel import FlatGradAllReduce, flat_layout, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                    # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    sync = FlatGradAllReduce(net.parameters())
    g = torch.Generator().manual_seed(100)
    x = torch.randn(8, 5, generator=g)                      # the global batch; each rank takes its shard
    b, e = shard_range(8, rank, world)
    sync.zero_()
    net(x[b:e]).pow(2).sum().backward()                     # sum-loss: mean over ranks == global grad / world
    assert all(p.grad.data_ptr() >= sync.bucket.data_ptr() for p in net.parameters())   # grads live in the bucket
    sync.allreduce_mean_()
    assert all(p.grad.data_ptr() % 16 == 0 for p in net.parameters())                   # every view is 16-byte aligned
    flat = torch.cat([p.grad.flatten() for p in net.parameters()])
    # single-process reference on the whole batch
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    ref.load_state_dict(net.state_dict())
    ref(x).pow(2).sum().backward()
    ref_flat = torch.cat([p.grad.flatten() for p in ref.parameters()]) / world
    out[rank] = float((flat - ref_flat).abs().max())
    dist.destroy_process_group()


def test_flat_grad_allreduce_equals_single_process_gradient():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world and max(out.values()) < 1e-6, dict(out)


def test_flat_layout_pads_to_16_bytes():
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (6, 27, 4, 1, 128)]
    offs, total = flat_layout(ps)
    assert offs == [0, 8, 36, 40, 44] and total == 172
    assert all(o % 4 == 0 for o in offs)
