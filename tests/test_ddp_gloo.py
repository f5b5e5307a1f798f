"""world_size-2 gloo test of the data-parallel step's only collective: the flat gradient buffer
all-reduce (demf_b200/engine.py FlatGradients), plus rank-sharded synthetic data."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from demf_b200 import engine


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    net[0].bias.requires_grad = False  # frozen parameters never enter the buffer
    flat = engine.FlatGradients(net.parameters())
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(100 + rank))
    flat.zero()
    net(x).pow(2).sum().backward()
    local = flat.buffer.clone()
    work = flat.all_reduce_mean()
    assert work is not None
    work.wait()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = torch.stack(gathered).mean(0)
    ok = torch.allclose(flat.buffer, expect, atol=1e-6)
    views_ok = all(p.grad.data_ptr() >= flat.buffer.data_ptr() for p in flat.params)
    # every gradient starts on a 16-byte boundary of the buffer (padding elements stay zero)
    n_expected = sum((p.numel() + 3) // 4 * 4 for p in net.parameters() if p.requires_grad)
    off = flat.offsets[[id(p) for p in flat.params].index(id(net[2].weight))]
    padding_zero = all(float(flat.buffer[o + p.numel():o + (p.numel() + 3) // 4 * 4].abs().sum()) == 0.0
                       for p, o in zip(flat.params, flat.offsets))
    out[rank] = (ok, views_ok, flat.buffer.numel() == n_expected and off % 4 == 0 and padding_zero,
                 torch.equal(net[2].weight.grad.flatten(), flat.buffer[off:off + 15]))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        assert all(out[rank]), (rank, out[rank])


def test_clip_norm():
    net = torch.nn.Linear(4, 4)
    flat = engine.FlatGradients(net.parameters())
    flat.buffer.fill_(3.0)
    norm = flat.clip_norm_(1.0)
    assert abs(norm.item() - 3.0 * (20 ** 0.5)) < 1e-4
    assert abs(torch.linalg.vector_norm(flat.buffer).item() - 1.0) < 1e-4
    assert torch.equal(net.weight.grad.flatten(), flat.buffer[:16])
