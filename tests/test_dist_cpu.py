"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: pair sharding and the single flattened
gradient all-reduce used by the training step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fepe_b200 import dist as fd


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 256, 257, 1000):
        for world in (1, 2, 3, 8):
            spans = [fd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
        data = torch.arange(64, dtype=torch.float32).reshape(8, 8) / 64.0      # the global batch of 8 "pairs"
        lo, hi = fd.shard_range(8, rank, world)
        loss = net(data[lo:hi]).pow(2).sum() / 8.0                            # each rank: its share of the mean
        loss.backward()
        for p in net.parameters():
            p.grad *= world                      # allreduce_mean divides by world; the global mean is the SUM of shares
        metric = fd.allreduce_mean_grads_(net.parameters(), extra=torch.tensor([float(rank + 1)]))
        tmax = fd.max_over_ranks(0.5 + rank)
        torch.save({"grads": [p.grad.clone() for p in net.parameters()], "metric": metric, "tmax": tmax},
                   os.path.join(out, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_allreduce_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    data = torch.arange(64, dtype=torch.float32).reshape(8, 8) / 64.0
    (net(data).pow(2).sum() / 8.0).backward()
    ref = [p.grad for p in net.parameters()]
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        for a, b in zip(got["grads"], ref):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
        assert abs(float(got["metric"]) - 1.5) < 1e-6
        assert got["tmax"] == 1.5


def _worker_flat(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
        unused = torch.nn.Linear(4, 4)                  # a sub-network that gets a gradient on rank 0 only
        fg = fd.FlatGradients(list(net.parameters()) + list(unused.parameters()), extra_numel=2)
        data = torch.arange(64, dtype=torch.float32).reshape(8, 8) / 64.0
        lo, hi = fd.shard_range(8, rank, world)
        for _step in range(2):                           # the views must survive zero_() and a second backward
            fg.zero_()
            loss = net(data[lo:hi]).pow(2).sum() / 8.0 * world
            if rank == 0:
                loss = loss + unused(torch.ones(1, 4)).sum() * world
            loss.backward()
            assert fg.check_views()
            fg.extra[0], fg.extra[1] = float(rank + 1), 10.0
            extra = fg.allreduce_mean_()
        torch.save({"grads": [p.grad.clone() for p in net.parameters()],
                    "unused": [p.grad.clone() for p in unused.parameters()], "extra": extra.clone()},
                   os.path.join(out, f"f{rank}.pt"))
        # the list-based variant with a None gradient on one rank: same message layout on both ranks
        for p in unused.parameters():
            p.grad = None
        net.zero_grad(set_to_none=True)
        loss = net(data[lo:hi]).pow(2).sum() / 8.0 * world
        if rank == 0:
            loss = loss + unused(torch.ones(1, 4)).sum() * world
        loss.backward()
        fd.allreduce_mean_grads_(list(net.parameters()) + list(unused.parameters()))
        torch.save({"grads": [p.grad.clone() for p in net.parameters()],
                    "unused": [p.grad.clone() for p in unused.parameters()]}, os.path.join(out, f"l{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_gradients_one_collective_and_missing_grads(tmp_path):
    """FlatGradients: .grad views of one buffer, a single all-reduce, identical layout on every rank even when a
    sub-network receives no gradient on one of them (the list-based reduce could hang or mix parameters)."""
    world = 2
    mp.spawn(_worker_flat, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    unused = torch.nn.Linear(4, 4)
    data = torch.arange(64, dtype=torch.float32).reshape(8, 8) / 64.0
    ((net(data).pow(2).sum() / 8.0) + unused(torch.ones(1, 4)).sum()).backward()
    for tag in ("f", "l"):
        for r in range(world):
            got = torch.load(os.path.join(str(tmp_path), f"{tag}{r}.pt"))
            for a, b in zip(got["grads"], [p.grad for p in net.parameters()]):
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), tag
            for a, b in zip(got["unused"], [p.grad for p in unused.parameters()]):
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), tag
            if tag == "f":
                assert torch.allclose(got["extra"], torch.tensor([1.5, 10.0]))
