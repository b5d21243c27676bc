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
