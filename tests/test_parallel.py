"""Host logic of the multi-GPU path (pathpyg_b200/parallel.py) with gloo at world size 2 on CPU.

The local compute is injected: an object with the functions of ``pathpyg_b200.ops`` that the distributed lift calls,
written in plain torch + the oracle's lift (tests may use the oracle; the product path cannot, and on the GPU box
the default ``local_ops`` is the CUDA library).  What is tested here is the partitioning, the ghost-zone exchange,
the pruning of the line graphs to what later orders need, the record routing and the global ids: the gathered
result must be IDENTICAL to the single-process oracle model."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mom
from pathpyg_b200 import parallel
from torch_local_ops import TorchOps


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def _stream(seed, n, m, horizon):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, m), generator=g)
    t = torch.sort(torch.randint(0, horizon, (m,), generator=g)).values
    w = torch.randint(1, 4, (m,), generator=g).float()
    return ei, t, w


def _check_temporal(rank, world, seed, n, m, horizon, delta, K, weighted, split):
    ei, t, w = _stream(seed, n, m, horizon)
    want = mom.from_temporal_graph(ei, t, n, delta=delta, max_order=K, edge_weight=w if weighted else None)
    if split == "even":
        lo, hi = parallel.partition_stream(m, rank, world)
    else:  # a very uneven split, including an empty range on the last rank
        cuts = [0, m // 5, m] if world == 2 else [0] + [m] * world
        lo, hi = cuts[rank], cuts[rank + 1]
    got = parallel.distributed_temporal_layers(ei[:, lo:hi].contiguous(), t[lo:hi].contiguous(), n, delta, K,
                                               edge_weight=w[lo:hi].contiguous() if weighted else None, local_ops=TorchOps)
    assert sorted(got) == sorted(want)
    for k, layer in want.items():
        full = got[k].gather()
        assert full.num_nodes == layer.num_nodes, (k, full.num_nodes, layer.num_nodes)
        assert torch.equal(full.node_sequence, layer.node_sequence), k
        assert torch.equal(full.edge_index, layer.edge_index), k
        assert torch.equal(full.edge_weight, layer.edge_weight), k
        # ownership: rows of the owned edges lie in the owned id range
        own = got[k]
        if own.edge_index.size(1):
            assert int(own.edge_index[0].min()) >= own.row_offset
            assert int(own.edge_index[0].max()) < own.row_offset + own.node_sequence.size(0)


@pytest.mark.parametrize("seed,n,m,horizon,delta,K,weighted,split", [
    (0, 20, 300, 60, 3, 2, False, "even"),
    (1, 15, 400, 50, 2, 3, True, "even"),
    (2, 12, 300, 40, 2, 4, True, "uneven"),
    (3, 40, 500, 30, 4, 2, False, "uneven"),
    (4, 15, 300, 40, 2, 5, True, "even"),
])
def test_distributed_temporal_layers_world2(seed, n, m, horizon, delta, K, weighted, split):
    _spawn(_check_temporal, 2, seed, n, m, horizon, delta, K, weighted, split)


@pytest.mark.parametrize("seed,n,m,horizon,delta,K,weighted", [(7, 9, 260, 35, 2, 5, True), (8, 30, 200, 25, 3, 3, False)])
def test_distributed_temporal_layers_world3(seed, n, m, horizon, delta, K, weighted):
    """Three ranks (the middle one's ghost zone spans into the third), orders up to 5."""
    _spawn(_check_temporal, 3, seed, n, m, horizon, delta, K, weighted, "even")


def _check_ghost(rank, world):
    ei, t, w = _stream(5, 10, 200, 40)
    lo, hi = parallel.partition_stream(200, rank, world)
    ext_ei, ext_t, ext_w = parallel.exchange_ghost_zone(ei[:, lo:hi].contiguous(), t[lo:hi].contiguous(), w[lo:hi].contiguous(), 6)
    limit = int(t[hi - 1]) + 6
    stop = int(torch.searchsorted(t, torch.tensor(limit), right=True))
    assert torch.equal(ext_ei, ei[:, lo:stop]) and torch.equal(ext_t, t[lo:stop]) and torch.equal(ext_w, w[lo:stop])


def test_ghost_zone_world3():
    _spawn(_check_ghost, 3)


def _check_allreduce(rank, world):
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    parallel.broadcast_parameters(net)
    x = torch.full((5, 4), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    parallel.allreduce_gradients(net, average=True)
    gathered = [[torch.empty_like(g) for _ in range(world)] for g in local]
    for g, buf in zip(local, gathered):
        dist.all_gather(buf, g)
    for p, buf in zip(net.parameters(), gathered):
        assert torch.allclose(p.grad, torch.stack(buf).mean(0), rtol=1e-6, atol=1e-7)


def test_allreduce_gradients_world2():
    _spawn(_check_allreduce, 2)


def test_shard_walks():
    lengths = torch.tensor([3, 9, 2, 2, 8, 4, 4])
    parts = [parallel.shard_walks(lengths, r, 3) for r in range(3)]
    assert parts[0][0] == 0 and parts[-1][1] == lengths.numel()
    assert all(parts[i][1] == parts[i + 1][0] for i in range(2))
    assert all(lo <= hi for lo, hi in parts)
