"""The multi-GPU path on real devices (NCCL): needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a 1-GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, fn, args):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world, *args):
    mp.spawn(_entry, args=(world, _free_port(), fn, args), nprocs=world, join=True)


def _need(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} CUDA devices")


def _check_temporal(rank, world, n, m, horizon, delta, K):
    import pathpyg_b200 as pp
    from pathpyg_b200 import parallel

    dev = torch.device("cuda", rank)
    g = torch.Generator().manual_seed(3)
    ei = torch.randint(0, n, (2, m), generator=g)
    t = torch.sort(torch.randint(0, horizon, (m,), generator=g)).values
    w = torch.randint(1, 4, (m,), generator=g).float()
    tg = pp.TemporalGraph.from_tensors(ei.to(dev), t.to(dev), n, edge_weight=w.to(dev))
    want = pp.MultiOrderModel.from_temporal_graph(tg, delta=delta, max_order=K)      # single-device result
    lo, hi = parallel.partition_stream(m, rank, world)
    got = parallel.distributed_temporal_layers(ei[:, lo:hi].to(dev), t[lo:hi].to(dev), n, delta, K, edge_weight=w[lo:hi].to(dev))
    for k, layer in want.layers.items():
        full = got[k].gather()
        assert full.num_nodes == layer.n, k
        assert torch.equal(full.node_sequence, layer.data.node_sequence), k
        assert torch.equal(full.edge_index, layer.data.edge_index.as_tensor()), k
        assert torch.equal(full.edge_weight, layer.data.edge_weight), k


@pytest.mark.parametrize("n,m,horizon,delta,K", [(50, 4000, 400, 3, 3), (2000, 200_000, 2000, 40, 2)])
def test_distributed_temporal_layers_nccl(n, m, horizon, delta, K):
    _need(2)
    _spawn(_check_temporal, 2, n, m, horizon, delta, K)


def _check_dp_training(rank, world):
    import pathpyg_b200 as pp
    from pathpyg_b200 import parallel

    dev = torch.device("cuda", rank)
    g = torch.Generator().manual_seed(11)
    n_nodes, walks = 200, 4000
    lengths = torch.randint(3, 9, (walks,), generator=g)
    flat = torch.randint(0, n_nodes, (int(lengths.sum()),), generator=g)
    flat[:n_nodes] = torch.arange(n_nodes)  # every node occurs in every shard's first walks? -> ensure below per shard
    lo, hi = parallel.shard_walks(lengths, rank, world)
    starts = torch.cumsum(lengths, 0) - lengths
    a, b = int(starts[lo]), int(starts[hi - 1] + lengths[hi - 1])
    my_flat, my_len = flat[a:b].clone(), lengths[lo:hi].clone()
    # a shard must cover all first-order nodes (lift_order.py:133-143): append one covering walk
    my_flat = torch.cat([my_flat, torch.arange(n_nodes)])
    my_len = torch.cat([my_len, torch.tensor([n_nodes])])
    p = pp.PathData()
    p.append_index_walks(my_flat, my_len, torch.ones(my_len.numel()))
    model = pp.MultiOrderModel.from_path_data(p.to(dev), max_order=2)
    H = 32
    model.layers[1].data.x = torch.randn(n_nodes, H, generator=torch.Generator().manual_seed(1)).to(dev)
    x_h = torch.randn(model.layers[2].n, H, generator=torch.Generator().manual_seed(2 + rank)).to(dev)
    data = model.to_dbgnn_data(max_order=2, x_h=x_h)
    y = (torch.arange(n_nodes) % 16).to(dev)
    torch.manual_seed(100 + rank)  # different initial weights per rank: the broadcast must fix that
    net = pp.nn.DBGNN(num_classes=16, num_features=(H, H), hidden_dims=[H, H, H]).to(dev)
    parallel.broadcast_parameters(net)
    opt = torch.optim.Adam(net.parameters(), lr=0.01)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(net(data), y)
        loss.backward()
        parallel.allreduce_gradients(net)
        opt.step()
        losses.append(float(loss.detach()))
    flat_w = torch.cat([q.detach().reshape(-1) for q in net.parameters()])
    gathered = [torch.empty_like(flat_w) for _ in range(world)]
    dist.all_gather(gathered, flat_w)
    assert all(torch.equal(gathered[0], o) for o in gathered), "replicas diverged"
    assert losses[-1] < losses[0]


def test_data_parallel_training_nccl():
    _need(2)
    _spawn(_check_dp_training, 2)
