"""a7 / a9 bookkeeping kernels (csrc/walks.cu) against the torch expressions of the reference
(core/multi_order_model.py:217-224,335,354-361,402-405; core/path_data.py:139-159): bit-exact."""
import pytest
import torch

import pathpyg_b200 as pp
from pathpyg_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


@pytest.mark.parametrize("n,hi", [(0, 5), (1, 1), (7, 4), (5000, 9), (300_000, 12)])
def test_counts_to_offsets_and_repeat(cuda, n, hi):
    g = torch.Generator().manual_seed(n + hi)
    counts = torch.randint(0, hi, (n,), generator=g)          # zero counts included
    offsets, total = ops.counts_to_offsets(counts.to(cuda))
    want = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(counts, 0)])
    assert torch.equal(offsets.cpu(), want) and total == int(counts.sum())
    for values in (torch.rand(n, generator=g), torch.randint(-5, 5, (n,), generator=g), torch.rand(n, generator=g, dtype=torch.float64)):
        got = ops.repeat_by_count(values.to(cuda), counts.to(cuda))
        assert got.dtype == values.dtype and torch.equal(got.cpu(), values.repeat_interleave(counts))


def test_negative_count_and_empty_walk_raise(cuda):
    with pytest.raises(ValueError):
        ops.counts_to_offsets(torch.tensor([3, -1, 2], device=cuda))
    with pytest.raises(ValueError):
        ops.walk_chain(torch.tensor([3, 0, 2], device=cuda))


@pytest.mark.parametrize("walks", [1, 6, 40_000])
def test_walk_chain_matches_the_masked_arange(cuda, walks):
    g = torch.Generator().manual_seed(walks)
    lengths = torch.randint(1, 12, (walks,), generator=g)     # walks of a single node included
    total = int(lengths.sum())
    pos = torch.arange(total)
    chain = torch.stack([pos[:-1], pos[1:]])
    keep = torch.ones(chain.size(1), dtype=torch.bool)
    keep[torch.cumsum(lengths, 0)[:-1] - 1] = False
    got = ops.walk_chain(lengths.to(cuda), base=17)
    assert torch.equal(got.cpu(), chain[:, keep] + 17)


def test_bincount(cuda):
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 1000, (200_000,), generator=g)
    assert torch.equal(ops.bincount(ids.to(cuda)).cpu(), torch.bincount(ids))
    assert torch.equal(ops.bincount(ids.to(cuda), 1500).cpu(), torch.bincount(ids, minlength=1500))
    assert ops.bincount(torch.empty(0, dtype=torch.long, device=cuda)).numel() == 0
    with pytest.raises(ValueError):
        ops.bincount(ids.to(cuda), 10)
    with pytest.raises(ValueError):
        ops.bincount(torch.tensor([1, -2], device=cuda), 5)


def test_append_index_walks_on_device_matches_host_container(cuda):
    """The device container (one scan + one kernel) and the host container (index plumbing) hold the same walks, also
    when walks are appended twice (node offset of the container)."""
    g = torch.Generator().manual_seed(9)
    lengths = torch.randint(1, 9, (500,), generator=g)
    flat = torch.randint(0, 50, (int(lengths.sum()),), generator=g)
    weights = torch.rand(500, generator=g)
    host, dev = pp.PathData(pp.IndexMap(list(range(50)))), pp.PathData(pp.IndexMap(list(range(50))), device=cuda)
    for _ in range(2):
        host.append_index_walks(flat, lengths, weights)
        dev.append_index_walks(flat.to(cuda), lengths.to(cuda), weights.to(cuda))
    for name in ("edge_index", "node_sequence", "dag_weight", "dag_num_edges", "dag_num_nodes"):
        a, b = getattr(host.data, name), getattr(dev.data, name)
        assert torch.equal(torch.as_tensor(a), torch.as_tensor(b).cpu()), name
    assert int(host.data.num_nodes) == int(dev.data.num_nodes)
