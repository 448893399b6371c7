"""The radix sort underneath a1 / a4 / the CSC build, through its C-ABI entry (ppg_sort_pairs_u64), against torch's
stable sort: uniform, skewed and degenerate key distributions, sizes around the tile-size switch, 2 to 8 digit passes."""
import pytest
import torch

from pathpyg_b200 import ops

pytestmark = pytest.mark.gpu


def check(keys: torch.Tensor, bits: int, cuda):
    want = torch.sort(keys, stable=True)
    k = keys.to(cuda)
    perm, _ = ops.sort_pairs_u64(k, bits)
    assert torch.equal(k.cpu(), want.values)
    assert torch.equal(perm.cpu().long(), want.indices)          # stable: equal keys keep their input order


@pytest.mark.parametrize("n", [1, 1000, 32 << 10, 100_003, 1_000_000, 1_800_000, 2_097_152, 2_097_153, 5_000_000])
@pytest.mark.parametrize("bits", [17, 34, 40, 62])
def test_uniform_keys(cuda, n, bits):
    gen = torch.Generator().manual_seed(n % 1000 + bits)
    check(torch.randint(0, 1 << bits, (n,), generator=gen), bits, cuda)


@pytest.mark.parametrize("bits", [32, 40, 50])
def test_skewed_and_degenerate_keys(cuda, bits):
    gen = torch.Generator().manual_seed(bits)
    n = 600_000
    top = 1 << (bits - 3)
    # every key shares its top digit
    check(torch.randint(0, top >> 6, (n,), generator=gen), bits, cuda)
    # one huge bucket next to many small ones
    mixed = torch.cat([torch.randint(0, top >> 6, (300_000,), generator=gen), torch.randint(0, 1 << bits, (300_000,), generator=gen)])
    check(mixed[torch.randperm(n, generator=gen)], bits, cuda)
    # a block of 8192 / 8193 keys sharing one top digit among spread-out keys
    for size in (8192, 8193):
        lo = torch.randint(0, 1 << (bits - 8), (size,), generator=gen) | (5 << (bits - 8)) if bits % 8 == 0 else \
            torch.randint(0, 1 << ((bits - 1) // 8 * 8), (size,), generator=gen) | (1 << ((bits - 1) // 8 * 8))
        rest = torch.randint(0, 1 << ((bits - 1) // 8 * 8), (40_000,), generator=gen)
        check(torch.cat([lo, rest])[torch.randperm(size + 40_000, generator=gen)], bits, cuda)
    # few distinct values (long runs of equal keys: stability), all equal, already sorted, reversed
    check(torch.randint(0, 7, (n,), generator=gen) << (bits - 4), bits, cuda)
    check(torch.full((n,), (1 << bits) - 1), bits, cuda)
    asc = torch.arange(n) * ((1 << bits) // n)
    check(asc, bits, cuda)
    check(asc.flip(0), bits, cuda)
