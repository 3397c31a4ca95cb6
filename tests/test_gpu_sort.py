"""Stage-level sort plug-in (vkgsb_sort_key_value_indirect, the vrdxCmdSortKeyValueIndirect equivalent) against the
reference's own criterion: equality with std::stable_sort by key applied to keys AND values
(third_party/vulkan_radix_sort/bench/bench.cc:69-129, cpu_benchmark.cc:29-51).  Integer work: bit-exact."""
import numpy as np
import pytest

import vkgs_b200
from oracle import oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def gpu_sort(keys, vals, count=None, max_n=None):
    n = len(keys)
    max_n = n if max_n is None else max_n
    count = n if count is None else count
    dev = torch.device("cuda:0")
    k = torch.from_numpy(keys.view(np.int32).copy()).to(dev)
    v = torch.from_numpy(vals.view(np.int32).copy()).to(dev)
    c = torch.tensor([count], dtype=torch.int32, device=dev)
    storage = torch.empty(vkgs_b200.sort_storage_bytes(max(max_n, 1)), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    vkgs_b200.sort_key_value_indirect(s, max_n, c.data_ptr(), k.data_ptr(), v.data_ptr(), storage.data_ptr())
    torch.cuda.synchronize()
    return k.cpu().numpy().view(np.uint32), v.cpu().numpy().view(np.uint32)


def stable_ref(keys, vals):
    order = np.argsort(keys, kind="stable")
    return keys[order], vals[order]


@pytest.mark.parametrize("n", [0, 1, 2, 33, 4095, 4096, 4097, 8191, 100_000, 1 << 20, (1 << 22) + 12345])
def test_uniform_keys(n):
    rng = np.random.default_rng(n + 7)                      # the reference bench is unseeded (bench.cc:24 TODO); ours is not
    keys = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    vals = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    k, v = gpu_sort(keys, vals)
    rk, rv = stable_ref(keys, vals)
    assert np.array_equal(k, rk) and np.array_equal(v, rv)


@pytest.mark.parametrize("bits", [1, 4, 8, 13, 24])
def test_few_distinct_keys_is_stable(bits):
    """data_generator.cc:12-27 `bits` knob: many ties, so stability is what is being checked."""
    n = 300_001
    rng = np.random.default_rng(bits)
    keys = rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    k, v = gpu_sort(keys, vals)
    rk, rv = stable_ref(keys, vals)
    assert np.array_equal(k, rk) and np.array_equal(v, rv)


def test_all_equal_and_extreme_keys():
    n = 70_000
    vals = np.arange(n, dtype=np.uint32)
    for fill in (0, 0xFFFFFFFF, 0x80000000):
        keys = np.full(n, fill, np.uint32)
        k, v = gpu_sort(keys, vals)
        assert np.array_equal(k, keys) and np.array_equal(v, vals)
    keys = np.where(np.arange(n) % 2 == 0, 0xFFFFFFFF, 0).astype(np.uint32)
    k, v = gpu_sort(keys, vals)
    rk, rv = stable_ref(keys, vals)
    assert np.array_equal(k, rk) and np.array_equal(v, rv)


def test_indirect_count_smaller_than_capacity():
    """vrdxCmdSortKeyValueIndirect: launch sized for maxElementCount, count read on the device; the tail is untouched."""
    n, count = 200_000, 123_457
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    k, v = gpu_sort(keys, vals, count=count, max_n=n)
    rk, rv = stable_ref(keys[:count], vals[:count])
    assert np.array_equal(k[:count], rk) and np.array_equal(v[:count], rv)
    assert np.array_equal(k[count:], keys[count:]) and np.array_equal(v[count:], vals[count:])


def test_depth_like_keys_match_oracle_and_reference_cpu_sort():
    """Keys shaped like bits(1 - z): narrow live range, low byte mostly zero, ~14 splats per key."""
    rng = np.random.default_rng(11)
    n = 2_000_000
    depth = rng.uniform(1.0, 10.0, n).astype(np.float32)
    z = (np.float32(1.0001) - np.float32(0.010001) / depth).astype(np.float32)
    keys = (np.float32(1.0) - z).view(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    k, v = gpu_sort(keys, vals)
    ok, ov = O.sort_pairs(keys, vals)
    assert np.array_equal(k, ok) and np.array_equal(v, ov)
    from oracle import ref as R
    if R.available():
        rk, rv = R.sort_key_value(keys[:300_000], vals[:300_000])        # the author's std::stable_sort statement
        k2, v2 = gpu_sort(keys[:300_000], vals[:300_000])
        assert np.array_equal(k2, rk) and np.array_equal(v2, rv)


def test_sortedness_and_permutation_at_2pow25():
    """Size-independent properties at the reference bench's headline size (README: 2^25 key-value)."""
    n = 1 << 25
    rng = np.random.default_rng(25)
    keys = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    k, v = gpu_sort(keys, vals)
    assert np.all(k[1:] >= k[:-1])                                        # sortedness
    assert np.array_equal(keys[v], k)                                     # values still travel with their keys
    seen = np.zeros(n, np.bool_); seen[v] = True
    assert seen.all()                                                     # permutation
    ties = k[1:] == k[:-1]
    assert np.all(v[1:][ties] > v[:-1][ties])                             # stability
