"""Parity of the merge sort (block bitonic sort + merge-path passes) against the oracle.  Bit-exact."""
import numpy as np
import pytest

import oracle
from b200sort import datagen
from b200sort._lib import ALGO_MERGE, check, lib
from helpers import assert_bit_exact, gpu_sort, stream_ptr, to_device, workspace

pytestmark = pytest.mark.gpu


def test_block_sort_sorts_every_tile():
    """Every compiled block sort (4096 / 8192-key tiles, warp-register rounds on / off)."""
    import torch
    L = lib()
    try:
        for v in [v for v in (0, 2, 4, 6) if v < L.b200sort_merge_num_variants()]:   # 2, 4, 6: make EXPERIMENTS=1
            assert L.b200sort_merge_set_variant(v) == 0
            name = L.b200sort_merge_variant_name(v).decode()
            T = L.b200sort_block_sort_tile()
            for dist, n in (("uniform", 10 * T), ("edge_mix", 3 * T + 17), ("descending", T), ("uniform", 5),
                            ("and3", 2 * T - 1), ("all_equal", 2 * T), ("lab_rand100", T + 1000)):
                keys = datagen.make(dist, n, 6)
                d_in = to_device(keys); d_out = torch.empty_like(d_in)
                check(L.b200sort_block_sort_i32(d_in.data_ptr(), d_out.data_ptr(), n, stream_ptr()))
                torch.cuda.synchronize()
                got = d_out.cpu().numpy()
                for base in range(0, n, T):
                    assert_bit_exact(got[base:base + T], np.sort(keys[base:base + T]), f"{name} {dist} tile@{base}")
                # in place
                check(L.b200sort_block_sort_i32(d_in.data_ptr(), d_in.data_ptr(), n, stream_ptr()))
                torch.cuda.synchronize()
                assert_bit_exact(d_in.cpu().numpy(), got, f"{name} in place")
    finally:
        L.b200sort_merge_set_variant(0)


def test_partition_points_match_cpu_merge_path_and_pass_merges():
    import torch
    T = lib().b200sort_merge_tile()
    for n, run, dist in ((8 * T, T, "uniform"), (8 * T, 2 * T, "and3"), (5 * T + 100, T, "uniform"),
                         (6 * T + 1, 4 * T, "edge_mix"), (4 * T, 2 * T, "ascending"), (4 * T, 2 * T, "descending"),
                         (7 * T + 5, 3 * T, "uniform"), (600 * T + 77, 8 * T, "uniform")):
        keys = datagen.make(dist, n, 9)
        runs = keys.copy()
        for base in range(0, n, run):
            runs[base:base + run] = np.sort(runs[base:base + run])
        d_in = to_device(runs); d_out = torch.empty_like(d_in)
        tiles = (n + T - 1) // T
        splits = torch.zeros(tiles + 1, dtype=torch.int32, device="cuda")
        check(lib().b200sort_merge_partition_i32(d_in.data_ptr(), n, run, splits.data_ptr(), stream_ptr()))
        check(lib().b200sort_merge_pass_i32(d_in.data_ptr(), d_out.data_ptr(), n, run, splits.data_ptr(), stream_ptr()))
        torch.cuda.synchronize()
        sp = splits.cpu().numpy()
        want = runs.copy()
        for t in range(tiles):
            g = t * T
            base = g // (2 * run) * (2 * run)
            a = runs[base:base + run]; b = runs[base + run:base + 2 * run]
            assert sp[t] == oracle.merge_path(a, b, g - base), (n, run, dist, t)
        for base in range(0, n, 2 * run):
            want[base:base + 2 * run] = oracle.rank_merge(runs[base:base + run], runs[base + run:base + 2 * run])
        assert_bit_exact(d_out.cpu().numpy(), want, f"n={n} run={run} {dist}")


@pytest.mark.parametrize("dist", sorted(datagen.DISTRIBUTIONS) + ["lab_rand100"])
@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 256, 4095, 4096, 4097, 8192, 65536, 100000, 1 << 20])
def test_merge_sort_bit_exact(dist, n):
    if dist == "lab_rand100" and n > 65536:
        pytest.skip("slow libc loop")
    keys = datagen.make(dist, n, 13)
    assert_bit_exact(gpu_sort(keys, ALGO_MERGE), oracle.order_array(keys) if n <= 1 << 16 else oracle.radix_sort(keys),
                     f"{dist} n={n}")


def test_every_merge_pass_kernel_bit_exact():
    """All compiled merge-pass kernels (b200sort_merge_set_variant), incl. ragged tails, unaligned
    output (scalar-store path), INT_MAX keys next to the sentinels and runs of equal keys."""
    import torch
    L = lib()
    cases = [(d, n) for d in ("uniform", "edge_mix", "all_equal", "and3", "descending", "skewed90")
             for n in (4097, 8192, 40000, 100001, (1 << 20) + 4099)]
    int_max = np.full(3 * 4096 + 5, np.iinfo(np.int32).max, np.int32); int_max[::7] = 5
    try:
        for v in [v for v in list(range(8)) + [8, 16] if v < L.b200sort_merge_num_variants()]:
            assert L.b200sort_merge_set_variant(v) == 0
            name = L.b200sort_merge_variant_name(v).decode()
            for dist, n in cases:
                keys = datagen.make(dist, n, 31)
                assert_bit_exact(gpu_sort(keys, ALGO_MERGE), oracle.radix_sort(keys), f"{name} {dist} n={n}")
            assert_bit_exact(gpu_sort(int_max, ALGO_MERGE), oracle.radix_sort(int_max), f"{name} INT_MAX")
            # unaligned buffers: keys and scratch start 4 bytes past a 16-byte boundary
            keys = datagen.uniform(50001, 3)
            buf = to_device(np.concatenate([np.zeros(1, np.int32), keys])); d = buf[1:]
            tmp = torch.empty(50002, dtype=torch.int32, device="cuda")[1:]
            ws, ptr, nbytes = workspace(50001, ALGO_MERGE)
            check(L.b200sort_merge_i32(d.data_ptr(), tmp.data_ptr(), 50001, ptr, nbytes, stream_ptr()))
            torch.cuda.synchronize()
            assert_bit_exact(d.cpu().numpy(), oracle.radix_sort(keys), f"{name} unaligned")
    finally:
        L.b200sort_merge_set_variant(0)


def test_golden_vectors_from_the_reference(golden_small, golden_mixed):
    for name, (keys, ref_out) in golden_small.items():
        assert_bit_exact(gpu_sort(keys, ALGO_MERGE), ref_out, name)
    for name, (keys, ref_out) in golden_mixed.items():
        assert_bit_exact(gpu_sort(keys, ALGO_MERGE), np.roll(ref_out, int((keys < 0).sum())), name)


@pytest.mark.parametrize("dist", ["uniform", "descending"])
def test_full_size_properties_2_28(dist):
    import torch
    n = 1 << 28
    if dist == "uniform":
        g = torch.Generator(device="cuda"); g.manual_seed(6)
        d = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    else:
        d = (n // 2 - 1 - torch.arange(n, dtype=torch.int64, device="cuda")).to(torch.int32)
    before_sum = int(d.sum(dtype=torch.int64).item())
    tmp = torch.empty_like(d)
    ws, ptr, nbytes = workspace(n, ALGO_MERGE)
    check(lib().b200sort_merge_i32(d.data_ptr(), tmp.data_ptr(), n, ptr, nbytes, stream_ptr()))
    torch.cuda.synchronize()
    assert bool((d[1:] >= d[:-1]).all().item()), "not sorted"
    assert int(d.sum(dtype=torch.int64).item()) == before_sum


# ---- the assignment's staged pipeline as a third algorithm (B200SORT_ALGO_LAB) ----------------------

def test_lab_tile_sort_sorts_every_tile():
    import torch
    T = lib().b200sort_merge_tile()
    for dist, n in (("uniform", 6 * T), ("edge_mix", 2 * T + 33), ("descending", T), ("lab_rand100", 3000), ("all_equal", 70)):
        keys = datagen.make(dist, n, 6)
        d_in = to_device(keys); d_out = torch.empty_like(d_in)
        check(lib().b200sort_lab_tile_sort_i32(d_in.data_ptr(), d_out.data_ptr(), n, stream_ptr()))
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        for base in range(0, n, T):
            assert_bit_exact(got[base:base + T], oracle.order_array(keys[base:base + T]), f"{dist} tile@{base}")


@pytest.mark.parametrize("dist", ["uniform", "uniform_nonneg", "lab_rand100", "edge_mix", "and3", "descending", "all_equal", "skewed90"])
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 256, 4096, 4097, 65536, 100001, 1 << 20])
def test_lab_pipeline_bit_exact(dist, n):
    from b200sort._lib import ALGO_LAB
    if dist == "lab_rand100" and n > 65536:
        pytest.skip("slow libc loop")
    keys = datagen.make(dist, n, 29)
    want = oracle.order_array(keys) if n <= 1 << 16 else oracle.radix_sort(keys)
    assert_bit_exact(gpu_sort(keys, ALGO_LAB), want, f"{dist} n={n}")


def test_lab_pipeline_golden_vectors(golden_small):
    from b200sort._lib import ALGO_LAB
    for name, (keys, ref_out) in golden_small.items():
        assert_bit_exact(gpu_sort(keys, ALGO_LAB), ref_out, name)
