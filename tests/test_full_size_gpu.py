"""BASELINE.json's full sizes, byte for byte: the CUDA path against the oracle (`memcmp`) at n = 2^28, and
n = 2^30 (the largest size of `configs` and B200SORT_MAX_N) by fingerprints plus the cases the status-word
width makes delicate.  Everything goes through the C-ABI."""
import numpy as np
import pytest

import oracle
from b200sort import datagen
from b200sort._lib import ALGO_MERGE, ALGO_RADIX, check, lib
from helpers import stream_ptr, workspace

pytestmark = pytest.mark.gpu


def _device_sort(keys_np, algo):
    import torch
    d = torch.from_numpy(keys_np).cuda()
    tmp = torch.empty_like(d)
    ws, ptr, nbytes = workspace(d.numel(), algo)
    check(lib().b200sort_sort_i32(algo, d.data_ptr(), tmp.data_ptr(), d.numel(), ptr, nbytes, stream_ptr()))
    torch.cuda.synchronize()
    del tmp, ws
    return d.cpu().numpy()


@pytest.mark.parametrize("dist", ["uniform", "skewed90"])
def test_2_28_is_bit_exact_for_radix_and_merge(dist):
    """n = 2^28 (the metric's size): both sorts `memcmp`-equal to the oracle's LSD byte radix sort
    (oracle/oracle_sort.c, pinned to the reference's own CPU path by tests/test_oracle.py)."""
    n = 1 << 28
    keys = datagen.make(dist, n, 77)
    want = oracle.radix_sort(keys)
    for algo, name in ((ALGO_RADIX, "radix"), (ALGO_MERGE, "merge")):
        got = _device_sort(keys, algo)
        assert got.tobytes() == want.tobytes(), f"{name} sort of 2^28 {dist} keys differs from the oracle"
        del got


def _fingerprint(t):
    """Order-independent 64-bit fingerprints of an int32 CUDA tensor: sum, and sum of a per-key mix (so that two
    multisets with equal sums still differ)."""
    import torch
    x = t.to(torch.int64)
    s1 = int(x.sum().item())
    y = (x * 0x9E3779B1 + 0x7F4A7C15) & 0xFFFFFFFF
    y = (y ^ (y >> 15)) * 0x2C1B3C6D & 0xFFFFFFFF
    s2 = int(y.sum().item())
    return s1, s2


def _sorted_on_device(t):
    return bool((t[1:] >= t[:-1]).all().item())


def test_2_30_radix_uniform_fingerprint():
    import torch
    n = 1 << 30
    g = torch.Generator(device="cuda"); g.manual_seed(30)
    parts = [torch.randint(-2**31, 2**31, (n // 4,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32) for _ in range(4)]
    d = torch.cat(parts); del parts
    before = _fingerprint(d)
    tmp = torch.empty_like(d)
    ws, ptr, nbytes = workspace(n, ALGO_RADIX)
    check(lib().b200sort_radix_i32(d.data_ptr(), tmp.data_ptr(), n, ptr, nbytes, stream_ptr()))
    torch.cuda.synchronize()
    del tmp
    assert _sorted_on_device(d), "2^30 keys: not sorted"
    assert _fingerprint(d) == before, "2^30 keys: not a permutation of the input"


def test_2_30_all_equal_and_the_skip_switch():
    """One bin holds all 2^30 keys in every pass: a digit count of exactly 2^30 does not fit the 30-bit status
    words.  With pass skipping (the default) every pass is the identity and the sort is a copy; with skipping
    switched off the size is refused instead of corrupting the look-back."""
    import torch
    n = 1 << 30
    L = lib()
    d = torch.full((n,), -123456789, dtype=torch.int32, device="cuda")
    tmp = torch.empty_like(d)
    ws, ptr, nbytes = workspace(n, ALGO_RADIX)
    check(L.b200sort_radix_i32(d.data_ptr(), tmp.data_ptr(), n, ptr, nbytes, stream_ptr()))
    torch.cuda.synchronize()
    assert bool((d == -123456789).all().item())
    try:
        L.b200sort_radix_set_skip(0)
        assert L.b200sort_radix_i32(d.data_ptr(), tmp.data_ptr(), n, ptr, nbytes, stream_ptr()) == 1   # B200SORT_ERR_INVALID
        # one key fewer is accepted with skipping off: counts stay below 2^30
        check(L.b200sort_radix_i32(d.data_ptr(), tmp.data_ptr(), n - 1, ptr, nbytes, stream_ptr()))
        torch.cuda.synchronize()
        assert bool((d == -123456789).all().item())
    finally:
        L.b200sort_radix_set_skip(1)


def test_2_30_two_valued_keys_keep_their_counts():
    """Half the keys 0, half -1, interleaved: digit counts of 2^29 in every pass (the largest count that can
    occur without a skipped pass) and the sign flip of the top digit."""
    import torch
    n = 1 << 30
    d = (torch.arange(n, dtype=torch.int32, device="cuda") & 1) - 1           # -1, 0, -1, 0, ...
    tmp = torch.empty_like(d)
    ws, ptr, nbytes = workspace(n, ALGO_RADIX)
    check(lib().b200sort_radix_i32(d.data_ptr(), tmp.data_ptr(), n, ptr, nbytes, stream_ptr()))
    torch.cuda.synchronize()
    assert bool((d[: n // 2] == -1).all().item()) and bool((d[n // 2:] == 0).all().item())
