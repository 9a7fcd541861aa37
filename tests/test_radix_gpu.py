"""Parity of the onesweep radix sort (CUDA, through the C-ABI) against the oracle.  Bit-exact."""
import hashlib

import numpy as np
import pytest

import oracle
from b200sort import datagen
from b200sort._lib import ALGO_RADIX, check, lib
from helpers import assert_bit_exact, gpu_sort, stream_ptr, to_device, workspace

pytestmark = pytest.mark.gpu


def test_digit_histograms_match_cpu_counts():
    import torch
    for dist, n in (("uniform", 1 << 20), ("and3", 100003), ("ascending", 4097), ("all_equal", 5),
                    ("edge_mix", 1 << 16), ("uniform", 3)):
        keys = datagen.make(dist, n, 2)
        d = to_device(keys)
        hist = torch.zeros(4 * 256, dtype=torch.int32, device="cuda")
        check(lib().b200sort_radix_histogram_i32(d.data_ptr(), n, hist.data_ptr(), stream_ptr()))
        got = hist.cpu().numpy().astype(np.uint64).reshape(4, 256)
        assert (got == oracle.digit_histograms(keys)).all(), (dist, n)


def test_histogram_handles_unaligned_pointers():
    import torch
    keys = datagen.uniform(10007, 4)
    buf = to_device(np.concatenate([np.zeros(3, np.int32), keys]))
    for off in (1, 2, 3):
        view = buf[off:off + 10000]
        hist = torch.zeros(4 * 256, dtype=torch.int32, device="cuda")
        check(lib().b200sort_radix_histogram_i32(view.data_ptr(), 10000, hist.data_ptr(), stream_ptr()))
        want = oracle.digit_histograms(view.cpu().numpy())
        assert (hist.cpu().numpy().astype(np.uint64).reshape(4, 256) == want).all(), off


@pytest.mark.parametrize("digit_pass", [0, 1, 2, 3])
def test_single_pass_is_the_stable_partition(digit_pass):
    import torch
    for dist, n in (("uniform", 1 << 18), ("and3", 50001), ("edge_mix", 20000), ("uniform", 17)):
        keys = datagen.make(dist, n, 3)
        d_in = to_device(keys)
        d_out = torch.empty_like(d_in)
        ws, ptr, nbytes = workspace(n, ALGO_RADIX)
        check(lib().b200sort_radix_pass_i32(d_in.data_ptr(), d_out.data_ptr(), n, digit_pass, ptr, nbytes,
                                            stream_ptr()))
        torch.cuda.synchronize()
        assert_bit_exact(d_out.cpu().numpy(), oracle.radix_pass(keys, digit_pass), f"{dist} n={n}")


@pytest.mark.parametrize("dist", sorted(datagen.DISTRIBUTIONS) + ["lab_rand100"])
@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 256, 4096, 8191, 8192, 8193, 65536, 100000, 1 << 20])
def test_radix_sort_bit_exact(dist, n):
    if dist == "lab_rand100" and n > 65536:
        pytest.skip("slow libc loop")
    keys = datagen.make(dist, n, 11)
    assert_bit_exact(gpu_sort(keys, ALGO_RADIX), oracle.radix_sort(keys), f"{dist} n={n}")


def test_radix_sort_all_tile_shapes():
    L = lib()
    keys = datagen.uniform((1 << 20) + 12345, 21)
    want = oracle.radix_sort(keys)
    low = datagen.masked(300001, 5, 0x00FF00FF)
    want_low = oracle.radix_sort(low)
    try:
        shapes = [v for v in range(L.b200sort_radix_num_variants())
                  if not L.b200sort_radix_variant_name(v).startswith(b"TIMING_")]     # the probes are the same kernels
        assert 3 <= len(shapes)
        for v in shapes:
            assert L.b200sort_radix_set_variant(v) == 0
            name = L.b200sort_radix_variant_name(v).decode()
            assert_bit_exact(gpu_sort(keys, ALGO_RADIX), want, name)
            assert_bit_exact(gpu_sort(low, ALGO_RADIX), want_low, name + " low-entropy")
    finally:
        L.b200sort_radix_set_variant(0)


def _gpu_sort_pairs(keys, vals, copy_form=False):
    import torch
    L = lib()
    n = keys.size
    dk, dv = to_device(keys), to_device(vals)
    tk = torch.empty(max(n, 1), dtype=torch.int32, device="cuda"); tv = torch.empty_like(tk)
    ws, ptr, nbytes = workspace(n, ALGO_RADIX)
    if copy_form:
        ok = torch.full((max(n, 1),), -1, dtype=torch.int32, device="cuda"); ov = torch.full_like(ok, -1)
        check(L.b200sort_radix_pairs_copy_i32(dk.data_ptr(), dv.data_ptr(), ok.data_ptr(), ov.data_ptr(),
                                              tk.data_ptr(), tv.data_ptr(), n, ptr, nbytes, stream_ptr()))
        torch.cuda.synchronize()
        assert_bit_exact(dk.cpu().numpy(), keys, "pairs copy form: keys modified")
        assert_bit_exact(dv.cpu().numpy(), vals, "pairs copy form: values modified")
        return ok[:n].cpu().numpy(), ov[:n].cpu().numpy()
    check(L.b200sort_radix_pairs_i32(dk.data_ptr(), dv.data_ptr(), tk.data_ptr(), tv.data_ptr(), n, ptr, nbytes,
                                     stream_ptr()))
    torch.cuda.synchronize()
    return dk.cpu().numpy(), dv.cpu().numpy()


@pytest.mark.parametrize("dist", ["uniform", "lab_rand100", "and3", "all_equal", "mask_0000ffff", "mask_00ff00ff",
                                  "skewed90", "descending", "edge_mix"])
@pytest.mark.parametrize("n", [0, 1, 2, 33, 5119, 5120, 5121, 100003, (1 << 20) + 17])
def test_sort_by_key_is_stable_and_bit_exact(dist, n):
    """Sort-by-key: keys AND values equal to the stable oracle's, in place and copy form.
    Values are the input positions, so the value array is the stable sorting permutation."""
    if dist == "lab_rand100" and n > 200000:
        pytest.skip("slow libc loop")
    keys = datagen.make(dist, n, 37)
    vals = np.arange(n, dtype=np.int32)
    want_k, want_v = oracle.sort_pairs(keys, vals)
    for copy_form in (False, True):
        got_k, got_v = _gpu_sort_pairs(keys, vals, copy_form)
        assert_bit_exact(got_k, want_k, f"{dist} n={n} keys copy={copy_form}")
        assert_bit_exact(got_v, want_v, f"{dist} n={n} values copy={copy_form}")


def test_sort_by_key_full_size_2_28():
    """n = 2^28 pairs: keys sorted, values a permutation that maps back to the input keys, stable."""
    import torch
    n = 1 << 28
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    keys = torch.randint(-2**15, 2**15, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)   # many ties
    orig = keys.clone()
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    tk = torch.empty_like(keys); tv = torch.empty_like(vals)
    ws, ptr, nbytes = workspace(n, ALGO_RADIX)
    check(lib().b200sort_radix_pairs_i32(keys.data_ptr(), vals.data_ptr(), tk.data_ptr(), tv.data_ptr(), n, ptr, nbytes,
                                         stream_ptr()))
    torch.cuda.synchronize()
    del tk, tv
    assert bool((keys[1:] >= keys[:-1]).all().item()), "keys not sorted"
    assert bool((orig[vals.long()] == keys).all().item()), "values do not point at their keys"
    ties = keys[1:] == keys[:-1]
    assert bool((vals[1:][ties] > vals[:-1][ties]).all().item()), "not stable"


def test_pass_skipping_on_and_off_agree():
    L = lib()
    try:
        for dist in ("mask_0000ffff", "mask_00ff00ff", "all_equal", "lab_rand100", "uniform"):
            keys = datagen.make(dist, 50000, 8)
            want = oracle.radix_sort(keys)
            for skip in (1, 0):
                L.b200sort_radix_set_skip(skip)
                assert_bit_exact(gpu_sort(keys, ALGO_RADIX), want, f"{dist} skip={skip}")
    finally:
        L.b200sort_radix_set_skip(1)


def test_golden_vectors_from_the_reference(golden_small, golden_large):
    for name, (keys, ref_out) in golden_small.items():
        assert_bit_exact(gpu_sort(keys, ALGO_RADIX), ref_out, name)
    for name, fx in golden_large.items():
        keys = datagen.make(fx["dist"], fx["n"], fx["seed"])
        got = gpu_sort(keys, ALGO_RADIX)
        assert hashlib.sha256(got.tobytes()).hexdigest() == fx["sha256_out"], name


def test_mixed_sign_is_the_reference_output_rotated(golden_mixed):
    for name, (keys, ref_out) in golden_mixed.items():
        assert_bit_exact(gpu_sort(keys, ALGO_RADIX), np.roll(ref_out, int((keys < 0).sum())), name)


def test_workspace_is_reusable_and_sort_is_idempotent():
    import torch
    n = 300000
    ws, ptr, nbytes = workspace(n, ALGO_RADIX)
    tmp = torch.empty(n, dtype=torch.int32, device="cuda")
    for seed in range(4):
        keys = datagen.uniform(n - seed * 1000, seed)
        d = to_device(keys)
        for _ in range(2):   # second call sorts sorted data: idempotence
            check(lib().b200sort_radix_i32(d.data_ptr(), tmp.data_ptr(), d.numel(), ptr, nbytes, stream_ptr()))
        torch.cuda.synchronize()
        assert_bit_exact(d.cpu().numpy(), oracle.radix_sort(keys), f"seed {seed}")


@pytest.mark.parametrize("dist", ["uniform", "ascending", "descending", "skewed90", "and3"])
def test_full_size_properties_2_28(dist):
    """BASELINE config sizes: sortedness + multiset fingerprint (size-independent properties)."""
    import torch
    n = 1 << 28
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    if dist == "uniform":
        d = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    elif dist == "ascending":
        d = (torch.arange(n, dtype=torch.int64, device="cuda") - n // 2).to(torch.int32)
    elif dist == "descending":
        d = (n // 2 - 1 - torch.arange(n, dtype=torch.int64, device="cuda")).to(torch.int32)
    elif dist == "skewed90":
        d = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
        hot = torch.rand(n, device="cuda", generator=g) < 0.9
        d = torch.where(hot, (d & 0x00FFFFFF) | 0x40000000, d)
    else:
        d = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
        for _ in range(2):
            d &= torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    before_sum = int(d.sum(dtype=torch.int64).item())
    before_xor = int(_xor_reduce(d))
    tmp = torch.empty_like(d)
    ws, ptr, nbytes = workspace(n, ALGO_RADIX)
    check(lib().b200sort_radix_i32(d.data_ptr(), tmp.data_ptr(), n, ptr, nbytes, stream_ptr()))
    torch.cuda.synchronize()
    assert bool((d[1:] >= d[:-1]).all().item()), "not sorted"
    assert int(d.sum(dtype=torch.int64).item()) == before_sum
    assert int(_xor_reduce(d)) == before_xor
    if dist in ("ascending", "descending"):
        want = (torch.arange(n, dtype=torch.int64, device="cuda") - n // 2).to(torch.int32)
        assert bool((d == want).all().item())


def _xor_reduce(d):
    import torch
    x = d.clone()
    while x.numel() > 1:
        if x.numel() % 2:
            x = torch.cat([x, torch.zeros(1, dtype=x.dtype, device=x.device)])
        x = x[: x.numel() // 2] ^ x[x.numel() // 2:]
    return x[0].item()


@pytest.mark.parametrize("algo_name", ["radix", "merge"])
def test_out_of_place_form_leaves_the_input_untouched(algo_name):
    import torch
    from helpers import ALGOS
    algo = ALGOS[algo_name]
    for dist in ("uniform", "mask_0000ffff", "all_equal", "mask_00ff00ff"):
        for n in (1, 1000, 70001):
            keys = datagen.make(dist, n, 23)
            d_in = to_device(keys); d_out = torch.full_like(d_in, -1); tmp = torch.empty_like(d_in)
            ws, ptr, nbytes = workspace(n, algo)
            check(lib().b200sort_sort_copy_i32(algo, d_in.data_ptr(), d_out.data_ptr(), tmp.data_ptr(), n,
                                               ptr, nbytes, stream_ptr()))
            torch.cuda.synchronize()
            assert_bit_exact(d_in.cpu().numpy(), keys, f"{dist} n={n}: input modified")
            assert_bit_exact(d_out.cpu().numpy(), oracle.radix_sort(keys), f"{dist} n={n}")


def test_safe_rank_fallback_is_selected_by_env_and_sorts():
    """B200SORT_RANK_SAFE=1 makes the library launch the ballot-ranked shape (documented behaviour
    only) instead of the atomicAdd-ranked one; the result is the same bytes."""
    import os
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')\n"
        "import oracle\n"
        "from b200sort import datagen\n"
        "from b200sort._lib import lib, ALGO_RADIX\n"
        "from helpers import gpu_sort\n"
        "L = lib()\n"
        "assert L.b200sort_radix_atomic_order_ok() == 0\n"
        "assert b'Ballot' in L.b200sort_radix_effective_variant_name()\n"
        "k = datagen.skewed(300001, 3)\n"
        "assert gpu_sort(k, ALGO_RADIX).tobytes() == oracle.radix_sort(k).tobytes()\n"
        "print('ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env={**os.environ, "B200SORT_RANK_SAFE": "1"},
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_default_pass_kernel_handles_partial_single_tiles_with_the_small_path_off():
    """n <= 8192 normally goes to the one-CTA kernel; with B200SORT_RADIX_SMALL=0 the default pass kernel itself has
    to sort a single, partially filled tile (and the histogram / plan kernels run on tiny inputs)."""
    import os
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')\n"
        "import oracle\n"
        "from b200sort import datagen\n"
        "from b200sort._lib import lib, ALGO_RADIX\n"
        "from helpers import gpu_sort\n"
        "L = lib()\n"
        "before = L.b200sort_launch_count()\n"
        "for dist in ('uniform', 'edge_mix', 'all_equal', 'mask_00ff00ff'):\n"
        "    for n in (2, 31, 32, 33, 1000, 8191, 8192):\n"
        "        k = datagen.make(dist, n, 3)\n"
        "        assert gpu_sort(k, ALGO_RADIX).tobytes() == oracle.radix_sort(k).tobytes(), (dist, n)\n"
        "assert L.b200sort_launch_count() - before >= 28 * 5, 'the one-CTA kernel ran instead of the pipeline'\n"
        "print('ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env={**os.environ, "B200SORT_RADIX_SMALL": "0"},
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_sort_by_key_falls_back_to_the_ballot_ranked_shape():
    """With B200SORT_RANK_SAFE=1 sort-by-key runs the ballot-ranked shape (round 1 returned an error here)."""
    import os
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')\n"
        "import numpy as np, torch\n"
        "import oracle\n"
        "from b200sort import datagen\n"
        "from b200sort._lib import lib, ALGO_RADIX, check\n"
        "from helpers import to_device, workspace, stream_ptr\n"
        "L = lib()\n"
        "assert L.b200sort_radix_atomic_order_ok() == 0\n"
        "for dist, n in (('lab_rand100', 70000), ('uniform', 300001), ('skewed90', 1 << 20), ('all_equal', 5000)):\n"
        "    k = datagen.make(dist, n, 41); v = np.arange(n, dtype=np.int32)\n"
        "    wk, wv = oracle.sort_pairs(k, v)\n"
        "    dk, dv = to_device(k), to_device(v); tk = torch.empty_like(dk); tv = torch.empty_like(dv)\n"
        "    ws, ptr, nb = workspace(n, ALGO_RADIX)\n"
        "    check(L.b200sort_radix_pairs_i32(dk.data_ptr(), dv.data_ptr(), tk.data_ptr(), tv.data_ptr(), n, ptr, nb, stream_ptr()))\n"
        "    torch.cuda.synchronize()\n"
        "    assert dk.cpu().numpy().tobytes() == wk.tobytes() and dv.cpu().numpy().tobytes() == wv.tobytes(), (dist, n)\n"
        "print('ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env={**os.environ, "B200SORT_RANK_SAFE": "1"},
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_atomic_order_selftest_reports():
    assert lib().b200sort_radix_atomic_order_ok() in (0, 1)
    name = lib().b200sort_radix_effective_variant_name().decode()
    assert ("kRankAdd" in name) == bool(lib().b200sort_radix_atomic_order_ok())
