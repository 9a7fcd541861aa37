"""The oracle (oracle/oracle_sort.c) pinned against the reference's own outputs and against numpy.

CPU only.  Golden vectors come from the reference's CPU path, order_with_trust (SRM/lab.cu:404-406),
run by tests/golden/make_golden.py."""
import hashlib

import numpy as np
import pytest

import oracle
from b200sort import datagen


def test_pipeline_matches_reference_outputs(golden_small):
    assert len(golden_small) >= 20
    for name, (keys, ref_out) in golden_small.items():
        assert oracle.order_array(keys).tobytes() == ref_out.tobytes(), name


def test_radix_leg_matches_reference_outputs(golden_small):
    for name, (keys, ref_out) in golden_small.items():
        assert oracle.radix_sort(keys).tobytes() == ref_out.tobytes(), name


def test_large_fixtures_by_hash(golden_large):
    for name, fx in golden_large.items():
        keys = datagen.make(fx["dist"], fx["n"], fx["seed"])
        assert hashlib.sha256(keys.tobytes()).hexdigest() == fx["sha256_in"], name
        assert hashlib.sha256(oracle.radix_sort(keys).tobytes()).hexdigest() == fx["sha256_out"], name
        if fx["n"] <= 1 << 16:
            assert hashlib.sha256(oracle.order_array(keys).tobytes()).hexdigest() == fx["sha256_out"], name


def test_mixed_sign_reference_is_unsigned_order_and_oracle_is_its_rotation(golden_mixed):
    """The reference's CPU path orders mixed-sign keys as unsigned (negatives after positives);
    north_star fixes signed order.  Both orders hold the same two sorted runs, so the oracle's
    output must be the reference's output rotated by the number of negative keys."""
    for name, (keys, ref_out) in golden_mixed.items():
        unsigned = np.sort(keys.view(np.uint32)).view(np.int32)
        assert ref_out.tobytes() == unsigned.tobytes(), name
        n_neg = int((keys < 0).sum())
        want = np.roll(ref_out, n_neg)
        assert oracle.order_array(keys).tobytes() == want.tobytes(), name
        assert oracle.radix_sort(keys).tobytes() == want.tobytes(), name
        assert want.tobytes() == np.sort(keys).tobytes(), name


@pytest.mark.skipif(not oracle.ref.available, reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_against_live_reference_cpu_path():
    for dist in ("uniform_nonneg", "lab_rand100", "lab_rand1000", "mask_0000ffff", "all_equal"):
        for n in (32, 64, 1024, 1 << 15):
            keys = datagen.make(dist, n, 3)
            ref_out = oracle.ref.order_with_trust(keys)
            assert oracle.radix_sort(keys).tobytes() == ref_out.tobytes(), (dist, n)
            assert oracle.order_array(keys).tobytes() == ref_out.tobytes(), (dist, n)


@pytest.mark.parametrize("dist", sorted(datagen.DISTRIBUTIONS))
@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 100, 1024, 5000])
def test_oracle_equals_numpy_signed_sort(dist, n):
    keys = datagen.make(dist, n, 5)
    want = np.sort(keys, kind="stable")
    assert oracle.order_array(keys).tobytes() == want.tobytes()
    assert oracle.radix_sort(keys).tobytes() == want.tobytes()


def test_split_tile_is_the_assignment_split_primitive():
    # letra.pdf p.2: zeros keep order in front, ones keep order behind (stable), bit by bit
    tile, iters = oracle.split_tile32(np.array([5, 1, 4, 0, 7, 2, 6, 3], dtype=np.int32))
    assert tile.tolist() == [0, 1, 2, 3, 4, 5, 6, 7] and iters == 3
    tile, iters = oracle.split_tile32(np.arange(32, dtype=np.int32))
    assert iters == 0                                   # early exit, SRM/lab.cu:61
    tile, iters = oracle.split_tile32(np.array([3, -1, 2, -5, 0, 7, -2, 1], dtype=np.int32))
    assert tile.tolist() == [-5, -2, -1, 0, 1, 2, 3, 7] and iters == 32   # needs the sign bit


def test_rank_and_rank_merge_tie_rule():
    run = np.array([1, 3, 3, 3, 9], dtype=np.int32)
    assert oracle.rank(run, 3, True) == 1 and oracle.rank(run, 3, False) == 4   # SRM/lab.cu:126-130
    assert oracle.rank(run, 0, True) == 0 and oracle.rank(run, 10, False) == 5
    a = np.array([1, 3, 3, 8], dtype=np.int32)
    b = np.array([3, 3, 4], dtype=np.int32)
    assert oracle.rank_merge(a, b).tolist() == [1, 3, 3, 3, 3, 4, 8]
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = np.sort(rng.integers(-5, 5, rng.integers(0, 40)).astype(np.int32))
        b = np.sort(rng.integers(-5, 5, rng.integers(0, 40)).astype(np.int32))
        assert oracle.rank_merge(a, b).tolist() == sorted(a.tolist() + b.tolist())


def test_stage_helpers_against_numpy():
    keys = datagen.uniform(10000, 9)
    h = oracle.digit_histograms(keys)
    bits = keys.view(np.uint32) ^ np.uint32(0x80000000)
    for p in range(4):
        want = np.bincount(((bits >> np.uint32(8 * p)) & np.uint32(255)).astype(np.int64), minlength=256)
        assert (h[p] == want).all()
        out = oracle.radix_pass(keys, p)
        order = np.argsort(((bits >> np.uint32(8 * p)) & np.uint32(255)), kind="stable")
        assert out.tobytes() == keys[order].tobytes()
    a = np.sort(datagen.uniform(300, 1)); b = np.sort(datagen.uniform(200, 2))
    merged = np.sort(np.concatenate([a, b]), kind="stable")
    for diag in (0, 1, 100, 250, 499, 500):
        i = oracle.merge_path(a, b, diag)
        assert np.array_equal(np.sort(np.concatenate([a[:i], b[:diag - i]])), merged[:diag])
    assert oracle.is_sorted(merged) and not oracle.is_sorted(keys)
    assert oracle.multiset_fingerprint(keys) == oracle.multiset_fingerprint(np.sort(keys))
    assert oracle.multiset_fingerprint(keys) != oracle.multiset_fingerprint(keys + 1)


def test_sort_pairs_is_the_stable_order_the_rank_merge_tie_rule_gives():
    """oracle.sort_pairs against an independent statement: merge sort by rank_merge's tie rule (equal keys
    of the left run first, SRM/lab.cu:163-170) on keys tagged with their input position."""
    rng = np.random.default_rng(5)
    for n in (1, 2, 33, 1000, 4097):
        keys = rng.integers(-5, 5, n).astype(np.int32)                 # many ties
        vals = np.arange(n, dtype=np.int32)
        got_k, got_v = oracle.sort_pairs(keys, vals)
        # independent: bottom-up merges, ties take from the left run
        runs = [[(int(k), int(v))] for k, v in zip(keys, vals)]
        while len(runs) > 1:
            nxt = []
            for i in range(0, len(runs), 2):
                if i + 1 == len(runs):
                    nxt.append(runs[i]); continue
                a, b, out, x, y = runs[i], runs[i + 1], [], 0, 0
                while x < len(a) or y < len(b):
                    if y == len(b) or (x < len(a) and a[x][0] <= b[y][0]):
                        out.append(a[x]); x += 1
                    else:
                        out.append(b[y]); y += 1
                nxt.append(out)
            runs = nxt
        assert [p[0] for p in runs[0]] == got_k.tolist() and [p[1] for p in runs[0]] == got_v.tolist()
        assert got_k.tobytes() == oracle.radix_sort(keys).tobytes()
