#!/usr/bin/env python
"""Generates tests/golden/*.npz by running THE REFERENCE ITSELF on seeded inputs.

The reference ships no golden vectors, so these are produced here from its
own code, compiled unmodified by ``make -C oracle ref``:

  ref_cpu   = order_with_trust   (SRM/lab.cu:404-406, Thrust sequential host sort) -- runs on CPU.

Run in the build container (needs /root/reference and oracle/_ref/libreflab.so):

    python tests/golden/make_golden.py

Files written
  lab_small.npz    inputs + reference outputs for the reference's own recipes (rand()%100 as
                   SRM/main.cpp:10, rand()%1000 as SRM/performanceTest.cpp:35 drawn back to back
                   from the default glibc seed) and for non-negative seeded distributions, at sizes
                   the fixture can hold verbatim (n <= 4096).
  mixed_sign.npz   inputs + reference outputs on mixed-sign keys.  The reference's CPU path sorts
                   these in UNSIGNED order (Thrust 2.8.2 RadixEncoder<int> widens to 64 bits before
                   flipping bit 31 on LP64); recorded as evidence, see tests/test_oracle.py.
  large.json       sha256 of the reference output for seeded inputs too large to store
                   (n = 2^16, 2^20; the input is regenerated from b200sort.datagen).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from b200sort import datagen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main() -> None:
    if not oracle.ref.available:
        raise SystemExit("oracle/_ref/libreflab.so missing: run `make -C oracle ref` first")
    ref = oracle.ref.order_with_trust

    small_in, small_out = {}, {}

    def add(name: str, keys: np.ndarray) -> None:
        small_in[name] = keys.astype(np.int32)
        small_out[name] = ref(keys)

    # the reference's own recipes
    for n in (256, 512, 1024, 2048, 4096):
        add(f"main_rand100_seed1_n{n}", datagen.lab_rand(n, 100, seed=1))
    datagen.lab_rand(0, 1000, seed=1)            # performanceTest: default seed, sizes back to back
    first = True
    for n in (256, 512, 1024, 2048, 4096):
        add(f"perftest_rand1000_seq_n{n}", datagen.lab_rand(n, 1000, seed=1 if first else None))
        first = False
    # non-negative seeded distributions (the sign domain the reference was exercised on)
    for n in (0, 1, 2, 31, 32, 33, 1000, 4096):
        add(f"uniform_nonneg_seed1_n{n}", datagen.uniform_nonneg(n, 1))
    add("all_equal_n1024", datagen.all_equal(1024, 7))
    add("ascending_nonneg_n1024", np.arange(1024, dtype=np.int32))
    add("descending_nonneg_n1024", np.arange(1024, dtype=np.int32)[::-1].copy())
    add("and3_nonneg_n4096", datagen.and_k(4096, 1, 3) & np.int32(0x7FFFFFFF))
    add("mask_0000ffff_n4096", datagen.masked(4096, 1, 0x0000FFFF))
    np.savez_compressed(os.path.join(HERE, "lab_small.npz"),
                        **{f"in__{k}": v for k, v in small_in.items()},
                        **{f"out__{k}": v for k, v in small_out.items()})

    mixed = {}
    for name, keys in (("uniform_seed1_n4096", datagen.uniform(4096, 1)),
                       ("edge_mix_seed1_n1024", datagen.edge_mix(1024, 1)),
                       ("ascending_n1024", datagen.ascending(1024)),
                       ("descending_n1024", datagen.descending(1024))):
        mixed[f"in__{name}"] = keys
        mixed[f"out__{name}"] = ref(keys)
    np.savez_compressed(os.path.join(HERE, "mixed_sign.npz"), **mixed)

    large = {}
    for dist, n in (("uniform_nonneg", 1 << 16), ("uniform_nonneg", 1 << 20),
                    ("lab_rand100", 1 << 16), ("mask_0000ffff", 1 << 20)):
        keys = datagen.make(dist, n, 1)
        out = ref(keys)
        large[f"{dist}_seed1_n{n}"] = {
            "dist": dist, "seed": 1, "n": n,
            "sha256_in": hashlib.sha256(keys.tobytes()).hexdigest(),
            "sha256_out": hashlib.sha256(out.tobytes()).hexdigest(),
        }
    with open(os.path.join(HERE, "large.json"), "w") as f:
        json.dump(large, f, indent=1, sort_keys=True)
    print("wrote", len(small_in), "small,", len(mixed) // 2, "mixed-sign,", len(large), "large fixtures")


if __name__ == "__main__":
    main()
