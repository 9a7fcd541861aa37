#!/usr/bin/env python
"""What compute-sanitizer runs (tools/sanitize.sh): every kernel family once at smoke sizes, each result checked
against the oracle.  `python tools/sanitize_target.py [big]` -- `big` adds one 2^20-key radix sort."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import oracle
from b200sort import datagen
from b200sort import dist as b200dist
from b200sort._lib import ALGO_LAB, ALGO_MERGE, ALGO_RADIX, check, lib
from helpers import gpu_sort, stream_ptr, to_device, workspace

L = lib()
check(L.b200sort_device_check())
big = len(sys.argv) > 1 and sys.argv[1] == "big"
n = 1 << 16


def same(a, b, what):
    assert a.tobytes() == b.tobytes(), what
    print("  ok", what, flush=True)


keys = datagen.uniform(n, 3)
want = oracle.radix_sort(keys)
skew = datagen.skewed(n + 123, 4)
for v in range(L.b200sort_radix_num_variants()):
    name = L.b200sort_radix_variant_name(v).decode()
    if name.startswith("TIMING_"):
        continue
    check(L.b200sort_radix_set_variant(v))
    same(gpu_sort(keys, ALGO_RADIX), want, f"radix shape {v} {name} n=2^16 uniform")
    same(gpu_sort(skew, ALGO_RADIX), oracle.radix_sort(skew), f"radix shape {v} n=2^16+123 skewed90 (hot-digit path, ragged tile)")
check(L.b200sort_radix_set_variant(0))
if big:
    k20 = datagen.uniform(1 << 20, 5)
    same(gpu_sort(k20, ALGO_RADIX), oracle.radix_sort(k20), "radix default n=2^20 uniform")
small = datagen.uniform(5000, 6)
same(gpu_sort(small, ALGO_RADIX), oracle.radix_sort(small), "one-CTA kernel n=5000")
same(gpu_sort(keys, ALGO_MERGE), want, "merge sort n=2^16")
same(gpu_sort(skew, ALGO_MERGE), oracle.radix_sort(skew), "merge sort n=2^16+123")
same(gpu_sort(keys, ALGO_LAB), want, "lab pipeline n=2^16")
# sort-by-key
vals = np.arange(n, dtype=np.int32)
ties = datagen.lab_rand(n, 100, seed=2)
wk, wv = oracle.sort_pairs(ties, vals)
dk, dv = to_device(ties), to_device(vals)
tk, tv = torch.empty_like(dk), torch.empty_like(dv)
ws, ptr, nb = workspace(n, ALGO_RADIX)
check(L.b200sort_radix_pairs_i32(dk.data_ptr(), dv.data_ptr(), tk.data_ptr(), tv.data_ptr(), n, ptr, nb, stream_ptr()))
torch.cuda.synchronize()
same(dk.cpu().numpy(), wk, "pairs: keys"); same(dv.cpu().numpy(), wv, "pairs: values (stable)")
# multi-GPU kernels with simulated ranks on one device
world, bits = 4, 12
srcs = [datagen.make("uniform", 40000 + 100 * r, seed=20 + r) for r in range(world)]
all_hist = np.stack([b200dist.host_histogram(k, bits) for k in srcs])
plans = [b200dist.plan(all_hist, r, bits) for r in range(world)]
owner, recv = plans[0][0], plans[0][1]
bufs = [torch.full((max(int(recv[r]), 1),), -7, dtype=torch.int32, device="cuda") for r in range(world)]
base = (ctypes.c_void_p * world)(*[b.data_ptr() for b in bufs])
owner_dev = to_device(owner)
pws = torch.zeros(512, dtype=torch.uint8, device="cuda")
pws_ptr = pws.data_ptr() + (-pws.data_ptr()) % 256
for s in range(world):
    d = to_device(srcs[s])
    hist = torch.zeros(1 << bits, dtype=torch.int64, device="cuda")
    check(L.b200sort_dist_histogram_i32(d.data_ptr(), d.numel(), bits, hist.data_ptr(), stream_ptr()))
    assert (hist.cpu().numpy().astype(np.uint64) == all_hist[s]).all()
    check(L.b200sort_dist_partition_i32(d.data_ptr(), d.numel(), bits, world, base, owner_dev.data_ptr(),
                                        plans[s][3].ctypes.data, pws_ptr, 256, stream_ptr()))
    torch.cuda.synchronize()
everything = np.concatenate(srcs)
top = (everything.view(np.uint32) ^ np.uint32(0x80000000)) >> np.uint32(32 - bits)
dest = owner[top.astype(np.int64)]
for r in range(world):
    same(np.sort(bufs[r].cpu().numpy()[:int(recv[r])]), np.sort(everything[dest == r]), f"dist partition: destination {r}")
# host operator
h = datagen.lab_rand(4096, 100, seed=1)
wh = oracle.order_array(h)
a = h.copy(); check(L.b200sort_order_array_host(a.ctypes.data, a.size, ALGO_RADIX)); same(a, wh, "order_array host operator")
L.b200sort_host_release()
if L.b200sort_debug_checked_build():
    # full-size runs too: the checks cost little
    for dist_name in ("uniform", "skewed90", "ascending"):
        k = datagen.make(dist_name, (1 << 24) + 4321, 8)
        same(gpu_sort(k, ALGO_RADIX), oracle.radix_sort(k), f"radix default n=2^24+4321 {dist_name}")
    check(L.b200sort_radix_set_variant(1))
    k = datagen.uniform((1 << 24) + 4321, 9)
    same(gpu_sort(k, ALGO_RADIX), oracle.radix_sort(k), "radix small-tile shape n=2^24+4321 uniform")
    check(L.b200sort_radix_set_variant(0))
    for dist_name in ("zipf16", "and3", "edge_mix"):
        k = datagen.make(dist_name, (1 << 23) + 77, 10)
        same(gpu_sort(k, ALGO_RADIX), oracle.radix_sort(k), f"radix default n=2^23+77 {dist_name}")
    sites = (ctypes.c_ulonglong * 16)()
    fails = int(L.b200sort_debug_check_failures_by_site(ctypes.cast(sites, ctypes.c_void_p)))
    print('per check site:', list(sites), flush=True)
    print(f"CHECKED BUILD: {fails} violated kernel invariants (staged positions, destination indices, bulk-copy alignment)", flush=True)
    assert fails == 0
print("SANITIZE TARGET PASSED", flush=True)
