#!/usr/bin/env python
"""Whole-sort time (histogram + passes, CUDA events inside the library) of compiled onesweep shapes over sizes and
distributions:  python tools/size_sweep.py V [V ...]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from b200sort import datagen
from b200sort._lib import ALGO_RADIX, check, lib
from helpers import stream_ptr, workspace

L = lib()
variants = [int(a) for a in sys.argv[1:]] or [0]


def timed(v, src):
    n = src.numel()
    check(L.b200sort_radix_set_variant(v))
    out = torch.empty_like(src); tmp = torch.empty_like(src)
    ws, ptr, nbytes = workspace(n, ALGO_RADIX)
    ms = (ctypes.c_float * 8)()
    acc = []
    for rep in range(8):
        check(L.b200sort_sort_timed_i32(ALGO_RADIX, src.data_ptr(), out.data_ptr(), tmp.data_ptr(), n, ptr, nbytes, stream_ptr(),
                                        ctypes.cast(ms, ctypes.c_void_p)))
        if rep >= 3:
            acc.append(sum(list(ms)[:6]))
    ok = bool((out[1:] >= out[:-1]).all().item())
    return float(np.mean(acc)), ok


print("shapes:", {v: L.b200sort_radix_variant_name(v).decode() for v in variants})
for log2n in (16, 18, 20, 22, 24, 26, 28):
    src = torch.from_numpy(datagen.uniform(1 << log2n, 3)).cuda()
    row = [timed(v, src) for v in variants]
    print(f"uniform   2^{log2n}: " + "  ".join(f"v{v} {t:8.4f} ms{'' if ok else ' WRONG'}" for v, (t, ok) in zip(variants, row)), flush=True)
for dist in ("skewed90", "and3", "ascending", "zipf16", "all_equal"):
    src = torch.from_numpy(datagen.make(dist, 1 << 28, 3)).cuda()
    row = [timed(v, src) for v in variants]
    print(f"{dist:9s} 2^28: " + "  ".join(f"v{v} {t:8.4f} ms{'' if ok else ' WRONG'}" for v, (t, ok) in zip(variants, row)), flush=True)
