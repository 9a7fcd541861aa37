#!/usr/bin/env python
"""Parity of one compiled onesweep shape against the oracle, then its per-pass time.
    python tools/variant_check.py VARIANT [log2n]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import oracle
from b200sort import datagen
from b200sort._lib import ALGO_RADIX, check, lib
from helpers import gpu_sort, stream_ptr, to_device, workspace

v = int(sys.argv[1])
log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
L = lib()
check(L.b200sort_radix_set_variant(v))
name = L.b200sort_radix_variant_name(v).decode()
print("variant", v, name, "atomic order ok:", L.b200sort_radix_atomic_order_ok(), flush=True)
bad = 0
for dist, n in (("uniform", 1 << 20), ("uniform", 10240), ("uniform", 10241), ("uniform", 333), ("uniform", (1 << 22) + 12345),
                ("and3", 300001), ("skewed90", 1 << 21), ("ascending", 1 << 20), ("descending", 777777), ("all_equal", 50000),
                ("mask_00ff00ff", 400000), ("edge_mix", 123457), ("zipf16", 1 << 20), ("uniform", 1 << 24)):
    keys = datagen.make(dist, n, 5)
    got = gpu_sort(keys, ALGO_RADIX)
    ok = got.tobytes() == oracle.radix_sort(keys).tobytes()
    print(f"  {dist:14s} n={n:9d} {'ok' if ok else 'MISMATCH'}", flush=True)
    bad += (not ok)
# unaligned output buffer (the bulk copies need 16-byte aligned destinations: the kernel must cope)
keys = datagen.uniform(500000, 9)
buf = torch.empty(500000 + 8, dtype=torch.int32, device="cuda")
tmpb = torch.empty(500000 + 8, dtype=torch.int32, device="cuda")
for off in (1, 2, 3):
    d = buf[off:off + 500000]; d.copy_(torch.from_numpy(keys)); t = tmpb[(off + 1) % 4:(off + 1) % 4 + 500000]
    ws, ptr, nbytes = workspace(500000, ALGO_RADIX)
    check(L.b200sort_radix_i32(d.data_ptr(), t.data_ptr(), 500000, ptr, nbytes, stream_ptr()))
    torch.cuda.synchronize()
    ok = d.cpu().numpy().tobytes() == oracle.radix_sort(keys).tobytes()
    print(f"  unaligned buffers off={off} {'ok' if ok else 'MISMATCH'}", flush=True)
    bad += (not ok)
print("PARITY", "PASSED" if bad == 0 else f"FAILED ({bad})", flush=True)

n = 1 << log2n
g = torch.Generator(device="cuda"); g.manual_seed(1)
src = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
out = torch.empty_like(src); tmp = torch.empty_like(src)
ws, ptr, nbytes = workspace(n, ALGO_RADIX)
ms = (ctypes.c_float * 8)()
acc = np.zeros(8)
for rep in range(7):
    check(L.b200sort_sort_timed_i32(ALGO_RADIX, src.data_ptr(), out.data_ptr(), tmp.data_ptr(), n, ptr, nbytes, stream_ptr(),
                                    ctypes.cast(ms, ctypes.c_void_p)))
    if rep >= 2:
        acc += np.array(list(ms)) / 5
ok = bool((out[1:] >= out[:-1]).all().item()) and int(out.sum(dtype=torch.int64).item()) == int(src.sum(dtype=torch.int64).item())
print(f"TIMING n=2^{log2n} uniform: hist {acc[0]:.4f} ms, passes {acc[1]:.4f} {acc[2]:.4f} {acc[3]:.4f} {acc[4]:.4f} ms, "
      f"sum {acc[:6].sum():.4f} ms = {n / acc[:6].sum() / 1e6:.1f} Gkeys/s; pass frac of 6544 GB/s = {8.0 * n / (acc[1:5].mean() / 1e3) / 1e9 / 6544.3:.3f}; "
      f"sorted+sum {'ok' if ok else 'WRONG'}", flush=True)
sys.exit(1 if bad or not ok else 0)
