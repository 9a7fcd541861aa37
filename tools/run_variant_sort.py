#!/usr/bin/env python
"""A few sorts of 2^log2n uniform keys with one compiled onesweep shape (profiling target).
    python tools/run_variant_sort.py VARIANT [log2n] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200sort._lib import ALGO_RADIX, check, lib
v = int(sys.argv[1]); log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 28; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L = lib(); check(L.b200sort_radix_set_variant(v)); n = 1 << log2n
g = torch.Generator(device="cuda"); g.manual_seed(1)
src = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
out = torch.empty_like(src); tmp = torch.empty_like(src)
wsb = L.b200sort_workspace_bytes(n, ALGO_RADIX)
ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda"); wp = ws.data_ptr() + (-ws.data_ptr()) % 256
for _ in range(reps):
    check(L.b200sort_sort_copy_i32(ALGO_RADIX, src.data_ptr(), out.data_ptr(), tmp.data_ptr(), n, wp, wsb, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
assert bool((out[1:] >= out[:-1]).all().item())
print("ok", L.b200sort_radix_variant_name(v).decode())
