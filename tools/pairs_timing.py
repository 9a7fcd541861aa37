#!/usr/bin/env python
"""Sort-by-key timing: b200sort_radix_pairs_copy_i32 at n = 2^LOG2N (default 28), uniform keys, values = positions.
CUDA events on the launching stream, inputs resident, the same pristine input every step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from b200sort._lib import ALGO_RADIX, check, lib

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = 1 << log2n
L = lib()
g = torch.Generator(device="cuda"); g.manual_seed(1)
keys = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
vals = torch.arange(n, dtype=torch.int32, device="cuda")
ok, ov, tk, tv = (torch.empty_like(keys) for _ in range(4))
nbytes = L.b200sort_workspace_bytes(n, ALGO_RADIX)
ws = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
ptr = ws.data_ptr() + (-ws.data_ptr()) % 256
s = torch.cuda.current_stream().cuda_stream


def run():
    check(L.b200sort_radix_pairs_copy_i32(keys.data_ptr(), vals.data_ptr(), ok.data_ptr(), ov.data_ptr(),
                                          tk.data_ptr(), tv.data_ptr(), n, ptr, nbytes, s))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
assert bool((ok[1:] >= ok[:-1]).all().item()) and bool((keys[ov.long()] == ok).all().item())
# 4 (histogram) + 4 passes x 16 B/pair
print(f"pairs n=2^{log2n}: {ms:.3f} ms/sort = {n / ms / 1e6:.1f} Gpairs/s; "
      f"{(4 + 64) * n / ms / 1e6:.0f} GB/s of algorithmic traffic (68 B/pair)")
