#!/usr/bin/env python
"""Where does a tile's lifetime go?  Runs one onesweep pass with the TIMING shape and prints the
mean time (SM cycles) between the phase stamps of group A (warp 0) and group B (warp 8)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from b200sort._lib import lib, check, ALGO_RADIX
L = lib()
n = 1 << 28
want = sys.argv[1].encode() if len(sys.argv) > 1 else b""
v = [i for i in range(L.b200sort_radix_num_variants())
     if L.b200sort_radix_variant_name(i).startswith(b"TIMING") and want in L.b200sort_radix_variant_name(i)][-1]
print("shape:", L.b200sort_radix_variant_name(v).decode())
check(L.b200sort_radix_set_variant(v))
tile = L.b200sort_radix_tile(); tiles = (n + tile - 1) // tile
g = torch.Generator(device="cuda"); g.manual_seed(1)
src = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
out = torch.empty_like(src)
wsb = L.b200sort_workspace_bytes(n, ALGO_RADIX)
ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda"); wp = ws.data_ptr() + (-ws.data_ptr()) % 256
dbg = torch.zeros(tiles * 2 * 16, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    check(L.b200sort_radix_pass_i32(src.data_ptr(), out.data_ptr(), n, 1, wp, wsb, s))   # warm
torch.cuda.synchronize()
check(L.b200sort_debug_set_phase_buffer(dbg.data_ptr()))
check(L.b200sort_radix_pass_i32(src.data_ptr(), out.data_ptr(), n, 1, wp, wsb, s))
torch.cuda.synchronize()
check(L.b200sort_debug_set_phase_buffer(None))
d = dbg.cpu().numpy().reshape(tiles, 2, 16).astype(np.float64)
names = ["start", "ticket+zero sync", "loads landed", "ranked", "post-rank sync", "digit group done", "staged", "pre-output sync", "written"]
mid = d[tiles // 8: tiles * 7 // 8]                      # steady state
for grp, label in ((0, "group A (warp 0)"), (1, "group B (warp 8)")):
    print(label)
    for i in range(1, 9):
        dt = mid[:, grp, i] - mid[:, grp, i - 1]
        print(f"  {names[i-1]:>18s} -> {names[i]:<18s} mean {dt.mean():8.0f} cyc   p50 {np.median(dt):8.0f}   p90 {np.percentile(dt, 90):8.0f}")
    tot = mid[:, grp, 8] - mid[:, grp, 0]
    print(f"  tile lifetime mean {tot.mean():.0f} cyc ({tot.mean()/1.965e3:.2f} us)")

if mid[:, 1, 10].max() > 0:
    print("group B detail (two-level shapes)")
    for a, b2, label in ((4, 10, "post-rank sync -> staged (waits for group A's positions, stages)"),
                         (10, 11, "level-1 walk (earlier tiles of my group)"),
                         (11, 12, "level-2 walk (earlier groups)")):
        dt = mid[:, 1, b2] - mid[:, 1, a]
        print(f"  {label:<66s} mean {dt.mean():8.0f} cyc   p50 {np.median(dt):8.0f}   p90 {np.percentile(dt, 90):8.0f}")
    r = (d[tiles // 8: tiles * 7 // 8, 1, 9] % 32).astype(int)
    dt1 = mid[:, 1, 11] - mid[:, 1, 10]
    for lo, hi in ((0, 1), (1, 8), (8, 16), (16, 31), (31, 32)):
        m = (r >= lo) & (r < hi)
        if m.any(): print(f"    level-1 walk for r in [{lo},{hi}): mean {dt1[m].mean():8.0f} cyc")
