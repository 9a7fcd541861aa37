#!/usr/bin/env python
"""Phase timing of the tensor-memory / bulk-store pass kernel (per tile iteration; clock64 stamps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from b200sort._lib import lib, check, ALGO_RADIX
L = lib()
n = 1 << 28
want = sys.argv[1].encode() if len(sys.argv) > 1 else b"TIMING_tma"
v = [i for i in range(L.b200sort_radix_num_variants()) if L.b200sort_radix_variant_name(i).startswith(want)][-1]
print("shape:", L.b200sort_radix_variant_name(v).decode())
check(L.b200sort_radix_set_variant(v))
tile = L.b200sort_radix_tile(); tiles = (n + tile - 1) // tile
g = torch.Generator(device="cuda"); g.manual_seed(1)
src = torch.randint(-2**31, 2**31, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
out = torch.empty_like(src)
wsb = L.b200sort_workspace_bytes(n, ALGO_RADIX)
ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda"); wp = ws.data_ptr() + (-ws.data_ptr()) % 256
dbg = torch.zeros((tiles + 1) * 2 * 16, dtype=torch.int64, device="cuda")   # + a dummy row
s = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    check(L.b200sort_radix_pass_i32(src.data_ptr(), out.data_ptr(), n, 1, wp, wsb, s))
torch.cuda.synchronize()
check(L.b200sort_debug_set_phase_buffer(dbg.data_ptr()))
check(L.b200sort_radix_pass_i32(src.data_ptr(), out.data_ptr(), n, 1, wp, wsb, s))
torch.cuda.synchronize()
check(L.b200sort_debug_set_phase_buffer(None))
d = dbg.cpu().numpy().reshape(tiles + 1, 2, 16)[:tiles].astype(np.float64)
mid = d[tiles // 8: tiles * 7 // 8]
mid = mid[(mid[:, 0, 7] > 0) & (mid[:, 1, 7] > 0) & (mid[:, 1, 10] > 0)]
if b"tma3" in want:
    # the staggered schedule: the two groups name their phases differently
    if b"tma3a" in want:
        na = ["start", "A counted all", "A published", "L passed", "-", "A staged", "Y passed", "loads issued"]
        nb = ["start", "B resolved", "-", "barriers passed", "-", "B staged", "Y passed", "written"]
        for grp, label, nm, idx in ((0, "group A (warp 0)", na, (0, 1, 2, 3, 5, 6, 7)), (1, "group B (warp 8)", nb, (0, 1, 3, 5, 6, 7))):
            print(label)
            for a, b2 in zip(idx[:-1], idx[1:]):
                dt = mid[:, grp, b2] - mid[:, grp, a]
                print(f"  {nm[a]:>18s} -> {nm[b2]:<18s} mean {dt.mean():8.0f} cyc   p50 {np.median(dt):8.0f}   p90 {np.percentile(dt, 90):8.0f}")
            tot = mid[:, grp, 7] - mid[:, grp, 0]
            print(f"  iteration mean {tot.mean():.0f} cyc ({tot.mean()/1.965e3:.2f} us)")
        for a, b2, label in ((0, 11, "start -> fetched rows landed (mbarrier)"), (11, 10, "look-back sums"), (10, 1, "start words added to the counters")):
            dt = mid[:, 1, b2] - mid[:, 1, a]
            print(f"group B: {label:<48s} mean {dt.mean():8.0f} cyc   p50 {np.median(dt):8.0f}   p90 {np.percentile(dt, 90):8.0f}   max {dt.max():8.0f}")
        sys.exit(0)
    if b"tma3i" in want:
        na = ["start", "A counted", "X passed", "A published", "L passed", "A staged", "Y passed", "written"]
        nb = ["start", "B counted", "rows requested", "B resolved", "barriers passed", "B staged", "Y passed", "tails written"]
        for grp, label, nm in ((0, "group A (warp 0)", na), (1, "group B (warp 8)", nb)):
            print(label)
            for i in range(1, 8):
                dt = mid[:, grp, i] - mid[:, grp, i - 1]
                print(f"  {nm[i-1]:>18s} -> {nm[i]:<18s} mean {dt.mean():8.0f} cyc   p50 {np.median(dt):8.0f}   p90 {np.percentile(dt, 90):8.0f}")
            tot = mid[:, grp, 7] - mid[:, grp, 0]
            print(f"  iteration mean {tot.mean():.0f} cyc ({tot.mean()/1.965e3:.2f} us)")
        for a, b2, label in ((2, 11, "rows requested -> landed (mbarrier)"), (11, 10, "look-back sums"), (10, 3, "start words added to the counters")):
            dt = mid[:, 1, b2] - mid[:, 1, a]
            print(f"group B: {label:<48s} mean {dt.mean():8.0f} cyc   p50 {np.median(dt):8.0f}   p90 {np.percentile(dt, 90):8.0f}   max {dt.max():8.0f}")
        sys.exit(0)
    na = ["start", "A counted", "L passed", "A staged", "X passed", "A published", "Y passed", "written"]
    nb = ["start", "B resolved", "L arrived", "B counted", "X passed", "B staged", "Y passed", "rows requested"]
    for grp, label, nm in ((0, "group A (warp 0)", na), (1, "group B (warp 8)", nb)):
        print(label)
        for i in range(1, 8):
            dt = mid[:, grp, i] - mid[:, grp, i - 1]
            print(f"  {nm[i-1]:>18s} -> {nm[i]:<18s} mean {dt.mean():8.0f} cyc   p50 {np.median(dt):8.0f}   p90 {np.percentile(dt, 90):8.0f}")
        tot = mid[:, grp, 7] - mid[:, grp, 0]
        print(f"  iteration mean {tot.mean():.0f} cyc ({tot.mean()/1.965e3:.2f} us)")
    for a, b2, label in ((0, 11, "start -> fetched rows landed (mbarrier)"), (11, 12, "sums over the fetched rows"), (12, 10, "walks beyond the fetched rows (global)"), (10, 1, "start words added to the counters")):
        dt = mid[:, 1, b2] - mid[:, 1, a]
        print(f"group B: {label:<48s} mean {dt.mean():8.0f} cyc   p50 {np.median(dt):8.0f}   p90 {np.percentile(dt, 90):8.0f}   max {dt.max():8.0f}")
    sys.exit(0)
names = {0: "keys in regs", 1: "ranked+parked", 2: "SYNC1", 3: "digit group done", 4: "SYNC2", 5: "staged", 6: "SYNC3", 7: "prev written"}
for grp, label in ((0, "group A (warp 0)"), (1, "group B (warp 8)")):
    print(label)
    for i in range(1, 8):
        dt = mid[:, grp, i] - mid[:, grp, i - 1]
        print(f"  {names[i-1]:>18s} -> {names[i]:<18s} mean {dt.mean():8.0f} cyc   p50 {np.median(dt):8.0f}   p90 {np.percentile(dt, 90):8.0f}")
    tot = mid[:, grp, 7] - mid[:, grp, 0]
    print(f"  iteration (keys in regs -> prev written) mean {tot.mean():.0f} cyc ({tot.mean()/1.965e3:.2f} us)")

for a, b2, label in ((2, 11, "SYNC1 -> fetched rows landed (mbarrier)"), (11, 12, "sum over the tile rows"), (12, 10, "sum over the group rows"), (10, 3, "staging layout + counters -> positions")):
    dt = mid[:, 1, b2] - mid[:, 1, a]
    print(f"group B: {label:<48s} mean {dt.mean():8.0f} cyc   p50 {np.median(dt):8.0f}   p90 {np.percentile(dt, 90):8.0f}   max {dt.max():8.0f}")
