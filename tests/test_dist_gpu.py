"""The multi-GPU path's two CUDA kernels on ONE GPU: the receive buffers of all simulated ranks are
local allocations, so the multisplit's per-destination blocks can be checked key for key."""
import ctypes

import numpy as np
import pytest

import oracle

from b200sort import datagen
from b200sort import dist as b200dist
from b200sort._lib import check, lib
from helpers import stream_ptr, to_device

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits", [4, 8, 12, 14])
def test_msd_histogram_matches_numpy(bits):
    import torch
    for dist_name, n in (("uniform", 1 << 20), ("skewed90", 300001), ("edge_mix", 5000), ("all_equal", 77)):
        keys = datagen.make(dist_name, n, 3)
        d = to_device(keys)
        hist = torch.zeros(1 << bits, dtype=torch.int64, device="cuda")
        check(lib().b200sort_dist_histogram_i32(d.data_ptr(), n, bits, hist.data_ptr(), stream_ptr()))
        assert (hist.cpu().numpy().astype(np.uint64) == b200dist.host_histogram(keys, bits)).all(), (dist_name, n)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("dist_name", ["uniform", "skewed90", "all_equal", "ascending"])
def test_partition_kernel_with_simulated_ranks(world, dist_name):
    """Every simulated source rank scatters into the (local) receive buffers of all destinations at
    the planner's offsets; afterwards destination r holds exactly the keys of its value range."""
    import torch
    bits, n = 12, 200000 if world != 4 else 6000000     # the larger one gives every CTA several tiles
    srcs = [datagen.make(dist_name, n + 1000 * r, seed=20 + r) for r in range(world)]
    all_hist = np.stack([b200dist.host_histogram(k, bits) for k in srcs])
    plans = [b200dist.plan(all_hist, r, bits) for r in range(world)]
    owner, recv = plans[0][0], plans[0][1]
    bufs = [torch.full((max(int(recv[r]), 1),), -7, dtype=torch.int32, device="cuda") for r in range(world)]
    base = (ctypes.c_void_p * world)(*[b.data_ptr() for b in bufs])
    owner_dev = to_device(owner)
    ws = torch.zeros(512, dtype=torch.uint8, device="cuda")
    ws_ptr = ws.data_ptr() + (-ws.data_ptr()) % 256
    for s in range(world):
        d = to_device(srcs[s])
        offs = plans[s][3]
        check(lib().b200sort_dist_partition_i32(d.data_ptr(), d.numel(), bits, world, base, owner_dev.data_ptr(),
                                                offs.ctypes.data, ws_ptr, 256, stream_ptr()))
        torch.cuda.synchronize()
    everything = np.concatenate(srcs)
    top = (everything.view(np.uint32) ^ np.uint32(0x80000000)) >> np.uint32(32 - bits)
    dest = owner[top.astype(np.int64)]
    for r in range(world):
        got = np.sort(bufs[r].cpu().numpy()[:int(recv[r])])
        want = np.sort(everything[dest == r])
        assert got.tobytes() == want.tobytes(), (world, dist_name, r)


PLAN_BYTES, PLAN_M_OFFSET, PLAN_TOP_OFFSET = 1424, 384, 400


def _device_plan(all_hist, world, rank, bits, cap):
    """(owner int32[nbins], recv, send, offs uint64[world], m, error) from b200sort_dist_plan_device."""
    import torch
    L = lib()
    d_hist = torch.from_numpy(all_hist.astype(np.int64).reshape(-1)).cuda()
    owner = torch.full((1 << bits,), -1, dtype=torch.int32, device="cuda")
    rec = torch.zeros(PLAN_BYTES // 8, dtype=torch.int64, device="cuda")
    nb = L.b200sort_dist_workspace_bytes(0, bits)
    ws = torch.zeros(nb + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = ws.data_ptr() + (-ws.data_ptr()) % 256
    check(L.b200sort_dist_plan_device(d_hist.data_ptr(), world, rank, bits, cap, owner.data_ptr(), rec.data_ptr(), ws_ptr, nb, stream_ptr()))
    torch.cuda.synchronize()
    r = rec.cpu().numpy()
    u64, u32 = r.view(np.uint64), r.view(np.uint32)
    return (owner.cpu().numpy(), u64[:world].copy(), u64[16:16 + world].copy(), u64[32:32 + world].copy(),
            int(u32[PLAN_M_OFFSET // 4]), int(u32[PLAN_M_OFFSET // 4 + 1]), owner, rec)


@pytest.mark.parametrize("bits", [4, 8, 12, 14])
@pytest.mark.parametrize("world", [1, 2, 3, 8, 16])
def test_device_planner_equals_host_planner(world, bits):
    """The planner that runs on the GPU (no host synchronisation in a sort) and the pure host planner the CPU tests
    exercise place the same boundaries, bit for bit, and agree on every count and offset."""
    for dist_name in ("uniform", "skewed90", "all_equal", "ascending", "edge_mix"):
        srcs = [datagen.make(dist_name, 20000 + 777 * r, seed=40 + r) for r in range(world)]
        all_hist = np.stack([b200dist.host_histogram(k, bits) for k in srcs])
        for rank in sorted({0, world // 2, world - 1}):
            owner_h, recv_h, send_h, offs_h = b200dist.plan(all_hist, rank, bits)
            owner_d, recv_d, send_d, offs_d, m, err, _, _ = _device_plan(all_hist, world, rank, bits, 1 << 40)
            assert (owner_d == owner_h).all(), (dist_name, world, bits, rank)
            assert (recv_d == recv_h).all() and (send_d == send_h).all() and (offs_d == offs_h).all()
            assert m == int(recv_h[rank]) and err == 0
    # a receive buffer that is too small is reported, not overrun
    *_, err, _, _ = _device_plan(all_hist, world, 0, bits, 10)
    assert err == 1


@pytest.mark.parametrize("shape", [0, 3])                  # 3: the large-array pass kernel whatever the size
@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("dist_name", ["uniform", "skewed90", "all_equal"])
def test_sync_free_path_with_simulated_ranks(world, dist_name, shape):
    """Device plan -> planned partition (bulk copies) -> local sort with the key count read from the device record:
    the whole multi-GPU sequence on one GPU, every destination compared with np.sort of the keys it owns."""
    import torch
    L = lib()
    if shape >= L.b200sort_radix_num_variants():
        pytest.skip("shape not compiled")
    check(L.b200sort_radix_set_variant(shape))
    try:
        _sync_free_path(L, torch, world, dist_name)
    finally:
        L.b200sort_radix_set_variant(0)


def _sync_free_path(L, torch, world, dist_name):
    bits, n = 14, 300000
    srcs = [datagen.make(dist_name, n + 1000 * r, seed=60 + r) for r in range(world)]
    all_hist = np.stack([b200dist.host_histogram(k, bits) for k in srcs])
    owner_h, recv_h, _, _ = b200dist.plan(all_hist, 0, bits)
    cap = int(recv_h.max()) + 64
    bufs = [torch.full((cap,), -7, dtype=torch.int32, device="cuda") for _ in range(world)]
    base = (ctypes.c_void_p * world)(*[b.data_ptr() for b in bufs])
    nb = L.b200sort_dist_workspace_bytes(0, bits)
    ws = torch.zeros(nb + 256, dtype=torch.uint8, device="cuda")
    ws_ptr = ws.data_ptr() + (-ws.data_ptr()) % 256
    recs, src_hists = [], []
    for s in range(world):
        *_, owner_t, rec_t = _device_plan(all_hist, world, s, bits, cap)
        recs.append(rec_t)
        d = to_device(srcs[s])
        src_hists.append(torch.full((world * 1024,), -1, dtype=torch.int32, device="cuda"))
        check(L.b200sort_dist_partition_planned_i32(d.data_ptr(), d.numel(), bits, world, base, owner_t.data_ptr(),
                                                    rec_t.data_ptr(), src_hists[-1].data_ptr(),
                                                    ws_ptr, nb, stream_ptr()))
        torch.cuda.synchronize()
    everything = np.concatenate(srcs)
    top = (everything.view(np.uint32) ^ np.uint32(0x80000000)) >> np.uint32(32 - bits)
    dest = owner_h[top.astype(np.int64)]
    from b200sort._lib import ALGO_RADIX
    wsb = L.b200sort_workspace_bytes(cap, ALGO_RADIX)
    sws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda")
    sws_ptr = sws.data_ptr() + (-sws.data_ptr()) % 256
    for r in range(world):
        m = int(recv_h[r])
        want = np.sort(everything[dest == r])
        # what a reduce-scatter over the ranks would hand rank r: the digit histograms of exactly the keys it received
        mine = torch.stack([h.view(world, 1024)[r] for h in src_hists]).sum(0).to(torch.int32).contiguous()
        assert int(mine[768:].abs().sum().item()) == 0              # the top byte is not counted at the source ...
        mine[768:] = recs[r].view(torch.int32)[PLAN_TOP_OFFSET // 4:PLAN_TOP_OFFSET // 4 + 256]   # ... it follows from the plan
        assert (mine.cpu().numpy().astype(np.uint64).reshape(4, 256) == oracle.digit_histograms(want)).all(), (world, dist_name, r)
        for d_hist in (None, mine.data_ptr()):            # the local sort with its own histogram kernel, and without
            out = torch.full((cap,), -9, dtype=torch.int32, device="cuda"); tmp = torch.empty_like(out)
            check(L.b200sort_radix_copy_devn_i32(bufs[r].data_ptr(), out.data_ptr(), tmp.data_ptr(), cap,
                                                 recs[r].data_ptr() + PLAN_M_OFFSET, d_hist, sws_ptr, wsb, stream_ptr()))
            torch.cuda.synchronize()
            assert out.cpu().numpy()[:m].tobytes() == want.tobytes(), (world, dist_name, r, d_hist is not None)
