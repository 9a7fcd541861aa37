"""The multi-GPU path's two CUDA kernels on ONE GPU: the receive buffers of all simulated ranks are
local allocations, so the multisplit's per-destination blocks can be checked key for key."""
import ctypes

import numpy as np
import pytest

from b200sort import datagen
from b200sort import dist as b200dist
from b200sort._lib import check, lib
from helpers import stream_ptr, to_device

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits", [4, 8, 12])
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
