"""Host-side logic of the multi-GPU path on CPU: the planner (through the C-ABI) with simulated
ranks, and a world_size-2 gloo run of the whole protocol with numpy standing in for the two CUDA
kernels (histogram, multisplit) -- counts, offsets, ownership and the global order are what is
under test here; the kernels themselves are covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from b200sort import datagen  # noqa: E402
from b200sort import dist as b200dist  # noqa: E402


def _simulate(keys_per_rank, bits):
    world = len(keys_per_rank)
    all_hist = np.stack([b200dist.host_histogram(k, bits) for k in keys_per_rank])
    plans = [b200dist.plan(all_hist, r, bits) for r in range(world)]
    owner = plans[0][0]
    recv_bufs = [np.full(int(plans[0][1][r]), np.int32(-7), dtype=np.int32) for r in range(world)]
    filled = [np.zeros(len(b), dtype=bool) for b in recv_bufs]
    for s in range(world):
        o, recv, send, offs = plans[s]
        assert (o == owner).all() and (recv == plans[0][1]).all()
        top = (keys_per_rank[s].view(np.uint32) ^ np.uint32(0x80000000)) >> np.uint32(32 - bits)
        dest = owner[top.astype(np.int64)]
        for r in range(world):
            block = keys_per_rank[s][dest == r]
            assert len(block) == int(send[r])
            lo = int(offs[r])
            assert not filled[r][lo:lo + len(block)].any(), "blocks overlap"
            recv_bufs[r][lo:lo + len(block)] = block
            filled[r][lo:lo + len(block)] = True
    assert all(f.all() for f in filled), "receive buffers have holes"
    return owner, [np.sort(b) for b in recv_bufs]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("dist_name", ["uniform", "skewed90", "and3", "all_equal", "ascending", "edge_mix", "lab_rand100"])
def test_planner_with_simulated_ranks(world, dist_name):
    bits = 8
    n = 20000 if dist_name != "lab_rand100" else 3000
    keys = [datagen.make(dist_name, n + 17 * r, seed=10 + r) for r in range(world)]
    owner, outs = _simulate(keys, bits)
    assert (np.diff(owner) >= 0).all() and owner.min() >= 0 and owner.max() < world
    merged = np.concatenate(outs)
    assert merged.tobytes() == np.sort(np.concatenate(keys)).tobytes()       # rank order == global order
    if dist_name == "uniform" and world > 1:
        sizes = np.array([len(o) for o in outs], dtype=np.float64)
        assert sizes.max() / sizes.mean() < 1.15                               # balanced


def test_planner_rejects_bad_arguments():
    from b200sort._lib import lib
    h = np.zeros(2 * 256, dtype=np.uint64)
    o = np.zeros(256, dtype=np.int32)
    L = lib()
    assert L.b200sort_dist_plan(h.ctypes.data, 2, 5, 8, o.ctypes.data, None, None, None) == 1   # rank >= world
    assert L.b200sort_dist_plan(h.ctypes.data, 99, 0, 8, o.ctypes.data, None, None, None) == 1  # world too large
    assert L.b200sort_dist_plan(h.ctypes.data, 2, 0, 2, o.ctypes.data, None, None, None) == 1   # bits too small
    assert L.b200sort_dist_plan(h.ctypes.data, 2, 0, 8, o.ctypes.data, None, None, None) == 0


def _gloo_worker(rank, world, port, bits, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    keys = datagen.skewed(30000 + 1000 * rank, seed=50 + rank)
    hist = torch.from_numpy(b200dist.host_histogram(keys, bits).astype(np.int64))
    gathered = [torch.zeros_like(hist) for _ in range(world)]
    dist.all_gather(gathered, hist)                                          # phase 2
    all_hist = np.stack([g.numpy().astype(np.uint64) for g in gathered])
    owner, recv, send, offs = b200dist.plan(all_hist, rank, bits)
    top = (keys.view(np.uint32) ^ np.uint32(0x80000000)) >> np.uint32(32 - bits)
    dest = owner[top.astype(np.int64)]
    recv_buf = np.zeros(int(recv[rank]), dtype=np.int32)
    # phase 3 stand-in: every block travels to its destination at the planner's offset
    for s in range(world):
        for r in range(world):
            if s == r and s == rank:
                block = keys[dest == r]
                recv_buf[int(offs[r]):int(offs[r]) + len(block)] = block
            elif s == rank:
                block = torch.from_numpy(np.ascontiguousarray(keys[dest == r]))
                meta = torch.tensor([int(offs[r]), block.numel()], dtype=torch.int64)
                dist.send(meta, dst=r)
                if block.numel():
                    dist.send(block, dst=r)
            elif r == rank:
                meta = torch.zeros(2, dtype=torch.int64)
                dist.recv(meta, src=s)
                if int(meta[1]):
                    block = torch.zeros(int(meta[1]), dtype=torch.int32)
                    dist.recv(block, src=s)
                    recv_buf[int(meta[0]):int(meta[0]) + int(meta[1])] = block.numpy()
    out = np.sort(recv_buf)                                                  # phase 4 stand-in
    np.save(os.path.join(out_dir, f"in{rank}.npy"), keys)
    np.save(os.path.join(out_dir, f"out{rank}.npy"), out)
    dist.barrier()
    dist.destroy_process_group()


def test_protocol_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    world, bits = 2, 8
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(world, port, bits, str(tmp_path)), nprocs=world, join=True)
    ins = [np.load(tmp_path / f"in{r}.npy") for r in range(world)]
    outs = [np.load(tmp_path / f"out{r}.npy") for r in range(world)]
    assert np.concatenate(outs).tobytes() == np.sort(np.concatenate(ins)).tobytes()
