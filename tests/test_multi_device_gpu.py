"""Paths that need more than one GPU (skipped on a one-GPU box): the real distributed sort under torchrun,
checked against a gathered host sort, and one process driving two devices."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from b200sort import datagen
from b200sort._lib import ALGO_MERGE, ALGO_RADIX, check, lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_sort_matches_gathered_host_sort(world):
    """torchrun --nproc-per-node WORLD tests/dist_check.py: every exchange mode x distribution, the ranks'
    outputs gathered and compared byte for byte with np.sort of the gathered inputs."""
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = {**os.environ, "MASTER_ADDR": "127.0.0.1"}
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
                        os.path.join(ROOT, "tests", "dist_check.py")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_one_process_two_devices():
    """Function attributes, the lane-order self-test and scratch allocations are per device: the same process
    sorts on cuda:0, then on cuda:1, then on cuda:0 again (device arrays and the host operator)."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import torch
    L = lib()
    keys = datagen.uniform((1 << 21) + 777, 91)
    want = oracle.radix_sort(keys)
    small = datagen.uniform(5000, 92)
    try:
        for dev in (0, 1, 0):
            torch.cuda.set_device(dev)
            check(L.b200sort_device_check())
            for algo in (ALGO_RADIX, ALGO_MERGE):
                d = torch.from_numpy(keys).to(f"cuda:{dev}")
                tmp = torch.empty_like(d)
                nbytes = L.b200sort_workspace_bytes(d.numel(), algo)
                ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=f"cuda:{dev}")
                ptr = ws.data_ptr() + (-ws.data_ptr()) % 256
                check(L.b200sort_sort_i32(algo, d.data_ptr(), tmp.data_ptr(), d.numel(), ptr, nbytes,
                                          torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
                assert d.cpu().numpy().tobytes() == want.tobytes(), (dev, algo)
            hist = torch.zeros(1024, dtype=torch.int32, device=f"cuda:{dev}")
            d = torch.from_numpy(small).to(f"cuda:{dev}")
            check(L.b200sort_radix_histogram_i32(d.data_ptr(), d.numel(), hist.data_ptr(), torch.cuda.current_stream().cuda_stream))
            assert (hist.cpu().numpy().astype(np.uint64).reshape(4, 256) == oracle.digit_histograms(small)).all()
            h = small.copy()
            check(L.b200sort_order_array_host(h.ctypes.data, h.size, ALGO_RADIX))
            assert h.tobytes() == oracle.radix_sort(small).tobytes(), dev
    finally:
        L.b200sort_host_release()
        torch.cuda.set_device(0)
