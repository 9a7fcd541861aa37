"""Shared plumbing of the GPU parity tests: everything goes through the C-ABI (ctypes), torch only
owns the device memory."""
from __future__ import annotations

import numpy as np

import b200sort
from b200sort._lib import ALGO_LAB, ALGO_MERGE, ALGO_RADIX, check, lib

ALGOS = {"radix": ALGO_RADIX, "merge": ALGO_MERGE, "lab": ALGO_LAB}


def to_device(a: np.ndarray):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def workspace(n: int, algo: int):
    import torch
    nbytes = lib().b200sort_workspace_bytes(n, algo)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
    ptr = ws.data_ptr() + (-ws.data_ptr()) % 256
    return ws, ptr, nbytes


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def gpu_sort(keys: np.ndarray, algo: int) -> np.ndarray:
    """Sorted copy of ``keys`` computed by the CUDA path (device-array C-ABI)."""
    import torch
    d = to_device(keys)
    tmp = torch.empty_like(d) if d.numel() else torch.empty(1, dtype=torch.int32, device="cuda")
    ws, ptr, nbytes = workspace(d.numel(), algo)
    check(lib().b200sort_sort_i32(algo, d.data_ptr(), tmp.data_ptr(), d.numel(), ptr, nbytes, stream_ptr()))
    torch.cuda.synchronize()
    return d.cpu().numpy()


def assert_bit_exact(got: np.ndarray, want: np.ndarray, what: str = "") -> None:
    assert got.dtype == want.dtype == np.int32
    assert got.shape == want.shape, what
    if got.tobytes() != want.tobytes():
        bad = np.flatnonzero(got != want)
        raise AssertionError(f"{what}: {bad.size} of {got.size} keys differ, first at {bad[0]}: "
                             f"got {got[bad[0]]}, want {want[bad[0]]}")
