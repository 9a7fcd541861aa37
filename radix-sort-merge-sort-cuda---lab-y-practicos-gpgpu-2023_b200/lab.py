"""Host-side mirror of the lab's operator interface (SRM/include/lab.h:9-10).

``order_array(keys)`` / ``order_with_trust(keys)`` take a host int32 array, sort it IN PLACE on
the GPU and block until the result is back -- the same argument meaning as the reference's
``order_array(int* srcCpu, int length)`` (SRM/lab.cu:303-402) and ``order_with_trust``
(SRM/lab.cu:404-406).  Errors raise ``B200SortError`` instead of calling ``exit`` (the exported
C++ symbols keep the reference's print-and-exit convention).

The device-array layer (what is measured against the HBM roofline) is ``radix_sort_`` /
``merge_sort_`` on torch CUDA tensors; torch only supplies device memory and streams.
"""
from __future__ import annotations

import numpy as np

from ._lib import ALGO_MERGE, ALGO_RADIX, MAX_N, check, lib


def _host_i32(keys) -> np.ndarray:
    if not isinstance(keys, np.ndarray) or keys.dtype != np.int32 or not keys.flags.c_contiguous \
            or not keys.flags.writeable or keys.ndim != 1:
        raise TypeError("keys must be a writable, contiguous, 1-D numpy int32 array (sorted in place)")
    if keys.size > MAX_N:
        raise ValueError(f"length {keys.size} exceeds {MAX_N}")
    return keys


def order_array(keys: np.ndarray, algo: int = ALGO_RADIX) -> None:
    """Sort ``keys`` (host int32) ascending in place with the onesweep radix sort."""
    k = _host_i32(keys)
    check(lib().b200sort_order_array_host(k.ctypes.data, k.size, algo))


def order_with_trust(keys: np.ndarray) -> None:
    """The drivers' second column: same contract, served by the merge sort."""
    k = _host_i32(keys)
    check(lib().b200sort_order_with_trust_host(k.ctypes.data, k.size))


# ---- device-array layer ----------------------------------------------------------------------------

class DeviceSorter:
    """Owns the scratch buffer and workspace for sorting device arrays of up to ``capacity`` keys."""

    def __init__(self, capacity: int, device=None):
        import torch
        self.torch = torch
        self.device = torch.device("cuda" if device is None else device)
        self.capacity = int(capacity)
        L = lib()
        ws = max(L.b200sort_workspace_bytes(self.capacity, ALGO_RADIX),
                 L.b200sort_workspace_bytes(self.capacity, ALGO_MERGE))
        self.tmp = torch.empty(max(self.capacity, 1), dtype=torch.int32, device=self.device)
        self.ws = torch.empty(ws + 256, dtype=torch.uint8, device=self.device)
        off = (-self.ws.data_ptr()) % 256
        self.ws_ptr = self.ws.data_ptr() + off
        self.ws_bytes = ws

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def sort_(self, keys, algo: int = ALGO_RADIX) -> None:
        """Sort the int32 CUDA tensor ``keys`` in place (stream-ordered on torch's current stream)."""
        t = self.torch
        if keys.dtype != t.int32 or not keys.is_cuda or not keys.is_contiguous() or keys.dim() != 1:
            raise TypeError("keys must be a contiguous 1-D int32 CUDA tensor")
        if keys.numel() > self.capacity:
            raise ValueError("keys longer than this sorter's capacity")
        check(lib().b200sort_sort_i32(algo, keys.data_ptr(), self.tmp.data_ptr(), keys.numel(),
                                      self.ws_ptr, self.ws_bytes, self._stream()))

    def radix_sort_(self, keys) -> None:
        self.sort_(keys, ALGO_RADIX)

    def merge_sort_(self, keys) -> None:
        self.sort_(keys, ALGO_MERGE)
