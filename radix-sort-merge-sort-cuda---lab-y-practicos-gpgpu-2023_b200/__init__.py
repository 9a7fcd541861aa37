"""b200sort -- B200-native (sm_100a) radix sort and merge sort of 32-bit signed keys behind the
GPGPU-2023 lab's operator boundary (``order_array`` / ``order_with_trust``,
/root/reference/Sord Radix y Merge/include/lab.h:9-10).

The product is ``libb200sort.so`` (hand-written CUDA + a C-ABI, see include/b200sort.h); this
package is the thin host-side mirror of that interface used by the tests, the drivers and
bench.py.  There is no CPU fallback anywhere: every sort call fails loudly without the CUDA
library or without an sm_100 device.
"""
from ._lib import (ALGO_LAB, ALGO_MERGE, ALGO_RADIX, B200SortError, lib, lib_path, check)  # noqa: F401
from .lab import order_array, order_with_trust  # noqa: F401
from . import datagen  # noqa: F401

__all__ = ["order_array", "order_with_trust", "lib", "lib_path", "B200SortError",
           "ALGO_RADIX", "ALGO_MERGE", "ALGO_LAB", "datagen"]
