"""ctypes front end of the CPU oracle (oracle/oracle_sort.c).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package never imports this module.

``ref`` (below) wraps the reference's own code, compiled unmodified from
/root/reference/Sord Radix y Merge/lab.cu into oracle/_ref/libreflab.so by ``make -C oracle ref``:
``ref.order_with_trust`` is the reference's CPU path (SRM/lab.cu:404-406) and runs anywhere;
``ref.order_array`` is the reference's GPU path (SRM/lab.cu:303-402) and needs a GPU.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libreflab.so")


def build(ref: bool = True) -> None:
    """Compile liboracle.so (gcc) and, when /root/reference is present, oracle/_ref/."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    if ref:
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def _load() -> ctypes.CDLL:
    if not os.path.exists(_LIB_PATH):
        build(ref=False)
    lib = ctypes.CDLL(_LIB_PATH)
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.oracle_split_tile32.argtypes = [vp, i32]
    lib.oracle_split_tile32.restype = i32
    lib.oracle_rank.argtypes = [vp, sz, ctypes.c_int32, i32]
    lib.oracle_rank.restype = sz
    lib.oracle_rank_merge.argtypes = [vp, sz, vp, sz, vp]
    lib.oracle_rank_merge.restype = None
    lib.oracle_order_array.argtypes = [vp, sz]
    lib.oracle_order_array.restype = i32
    lib.oracle_radix_sort_i32.argtypes = [vp, sz]
    lib.oracle_radix_sort_i32.restype = i32
    lib.oracle_digit_histograms.argtypes = [vp, sz, vp]
    lib.oracle_digit_histograms.restype = None
    lib.oracle_radix_pass.argtypes = [vp, vp, sz, i32]
    lib.oracle_radix_pass.restype = None
    lib.oracle_merge_path.argtypes = [vp, sz, vp, sz, sz]
    lib.oracle_merge_path.restype = sz
    lib.oracle_is_sorted.argtypes = [vp, sz]
    lib.oracle_is_sorted.restype = i32
    lib.oracle_multiset_fingerprint.argtypes = [vp, sz, vp]
    lib.oracle_multiset_fingerprint.restype = None
    return lib


_lib = _load()


def _i32(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


def split_tile32(tile) -> tuple[np.ndarray, int]:
    """SRM/lab.cu:47-87 on one tile (len <= 32): (sorted tile, split iterations run)."""
    t = _i32(tile).copy()
    assert t.size <= 32
    it = _lib.oracle_split_tile32(t.ctypes.data, t.size)
    return t, it


def rank(run, x: int, before_equals: bool) -> int:
    """SRM/lab.cu:102-132."""
    r = _i32(run)
    return int(_lib.oracle_rank(r.ctypes.data, r.size, int(x), 1 if before_equals else 0))


def rank_merge(a, b) -> np.ndarray:
    """SRM/lab.cu:144-182."""
    a, b = _i32(a), _i32(b)
    out = np.empty(a.size + b.size, dtype=np.int32)
    _lib.oracle_rank_merge(a.ctypes.data, a.size, b.ctypes.data, b.size, out.ctypes.data)
    return out


def order_array(keys) -> np.ndarray:
    """SRM/lab.cu:303-402 restated: returns the sorted copy."""
    k = _i32(keys).copy()
    if _lib.oracle_order_array(k.ctypes.data, k.size) != 0:
        raise MemoryError("oracle_order_array")
    return k


def radix_sort(keys) -> np.ndarray:
    """The library-sort leg (SRM/lab.cu:404-406), signed order: returns the sorted copy."""
    k = _i32(keys).copy()
    if _lib.oracle_radix_sort_i32(k.ctypes.data, k.size) != 0:
        raise MemoryError("oracle_radix_sort_i32")
    return k


def radix_sort_inplace(keys: np.ndarray) -> None:
    assert keys.dtype == np.int32 and keys.flags.c_contiguous
    if _lib.oracle_radix_sort_i32(keys.ctypes.data, keys.size) != 0:
        raise MemoryError("oracle_radix_sort_i32")


def digit_histograms(keys) -> np.ndarray:
    k = _i32(keys)
    h = np.zeros(4 * 256, dtype=np.uint64)
    _lib.oracle_digit_histograms(k.ctypes.data, k.size, h.ctypes.data)
    return h.reshape(4, 256)


def radix_pass(keys, digit_pass: int) -> np.ndarray:
    k = _i32(keys)
    out = np.empty_like(k)
    _lib.oracle_radix_pass(k.ctypes.data, out.ctypes.data, k.size, int(digit_pass))
    return out


def sort_pairs(keys, vals) -> tuple[np.ndarray, np.ndarray]:
    """(key, value) pairs ordered by key, ascending signed, STABLE: equal keys keep their input order.
    That is the reference's tie rule (A's equal elements before B's, SRM/lab.cu:163-170) applied at every
    merge level of SRM/lab.cu:303-402, i.e. what carrying the "positions" array of the doc comment at
    SRM/lab.cu:44-45 through the lab pipeline would return.  numpy's stable argsort states it directly."""
    k, v = _i32(keys), _i32(vals)
    assert k.shape == v.shape
    order = np.argsort(k, kind="stable")
    return k[order], v[order]


def merge_path(a, b, diag: int) -> int:
    a, b = _i32(a), _i32(b)
    return int(_lib.oracle_merge_path(a.ctypes.data, a.size, b.ctypes.data, b.size, int(diag)))


def is_sorted(keys) -> bool:
    k = _i32(keys)
    return bool(_lib.oracle_is_sorted(k.ctypes.data, k.size))


def multiset_fingerprint(keys) -> tuple[int, int, int]:
    k = _i32(keys)
    out = np.zeros(3, dtype=np.uint64)
    _lib.oracle_multiset_fingerprint(k.ctypes.data, k.size, out.ctypes.data)
    return tuple(int(v) for v in out)


class _Ref:
    """The reference itself (oracle/_ref/libreflab.so); C++-mangled symbols of SRM/include/lab.h."""

    def __init__(self) -> None:
        self._lib = None

    @property
    def available(self) -> bool:
        return os.path.exists(_REF_PATH)

    def _get(self) -> ctypes.CDLL:
        if self._lib is None:
            if not self.available:
                raise FileNotFoundError(
                    f"{_REF_PATH} missing: run `make -C oracle ref` where /root/reference exists")
            lib = ctypes.CDLL(_REF_PATH)
            for name in ("_Z11order_arrayPii", "_Z16order_with_trustPii"):
                fn = getattr(lib, name)
                fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
                fn.restype = None
            self._lib = lib
        return self._lib

    def order_with_trust(self, keys) -> np.ndarray:
        """Reference CPU path, SRM/lab.cu:404-406 (Thrust sequential host sort)."""
        k = _i32(keys).copy()
        self._get()._Z16order_with_trustPii(k.ctypes.data, k.size)
        return k

    def order_with_trust_inplace(self, keys: np.ndarray) -> None:
        assert keys.dtype == np.int32 and keys.flags.c_contiguous
        self._get()._Z16order_with_trustPii(keys.ctypes.data, keys.size)

    def order_array(self, keys) -> np.ndarray:
        """Reference GPU path, SRM/lab.cu:303-402.  Needs a GPU; exits the process on CUDA
        errors (SRM/include/utils.h:19-25); launchable for n <= 2^17 only; hangs on mixed signs."""
        k = _i32(keys).copy()
        self._get()._Z11order_arrayPii(k.ctypes.data, k.size)
        return k


ref = _Ref()
