"""Seeded, recipe-named key generators shared by the tests, the drivers and bench.py.

Every distribution BASELINE.json's configs name is here with an exact recipe, so a result line
can quote ``name(seed)`` and be reproduced.  Pure numpy (host side); no GPU needed.

The reference generates its inputs with glibc ``rand() % 100`` (SRM/main.cpp:7-14) and
``rand() % 1000`` (SRM/performanceTest.cpp:34-37); ``lab_rand`` replays those through libc.
"""
from __future__ import annotations

import ctypes
import ctypes.util

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(n: int, seed: int, stream: int = 0) -> np.ndarray:
    """n 64-bit words of the splitmix64 sequence started at ``seed`` (stream offsets the state)."""
    with np.errstate(over="ignore"):
        base = np.uint64((seed * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF)
        z = base + (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform_u32(n: int, seed: int, stream: int = 0) -> np.ndarray:
    return (splitmix64(n, seed, stream) >> np.uint64(32)).astype(np.uint32)


def uniform(n: int, seed: int = 1) -> np.ndarray:
    """Uniform over the full signed 32-bit range (config 1, 2(i), 3)."""
    return uniform_u32(n, seed).view(np.int32)


def uniform_nonneg(n: int, seed: int = 1) -> np.ndarray:
    """Uniform over [0, 2^31): the sign domain of the reference's own inputs."""
    return (uniform_u32(n, seed) >> np.uint32(1)).view(np.int32)


def and_k(n: int, seed: int = 1, k: int = 3) -> np.ndarray:
    """Low entropy: bitwise AND of k independent uniform words (config 2(ii)); each bit is set
    with probability 2^-k."""
    w = uniform_u32(n, seed, 0)
    for s in range(1, k):
        w &= uniform_u32(n, seed, s)
    return w.view(np.int32)


def masked(n: int, seed: int = 1, mask: int = 0x0000FFFF) -> np.ndarray:
    """Low entropy: uniform word AND a fixed mask (config 2(ii)); constant digits let whole
    radix passes be skipped."""
    return (uniform_u32(n, seed) & np.uint32(mask)).view(np.int32)


def skewed(n: int, seed: int = 1, hot: float = 0.9) -> np.ndarray:
    """Skewed (config 4): a fraction ``hot`` of the keys share one top byte (0x40), the rest are
    uniform over the full range."""
    w = uniform_u32(n, seed, 0)
    pick = uniform_u32(n, seed, 1) < np.uint32(int(hot * 0xFFFFFFFF))
    w = np.where(pick, (w & np.uint32(0x00FFFFFF)) | np.uint32(0x40000000), w)
    return w.astype(np.uint32).view(np.int32)


def zipf16(n: int, seed: int = 1) -> np.ndarray:
    """Skewed (config 4): Zipf(s=1) over 2^16 distinct values, each value hashed to an int32."""
    u = (splitmix64(n, seed) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    r = np.floor(np.exp(u * np.log(65536.0))).astype(np.uint64)  # P(r) ~ 1/r on [1, 65536]
    with np.errstate(over="ignore"):
        z = r * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(29))) * np.uint64(0xBF58476D1CE4E5B9)
    return (z >> np.uint64(32)).astype(np.uint32).view(np.int32)


def ascending(n: int, seed: int = 0) -> np.ndarray:
    """Already sorted (config 4): i - n/2, so the run crosses zero."""
    return (np.arange(n, dtype=np.int64) - n // 2).astype(np.int32)


def descending(n: int, seed: int = 0) -> np.ndarray:
    """Reverse sorted (config 4)."""
    return ascending(n)[::-1].copy()


def all_equal(n: int, seed: int = 7) -> np.ndarray:
    return np.full(n, np.int32(seed), dtype=np.int32)


def edge_mix(n: int, seed: int = 1) -> np.ndarray:
    """INT_MIN / -1 / 0 / 1 / INT_MAX and neighbours, shuffled by the seed."""
    vals = np.array([-2**31, -2**31 + 1, -2, -1, 0, 1, 2, 2**31 - 2, 2**31 - 1], dtype=np.int64)
    idx = (splitmix64(n, seed) % np.uint64(len(vals))).astype(np.int64)
    return vals[idx].astype(np.int32)


def lab_rand(n: int, mod: int = 100, seed: int | None = 1) -> np.ndarray:
    """glibc ``rand() % mod`` exactly as SRM/main.cpp:10 (mod 100) and
    SRM/performanceTest.cpp:35 (mod 1000) draw it.  ``seed=None`` continues the current libc
    sequence (performanceTest never calls srand and draws its sizes back to back)."""
    libc = ctypes.CDLL(ctypes.util.find_library("c") or "libc.so.6")
    libc.rand.restype = ctypes.c_int
    if seed is not None:
        libc.srand(ctypes.c_uint(seed))
    out = np.empty(n, dtype=np.int32)
    rand = libc.rand
    for i in range(n):
        out[i] = rand() % mod
    return out


DISTRIBUTIONS = {
    "uniform": uniform,
    "uniform_nonneg": uniform_nonneg,
    "and2": lambda n, seed=1: and_k(n, seed, 2),
    "and3": lambda n, seed=1: and_k(n, seed, 3),
    "and4": lambda n, seed=1: and_k(n, seed, 4),
    "mask_0000ffff": lambda n, seed=1: masked(n, seed, 0x0000FFFF),
    "mask_00ff00ff": lambda n, seed=1: masked(n, seed, 0x00FF00FF),
    "skewed90": skewed,
    "zipf16": zipf16,
    "ascending": ascending,
    "descending": descending,
    "all_equal": all_equal,
    "edge_mix": edge_mix,
}


def make(name: str, n: int, seed: int = 1) -> np.ndarray:
    """Keys of distribution ``name`` (see DISTRIBUTIONS; also ``lab_rand100`` / ``lab_rand1000``)."""
    if name == "lab_rand100":
        return lab_rand(n, 100, seed)
    if name == "lab_rand1000":
        return lab_rand(n, 1000, seed)
    return np.ascontiguousarray(DISTRIBUTIONS[name](n, seed))
