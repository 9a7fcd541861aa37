"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/*.h declares.
No compute calls here (there is no GPU): only loading, symbol lookup and argument validation."""
import ctypes
import os
import re

import numpy as np
import pytest

import b200sort
from b200sort import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header: str) -> set[str]:
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(b200sort_[a-z0-9_]+)\s*\(", text))


def test_library_is_built_in_tree():
    assert os.path.exists(b200sort.lib_path()), "run `make` (or __graft_entry__.build())"
    assert os.path.dirname(b200sort.lib_path()).startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared_functions("b200sort.h")
    assert len(declared) >= 25
    handle = ctypes.CDLL(b200sort.lib_path())
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/b200sort.h but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))


def test_lab_h_cxx_symbols_are_exported():
    """include/lab.h: order_array(int*,int), order_with_trust(int*,int) with C++ linkage, the exact
    manglings the reference's drivers bind (SRM/include/lab.h:9-10)."""
    handle = ctypes.CDLL(b200sort.lib_path())
    for name in ("_Z11order_arrayPii", "_Z16order_with_trustPii"):
        assert hasattr(handle, name)
    text = open(os.path.join(ROOT, "include", "lab.h")).read()
    assert "void order_array(int *srcCpu, int length);" in text
    assert "void order_with_trust(int *src, int length);" in text


def test_no_thrust_cub_or_oracle_in_the_product():
    pkg = os.path.dirname(b200sort.lib_path())
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")):
                src = open(os.path.join(dirpath, f)).read()
                assert "#include <thrust" not in src and "#include <cub" not in src, f
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_status_strings_and_validation_without_a_gpu():
    L = b200sort.lib()
    assert L.b200sort_version().startswith(b"b200sort")
    assert L.b200sort_status_string(0) == b"ok"
    nv = L.b200sort_radix_num_variants()
    assert nv >= 4                                   # default, small-array shape, ballot fallback, default at any size
    names = [L.b200sort_radix_variant_name(v) for v in range(nv)]
    assert all(names) and len(set(names)) == nv and L.b200sort_radix_variant_name(nv) is None
    assert b"kRankBallot" in names[2]                # the documented-behaviour fallback keeps its index
    assert L.b200sort_radix_set_variant(-1) == 1 and L.b200sort_radix_set_variant(nv) == 1
    assert L.b200sort_radix_set_variant(0) == 0
    assert L.b200sort_radix_tile() >= 2048 and L.b200sort_block_sort_tile() >= 1024
    # argument validation happens before any device work
    assert L.b200sort_radix_i32(None, None, 10, None, 0, None) == 1            # NULL keys
    assert L.b200sort_radix_i32(None, None, (1 << 30) + 1, None, 0, None) == 1  # n too large
    assert L.b200sort_sort_i32(7, 8, 8, 10, 8, 1 << 20, None) == 1             # unknown algorithm
    assert L.b200sort_radix_i32(8, 8, 10, None, 0, None) == 2                  # no workspace
    assert L.b200sort_radix_i32(None, None, 0, None, 0, None) == 0             # empty input is fine
    assert L.b200sort_merge_i32(None, None, 1, None, 0, None) == 0
    for algo in (0, 1):
        w1, w2 = L.b200sort_workspace_bytes(1 << 20, algo), L.b200sort_workspace_bytes(1 << 24, algo)
        assert 0 < w1 <= w2


def test_host_operator_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    keys = np.array([3, 1, 2], dtype=np.int32)
    with pytest.raises(b200sort.B200SortError) as e:
        b200sort.order_array(keys)
    assert e.value.status == 4            # B200SORT_ERR_NO_DEVICE: no CPU fallback
    assert keys.tolist() == [3, 1, 2]
    with pytest.raises(TypeError):
        b200sort.order_array(np.array([1.0, 2.0]))
