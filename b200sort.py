"""Import shim: ``import b200sort`` loads the package that lives in the directory
``radix-sort-merge-sort-cuda---lab-y-practicos-gpgpu-2023_b200/`` (a name Python cannot import
directly because of the hyphens)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "radix-sort-merge-sort-cuda---lab-y-practicos-gpgpu-2023_b200")
_spec = importlib.util.spec_from_file_location(
    "b200sort", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200sort"] = _mod
_spec.loader.exec_module(_mod)
