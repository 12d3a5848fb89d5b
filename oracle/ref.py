"""oracle/ref.py -- TEST INFRASTRUCTURE.  Loader for the compiled, unmodified reference module
(`oracle/_ref/fast_sampler.so`, built from /root/reference/fast_sampler by oracle/build_ref.sh).

The module is loaded under its own name (pybind requires ``PyInit_fast_sampler``) but is NOT
registered in ``sys.modules``, so it never shadows the product's ``fast_sampler`` shim.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def ref_path(nopin: bool = False) -> str:
    return os.path.join(_HERE, "_ref", "nopin" if nopin else "", "fast_sampler.so")


def available(nopin: bool = False) -> bool:
    return os.path.exists(ref_path(nopin))


def load_reference(nopin: bool = False):
    """Returns the reference pybind module.  ``nopin=True`` gives the variant whose hard-wired
    ``pinned_memory(true)`` literals are switched off so that the distributed Session runs
    without a CUDA driver (fixture generation in the build container only)."""
    key = bool(nopin)
    if key not in _cache:
        import torch  # noqa: F401  (libtorch must be loaded first)
        path = ref_path(nopin)
        loader = importlib.machinery.ExtensionFileLoader("fast_sampler", path)
        spec = importlib.util.spec_from_file_location("fast_sampler", path, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _cache[key] = mod
    return _cache[key]
