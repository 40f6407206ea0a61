"""Test infrastructure only -- never imported by the product path.

Imports the *unmodified* reference package (pure python, read from
``/root/reference``) with its three native extension modules resolved to the
binaries that ``oracle/build_ref.sh`` compiled into ``oracle/_ref/``.

Only works where ``/root/reference`` exists (the build container); on the GPU
box ``load_reference()`` returns ``None`` and callers fall back to the committed
golden fixtures under ``tests/golden`` and to the numpy port in
``oracle/tensor_ops.py`` (which is pinned against this reference here).
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("REF_ROOT", "/root/reference")
REF_BIN = os.path.join(_HERE, "_ref")
_EXT_NAMES = ("interputils_cython", "tree_ext", "simplex_helpers")
_PREFIX = "discretize._extensions."


class _RefExtFinder(importlib.abc.MetaPathFinder):
    """Resolve discretize._extensions.<native> to oracle/_ref/<native>.so."""

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_PREFIX):
            return None
        short = fullname[len(_PREFIX):]
        if short not in _EXT_NAMES:
            return None
        suffix = sysconfig.get_config_var("EXT_SUFFIX")
        so = os.path.join(REF_BIN, short + suffix)
        if not os.path.exists(so):
            return None
        loader = importlib.machinery.ExtensionFileLoader(fullname, so)
        return importlib.util.spec_from_file_location(fullname, so, loader=loader)


_cached = None


def reference_available() -> bool:
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return os.path.isdir(os.path.join(REF_ROOT, "discretize")) and all(
        os.path.exists(os.path.join(REF_BIN, n + suffix)) for n in _EXT_NAMES
    )


def load_reference():
    """Return the reference ``discretize`` module, or None when unavailable."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        return None
    if not any(isinstance(f, _RefExtFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefExtFinder())
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import discretize  # noqa: F401  (the reference)

    _cached = discretize
    return discretize
