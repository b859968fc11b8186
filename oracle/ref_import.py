"""
TEST INFRASTRUCTURE ONLY.

Make the UNMODIFIED reference (``/root/reference/trtools``) importable in the
build container behind the shims in ``oracle/shims`` (cyvcf2 / statsmodels /
matplotlib / pysam are not installed here).  Used by
``tests/golden/make_golden.py`` and by the ``needs_reference`` tests; it is a
no-op-with-False on the GPU box, where ``/root/reference`` does not exist.
"""
import importlib.util
import os
import sys

REFERENCE_ROOT = os.environ.get("TRTOOLS_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "trtools"))


def enable() -> bool:
    """Put the reference and any needed shims on sys.path. Returns availability."""
    if not reference_available():
        return False
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)
    for name in ("cyvcf2", "statsmodels", "matplotlib", "pysam"):
        if name in sys.modules:
            continue
        shim_dir = os.path.join(_SHIMS, name)
        # step aside if the real package is importable
        saved = list(sys.path)
        try:
            sys.path = [p for p in sys.path if os.path.abspath(p) != _SHIMS]
            real = importlib.util.find_spec(name) is not None
        except (ImportError, ValueError):
            real = False
        finally:
            sys.path = saved
        if not real and os.path.isdir(shim_dir) and _SHIMS not in sys.path:
            sys.path.insert(0, _SHIMS)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return True
