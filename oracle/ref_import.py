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

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the read-only source tree (build container) or the unmodified copy installed by baseline/install_ref.py
# (git-ignored, travels to the GPU box; used ONLY by bench.py --impl reference)
INSTALLED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _pick_root():
    env = os.environ.get("TRTOOLS_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", INSTALLED_ROOT]:
        if cand and os.path.isdir(os.path.join(cand, "trtools")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    """The reference tree WITH its test fixtures (build container only): what the needs_reference tests want."""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "trtools", "testsupport"))


def reference_code_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "trtools"))


def enable() -> bool:
    """Put the reference and any needed shims on sys.path. Returns availability."""
    if not reference_code_available():
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
