"""CPU test: the C-ABI shared library loads and exports every symbol include/trtools_b200.h declares
(no compute call is made — there is no GPU in the build container)."""
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(REPO, "include", "trtools_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trt_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from trtools_b200 import build, _lib
    build.build()
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "libtrtools_b200.so does not export " + name
    assert sorted(_lib.EXPORTS) == declared, set(_lib.EXPORTS) ^ set(declared)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path fails loudly instead of falling back."""
    from trtools_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.TrtError) as e:
        _lib.Context(0)
    assert e.value.code == _lib.TRT_ENODEV
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under trtools_b200/ may import it."""
    pkg = os.path.join(REPO, "trtools_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
