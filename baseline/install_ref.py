#!/usr/bin/env python3
"""
Install the UNMODIFIED reference (gymrek-lab/TRTools, /root/reference) into baseline/_ref (git-ignored; it travels
to the GPU box with the repo snapshot) so that `bench.py --impl reference` can time the reference's own Python code
on the box's host cores.

1. Try the documented offline install:
       python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
   In this image it fails: the reference's build backend (poetry-core, pyproject.toml [build-system]) is neither
   installed nor in /opt/wheelhouse ("ModuleNotFoundError: No module named 'poetry'").
2. Fall back to what `pip --target` would have produced for this pure-Python package: the `trtools/` package tree
   (*.py only; the 79 MB of test fixtures under testsupport/ and the tests/ directories are not needed), byte for
   byte.  Nothing is edited; third-party dependencies missing from the image (cyvcf2, statsmodels, pysam,
   matplotlib) are provided at run time by oracle/shims (oracle/ref_import.py).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("TRTOOLS_REFERENCE_ROOT", "/root/reference")


def install(force=False) -> str:
    marker = os.path.join(DEST, "trtools", "__init__.py")
    if os.path.exists(marker) and not force:
        return "present"
    if not os.path.isdir(os.path.join(SRC, "trtools")):
        return "reference tree not available"
    os.makedirs(DEST, exist_ok=True)
    res = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
                          "/opt/wheelhouse", "--no-deps", "--target", DEST, SRC], capture_output=True, text=True)
    if res.returncode == 0 and os.path.exists(marker):
        return "pip"
    n = 0
    for root, dirs, files in os.walk(os.path.join(SRC, "trtools")):
        dirs[:] = [d for d in dirs if d not in ("testsupport", "tests", "__pycache__")]
        rel = os.path.relpath(root, SRC)
        for f in files:
            if f.endswith(".py"):
                os.makedirs(os.path.join(DEST, rel), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(DEST, rel, f))
                n += 1
    with open(os.path.join(DEST, "INSTALL_LOG.txt"), "w") as fh:
        fh.write("pip install failed (rc {}): {}\ncopied {} .py files of the unmodified trtools package from {}\n".format(
            res.returncode, (res.stderr or res.stdout).strip().splitlines()[-1:] or "", n, SRC))
    return "copied {} files (pip failed: no poetry build backend)".format(n)


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
