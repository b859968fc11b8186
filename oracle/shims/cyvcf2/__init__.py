"""
TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Stand-in for the third-party ``cyvcf2`` package (pinned >=0.30.27 in the
reference's pyproject.toml:19; not installed in this image) so that the
UNMODIFIED reference under /root/reference can be imported in this container
to (a) validate the oracle restatement and (b) generate the golden vectors
under tests/golden/.  It simply re-exports the text-VCF reader of the
product package; if the real cyvcf2 is importable this directory must not be
put on sys.path (oracle/ref_import.py checks).
"""
from trtools_b200.cyvcf2_compat import VCF, Variant, Writer  # noqa: F401

__version__ = "0.0-shim"
