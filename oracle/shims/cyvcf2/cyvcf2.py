"""TEST INFRASTRUCTURE ONLY: ``from cyvcf2 import cyvcf2`` (trtools/utils/mergeutils.py:10) names the extension module
of the real package; here it re-exports the shim's classes."""
from . import VCF, Variant, Writer  # noqa: F401
