"""
TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Stand-in for the third-party ``cyvcf2`` package (pinned >=0.30.27 in the
reference's pyproject.toml:19; not installed in this image) so that the
UNMODIFIED reference under /root/reference (or its copy under baseline/_ref) can
be imported to (a) validate the oracle restatement, (b) generate the golden
vectors under tests/golden/ and (c) serve as the CPU arm of bench.py.

The reader is pinned to the pure-Python TEXT reader (``TextVCF``): the checker
must never run through the C++ ingest it is used to check.  ``VCF("mem://key")``
opens an in-memory record source registered with :func:`register_memory_vcf`
(bench.py pushes synthetic cyvcf2-layout records through the reference's own
``load_trs`` that way, without writing 50 000-sample VCF text).  If the real
cyvcf2 is importable this directory must not be put on sys.path
(oracle/ref_import.py checks).
"""
from trtools_b200.cyvcf2_compat import TextVCF, Variant, Writer  # noqa: F401

__version__ = "0.0-shim"

_MEMORY = {}


def register_memory_vcf(key, records, raw_header, samples):
    """records: list of cyvcf2.Variant look-alikes (oracle.records.LocusAsVariant)."""
    _MEMORY[key] = (list(records), str(raw_header), list(samples))


class _MemoryVCF:
    """The part of the cyvcf2.VCF surface the reference's record loops touch."""

    def __init__(self, key):
        recs, self.raw_header, self.samples = _MEMORY[key]
        self._it = iter(recs)

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._it)

    def close(self):
        pass


class VCF(TextVCF):
    """``cyvcf2.VCF``: the text reader, or the registered in-memory source for ``mem://`` names."""

    def __new__(cls, fname, *args, **kwargs):
        if isinstance(fname, str) and fname.startswith("mem://"):
            return _MemoryVCF(fname[len("mem://"):])
        return super().__new__(cls)
