"""
GPU tests of the ingest hand-off: the drop-in CLIs must write byte-identical outputs whether the records come
from the C++ block reader (vcf_ingest.NativeVCF, the default) or from the pure-Python text reader
(TRTOOLS_B200_INGEST=python).  The golden-output CLI tests (test_gpu_statstr / test_gpu_dumpstr / test_gpu_assoc)
already run through the C++ reader; this file pins the two readers against each other on whole files.
"""
import os
import warnings

import numpy as np
import pytest

from test_gpu_dumpstr import dump_args
from test_gpu_statstr import _statstr_args

pytestmark = pytest.mark.gpu


def _both_modes(monkeypatch, fn):
    outs = []
    for mode in ("native", "python"):
        monkeypatch.setenv("TRTOOLS_B200_INGEST", mode)
        outs.append(fn(mode))
    return outs


@pytest.mark.parametrize("vcf,vcftype", [("trio_chr21_hipstr.sorted.vcf.gz", "hipstr"), ("many_samples.vcf.gz", "hipstr"),
                                         ("test_gangstr_head.vcf", "gangstr"), ("test_ExpansionHunter.vcf", "eh")])
def test_statstr_outputs_identical_for_both_readers(data_dir, tmp_path, monkeypatch, vcf, vcftype):
    from trtools_b200 import statSTR, cyvcf2_compat
    from trtools_b200.vcf_ingest import NativeVCF

    def run(mode):
        assert (cyvcf2_compat.VCF is NativeVCF) == (mode == "native")
        out = str(tmp_path / mode)
        assert statSTR.main(_statstr_args(os.path.join(data_dir, vcf), out, vcftype=vcftype, precision=6)) == 0
        return open(out + ".tab").read()

    a, b = _both_modes(monkeypatch, run)
    assert a == b and a.count("\n") >= 2


@pytest.mark.parametrize("vcf,kw", [
    ("trio_chr21_hipstr.sorted.vcf.gz", dict(vcftype="hipstr", hipstr_min_call_DP=10, hipstr_max_call_DP=1000,
                                             hipstr_min_call_Q=0.9, hipstr_max_call_flank_indel=0.15,
                                             hipstr_max_call_stutter=0.15, hipstr_min_supp_reads=2,
                                             min_locus_callrate=0.5, min_locus_hwep=1e-4, filter_hrun=True)),
    ("many_samples.vcf.gz", dict(vcftype="hipstr", hipstr_min_call_DP=15, hipstr_min_call_Q=0.95, min_locus_het=0.1,
                                 use_length=True, drop_filtered=True)),
    ("test_gangstr_head.vcf", dict(vcftype="gangstr", gangstr_min_call_DP=10, gangstr_min_call_Q=0.9,
                                   gangstr_expansion_prob_het=0.8, gangstr_filter_span_only=True)),
])
def test_dumpstr_outputs_identical_for_both_readers(data_dir, tmp_path, monkeypatch, vcf, kw):
    from trtools_b200 import dumpSTR

    def run(mode):
        out = str(tmp_path / mode)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert dumpSTR.main(dump_args(out, os.path.join(data_dir, vcf), **kw)) == 0
        return [open(out + ext).read() for ext in (".vcf", ".samplog.tab", ".loclog.tab")]

    a, b = _both_modes(monkeypatch, run)
    assert a == b and len(a[0]) > 1000


def test_synthetic_vcf_text_end_to_end(tmp_path):
    """BASELINE config 2's shape in miniature, as VCF text: file -> C++ reader -> GPU block -> statistics equal the
    statistics of the generator's own arrays uploaded directly."""
    from trtools_b200 import _lib, block as _block, synth
    from trtools_b200.vcf_ingest import NativeVCF
    L, S = 96, 3000
    loci = synth.make_loci(L)
    calls = synth.fill_calls(loci, S)
    path = str(tmp_path / "synth.vcf")
    synth.write_vcf(path, loci, calls)
    ctx = _lib.default_context()
    v = NativeVCF(path)
    v._prefetch = ("DP", "DFLANKINDEL", "Q")
    v._native_block_loci = L
    recs = list(v)
    blk = _block.build_block(ctx, "hipstr", recs, ("DP", "DFLANKINDEL", "Q"))
    assert np.array_equal(blk.gt, calls.gt)
    assert blk.gt_packed is not None and blk.gt_packed[0].base is not None     # the reader's packed slab itself
    st = blk.stats(False)
    # the same block from the generator's arrays, through the record-free upload path
    metas = [_block.record_meta("hipstr", r) for r in recs]
    blk2 = _block.Block(ctx, "hipstr", metas, calls.gt.copy(), {})
    st2 = blk2.stats(False)
    for k in st:
        a, b = np.asarray(st[k]), np.asarray(st2[k])
        assert a.shape == b.shape and np.array_equal(a, b, equal_nan=(a.dtype.kind == 'f')), k
