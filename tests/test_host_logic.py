"""
CPU tests of the host side of the drop-in (no GPU, no compute calls): caller detection and its error behaviour
(reference trtools/utils/tr_harmonizer.py:180-262), INFO validation of the per-caller harmonizers (:303-550), the
cyvcf2 decode conventions the ingest stage relies on, the utils known answers
(trtools/utils/tests/test_utils.py) and the locus sharding used by the multi-GPU path.
"""
import os

import numpy as np
import pytest

from trtools_b200 import block, cyvcf2_compat, tr_harmonizer as trh, utils


def _vcf(data_dir, name):
    return cyvcf2_compat.VCF(os.path.join(data_dir, name))


@pytest.mark.parametrize("fname,want,auto_ok", [
    ("test_gangstr_head.vcf", "gangstr", True), ("test_advntr.vcf", "advntr", True), ("test_popstr.vcf", "popstr", True),
    ("test_ExpansionHunter.vcf", "eh", True), ("many_samples.vcf.gz", "hipstr", True),
    ("test_longtr.vcf", "longtr", False),     # its header also names HipSTR: the reference's tests pass vcftype='longtr' too
])
def test_infer_vcftype_from_header(data_dir, fname, want, auto_ok):
    v = _vcf(data_dir, fname)
    if auto_ok:
        assert trh.InferVCFType(v).name == want
    else:
        with pytest.raises(TypeError):
            trh.InferVCFType(v)
    assert trh.InferVCFType(v, want).name == want            # a matching user choice is accepted
    other = "gangstr" if want != "gangstr" else "advntr"
    with pytest.raises(TypeError):                            # a contradicting one is not (:236-244)
        trh.InferVCFType(v, other)
    with pytest.raises(ValueError):                           # unknown type string (:49-66)
        trh.InferVCFType(v, "nonsense")


def test_infer_vcftype_ambiguous_header_needs_user_choice(data_dir):
    # the trio file's header mentions both HipSTR and a GangSTR-named bed: 'auto' must refuse (:228-235)
    v = _vcf(data_dir, "trio_chr21_hipstr.sorted.vcf.gz")
    with pytest.raises(TypeError):
        trh.InferVCFType(v)
    assert trh.InferVCFType(v, "hipstr") == trh.VcfTypes.hipstr
    assert not trh.IsBeagleVCF(v)


def test_vcftype_predicates():
    assert trh.MayHaveImpureRepeats("hipstr") and not trh.MayHaveImpureRepeats("gangstr")
    assert trh.HasLengthRefGenotype("eh") and not trh.HasLengthRefGenotype("popstr")
    assert trh.HasLengthAltGenotypes("popstr") and trh.HasLengthAltGenotypes("eh") and not trh.HasLengthAltGenotypes("advntr")
    with pytest.raises(ValueError):
        trh.MayHaveImpureRepeats("nonsense")


class _Rec:
    """Minimal cyvcf2.Variant look-alike (the reference's DummyCyvcf2Record, test_trharmonizer.py:18-50)."""

    def __init__(self, info, ref="ACACAC", alt=("ACAC",), pos=100, rid=None):
        self.CHROM, self.POS, self.ID, self.REF, self.ALT, self.INFO = "1", pos, rid, ref, list(alt), info
        self.FILTER, self.FORMAT = None, ["GT"]


@pytest.mark.parametrize("vcftype,info", [
    ("hipstr", {"START": 100, "END": 105}),                  # PERIOD missing (:349-354)
    ("longtr", {"START": 100, "PERIOD": 2}),
    ("gangstr", {}),                                          # RU missing (:316-321)
    ("gangstr", {"RU": "ac", "VID": "x"}),                   # an adVNTR record fed to the GangSTR harmonizer (:322-332)
    ("advntr", {"RU": "ac"}),                                 # VID missing (:424-428)
    ("popstr", {}),                                           # Motif missing (:486-489)
    ("eh", {"RU": "ac", "RL": 6}),                            # VARID missing (:528-532)
])
def test_missing_mandatory_info_raises_typeerror_before_any_device_work(vcftype, info):
    with pytest.raises(TypeError):
        block.record_meta(vcftype, _Rec(info))


def test_record_meta_scalars():
    m = block.record_meta("hipstr", _Rec({"START": 102, "END": 105, "PERIOD": 2}, rid="STR_1"))
    assert (m.start, m.end, m.period, m.record_id, m.quality_field, m.harmonized_pos) == (102, 105, 2, "STR_1", "Q", 102)
    m = block.record_meta("hipstr", _Rec({"START": 100, "END": 105, "PERIOD": 2, "IMP": True}))
    assert m.quality_field is None                            # Beagle-imputed: no Q (:406)
    m = block.record_meta("gangstr", _Rec({"RU": "ac"}))
    assert m.period == 2 and m.quality_field == "Q"


def test_cyvcf2_decode_conventions(data_dir):
    """genotype.array(): int16 [S, P+1], -1 = '.', -2 = ploidy pad, last column phased; Integer FORMAT -> int32 with
    INT32_MIN missing; Float FORMAT -> float32 with NaN missing; FILTER None for PASS and '.'."""
    assert cyvcf2_compat.parse_gt_column(["0|1", "1/1", ".", "./.", "2", "0/."]).tolist() == [
        [0, 1, 1], [1, 1, 0], [-1, -2, 0], [-1, -1, 0], [2, -2, 0], [0, -1, 0]]
    v = _vcf(data_dir, "many_samples.vcf.gz")
    rec = next(iter(v))
    gt = rec.genotype.array()
    assert gt.dtype == np.int16 and gt.shape == (len(v.samples), rec.ploidy + 1)
    dp = rec.format("DP")
    assert dp.dtype == np.int32 and dp.shape == (len(v.samples), 1)
    q = rec.format("Q")
    assert q.dtype == np.float32
    assert rec.FILTER is None
    assert isinstance(rec.INFO.get("START"), int) and rec.INFO.get("NOT_THERE") is None
    with pytest.raises(KeyError):
        rec.format("NOT_A_FIELD")


def test_utils_known_answers():
    """trtools/utils/tests/test_utils.py:17-100."""
    assert utils.GetHeterozygosity({0: 1}) == 0 and utils.GetHeterozygosity({0: 0.5, 1: 0.5}) == 0.5
    assert np.isnan(utils.GetHeterozygosity({0: 0.5, 1: 0.4}))        # ValidateAlleleFreqs tolerance 1e-3
    assert utils.GetMean({0: 0.5, 1: 0.5}) == 0.5 and utils.GetMode({0: 0.5, 1: 0.5}) == 0
    assert utils.GetVariance({0: 1}) == 0 and utils.GetVariance({0: 0.5, 1: 0.5}) == 0.25
    assert abs(utils.GetEntropy({0: 0.5, 1: 0.5}) - 1.0) < 1e-12
    assert utils.GetHomopolymerRun("AATAAAAAAAT") == 7 and utils.GetHomopolymerRun("") == 0
    assert utils.GetCanonicalOneStrand("TGCA") == "ATGC"
    assert utils.FabricateAllele("ACG", 2.34) == "ACGACGA" and utils.FabricateAllele("AC", 3.5) == "ACACAC"   # utils.py:566-602 (values from the reference)
    gtc = {(0, 1): 10, (0, 0): 20, (1, 1): 70}
    p = utils.GetHardyWeinbergBinomialTest({0: 0.25, 1: 0.75}, gtc)
    assert 0 <= p < 1e-5
    assert np.isnan(utils.GetHardyWeinbergBinomialTest({0: 0.5, 1: 0.5}, {(0, 3): 4}))   # allele not in the freqs


def test_locus_shard_covers_block_in_vcf_order():
    from trtools_b200 import dist
    for L, W in [(100000, 8), (7, 3), (2, 4), (0, 2)]:
        spans = [dist.locus_shard(L, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == L
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
