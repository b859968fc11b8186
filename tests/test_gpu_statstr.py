"""
GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C-ABI /
the drop-in Python API, against (a) golden vectors produced by the unmodified reference and
(b) the oracle on seeded synthetic inputs.  Integer counts bit-exact; floating-point statistics
within 1e-6 relative (BASELINE.json north_star), NaN == NaN.
"""
import argparse
import os

import numpy as np
import pytest

from helpers import assert_close, assert_close_list, REL_TOL

pytestmark = pytest.mark.gpu

FIXTURES = ["hipstr_many", "hipstr_trio", "gangstr", "popstr", "eh", "advntr", "longtr", "synth_small",
            "synth_wide", "edge"]
_cache = {}


def fixture(golden_dir, name):
    from oracle.records import load_loci
    if name not in _cache:
        _cache[name] = load_loci(os.path.join(golden_dir, name + ".npz"))
    return _cache[name]


@pytest.fixture(scope="module")
def ctx():
    from trtools_b200 import _lib
    return _lib.default_context()


def _blocks_from_loci(ctx, loci):
    """Group fixture loci into GPU blocks (same vcftype / sample count / ploidy)."""
    from oracle.records import LocusAsVariant
    from trtools_b200 import block
    runs, cur, key = [], [], None
    for i, l in enumerate(loci):
        k = (l.vcftype, None if l.gt is None else l.gt.shape)
        if cur and k != key:
            runs.append(cur)
            cur = []
        key = k
        cur.append(i)
    if cur:
        runs.append(cur)
    for run in runs:
        recs = [LocusAsVariant(loci[i]) for i in run]
        yield run, recs, loci[run[0]].vcftype


@pytest.mark.parametrize("name", FIXTURES)
def test_harmonize_kernel_matches_reference(golden_dir, ctx, name):
    from trtools_b200 import block, tr_harmonizer as trh
    loci, extra, _ = fixture(golden_dir, name)
    n_checked = 0
    for run, recs, vt in _blocks_from_loci(ctx, loci):
        refs = [extra["ref"][i] for i in run]
        if any("error" in r for r in refs):
            for rec, r in zip(recs, refs):
                if "error" in r:
                    with pytest.raises((TypeError, ValueError)) as e:
                        trh.HarmonizeRecord(vt, rec)
                    assert type(e.value).__name__ == r["error"]
            continue
        blk = block.build_block(ctx, vt, recs)
        for j, (rec, r) in enumerate(zip(recs, refs)):
            tr = trh.TRRecord._from_block(blk, j, rec)
            h = r["harm"]
            assert tr.ref_allele == h["ref_allele"], (name, run[j])
            assert tr.alt_alleles == h["alt_alleles"], (name, run[j])
            assert tr.motif == h["motif"], (name, run[j])
            assert tr.record_id == h["record_id"]
            assert tr.pos == h["pos"] and tr.end_pos == h["end_pos"]
            assert tr.ref_allele_length == h["ref_len"]          # float64 bit-exact
            assert tr.alt_allele_lengths == h["alt_lens"]
            assert tr.HasFullStringGenotypes() == h["has_full"]
            assert tr.quality_field == h["quality_field"]
            n_checked += 1
    assert n_checked > 0 or name == "edge"


def _afreq_from_ac(keys, ac, count):
    from trtools_b200.statSTR import _afreq_string
    return _afreq_string(keys, ac, count)


@pytest.mark.parametrize("name", FIXTURES)
def test_locus_stats_kernel_matches_reference(golden_dir, ctx, name):
    from trtools_b200 import block
    from trtools_b200.statSTR import _locus_keys
    loci, extra, _ = fixture(golden_dir, name)
    masks = [np.array(m, dtype=np.uint8) for m in extra.get("group_masks", [])]
    for run, recs, vt in _blocks_from_loci(ctx, loci):
        refs = [extra["ref"][i] for i in run]
        if any("error" in r for r in refs) or loci[run[0]].gt is None:
            continue
        blk = block.build_block(ctx, vt, recs)
        S = blk.S
        gm = None
        if masks and masks[0].shape[0] == S:
            gm = np.stack([np.ones(S, np.uint8)] + masks)
        for key, uselength in (("stats_len", True), ("stats_seq", False)):
            st = blk.stats(uselength, 0.01, gm)
            G = 1 if gm is None else gm.shape[0]
            for j, r in enumerate(refs):
                c = r["counts"]
                sl = blk.allele_slice(j)
                ac = st["ac"][0, sl]
                assert {str(k): int(v) for k, v in enumerate(ac) if v > 0} == c["ac_idx"], (name, run[j])
                assert int(st["n_called"][0, j]) == c["n_called"]
                assert int(st["n_called_nonstrict"][0, j]) == c["n_called_nonstrict"]
                if key not in r:
                    continue
                ref = r[key]
                keys = _locus_keys(blk, j, uselength)
                for g in range(G):
                    what = "{} locus {} {} group {}".format(name, run[j], key, g)
                    assert _afreq_from_ac(keys, st["ac"][g, sl], False) == ref["afreq"][g], what
                    assert _afreq_from_ac(keys, st["ac"][g, sl], True) == ref["acount"][g], what
                    assert int(st["nalleles"][g, j]) == ref["nalleles"][g], what
                    assert int(st["n_called"][g, j]) == ref["numcalled"][g], what
                    for stat in ("thresh", "hwep", "het", "entropy", "mean", "mode", "var"):
                        assert_close(st[stat][g, j], ref[stat][g], what + " " + stat, rel=REL_TOL, abs_tol=1e-300)


def _statstr_args(vcf, out, **kw):
    ns = argparse.Namespace(vcf=vcf, out=out, vcftype="auto", samples=None, sample_prefixes=None, region=None,
                            precision=3, nalleles_thresh=0.01, plot_afreq=False, use_length=False,
                            only_passing=False, block_size=512)
    for s in ("thresh", "afreq", "acount", "nalleles", "hwep", "het", "entropy", "mean", "mode", "var", "numcalled"):
        setattr(ns, s, True)
    for k, v in kw.items():
        setattr(ns, k, v)
    return ns


def _same_tab(got: str, want: str, float_cols_rel=1.01e-3):
    """Text equality; numeric cells may differ in the last printed digit (the reference's own
    comparator, test_statSTR.py:111-131, allows the same)."""
    g, w = got.splitlines(), want.splitlines()
    assert len(g) == len(w)
    assert g[0] == w[0]
    n_diff = 0
    for i, (a, b) in enumerate(zip(g, w)):
        if a == b:
            continue
        ca, cb = a.split("\t"), b.split("\t")
        assert len(ca) == len(cb), i
        for x, y in zip(ca, cb):
            if x == y:
                continue
            fx, fy = float(x), float(y)
            assert abs(fx - fy) <= float_cols_rel * max(abs(fx), abs(fy)), (i, x, y)
            n_diff += 1
    return n_diff


def test_statstr_cli_matches_reference_output(golden_dir, data_dir, tmp_path):
    from trtools_b200 import statSTR
    loci, extra, _ = fixture(golden_dir, "hipstr_many")
    vcf = os.path.join(data_dir, "many_samples.vcf.gz")
    cases = [("tab_all", {}), ("tab_all_uselength", {"use_length": True}),
             ("tab_strat", {"samples": os.path.join(data_dir, "many_samples_subsample1.txt") + "," +
                            os.path.join(data_dir, "many_samples_subsample2.txt"), "sample_prefixes": "1,2"})]
    for key, kw in cases:
        out = str(tmp_path / key)
        assert statSTR.main(_statstr_args(vcf, out, precision=4, **kw)) == 0
        n_diff = _same_tab(open(out + ".tab").read(), extra[key])
        assert n_diff <= 20, (key, n_diff)      # last-digit rounding of printed floats only


def test_statstr_config1_trio(golden_dir, data_dir, tmp_path):
    """BASELINE config 1: statSTR --afreq --mean --vcftype hipstr on trio_chr21_hipstr."""
    from trtools_b200 import statSTR
    _, extra, _ = fixture(golden_dir, "hipstr_trio")
    out = str(tmp_path / "c1")
    args = _statstr_args(os.path.join(data_dir, "trio_chr21_hipstr.sorted.vcf.gz"), out, vcftype="hipstr")
    for s in ("thresh", "acount", "nalleles", "hwep", "het", "entropy", "mode", "var", "numcalled"):
        setattr(args, s, False)
    assert statSTR.main(args) == 0
    assert _same_tab(open(out + ".tab").read(), extra["tab_c1"]) <= 5


def test_other_callers_cli(golden_dir, data_dir, tmp_path):
    from trtools_b200 import statSTR
    for name, fn, vt in [("gangstr", "test_gangstr_head.vcf", "gangstr"), ("popstr", "test_popstr.vcf", "popstr"),
                         ("eh", "test_ExpansionHunter.vcf", "eh"), ("advntr", "test_advntr.vcf", "advntr"),
                         ("longtr", "test_longtr.vcf", "longtr")]:
        _, extra, _ = fixture(golden_dir, name)
        for key, ul in (("tab", False), ("tab_uselength", True)):
            out = str(tmp_path / (name + key))
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                assert statSTR.main(_statstr_args(os.path.join(data_dir, fn), out, vcftype=vt, precision=4,
                                                  use_length=ul)) == 0
            assert _same_tab(open(out + ".tab").read(), extra[key]) <= 10, (name, key)


def test_synthetic_generator_twins_agree(ctx):
    """trt_synth_fill (CUDA) == trtools_b200.synth.fill_calls (numpy), bit for bit."""
    from trtools_b200 import synth, _lib
    for L, S, seed, off in [(37, 1000, 20261017, 0), (5, 4099, 3, 1000)]:
        sl = synth.make_loci(L, seed=seed, locus_offset=off)
        want = synth.fill_calls(sl, S)
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.synth_fill(seed, off, sl.cum_freq, sl.miss_thresh, sl.half_thresh, True)
        assert np.array_equal(ctx.block_get_gt(0, L), want.gt)
        assert np.array_equal(ctx.block_get_format(_lib.FMT_DP, 0, L, np.int32), want.dp)
        assert np.array_equal(ctx.block_get_format(_lib.FMT_DSTUTTER, 0, L, np.int32), want.dstutter)
        assert np.array_equal(ctx.block_get_format(_lib.FMT_DFLANKINDEL, 0, L, np.int32), want.dflankindel)
        q = ctx.block_get_format(_lib.FMT_Q, 0, L, np.float32)
        assert np.array_equal(np.isnan(q), np.isnan(want.q)) and np.array_equal(q[~np.isnan(q)], want.q[~np.isnan(want.q)])


@pytest.mark.parametrize("L,S", [(64, 10000), (24, 50000), (40, 2055), (16, 16385)])
def test_stats_vs_oracle_on_synthetic(ctx, L, S):
    """Fast TMA path (S >= 2048) vs the oracle on seeded synthetic blocks, both relations."""
    from oracle import stats as ostats, trh as otrh
    from oracle.records import synth_to_loci, LocusAsVariant
    from trtools_b200 import synth, block
    from trtools_b200.statSTR import _locus_keys, _afreq_string
    sl = synth.make_loci(L, seed=S)
    calls = synth.fill_calls(sl, S)
    loci = synth_to_loci(sl, calls, with_fmt=False)
    blk = block.build_block(ctx, "hipstr", [LocusAsVariant(l) for l in loci])
    rng = np.random.default_rng(S)
    gm = np.stack([np.ones(S, np.uint8), (rng.random(S) < 0.3).astype(np.uint8)])
    groups = [None, gm[1].astype(bool)]
    step = max(1, L // 8)
    for uselength in (True, False):
        st = blk.stats(uselength, 0.01, gm)
        for j in range(0, L, step):
            h = otrh.harmonize(loci[j])
            want = ostats.locus_stats(h, loci[j].gt, ostats.STAT_ORDER, groups, uselength=uselength)
            keys = _locus_keys(blk, j, uselength)
            sl_ = blk.allele_slice(j)
            for g in range(2):
                what = "L{} S{} locus {} ul {} g {}".format(L, S, j, uselength, g)
                assert _afreq_string(keys, st["ac"][g, sl_], True) == want["acount"][g], what
                assert int(st["n_called"][g, j]) == want["numcalled"][g], what
                assert int(st["nalleles"][g, j]) == want["nalleles"][g], what
                for stat in ("thresh", "hwep", "het", "entropy", "mean", "mode", "var"):
                    assert_close(st[stat][g, j], want[stat][g], what + " " + stat, abs_tol=1e-300)


@pytest.mark.parametrize("L,S,G", [(48, 10000, 5), (40, 2055, 3), (12, 40000, 4), (150, 4100, 7)])
def test_sample_groups_share_scan_passes(ctx, L, S, G):
    """statSTR --samples with several files: the scan counts up to three sample groups per read of the genotypes.
    Overlapping, full and empty groups against the oracle, and against one-group-at-a-time runs of the same block."""
    from oracle import stats as ostats, trh as otrh
    from oracle.records import synth_to_loci, LocusAsVariant
    from trtools_b200 import synth, block
    from trtools_b200.statSTR import _locus_keys, _afreq_string
    sl = synth.make_loci(L, seed=S + G)
    calls = synth.fill_calls(sl, S)
    loci = synth_to_loci(sl, calls, with_fmt=False)
    blk = block.build_block(ctx, "hipstr", [LocusAsVariant(l) for l in loci])
    rng = np.random.default_rng(S * 7 + G)
    gm = np.stack([(rng.random(S) < p).astype(np.uint8) for p in np.linspace(0.15, 0.9, G)])
    gm[1] = 1                                  # everyone
    if G > 3:
        gm[3] = 0                              # nobody
    st = blk.stats(True, 0.01, gm)
    for g in range(G):
        alone = blk.stats(True, 0.01, gm[g:g + 1])
        for key in ("ac", "n_called", "nalleles", "het", "hwep", "mean", "var", "mode", "entropy", "thresh"):
            a, b = np.asarray(st[key][g]), np.asarray(alone[key][0])
            assert np.array_equal(a, b, equal_nan=True), (key, g)
    groups = [gm[g].astype(bool) for g in range(G)]
    for j in range(0, L, max(1, L // 6)):
        h = otrh.harmonize(loci[j])
        want = ostats.locus_stats(h, loci[j].gt, ostats.STAT_ORDER, groups, uselength=True)
        keys = _locus_keys(blk, j, True)
        sl_ = blk.allele_slice(j)
        for g in range(G):
            what = "L{} S{} locus {} g {}".format(L, S, j, g)
            assert _afreq_string(keys, st["ac"][g, sl_], True) == want["acount"][g], what
            assert int(st["n_called"][g, j]) == want["numcalled"][g], what
            for stat in ("thresh", "hwep", "het", "entropy", "mean", "mode", "var"):
                assert_close(st[stat][g, j], want[stat][g], what + " " + stat, abs_tol=1e-300)


def test_many_alleles_and_odd_shapes(ctx):
    """Loci beyond the fast path's allele budget (generic kernel), haploid and triploid blocks."""
    from oracle import stats as ostats, trh as otrh
    from oracle.records import Locus, LocusAsVariant
    from trtools_b200 import block
    rng = np.random.default_rng(5)
    S = 3000
    # 150 alleles (> 96): falls back to the generic kernel inside a fast-path launch
    ref = "AC" * 20
    alts = ["AC" * k for k in range(1, 151) if k != 20][:149]
    gt = np.stack([rng.integers(0, 150, S), rng.integers(0, 150, S), np.ones(S, int)], axis=1).astype(np.int16)
    gt[rng.random(S) < 0.05, :2] = (-1, -2)
    big = Locus("hipstr", "1", 100, ref, alts, {"START": 100, "END": 139, "PERIOD": 2}, gt)
    small = Locus("hipstr", "1", 500, "ACAC", ["AC"], {"START": 500, "END": 503, "PERIOD": 2},
                  np.stack([rng.integers(0, 2, S), rng.integers(0, 2, S), np.zeros(S, int)], axis=1).astype(np.int16))
    blk = block.build_block(ctx, "hipstr", [LocusAsVariant(big), LocusAsVariant(small)])
    for ul in (True, False):
        st = blk.stats(ul)
        for j, l in enumerate((big, small)):
            h = otrh.harmonize(l)
            want = ostats.locus_stats(h, l.gt, ostats.STAT_ORDER, [None], uselength=ul)
            ac = otrh.allele_counts(h, l.gt, index=True)
            got = st["ac"][0, blk.allele_slice(j)]
            assert {int(k): int(v) for k, v in ac.items()} == {i: int(v) for i, v in enumerate(got) if v > 0}
            for stat in ("thresh", "hwep", "het", "entropy", "mean", "mode", "var"):
                assert_close(st[stat][0, j], want[stat][0], "{} {}".format(j, stat), abs_tol=1e-300)
    # triploid / haploid rows
    for P in (1, 3):
        g = rng.integers(-2, 3, size=(200, P)).astype(np.int16)
        g = np.concatenate([g, np.zeros((200, 1), np.int16)], axis=1)
        l = Locus("hipstr", "1", 100, "ACACAC", ["ACAC", "ACACACAC"], {"START": 100, "END": 105, "PERIOD": 2}, g)
        blk = block.build_block(ctx, "hipstr", [LocusAsVariant(l)])
        st = blk.stats(True)
        h = otrh.harmonize(l)
        ac = otrh.allele_counts(h, l.gt, index=True)
        assert {int(k): int(v) for k, v in ac.items()} == {i: int(v) for i, v in enumerate(st["ac"][0]) if v > 0}
        assert int(st["n_called"][0, 0]) == int(np.sum(otrh.called_samples(l.gt)))
        if P == 3:
            want = ostats.stat_hwep(h, l.gt, [None], True)[0]
            assert_close(st["hwep"][0, 0], want, "triploid hwep")


def test_trrecord_api_known_answers(ctx):
    """Known answers of the reference's own unit tests (trtools/utils/tests/test_trharmonizer.py:
    312-715) through the drop-in TRRecord constructor and accessors."""
    import types
    from trtools_b200 import tr_harmonizer as trh

    class DummyCyvcf2Record:      # same surface as the reference's test double (:18-50)
        def __init__(self, gts, ref, alt):
            self.POS, self.CHROM, self.FORMAT, self.INFO = 42, '1984', {}, {}
            self.ALT, self.REF = list(alt), ref
            if gts is not None:
                self.genotype = types.SimpleNamespace()
                self._gts = np.array(gts)
                if len(self._gts) > 0:
                    self._gts = np.concatenate((self._gts, np.zeros((self._gts.shape[0], 1))), axis=1)
                self.genotype.array = lambda: self._gts
                self.genotype.n_samples = len(gts)
                self.ploidy = self._gts.shape[1] - 1 if len(self._gts) else 2
            else:
                self.genotype = None

        def format(self, key):
            return self.FORMAT.get(key, None)

    gts = [[0, 1], [1, 1], [1, 1], [1, 2], [2, 2], [0, -1]]
    rec = DummyCyvcf2Record(gts, "CAGCAGCAG", ["CAGCAGCAGCAG", "CAGCAGCAGCAGCAGCAG"])
    tr = trh.TRRecord(rec, "CAGCAGCAG", ["CAGCAGCAGCAG", "CAGCAGCAGCAGCAGCAG"], "CAG", "STR1", None)
    assert tr.ref_allele_length == 3 and tr.alt_allele_lengths == [4, 6]
    # test_GetGenotypeCounts :441-508
    assert tr.GetGenotypeCounts() == {(3, 4): 1, (4, 4): 2, (4, 6): 1, (6, 6): 1}
    assert tr.GetGenotypeCounts(index=True) == {(0, 1): 1, (1, 1): 2, (1, 2): 1, (2, 2): 1}
    assert tr.GetGenotypeCounts(uselength=False) == {
        ("CAGCAGCAG", "CAGCAGCAGCAG"): 1, ("CAGCAGCAGCAG", "CAGCAGCAGCAG"): 2,
        ("CAGCAGCAGCAG", "CAGCAGCAGCAGCAGCAG"): 1, ("CAGCAGCAGCAGCAGCAG", "CAGCAGCAGCAGCAGCAG"): 1}
    assert tr.GetGenotypeCounts(include_nocalls=True)[(-1, 3)] == 1
    assert tr.GetGenotypeCounts(sample_index=[0, 1, 3]) == {(3, 4): 1, (4, 4): 1, (4, 6): 1}
    # test_GetAlleleCounts :511-558 / GetAlleleFreqs :561-633
    assert tr.GetAlleleCounts() == {3: 2, 4: 6, 6: 3}
    assert tr.GetAlleleCounts(index=True) == {0: 2, 1: 6, 2: 3}
    assert tr.GetAlleleCounts(uselength=False) == {"CAGCAGCAG": 2, "CAGCAGCAGCAG": 6, "CAGCAGCAGCAGCAGCAG": 3}
    assert tr.GetAlleleCounts(sample_index=[0, 5]) == {3: 2, 4: 1}
    fr = tr.GetAlleleFreqs()
    assert fr[3] == pytest.approx(2 / 11) and fr[4] == pytest.approx(6 / 11) and fr[6] == pytest.approx(3 / 11)
    assert tr.GetMaxAllele() == 6 and tr.GetMaxAllele(sample_index=[0, 5]) == 4
    # called samples / call rate / ploidies :666-715
    assert np.array_equal(tr.GetCalledSamples(), [True] * 5 + [False])
    assert np.array_equal(tr.GetCalledSamples(strict=False), [True] * 6)
    assert tr.GetCallRate() == pytest.approx(5 / 6) and tr.GetCallRate(strict=False) == 1
    lg = tr.GetLengthGenotypes()
    assert lg.dtype == np.float64 and np.array_equal(lg[:, :-1], [[3, 4], [4, 4], [4, 4], [4, 6], [6, 6], [3, -1]])
    assert np.array_equal(tr.GetDosages(), np.array([7, 8, 8, 10, 12, 3], dtype=np.float32))
    # accessors around a same-length sequence variant and a lower-ploidy sample; expected values produced by the
    # unmodified reference on the same record (string genotypes :963-1017, unique mappings :1049-1082, :1247-1273,
    # ploidies :899-919, normalised dosages :1191-1205)
    gts_b = [[0, 1], [1, 1], [1, 1], [1, 2], [2, 2], [0, -1], [2, -2]]
    rec_b = DummyCyvcf2Record(gts_b, "CAGCAGCAG", ["CAGCAGCAGCAG", "CAGCAACAGCAG"])
    trb = trh.TRRecord(rec_b, "CAGCAGCAG", ["CAGCAGCAGCAG", "CAGCAACAGCAG"], "CAG", "STR1", None)
    sg = trb.GetStringGenotypes()
    assert sg.tolist() == [['CAGCAGCAG', 'CAGCAGCAGCAG', '0'], ['CAGCAGCAGCAG', 'CAGCAGCAGCAG', '0'],
                           ['CAGCAGCAGCAG', 'CAGCAGCAGCAG', '0'], ['CAGCAGCAGCAG', 'CAGCAACAGCAG', '0'],
                           ['CAGCAACAGCAG', 'CAGCAACAGCAG', '0'], ['CAGCAGCAG', '.', '0'], ['CAGCAACAGCAG', ',', '0']]
    assert trb.UniqueStringGenotypeMapping() == {0: 0, 1: 1, 2: 2} and trb.UniqueLengthGenotypeMapping() == {0: 0, 1: 1, 2: 1}
    assert trb.UniqueStringGenotypes() == {0, 1, 2} and trb.UniqueLengthGenotypes() == {0, 1}
    assert trb.GetSamplePloidies().tolist() == [2, 2, 2, 2, 2, 2, 1] and trb.GetMaxPloidy() == 2 and trb.GetNumSamples() == 7
    assert np.array_equal(trb.GetDosages(), np.array([7, 8, 8, 8, 8, 3, 4], dtype=np.float32))
    assert np.array_equal(trb.GetDosages(trh.TRDosageTypes.bestguess_norm),
                          np.array([1, 2, 2, 2, 2, np.nan, np.nan], dtype=np.float32), equal_nan=True)
    assert not trb.HasQualityScores()
    assert trb.GetAlleleCounts() == {3: 2, 4: 10}
    assert trb.GetAlleleCounts(uselength=False) == {"CAGCAACAGCAG": 4, "CAGCAGCAG": 2, "CAGCAGCAGCAG": 6}
    assert trb.GetGenotypeCounts() == {(-2, 4): 1, (3, 4): 1, (4, 4): 4}
    assert np.array_equal(trb.GetCalledSamples(), [True] * 5 + [False, True])
    assert str(trb) == "STR1 CAG CAGCAGCAG CAGCAGCAGCAG,CAGCAACAGCAG"
    # fabricated alleles (length-only callers)
    rec2 = DummyCyvcf2Record(gts, "A", ["<STR4>", "<STR5.5>"])
    tr2 = trh.TRRecord(rec2, None, None, "CAG", "x", None, ref_allele_length=3, alt_allele_lengths=[4, 5.5])
    assert tr2.ref_allele == "CAGCAGCAG" and tr2.alt_alleles == ["CAGCAGCAGCAG", "CAGCAGCAGCAGCAGC"]
    assert tr2.GetAlleleCounts() == {3: 2, 4: 6, 5.5: 3}
    # no samples
    tr3 = trh.TRRecord(DummyCyvcf2Record(None, "CAGCAG", []), "CAGCAG", [], "CAG", "y", None)
    assert tr3.GetAlleleCounts() == {} and tr3.GetGenotypeCounts() == {} and tr3.GetGenotypeIndicies() is None
    # argument validation errors of the constructor (:720-731)
    with pytest.raises(ValueError):
        trh.TRRecord(rec, "CAG", ["CAG"], "CAG", "", None, alt_allele_lengths=[1])
    with pytest.raises(ValueError):
        trh.TRRecord(rec, "CAGCAGCAG", ["CAG"], "CAG", "", None)      # wrong number of alts


@pytest.mark.parametrize("L,S,seed", [(3000, 4100, 5), (1500, 9000, 6)])
def test_many_loci_per_cta(ctx, L, S, seed):
    """Persistent CTAs process many loci back to back (L >> 148): device-generated block, sampled loci
    checked against the oracle (catches cross-locus hazards in the thread-private tables)."""
    from oracle import stats as ostats, trh as otrh
    from oracle.records import synth_to_loci
    from trtools_b200 import synth
    sl = synth.make_loci(L, seed=seed)
    ctx.block_begin(L, S, 2, "hipstr")
    ctx.synth_fill(seed, 0, sl.cum_freq, sl.miss_thresh, sl.half_thresh, with_format=False)
    ctx.block_set_alleles(*synth.allele_tables(sl))
    ctx._current_block = None
    ctx.harmonize()
    rng = np.random.default_rng(seed)
    for uselength in (True, False):
        st = ctx.locus_stats(uselength, None, 0.01)
        picks = sorted(set(rng.integers(0, L, 10).tolist() + [0, L - 1, 147, 148, 149, 296]))
        for j in picks:
            calls = synth.fill_calls(sl, S, slice(j, j + 1))
            sub = synth.SynthLoci(seed=sl.seed, n_loci=1, chrom=sl.chrom[j:j + 1], pos=sl.pos[j:j + 1],
                                  start=sl.start[j:j + 1], end=sl.end[j:j + 1], period=sl.period[j:j + 1],
                                  ref=sl.ref[j:j + 1], alts=sl.alts[j:j + 1], n_alleles=sl.n_alleles[j:j + 1],
                                  cum_freq=sl.cum_freq[j:j + 1], locus_offset=j)
            locus = synth_to_loci(sub, calls, with_fmt=False)[0]
            h = otrh.harmonize(locus)
            want = ostats.locus_stats(h, locus.gt, ostats.STAT_ORDER, [None], uselength=uselength)
            ac = otrh.allele_counts(h, locus.gt, index=True)
            a0, a1 = int(ctx.locus_off[j]), int(ctx.locus_off[j + 1])
            got = st["ac"][0, a0:a1]
            assert {int(k): int(v) for k, v in ac.items()} == {i: int(v) for i, v in enumerate(got) if v > 0}, j
            assert int(st["n_called"][0, j]) == want["numcalled"][0], j
            for stat in ("thresh", "hwep", "het", "entropy", "mean", "mode", "var"):
                assert_close(st[stat][0, j], want[stat][0], "locus {} {}".format(j, stat), abs_tol=1e-300)


@pytest.mark.parametrize("L,S", [(60, 4104), (33, 2048), (25, 1001)])
def test_packed_length_genotype_tensor(ctx, monkeypatch, L, S):
    """Packed int16 [L][S][P] tensor (device-side GetLengthGenotypes, tr_harmonizer.py:1210-1245): entry = rank of the
    haplotype's length among the locus' distinct lengths, -1 / -2 sentinels kept.  Vectorised kernel vs the scalar one,
    vs a numpy restatement, and round-tripped through the oracle's allele lengths."""
    from oracle import trh as otrh
    from oracle.records import synth_to_loci
    from trtools_b200 import synth
    sl = synth.make_loci(L, seed=400 + L)
    calls = synth.fill_calls(sl, S)

    def run():
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.block_set_gt(calls.gt)
        ctx.block_set_alleles(*synth.allele_tables(sl))
        h = ctx.harmonize()
        return h, ctx.pack_length_genotypes()

    h, packed = run()
    monkeypatch.setenv("TRT_PACK_SCALAR", "1")
    _, packed_scalar = run()
    monkeypatch.delenv("TRT_PACK_SCALAR")
    assert packed.shape == (L, S, 2) and packed.dtype == np.int16
    assert np.array_equal(packed, packed_scalar)
    off = ctx.locus_off
    loci = synth_to_loci(sl, calls, with_fmt=False)
    for l in range(L):
        lens = h["allele_len"][off[l]:off[l + 1]]
        uniq = np.unique(lens)
        rank = np.searchsorted(uniq, lens).astype(np.int16)
        gt = calls.gt[l, :, :2]
        want = np.where(gt >= 0, rank[np.clip(gt, 0, len(lens) - 1)], gt)
        assert np.array_equal(packed[l], want), l
        if l % 7 == 0:      # length genotypes of the reference = table[packed]
            oh = otrh.harmonize(loci[l])
            ref_lens = np.array([oh.ref_allele_length] + list(oh.alt_allele_lengths))
            lg = np.where(gt >= 0, ref_lens[np.clip(gt, 0, len(lens) - 1)], gt.astype(float))
            got = np.where(packed[l] >= 0, uniq[np.clip(packed[l], 0, len(uniq) - 1)], packed[l].astype(float))
            assert np.array_equal(got, lg), l


def test_packed_gt_transfer_form_round_trip_and_same_statistics(ctx):
    """trt_block_set_gt_packed (2 bytes per call + phase bits across PCIe, expanded on the device) rebuilds exactly the
    rows trt_block_set_gt uploads — pads, no-calls, half calls, unphased calls, S not a multiple of 8 — the packed
    read-back equals numpy's packing, and the statistics are identical through both transfer paths."""
    from trtools_b200 import synth
    from trtools_b200.block import pack_gt, unpack_gt
    for L, S, seed in ((37, 3001, 5), (5, 8, 6), (64, 4096, 7)):
        sl = synth.make_loci(L, seed=seed)
        calls = synth.fill_calls(sl, S)
        gt = calls.gt.copy()
        rng = np.random.default_rng(seed)
        gt[rng.random((L, S)) < 0.05, 2] = 0                      # some unphased calls
        hap = rng.random((L, S)) < 0.03
        gt[hap, 1] = -2                                           # haploid calls (ploidy pad)
        g2, ph = pack_gt(gt)
        assert np.array_equal(unpack_gt(g2, ph), gt)
        tables = synth.allele_tables(sl)
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.block_set_gt(gt)
        ctx.block_set_alleles(*tables)
        ctx._current_block = None
        ctx.check(ctx.lib.trt_harmonize(ctx.h))
        want = {k: v.copy() for k, v in ctx.locus_stats(False, None, 0.01).items()}
        back2, backph = ctx.block_get_gt_packed(0, L, with_phase=True)
        assert np.array_equal(back2, g2) and np.array_equal(backph, ph)
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.block_set_gt_packed(g2, ph)
        ctx.block_set_alleles(*tables)
        ctx.check(ctx.lib.trt_harmonize(ctx.h))
        assert np.array_equal(ctx.block_get_gt(0, L), gt)
        got = ctx.locus_stats(False, None, 0.01)
        for k in want:
            assert np.array_equal(want[k], got[k], equal_nan=True), k
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.block_set_gt_packed(g2)                               # no phase plane: every call reads unphased
        ctx.block_set_alleles(*tables)
        unph = gt.copy()
        unph[:, :, 2] = 0
        assert np.array_equal(ctx.block_get_gt(0, L), unph)
        # the nibble form (one byte per call; the synthetic loci have <= 14 alleles)
        from trtools_b200.block import pack_gt4, unpack_gt4
        g4, ph4 = pack_gt4(gt)
        assert np.array_equal(unpack_gt4(g4, ph4), gt) and np.array_equal(ph4, ph)
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.block_set_gt_nibble(g4, ph4)
        ctx.block_set_alleles(*tables)
        ctx.check(ctx.lib.trt_harmonize(ctx.h))
        assert np.array_equal(ctx.block_get_gt(0, L), gt)
        back4, backph4 = ctx.block_get_gt_nibble(0, L, with_phase=True)
        assert np.array_equal(back4, g4) and np.array_equal(backph4, ph4)
        got = ctx.locus_stats(False, None, 0.01)
        for k in want:
            assert np.array_equal(want[k], got[k], equal_nan=True), k
    big = np.zeros((1, 16, 3), np.int16)
    big[0, 3, 0] = 300
    assert pack_gt(big) is None                                   # does not fit: callers use the int16 layout
    big[0, 3, 0] = 14
    from trtools_b200.block import pack_gt4
    assert pack_gt4(big) is None and pack_gt(big) is not None     # 14 fits two bytes, not a nibble
