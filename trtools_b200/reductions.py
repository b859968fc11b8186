"""
Block-level reductions behind the reference's qcSTR and compareSTR tools.  The tools themselves
(plots, argument handling) are outside the accelerated path; these functions are what their record loops reduce to
once records arrive as GPU blocks:

* :func:`qc_block` — ``qcSTR.main``'s loop body (reference trtools/qcSTR/qcSTR.py:523-570): calls per sample and per
  locus, the quality sums, and the per-locus length allele counts the reference-bias plot is built from;
* :func:`compare_blocks` — ``compareSTR.UpdateComparisonResults`` (trtools/compareSTR/compareSTR.py:508-643, without the
  FORMAT stratification) for two blocks holding the same loci of two call sets.
"""
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib, block as _block


def qc_block(blk: "_block.Block", sample_index: Optional[np.ndarray], sample_calls: np.ndarray,
             per_sample_total_qual: Optional[np.ndarray] = None, quality_key: Optional[str] = None,
             ignore_no_call: bool = False) -> Dict[str, np.ndarray]:
    """One ``trt_qc_reduce`` pass over the block.  ``sample_calls`` (int64 [S], whole sample axis) and, when a quality
    field is given, ``per_sample_total_qual`` (float64 [S]) are accumulated in place; returns the per-locus call counts
    (``chrom_calls`` increments), the per-locus mean quality, and the per-locus allele counts by index among the
    selected samples (lengths: ``blk.h['allele_len']``)."""
    blk._activate()
    mask = None if sample_index is None else np.ascontiguousarray(sample_index, dtype=np.uint8)
    slot = -1
    if quality_key is not None:
        if quality_key not in blk.fmt_slot:
            raise KeyError("quality field {} is not in the block".format(quality_key))
        slot = blk.fmt_slot[quality_key]
    res = blk.ctx.qc_reduce(sample_calls, per_sample_total_qual, mask, slot, ignore_no_call, blk.rec_ploidy)
    st = blk.ctx.locus_stats(True, None if mask is None else mask[None, :], 0.01, want=("ac",))
    res["allele_counts"] = st["ac"][0]
    return res


def compare_blocks(blk1: "_block.Block", blk2: "_block.Block", sample_idxs: Sequence[np.ndarray], ignore_phasing: bool,
                   sample_results: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """``trt_compare`` of two blocks holding the SAME loci (record i of one is record i of the other) of two call sets.
    ``sample_idxs``: the two index arrays of the shared samples; ``sample_results``: int64 arrays ``numcalls``,
    ``conc-seq-count``, ``conc-len-count`` over the shared samples, accumulated in place.  Returns per-locus ``numcalls``,
    ``conc_seq``, ``conc_len`` (counts), ``len_sums`` [L, 5] and raises the reference's ValueError for a locus whose
    calls differ in ploidy or mix phased and unphased genotypes."""
    if blk1.L != blk2.L or blk1.P != blk2.P:
        raise ValueError("the two blocks must hold the same loci with the same ploidy")
    L = blk1.L
    # per allele of set 2: its length and the sequence class of set 1 it equals (or an id of its own)
    seq_id2 = np.empty(int(blk2.locus_off[-1]), np.int32)
    reflen = np.empty(L, np.float64)
    for l in range(L):
        a1, a2 = blk1.trimmed_alleles(l), blk2.trimmed_alleles(l)
        s1 = blk1.allele_slice(l)
        cls1 = blk1.h["seq_class"][s1]
        first = {}
        for j, seq in enumerate(a1):
            first.setdefault(seq, int(cls1[j]))
        own = {}
        s2 = blk2.allele_slice(l)
        for j, seq in enumerate(a2):
            seq_id2[s2.start + j] = first[seq] if seq in first else -1 - own.setdefault(seq, len(own))
        motif = blk1.motif(l)
        reflen[l] = len(a1[0]) / len(motif)
    gt2 = blk2.gt
    blk1._activate()
    res = blk1.ctx.compare(gt2, sample_idxs[0], sample_idxs[1], blk2.locus_off, seq_id2, blk2.h["allele_len"], reflen,
                           ignore_phasing, sample_results["numcalls"], sample_results["conc-seq-count"],
                           sample_results["conc-len-count"])
    bad = np.nonzero(res["status"])[0]
    if len(bad):
        l = int(bad[0])
        m = blk1.metas[l]
        pos = m.harmonized_pos if m.harmonized_pos is not None else m.vcf_pos
        what = "of different ploidy" if int(res["status"][l]) == 1 else "with different phasedness"
        raise ValueError("Found sample(s) %s at %s:%s" % (what, m.chrom, pos))
    return res
