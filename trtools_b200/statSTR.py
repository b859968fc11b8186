"""
statSTR — per-locus statistics of a TR VCF (drop-in for reference trtools/statSTR/statSTR.py).

Same flags, same ``main(args) -> int`` contract, byte-compatible ``<out>.tab``.  The per-record
Python loop of the reference (statSTR.py:575-639, which re-derives allele counts with
``np.unique`` once per statistic) is replaced by one GPU pass per block of records:
``trt_harmonize`` + ``trt_locus_stats`` (one scan of the GT rows per sample group, then the FP64
epilogue incl. the exact HWE binomial test).  Python only formats the text.
"""
import argparse
import os
import sys
import time
from typing import Any, List

import numpy as np

from . import __version__, _lib, block as _block, common, dist as _dist
from . import tr_harmonizer as trh
from . import utils

MAXPLOTS = 10


def GetHeader(header, sample_prefixes):
    """reference statSTR.py:82-102."""
    if len(sample_prefixes) == 0:
        return [header]
    return [header + "-" + sp for sp in sample_prefixes]


# ---- per-record wrappers kept for API compatibility (reference statSTR.py:104-426) -----------------
# They accept a GPU-backed TRRecord and the reference's ``sample_indexes`` lists.
def _masks(trrecord, sample_indexes):
    S = trrecord._blk.S
    if all(si is None for si in sample_indexes):
        return None
    return np.stack([np.ones(S, np.uint8) if si is None else trh._sample_mask(si, S) for si in sample_indexes])


def _record_stats(trrecord, sample_indexes, uselength, nalleles_thresh=0.01):
    st = trrecord._blk.stats(uselength, nalleles_thresh, _masks(trrecord, sample_indexes))
    return st, trrecord._l


def GetThresh(trrecord, sample_indexes: List[Any] = [None]) -> List[float]:
    st, l = _record_stats(trrecord, sample_indexes, True)
    return [float(st["thresh"][g, l]) for g in range(len(sample_indexes))]


def GetAFreq(trrecord, sample_indexes: List[Any] = [None], count: bool = False, uselength: bool = True) -> List[str]:
    st, l = _record_stats(trrecord, sample_indexes, uselength)
    blk = trrecord._blk
    keys = _locus_keys(blk, l, uselength)
    sl = blk.allele_slice(l)
    return [_afreq_string(keys, st["ac"][g, sl], count) for g in range(len(sample_indexes))]


def GetNAlleles(trrecord, sample_indexes: List[Any] = [None], nalleles_thresh: float = 0.01,
                uselength: bool = True) -> List[int]:
    st, l = _record_stats(trrecord, sample_indexes, uselength, nalleles_thresh)
    return [int(st["nalleles"][g, l]) for g in range(len(sample_indexes))]


def _stat_list(name, trrecord, sample_indexes, uselength):
    st, l = _record_stats(trrecord, sample_indexes, uselength)
    return [float(st[name][g, l]) for g in range(len(sample_indexes))]


def GetHWEP(trrecord, sample_indexes: List[Any] = [None], uselength: bool = True) -> List[float]:
    return _stat_list("hwep", trrecord, sample_indexes, uselength)


def GetHet(trrecord, sample_indexes: List[Any] = [None], uselength: bool = True) -> List[float]:
    return _stat_list("het", trrecord, sample_indexes, uselength)


def GetEntropy(trrecord, sample_indexes: List[Any] = [None], uselength: bool = True) -> List[float]:
    return _stat_list("entropy", trrecord, sample_indexes, uselength)


def GetMean(trrecord, sample_indexes: List[Any] = [None]) -> List[float]:
    return _stat_list("mean", trrecord, sample_indexes, True)


def GetMode(trrecord, sample_indexes: List[Any] = [None]) -> List[float]:
    return _stat_list("mode", trrecord, sample_indexes, True)


def GetVariance(trrecord, sample_indexes: List[Any] = [None]) -> List[float]:
    return _stat_list("var", trrecord, sample_indexes, True)


def GetNumSamples(trrecord, sample_indexes=[None]):
    st, l = _record_stats(trrecord, sample_indexes, True)
    return [int(st["n_called"][g, l]) for g in range(len(sample_indexes))]


# ---- text helpers ---------------------------------------------------------------------------------
def _locus_keys(blk: "_block.Block", l: int, uselength: bool):
    sl = blk.allele_slice(l)
    if uselength:
        return [np.float64(x) for x in blk.h["allele_len"][sl]]
    return blk.trimmed_alleles(l)


def _afreq_string(keys, ac, count: bool) -> str:
    """statSTR.py:159-171: ``key:value`` over sorted keys of the alleles that were called."""
    folded = {}
    for k, c in zip(keys, ac):
        if c > 0:
            folded[k] = folded.get(k, 0) + int(c)
    if not folded:
        return "."
    if count:
        return ",".join(["%s:%i" % (a, folded[a]) for a in sorted(folded)])
    total = float(sum(folded.values()))
    return ",".join(["%s:%.3f" % (a, folded[a] / total) for a in sorted(folded)])


def format_nan_precision(precision_format, val):
    """reference statSTR.py:490-494."""
    if np.isnan(val):
        return "\tnan"
    return precision_format.format(val)


def getargs():  # pragma: no cover
    """reference statSTR.py:428-488 (same flags and defaults)."""
    parser = argparse.ArgumentParser(__doc__, formatter_class=utils.ArgumentDefaultsHelpFormatter)
    inout_group = parser.add_argument_group("Input/output")
    inout_group.add_argument("--vcf", help="Input STR VCF file", type=str, required=True)
    inout_group.add_argument("--out", help=("Output file prefix. Use stdout to print file to standard "
                                            "output. In addition, if not stdout then timing diagnostics are print to "
                                            "stdout."), type=str, required=True)
    inout_group.add_argument("--vcftype", help="Options=%s" % [str(item) for item in trh.VcfTypes.__members__],
                             type=str, default="auto")
    inout_group.add_argument("--precision", help="How much precision to use when printing decimals", type=int,
                             default=3)
    filter_group = parser.add_argument_group("Filtering group")
    filter_group.add_argument("--samples", help="File containing list of samples to include. Or a comma-separated "
                              "list of files to compute stats separate for each group of samples", type=str)
    filter_group.add_argument("--sample-prefixes", help="Prefixes to name output for each samples group. By default "
                              "uses 1,2,3 etc.", type=str)
    filter_group.add_argument("--region", help="Restrict to the region chrom:start-end. Requires file to bgzipped "
                              "and tabix indexed.", type=str)
    filter_group.add_argument("--only-passing", help="Only process records  where FILTER==PASS", action="store_true")
    stat_group_name = "Stats group"
    stat_group = parser.add_argument_group(stat_group_name)
    stat_group.add_argument("--thresh", help="Output threshold field (max allele size, used for GangSTR strinfo).",
                            action="store_true")
    stat_group.add_argument("--afreq", help="Output allele frequencies", action="store_true")
    stat_group.add_argument("--acount", help="Output allele counts", action="store_true")
    stat_group.add_argument("--nalleles", help="Output number of alleles with frequency exceeding a specified "
                            "threshold", action="store_true")
    stat_group.add_argument("--nalleles-thresh", help="The threshold for nalleles", type=float, default=0.01)
    stat_group.add_argument("--hwep", help="Output HWE p-values per loci.", action="store_true")
    stat_group.add_argument("--het", help="Output the heterozygosity of each locus.", action="store_true")
    stat_group.add_argument("--entropy", help="Output the entropy of each locus.", action="store_true")
    stat_group.add_argument("--mean", help="Output mean of the allele frequencies.", action="store_true")
    stat_group.add_argument("--mode", help="Output mode of the allele frequencies.", action="store_true")
    stat_group.add_argument("--var", help="Output variance of the allele frequencies.", action="store_true")
    stat_group.add_argument("--numcalled", help="Output number of samples called.", action="store_true")
    stat_group.add_argument("--use-length", help="Calculate per-locus stats (het, HWE) collapsing alleles by length. "
                            "This is implicitly true for genotypers which only emit length based genotypes.",
                            action="store_true")
    plot_group = parser.add_argument_group("Plotting group")
    plot_group.add_argument("--plot-afreq", help="Output allele frequency plot. Will only do for a maximum of 10 TRs.",
                            action="store_true")
    gpu_group = parser.add_argument_group("GPU")
    gpu_group.add_argument("--block-size", help="Records staged per GPU block", type=int, default=2048)
    ver_group = parser.add_argument_group("Version")
    ver_group.add_argument("--version", action="version", version='{version}'.format(version=__version__))
    args = parser.parse_args()
    stat_dict = {}
    for grp in parser._action_groups:
        if grp.title == stat_group_name:
            stat_dict = {a.dest: getattr(args, a.dest, None) for a in grp._group_actions}
    if not any(stat_dict.values()):
        common.WARNING("Error: Please use at least one of the flags in the Stats group. See statSTR --help for options.")
        return None
    return args


def _blocks(records, block_size):
    """Yield lists of consecutive records with equal ploidy, at most block_size long."""
    cur = []
    for rec in records:
        p = rec.ploidy if rec.genotype is not None else 0
        if cur and (len(cur) >= block_size or p != cur_p):
            yield cur
            cur = []
        if not cur:
            cur_p = p
        cur.append(rec)
    if cur:
        yield cur


def main(args):
    """reference statSTR.py:496-647."""
    if not os.path.exists(args.vcf):
        common.WARNING("Error: %s does not exist" % args.vcf)
        return 1
    if not os.path.exists(os.path.dirname(os.path.abspath(args.out))):
        common.WARNING("Error: The directory which contains the output location {} does"
                       " not exist".format(args.out))
        return 1
    if os.path.isdir(args.out) and args.out.endswith(os.sep):
        common.WARNING("Error: The output location {} is a directory".format(args.out))
        return 1
    checkgz = args.region is not None
    invcf = utils.LoadSingleReader(args.vcf, checkgz=checkgz)
    if invcf is None:
        return 1
    if args.vcftype != 'auto':
        vcftype = trh.VcfTypes[args.vcftype]
    else:
        vcftype = trh.InferVCFType(invcf)

    sample_prefixes = []
    sample_indexes = []
    if args.samples:
        all_samples = np.array(invcf.samples)
        sfiles = args.samples.split(",")
        if args.sample_prefixes:
            sample_prefixes = args.sample_prefixes.split(",")
        else:
            sample_prefixes = [str(item) for item in range(1, len(sfiles) + 1)]
        if len(sfiles) != len(sample_prefixes):
            common.WARNING("--sample-prefixes must be same length as --samples")
            return 1
        for sf in sfiles:
            sample_list = np.array([item.strip() for item in open(sf, "r").readlines()])
            if not np.any(np.isin(all_samples, sample_list)):
                common.WARNING("No samples from {} found in the VCF file".format(sf))
                return 1
            sample_indexes.append(np.isin(all_samples, sample_list))
        group_masks = np.stack(sample_indexes).astype(np.uint8)
    else:
        sample_indexes = [None]
        group_masks = None
    G = len(sample_indexes)

    header = ["chrom", "start", "end"]
    for flag in ("thresh", "afreq", "acount", "nalleles", "hwep", "het", "entropy", "mean", "mode", "var", "numcalled"):
        if getattr(args, flag):
            header.extend(GetHeader(flag, sample_prefixes))
    precision_format = "\t{:." + str(args.precision) + "}"
    if getattr(args, "plot_afreq", False):
        common.WARNING("--plot-afreq is not available in trtools_b200 (plotting is outside the accelerated path)")
        if args.out == "stdout":
            common.WARNING("Cannot use --out stdout when generating plots")
            return 1
    ctx = _lib.default_context()
    block_size = int(getattr(args, "block_size", 2048) or 2048)
    if hasattr(invcf, "_native_block_loci"):
        invcf._native_block_loci = block_size        # one native run per GPU block: zero-copy hand-off of the arrays
    # several GPUs (torchrun): blocks are dealt round-robin, rank 0 gathers the rows (NCCL) and writes the file
    comm = _dist.cli_comm(ctx)
    sharder = _dist.BlockSharder(comm)
    outf = None
    try:
        if sharder.rank == 0:
            outf = sys.stdout if args.out == "stdout" else open(args.out + ".tab", "w")
            outf.write("\t".join(header) + "\n")
        region = invcf(args.region) if args.region else invcf
        start_time = time.time()
        nrecords = 0
        use_length = bool(args.use_length)
        for recs in _blocks(region, block_size):
            if not sharder.mine():
                nrecords += len(recs)
                continue
            blk = _block.build_block(ctx, vcftype.name, recs)
            flags = blk.h["flags"]
            if np.any(flags & _lib.HF_MOTIF_NONACGT):
                bad = int(np.nonzero(flags & _lib.HF_MOTIF_NONACGT)[0][0])
                raise KeyError(blk.motif(bad))
            st = blk.stats(use_length, args.nalleles_thresh, group_masks) if blk.has_samples else None
            lines = []
            for l, record in enumerate(recs):
                nrecords += 1
                if args.only_passing and record.FILTER is not None:
                    continue
                sl = blk.allele_slice(l)
                ref_len_bp = int(blk.h["trim_len"][sl.start])
                row = [str(record.CHROM) + "\t" + str(record.POS) + "\t" + str(record.POS + ref_len_bp)]

                def fnum(name):
                    for g in range(G):
                        row.append(format_nan_precision(precision_format, st[name][g, l]) if st is not None
                                   else "\tnan")
                if args.thresh:
                    fnum("thresh")
                if args.afreq or args.acount:
                    keys = _locus_keys(blk, l, use_length)
                if args.afreq:
                    for g in range(G):
                        row.append("\t" + (_afreq_string(keys, st["ac"][g, sl], False) if st is not None else "."))
                if args.acount:
                    for g in range(G):
                        row.append("\t" + (_afreq_string(keys, st["ac"][g, sl], True) if st is not None else "."))
                if args.nalleles:
                    for g in range(G):
                        row.append("\t" + str(int(st["nalleles"][g, l]) if st is not None else 0))
                if args.hwep:
                    fnum("hwep")
                if args.het:
                    fnum("het")
                if args.entropy:
                    fnum("entropy")
                if args.mean:
                    fnum("mean")
                if args.mode:
                    fnum("mode")
                if args.var:
                    fnum("var")
                if args.numcalled:
                    for g in range(G):
                        row.append("\t" + str(int(st["n_called"][g, l]) if st is not None else 0))
                lines.append("".join(row) + "\n")
            if comm is None:
                outf.write("".join(lines))
                outf.flush()
            else:
                sharder.add("".join(lines))
            if args.out != "stdout" and nrecords and sharder.rank == 0:
                print("Finished {} records, time/record={:.5}sec".format(
                    nrecords, (time.time() - start_time) / nrecords), flush=True, end="\r")
        if comm is not None:
            merged = sharder.finish()
            if merged is not None:
                outf.write(b"".join(merged).decode("utf-8"))
    finally:
        if outf is not None and args.out != "stdout":
            outf.close()
        if comm is not None:
            comm.barrier()
            comm.close()
    if args.out != "stdout" and sharder.rank == 0:
        print("\nDone", flush=True)
    return 0


def run():  # pragma: no cover
    args = getargs()
    if args is None:
        sys.exit(1)
    sys.exit(main(args))


if __name__ == "__main__":  # pragma: no cover
    run()
