#!/usr/bin/env python3
"""
associaTR — per-locus association of a phenotype with TR length (drop-in for reference
trtools/associaTR/associaTR.py): same positional arguments / flags, ``main(args)``, same TSV.

The reference regresses one locus at a time with statsmodels (associaTR.py:246-291).  Here the covariate
design is standardised once (host, as in the reference :140-194) and uploaded; every block of records then
takes one ``trt_assoc_ols`` call: scan of the GT rows for the tested-sample allele counts, the tiled FP64
moments kernel, exact down-dates for each locus' uncalled rows, and a Cholesky solve + t-test per locus on
the GPU.  Python formats the rows.
"""
import argparse
import datetime
import shutil
import sys
import time

import numpy as np

from . import __version__, _lib, block as _block, dist as _dist
from . import load_and_filter_genotypes
from . import tr_harmonizer as trh
from . import utils

cyvcf2 = utils.cyvcf2
pval_precision = 2

_REASONS = {_lib.AF_NO_CALLED: 'No called samples', _lib.AF_ONE_ALLELE: 'Only one called allele',
            _lib.AF_NCOVARS: 'n covars >= n samples'}


def _merge_arrays(a, b):
    """Left outer join on the first column (reference associaTR.py:24-49)."""
    assert len(a.shape) == 2 and len(b.shape) == 2
    assert len(set(a[:, 0]).intersection(b[:, 0])) > 0
    assert len(set(a[:, 0])) == a.shape[0]
    assert len(set(b[:, 0])) == b.shape[0]
    b = b[np.isin(b[:, 0], a[:, 0])]
    matches = np.isin(a[:, 0], b[:, 0])
    a_sort = np.argsort(a[matches, 0])
    b_match_sorted = np.searchsorted(a[matches, 0], b[:, 0], sorter=a_sort)
    new_data = np.full((a.shape[0], b.shape[1] - 1), np.nan)
    new_data[matches, :] = b[np.argsort(b_match_sorted), 1:][np.argsort(a_sort), :]
    return np.concatenate((a, new_data), axis=1)


def prepare_design(all_samples, trait_fnames, same_samples, sample_fname):
    """Covariate preparation of perform_gwas_helper (reference associaTR.py:140-194).
    Returns (covars [n, K] standardised with col 1 = 1, outcome [n], pheno_std, sample_filter bool [S])."""
    print('{} samples in the VCF'.format(len(all_samples)), flush=True)
    if not same_samples:
        covars = np.load(trait_fnames[0])
        if np.sum(np.isin(np.array(all_samples, dtype=float), covars[:, 0])) < 3:
            print(all_samples, covars[:, 0])
            print('Less than 3 samples matched between the covars array and the VCF. '
                  'Prehaps you meant to run with --same-samples? Erroring out.')
            sys.exit(1)
        for trait_fname in trait_fnames[1:]:
            covars = _merge_arrays(covars, np.load(trait_fname))
        covars = _merge_arrays(np.array(all_samples, dtype=float).reshape(-1, 1), covars)
    else:
        arrays = []
        for trait_fname in trait_fnames:
            arrays.append(np.load(trait_fname))
            if not arrays[-1].shape[0] == len(all_samples):
                print("different number of samples in covariates file {trait_fname} than VCF, "
                      "and --same-samples was specified. Erroring out.")
                sys.exit(1)
        covars = np.hstack([np.full((arrays[0].shape[0], 1), -1), *arrays])
    if sample_fname:
        with open(sample_fname) as sample_file:
            sample_subset = [line.strip() for line in sample_file.readlines()]
        sample_filter = np.isin(all_samples, sample_subset)
        print(('{} samples remain after subsetting to samples from the file {}.\n'
               '{} samples from the sample file were not present in the VCF and were discarded.'
               ).format(np.sum(sample_filter), sample_fname, len(sample_subset) - np.sum(sample_filter)))
    else:
        sample_filter = np.array([True] * len(all_samples))
    prev_n = sum(sample_filter)
    sample_filter = sample_filter & ~np.any(np.isnan(covars), axis=1)
    cur_n = sum(sample_filter)
    print(('Removing {} samples which had missing phenotypes or covariates.\n'
           'Using {} for the regression.\n'
           'The number of samples used in each variant\'s regression will only be lower '
           'if that variant has missing calls.\n').format(prev_n - cur_n, cur_n))
    covars = covars[sample_filter, :].astype(float)
    pheno_std = np.std(covars[:, 1])
    covars = (covars - np.mean(covars, axis=0)) / np.std(covars, axis=0)
    outcome = covars[:, 1].copy()
    covars[:, 1] = 1
    return covars, outcome, pheno_std, sample_filter


def _header(phenotype_name, fields):
    return ("chrom\tpos\talleles\tn_samples_tested\tlocus_filtered\tp_{0}\tcoeff_{0}\t".format(phenotype_name)
            + 'se_{}\tregression_R^2\t'.format(phenotype_name) + '\t'.join(fields) + '\n')


def _write_block(outfile, blk, res, pheno_std, non_major_cutoff):
    """TSV rows of one GPU block (reference associaTR.py:252-304, lafg.py:216-221)."""
    h = blk.h
    lines = []
    for l in range(blk.L):
        sl = blk.allele_slice(l)
        lens = h["allele_len"][sl]
        rounded = [round(float(x), load_and_filter_genotypes.allele_len_precision) for x in lens]
        alleles = ','.join(list(np.unique(rounded).astype(str)))
        counts = res["ac_len"][sl]
        total = float(counts.sum())
        by_len = {}
        for key, c in zip([float(x) for x in lens], counts):
            if c > 0:
                by_len[key] = by_len.get(key, 0) + int(c)
        freq = load_and_filter_genotypes.clean_len_alleles({k: by_len[k] / total for k in sorted(by_len)})
        m = blk.metas[l]
        motif = blk.motif(l)
        details = [motif, str(len(motif)), str(round(float(lens[0]), load_and_filter_genotypes.allele_len_precision)),
                   load_and_filter_genotypes.dict_str({k: '{:.2g}'.format(v) for k, v in freq.items()})]
        pos = m.harmonized_pos if m.harmonized_pos is not None else m.vcf_pos
        head = "{}\t{}\t{}\t{}\t".format(m.chrom, pos, alleles, int(res["n_tested"][l]))
        code = int(res["filter_code"][l])
        if code != _lib.AF_OK:
            reason = _REASONS.get(code) or 'non-major allele count<{}'.format(non_major_cutoff)
            lines.append(head + '{}\tnan\tnan\tnan\tnan\t'.format(reason) + '\t'.join(details) + '\n')
            continue
        coef = res["coef"][l] * pheno_std
        se = res["se"][l] * pheno_std
        lines.append(head + 'False\t' + ("{:." + str(pval_precision) + "e}\t{}\t{}\t{}\t").format(
            res["p"][l], coef, se, res["r2"][l]) + '\t'.join(details) + '\n')
    outfile.write(''.join(lines))
    outfile.flush()


def dosage_classes(blk):
    """Per allele of a block: (class representative, python-rounded length, numpy-rounded length) — the keys the
    reference's dosage branch groups alleles by (lafg.py:171-172 ``round``) and compares best guesses with
    (lafg.py:199 ``np.around``)."""
    lens = blk.h["allele_len"]
    prec = load_and_filter_genotypes.allele_len_precision
    len_round = np.array([round(float(x), prec) for x in lens], dtype=np.float64)
    len_around = np.around(lens.astype(np.float64), prec)
    cls = np.zeros(len(lens), np.int32)
    for l in range(blk.L):
        sl = blk.allele_slice(l)
        first = {}
        for j, v in enumerate(len_round[sl]):
            cls[sl.start + j] = first.setdefault(v, j)
    return cls, len_round, len_around


def _r2(n, sx, sxx, sy, syy, sxy):
    """np.corrcoef(x, y)[0, 1] ** 2 from the sums over n entries (NaN when either side is constant)."""
    with np.errstate(divide='ignore', invalid='ignore'):
        n = np.float64(n)
        cov = np.float64(sxy) - np.float64(sx) * np.float64(sy) / n
        vx = np.float64(sxx) - np.float64(sx) * np.float64(sx) / n
        vy = np.float64(syy) - np.float64(sy) * np.float64(sy) / n
        if not (vx > 0 and vy > 0):
            return np.float64(np.nan)
        r = cov / np.sqrt(vx * vy)
        return np.float64(min(max(r, -1.0), 1.0)) ** 2


def _write_dosage_block(outfile, blk, res, meta, pheno_std, non_major_cutoff, sample_filter):
    """TSV rows of one GPU block with --beagle-dosages (reference associaTR.py:252-304, lafg.py:175-238)."""
    cls, len_round, _ = meta
    lafg = load_and_filter_genotypes
    lines = []
    for l in range(blk.L):
        sl = blk.allele_slice(l)
        lr = len_round[sl]
        n = int(res["n_tested"][l])
        alleles = ','.join(list(np.unique(lr).astype(str)))
        cs = res["class_stats"][sl]
        reps = [j for j in range(len(lr)) if cls[sl.start + j] == j]
        with np.errstate(divide='ignore', invalid='ignore'):
            freq = {float(lr[j]): np.float64(cs[j, 0]) / (2 * n) for j in sorted(reps, key=lambda j: lr[j])}
        r2 = {}
        for j in range(len(lr)):                 # dict order of the reference: first occurrence among the alleles
            key = float(lr[j])
            if key not in r2:
                c = int(cls[sl.start + j])
                r2[key] = _r2(2 * n, cs[c, 2], cs[c, 2], cs[c, 0], cs[c, 1], cs[c, 3])
        ls = res["length_stats"][l]
        length_r2 = _r2(2 * n, ls[0], ls[1], ls[2], ls[3], ls[4])
        if n > 0 and any(cs[j, 2] == 2 * n for j in reps):          # every best guess the same length: see lafg
            tr = trh.TRRecord._from_block(blk, l, blk._records[l])
            curr = sample_filter & tr.GetCalledSamples()
            length_r2 = lafg.flat_locus_length_r2(tr, curr, lafg.dosage_arrays(tr, curr, [float(x) for x in lr]))
        m = blk.metas[l]
        motif = blk.motif(l)
        details = [motif, str(len(motif)), str(round(float(blk.h["allele_len"][sl.start]), lafg.allele_len_precision)),
                   lafg.dict_str({k: '{:.2g}'.format(v) for k, v in freq.items()}),
                   lafg.dict_str(lafg.round_vals(r2, lafg.r2_precision)), str(round(length_r2, lafg.r2_precision))]
        pos = m.harmonized_pos if m.harmonized_pos is not None else m.vcf_pos
        head = "{}\t{}\t{}\t{}\t".format(m.chrom, pos, alleles, n)
        reason = lafg.locus_filter_reason(freq, n, non_major_cutoff, True)
        if not reason and int(res["ncovars_code"][l]) == _lib.AF_NCOVARS:
            reason = _REASONS[_lib.AF_NCOVARS]
        if reason:
            lines.append(head + '{}\tnan\tnan\tnan\tnan\t'.format(reason) + '\t'.join(details) + '\n')
            continue
        coef = res["coef"][l] * pheno_std
        se = res["se"][l] * pheno_std
        lines.append(head + 'False\t' + ("{:." + str(pval_precision) + "e}\t{}\t{}\t{}\t").format(
            res["p"][l], coef, se, res["r2"][l]) + '\t'.join(details) + '\n')
    outfile.write(''.join(lines))
    outfile.flush()


def perform_gwas_helper(outfile, all_samples, get_genotype_iter, phenotype_name, trait_fnames, same_samples,
                        sample_fname, beagle_dosages, plotting_phenotype_fname, paired_genotype_plot,
                        plot_phenotype_residuals, plotting_ci_alphas):
    """The reference's injected-generator seam (associaTR.py:117-422).  ``get_genotype_iter(sample_mask)`` must
    return an object with a ``blocks()`` generator of GPU blocks (see ``perform_gwas``) — the per-locus numpy
    generator protocol of the reference is served by ``load_and_filter_genotypes.load_trs`` for user scripts."""
    if plotting_phenotype_fname:
        raise NotImplementedError("the plotting-phenotype columns are outside the accelerated path")
    covars, outcome, pheno_std, sample_filter = prepare_design(all_samples, trait_fnames, same_samples, sample_fname)
    source = get_genotype_iter(sample_filter.copy())
    fields = list(source.detail_fields)
    if beagle_dosages:
        fields += ['dosage_estimated_r2_per_length_allele', 'r2_length_dosages_vs_best_guess_lengths']
    outfile.write(_header(phenotype_name, fields))
    ctx = source.ctx
    ctx.assoc_set_design(covars, outcome, np.nonzero(sample_filter)[0].astype(np.int32))
    n_loci = 0
    start = time.time()
    sharder = getattr(source, "sharder", None)
    multi = sharder is not None and sharder.comm is not None
    first_block = True
    for blk in source.blocks():
        blk._activate()
        if beagle_dosages:
            if first_block and "AP1" not in (blk._records[0].FORMAT or []):      # reference lafg.py:139-146
                print("--beagle-dosages specified, missing required field AP1 for the TR")
                if "GP" in (blk._records[0].FORMAT or []):
                    print("We could support the GP field, but currently only support the AP fields")
                print("Erroring out")
                sys.exit(1)
            blk.ensure_ap()
            meta = dosage_classes(blk)
            res = ctx.assoc_dosage_ols(*meta)
            write = lambda f: _write_dosage_block(f, blk, res, meta, pheno_std, source.non_major_cutoff, sample_filter)
        else:
            res = ctx.assoc_ols(source.non_major_cutoff)
            write = lambda f: _write_block(f, blk, res, pheno_std, source.non_major_cutoff)
        first_block = False
        if multi:                                   # several GPUs: rank 0 gathers every rank's rows at the end
            import io
            buf = io.StringIO()
            write(buf)
            sharder.add(buf.getvalue())
        else:
            write(outfile)
        n_loci += blk.L
    if multi:
        merged = sharder.finish()
        if merged is not None:
            outfile.write(b"".join(merged).decode("utf-8"))
            outfile.flush()
    total_time = time.time() - start
    if n_loci > 0:
        print("Done.\nTotal loci: {}\nTotal time: {}s\ntime/locus: {}s\n".format(n_loci, total_time, total_time / n_loci),
              flush=True)
    else:
        print("No variants found in the region being looked at\n", flush=True)


class _BlockSource:
    """Blocks of harmonized records of one VCF (the GPU counterpart of load_trs)."""
    detail_fields = ['motif', 'period', 'ref_len', 'allele_frequency']

    def __init__(self, tr_vcf, region, non_major_cutoff, vcftype, period_check, block_size, ctx=None, sharder=None):
        self.tr_vcf, self.region, self.non_major_cutoff = tr_vcf, region, non_major_cutoff
        self.vcftype, self.period_check, self.block_size = vcftype, period_check, block_size
        self.ctx = ctx or _lib.default_context()
        self.sharder = sharder          # several GPUs: only the blocks this rank owns are built

    def _build(self, vcftype, recs):
        if self.sharder is not None and not self.sharder.mine():
            return None
        return _block.build_block(self.ctx, vcftype, recs)

    def blocks(self):
        vcf = cyvcf2.VCF(self.tr_vcf)
        if hasattr(vcf, "_native_block_loci"):
            vcf._native_block_loci = self.block_size     # one native run per GPU block (zero-copy hand-off)
        inferred = trh.InferVCFType(vcf, self.vcftype if self.vcftype else 'auto')
        region_start = None
        it = vcf
        if self.region is not None:
            region_start = int(self.region.split(':')[1].split('-')[0])
            it = vcf(self.region)
        recs = []
        for record in it:
            if region_start is not None and record.POS < region_start:
                continue
            if self.period_check and record.INFO.get('PERIOD') is None:
                continue
            if recs and (len(recs) >= self.block_size or record.ploidy != recs[0].ploidy):
                blk = self._build(inferred.name, recs)
                if blk is not None:
                    yield blk
                recs = []
            recs.append(record)
        if recs:
            blk = self._build(inferred.name, recs)
            if blk is not None:
                yield blk


def perform_gwas(outfname, tr_vcf, phenotype_name, traits_fnames, vcftype, same_samples, sample_fname, region,
                 non_major_cutoff, beagle_dosages, plotting_phenotype_fname, paired_genotype_plot,
                 plot_phenotype_residuals, plotting_ci_alphas, imputed_ukb_strs_paper_period_check, block_size=512):
    """reference associaTR.py:424-470."""
    all_samples = cyvcf2.VCF(tr_vcf).samples
    ctx = _lib.default_context()
    comm = _dist.cli_comm(ctx)                 # several GPUs under torchrun: blocks dealt round-robin, rows gathered on rank 0
    sharder = _dist.BlockSharder(comm) if comm is not None else None
    get_genotype_iter = lambda samples: _BlockSource(tr_vcf, region, non_major_cutoff, vcftype,
                                                     imputed_ukb_strs_paper_period_check, block_size, ctx, sharder)
    rank0 = comm is None or comm.rank == 0
    if rank0:
        print("Writing output to {}.temp".format(outfname), flush=True)
    import io
    with (open(outfname + '.temp', 'w') if rank0 else io.StringIO()) as outfile:
        perform_gwas_helper(outfile, all_samples, get_genotype_iter, phenotype_name, traits_fnames, same_samples,
                            sample_fname, beagle_dosages, plotting_phenotype_fname, paired_genotype_plot,
                            plot_phenotype_residuals, plotting_ci_alphas)
    if rank0:
        print("Moving {}.temp to {}".format(outfname, outfname), flush=True)
        shutil.move(outfname + '.temp', outfname)
        print("Done.", flush=True)
    if comm is not None:
        comm.barrier()
        comm.close()


def run():  # pragma: no cover
    """reference associaTR.py:472-583 (same arguments)."""
    parser = argparse.ArgumentParser(__doc__, formatter_class=utils.ArgumentDefaultsHelpFormatter)
    parser.add_argument('outfile')
    parser.add_argument('tr_vcf')
    parser.add_argument('phenotype_name', help='name of the phenotype being regressed against')
    parser.add_argument('traits', nargs='+', help='.npy 2d float arrays of trait values; see the TRTools documentation')
    parser.add_argument('--vcftype', choices=[str(item) for item in trh.VcfTypes.__members__])
    parser.add_argument('--same-samples', default=False, action='store_true')
    parser.add_argument('--sample-list')
    parser.add_argument('--region', help="Restrict to \"chr:start-end\"")
    parser.add_argument('--non-major-cutoff', type=float, default=20)
    parser.add_argument('--beagle-dosages', action='store_true', default=False)
    parser.add_argument('--plotting-phenotype', help=argparse.SUPPRESS)
    parser.add_argument('--paired-genotype-plot', action='store_true', default=False, help=argparse.SUPPRESS)
    parser.add_argument('--plot-phenotype-residuals', action='store_true', default=False, help=argparse.SUPPRESS)
    parser.add_argument('--plotting-ci-alphas', type=float, nargs='*', default=[], help=argparse.SUPPRESS)
    parser.add_argument('--imputed-ukb-strs-paper-period-check', default=False, action='store_true', help=argparse.SUPPRESS)
    parser.add_argument('--block-size', type=int, default=512, help="Records staged per GPU block")
    parser.add_argument("--version", action="version", version='{}'.format(__version__))
    main(parser.parse_args())


def main(args):
    """reference associaTR.py:585-618."""
    today = datetime.datetime.now().strftime("%Y_%m_%d")
    print('-------Running AssociaTR (trtools_b200 v{}) ----------'.format(__version__))
    print("Run date: {}".format(today))
    print(args, flush=True)
    perform_gwas(args.outfile, args.tr_vcf, args.phenotype_name, args.traits, args.vcftype, args.same_samples,
                 args.sample_list, args.region, args.non_major_cutoff, args.beagle_dosages, args.plotting_phenotype,
                 args.paired_genotype_plot, args.plot_phenotype_residuals, args.plotting_ci_alphas,
                 args.imputed_ukb_strs_paper_period_check, getattr(args, "block_size", 512) or 512)


if __name__ == '__main__':  # pragma: no cover
    run()
