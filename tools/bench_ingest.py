#!/usr/bin/env python
"""
Ingest measurement: a synthetic HipSTR VCF written as text, then

  1. reader only — records -> GT + DP + DFLANKINDEL + Q arrays, pure-Python text reader (bounded sample of the
     records) vs the C++ block reader (whole file), arrays compared bit for bit on the sample;
  2. with a GPU — the statSTR drop-in end to end (file -> .tab, all statistics) through the C++ reader, and on a
     bounded prefix of the file through the text reader; the two .tab outputs must be identical on that prefix.

    python tools/bench_ingest.py [--loci 2000] [--samples 10000] [--text-sample 64] [--out gpurun_out/ingest.json]

Prints one JSON object.  Host code + (optionally) the CUDA statistics path; nothing here is a bench.py metric.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from trtools_b200 import _lib, cyvcf2_compat as cc, synth          # noqa: E402
from trtools_b200.vcf_ingest import NativeVCF                      # noqa: E402

KEYS = ("DP", "DFLANKINDEL", "Q")


def read_arrays(cls, path, limit=None, threads=None):
    t0 = time.time()
    v = cls(path) if cls is cc.TextVCF else cls(path, threads=threads)
    if cls is NativeVCF:
        v._prefetch = KEYS
    out = []
    for r in v:
        out.append((r.genotype.array(),) + tuple(r.format(k) for k in KEYS))
        if limit and len(out) >= limit:
            break
    return out, time.time() - t0


def statstr_args(vcf, out):
    ns = argparse.Namespace(vcf=vcf, out=out, vcftype="hipstr", samples=None, sample_prefixes=None, region=None,
                            precision=3, nalleles_thresh=0.01, plot_afreq=False, use_length=False,
                            only_passing=False, block_size=512)
    for s in ("thresh", "afreq", "acount", "nalleles", "hwep", "het", "entropy", "mean", "mode", "var", "numcalled"):
        setattr(ns, s, True)
    return ns


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=2000)
    ap.add_argument("--samples", type=int, default=10000)
    ap.add_argument("--text-sample", type=int, default=48, help="records the pure-Python reader is timed on")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="trt_ingest_")
    path = os.path.join(tmp, "synth.vcf")
    loci = synth.make_loci(a.loci)
    calls = synth.fill_calls(loci, a.samples)
    synth.write_vcf(path, loci, calls)
    nbytes = os.path.getsize(path)
    res = {"workload": "synthetic HipSTR VCF text, %d loci x %d samples, %.1f MB, keys GT+%s" %
                       (a.loci, a.samples, nbytes / 1e6, "+".join(KEYS)),
           "host_threads": os.cpu_count()}
    nat, t_nat = read_arrays(NativeVCF, path)
    txt, t_txt = read_arrays(cc.TextVCF, path, limit=a.text_sample)
    same = all(x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x.view(np.uint8), y.view(np.uint8))
               for ra, rb in zip(nat, txt) for x, y in zip(ra, rb))
    # the synthetic arrays themselves are the second witness
    gt_ok = all(np.array_equal(nat[i][0], calls.gt[i]) for i in range(a.loci))
    dp_ok = all(np.array_equal(nat[i][1][:, 0], np.where(calls.gt[i][:, 1] == -2, np.int32(-2147483648), calls.dp[i]))
                for i in range(a.loci))
    res["reader"] = {"native_loci_per_s": a.loci / t_nat, "native_MB_per_s": nbytes / 1e6 / t_nat,
                     "text_reader_loci_per_s": len(txt) / t_txt, "text_reader_sample": len(txt),
                     "speedup": (a.loci / t_nat) / (len(txt) / t_txt),
                     "identical_to_text_reader_on_sample": bool(same),
                     "gt_equals_generator": bool(gt_ok), "dp_equals_generator": bool(dp_ok)}
    if _lib.device_count() > 0:
        from trtools_b200 import statSTR
        # first run pays CUDA context creation and first-use allocations; the later runs are the steady state
        runs = {}
        for label, ra in (("cold", "1"), ("warm", "1"), ("warm_no_readahead", "0"), ("warm_again", "1")):
            os.environ["TRTOOLS_B200_INGEST_READAHEAD"] = ra
            t0 = time.time()
            assert statSTR.main(statstr_args(path, os.path.join(tmp, "native"))) == 0
            runs[label] = time.time() - t0
        os.environ["TRTOOLS_B200_INGEST_READAHEAD"] = "1"
        t_e2e = min(runs["warm"], runs["warm_again"])
        # the text reader on a prefix of the file
        prefix = os.path.join(tmp, "prefix.vcf")
        with open(path) as f, open(prefix, "w") as g:
            n = 0
            for line in f:
                g.write(line)
                n += not line.startswith("#")
                if n >= a.text_sample:
                    break
        os.environ["TRTOOLS_B200_INGEST"] = "python"
        t0 = time.time()
        assert statSTR.main(statstr_args(prefix, os.path.join(tmp, "text"))) == 0
        t_txt_e2e = time.time() - t0
        os.environ["TRTOOLS_B200_INGEST"] = "native"
        got = open(os.path.join(tmp, "native.tab")).read().splitlines()[:a.text_sample + 1]
        want = open(os.path.join(tmp, "text.tab")).read().splitlines()
        res["statSTR_e2e"] = {"native_loci_per_s": a.loci / t_e2e, "seconds": t_e2e,
                              "seconds_by_run": runs,
                              "text_reader_loci_per_s": a.text_sample / t_txt_e2e,
                              "tab_identical_on_prefix": got == want}
    line = json.dumps(res)
    print(line)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
