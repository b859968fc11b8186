#!/usr/bin/env python3
"""
bench.py — TR loci/sec of the hot path on a synthetic 100k-locus x 50k-sample HipSTR block.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--loci L] [--samples S]

A "step" is one pass of the hot path (harmonize kernel + GT scan + FP64 epilogue, results copied to
the host) over the whole synthetic block, which is generated directly in HBM by ``trt_synth_fill``
(30 GB of GT never crosses PCIe).  ``value`` = loci/s with the GT rows resident in HBM; ``e2e`` =
the same pass through the C-ABI with HOST buffers (pinned staging block -> H2D -> kernels -> D2H).
``--impl reference`` times the CPU restatement of the reference's algorithm (oracle port; the
reference itself is pure Python and cannot travel to the GPU box) on a bounded locus sample.
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

SEED = 20261017
STATS6 = ("afreq", "het", "hwep", "mean", "var", "entropy")      # BASELINE.json configs[1]
METRIC = "TR loci/sec (statSTR all; associaTR OLS) 100k×50k samp, 1/2/4/8 GPU"


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's statSTR path on a bounded locus sample
# ---------------------------------------------------------------------------------------------------
def _synth_sub(lo, hi, S, seed, with_fmt):
    from oracle.records import synth_to_loci
    from trtools_b200 import synth
    sl = synth.make_loci(hi, seed=seed)
    calls = synth.fill_calls(sl, S, slice(lo, hi))
    sub = synth.SynthLoci(seed=sl.seed, n_loci=hi - lo, chrom=sl.chrom[lo:hi], pos=sl.pos[lo:hi], start=sl.start[lo:hi],
                          end=sl.end[lo:hi], period=sl.period[lo:hi], ref=sl.ref[lo:hi], alts=sl.alts[lo:hi],
                          n_alleles=sl.n_alleles[lo:hi], cum_freq=sl.cum_freq[lo:hi], locus_offset=lo)
    return synth_to_loci(sub, calls, with_fmt=with_fmt)


def reference_kind():
    """"reference" when the unmodified TRTools code is importable (baseline/_ref, see baseline/install_ref.py, or
    /root/reference in the build container), else "port" (the oracle restatement)."""
    from oracle import ref_import
    return "reference" if ref_import.reference_code_available() else "port"


def _cpu_worker(args):
    lo, hi, S, seed, kind = args
    loci = _synth_sub(lo, hi, S, seed, False)
    rows = []
    if kind == "reference":
        # the reference's own per-record loop body (trtools/statSTR/statSTR.py:576-630) on cyvcf2-layout records
        from oracle import ref_import
        from oracle.records import LocusAsVariant
        ref_import.enable()
        import trtools.utils.tr_harmonizer as rtrh
        import trtools.statSTR.statSTR as rstat
        recs = [LocusAsVariant(l) for l in loci]
        t0 = time.perf_counter()
        for rec in recs:
            tr = rtrh.HarmonizeRecord(rtrh.VcfTypes.hipstr, rec)
            items = [rec.CHROM, rec.POS, rec.POS + len(tr.ref_allele)]
            items += rstat.GetAFreq(tr, [None], uselength=False)
            for fn in (rstat.GetHet, rstat.GetHWEP):
                items += fn(tr, [None], uselength=False)
            items += rstat.GetMean(tr, [None]) + rstat.GetVariance(tr, [None])
            items += rstat.GetEntropy(tr, [None], uselength=False)
            rows.append("\t".join(str(x) for x in items))
        return time.perf_counter() - t0, len(rows)
    from oracle import stats as ostats, trh as otrh
    t0 = time.perf_counter()
    for l in loci:
        h = otrh.harmonize(l)                                           # HarmonizeRecord
        vals = ostats.locus_stats(h, l.gt, STATS6, [None], uselength=False)   # statSTR stat wrappers
        rows.append(ostats.format_row(l.chrom, l.pos, h, vals))          # .tab row
    return time.perf_counter() - t0, len(rows)


def cpu_statstr(n_loci, S, cores, seed=SEED, kind=None):
    """loci/s of the reference's statSTR path with one process per core on disjoint loci (wall clock of the
    timed sections, data generation excluded)."""
    import multiprocessing as mp
    kind = kind or reference_kind()
    per = max(1, n_loci // cores)
    jobs = [(i * per, (i + 1) * per, S, seed, kind) for i in range(cores)]
    ctxmp = mp.get_context("fork")
    with ctxmp.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = max(r[0] for r in res)
    done = sum(r[1] for r in res)
    return done / wall, done, wall


def _cpu_other_worker(args):
    """oracle ports of the dumpSTR and associaTR per-locus paths on a few loci (seconds per locus on one core)."""
    lo, hi, S, seed = args
    from oracle import dumpstr as od, assoc as oassoc, trh as otrh
    loci = _synth_sub(lo, hi, S, seed, True)
    cf = [od.hipstr_flank_indels(0.15), od.min_value("HipSTRCallMinDepth", "DP", 20)]
    lf = [od.LocusFilter("hwe", 1e-4, False)]
    sinfo, linfo = od.new_sample_info(S, cf), od.new_loc_info(lf)
    t0 = time.perf_counter()
    for l in loci:
        h = otrh.harmonize(l)
        r = od.apply_call_filters(l, cf, sinfo)
        od.apply_locus_filters(l, h, r.gt, lf, linfo)
        od.recompute_info(h, r.gt, False)
    t_dump = time.perf_counter() - t0
    rng = np.random.default_rng(seed)
    traits = np.hstack([rng.standard_normal((S, 1)), rng.standard_normal((S, 10))])
    design = oassoc.prepare_design([traits], S, None)
    t0 = time.perf_counter()
    for l in loci:
        h = otrh.harmonize(l)
        oassoc.regress_locus(oassoc.load_locus(l, h, design.sample_filter.copy(), 20), design).to_text()
    t_assoc = time.perf_counter() - t0
    return t_dump, t_assoc, len(loci)


def cpu_other_tools(n_loci, S, cores, seed=SEED):
    import multiprocessing as mp
    per = max(1, n_loci // cores)
    jobs = [(i * per, (i + 1) * per, S, seed) for i in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_other_worker, jobs)
    done = sum(r[2] for r in res)
    return {"dumpSTR": {"value": done / max(r[0] for r in res), "unit": "loci/s", "cores": cores, "kind": "port",
                        "sample": "{} loci x {} samples".format(done, S)},
            "associaTR": {"value": done / max(r[1] for r in res), "unit": "loci/s", "cores": cores, "kind": "port",
                          "sample": "{} loci x {} samples".format(done, S)}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = reference_kind()
    L, S = args.loci, args.samples
    n_sub = cores * (1 if kind == "reference" else 4)      # ~1.5 s (reference) / ~0.9 s (port) per locus per core at S = 50k
    vals = []
    for _ in range(min(args.warmup, 1)):
        cpu_statstr(cores, S, cores, kind=kind)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, done, wall = cpu_statstr(n_sub, S, cores, kind=kind)
        vals.append(v)
    total = time.perf_counter() - t0
    value = float(np.mean(vals))
    what = ("the UNMODIFIED reference (trtools.utils.tr_harmonizer.HarmonizeRecord + trtools.statSTR.statSTR stat functions, "
            "baseline/_ref) on cyvcf2-layout records") if kind == "reference" else "oracle port of the reference's Python/numpy path"
    sample = "{} loci x {} samples per step (of the {}-locus workload), {} processes".format(n_sub, S, L, cores)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "loci/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "statSTR --afreq --het --hwep --mean --var --entropy, synthetic HipSTR {}x{}".format(L, S),
                   "loci": L, "samples": S, "note": what + "; VCF parsing excluded; loci/s extrapolates linearly "
                   "(loci are independent)"},
        "cpu_baseline": {"value": value, "unit": "loci/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from trtools_b200 import _lib, synth, dist as tdist
    numa_cpus = tdist.bind_to_gpu_numa_node(local_rank) if (world > 1 and not args.no_numa_bind) else None
    dist = tdist.init("nccl") if world > 1 else None

    ctx = _lib.Context(local_rank)
    info = ctx.device_info()
    L, S = args.loci, args.samples
    # weak scaling: every rank owns its own L loci (global locus ids rank*L .. rank*L+L-1)
    loci = synth.make_loci(L, seed=SEED, locus_offset=rank * L)
    tables = synth.allele_tables(loci)
    ctx.block_begin(L, S, 2, "hipstr")
    ctx.synth_fill(SEED, rank * L, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=False)
    ctx.block_set_alleles(*tables)

    def step():
        ctx.check(ctx.lib.trt_harmonize(ctx.h))
        return ctx.locus_stats(False, None, 0.01, pinned=True)      # results land in page-locked host arrays

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.synchronize()

    STAT_COLS = ("thresh", "het", "entropy", "mean", "mode", "var", "hwep")
    plan = tdist.GatherPlan(dist, len(STAT_COLS), L) if dist is not None else None

    def gather_rows(st):
        """NCCL gather of the fixed-width per-locus result table on rank 0 (north_star: the only collective)."""
        if plan is not None:
            plan.gather([st[k][0] for k in STAT_COLS], wait=False)     # in flight under the next step's kernels

    for _ in range(max(args.warmup, 0)):
        gather_rows(step())
    sampler = ClockSampler(local_rank)
    scan_ms = []
    barrier()
    sampler.start()
    launches0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    ctx.stopwatch_start()
    for _ in range(args.steps):
        st = step()
        scan_ms.append(ctx.last_scan_ms())
        gather_rows(st)
    ms = ctx.stopwatch_stop()
    if plan is not None:
        plan.wait()                                     # the last gather belongs to the timed region
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1000.0
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    # the device stopwatch only spans this rank's stream; use the larger of device and wall time,
    # then the max over ranks
    ms = max(ms, 0.0)
    step_ms = max(ms, wall_ms) / args.steps
    step_ms = tdist.max_over_ranks(dist, step_ms)
    value = world * L / (step_ms / 1000.0)

    # ---- roofline of the dominant kernel (GT scan): 6 algorithmic bytes per call ----------------
    peak, peak_src = load_peaks()
    scan = float(np.mean(scan_ms)) if scan_ms else float("nan")
    algo_bytes = 6.0 * L * S
    achieved = algo_bytes / (scan / 1000.0) / 1e9 if scan > 0 else float("nan")
    traffic = None
    tp = os.path.join(REPO, "profiles", "scan_traffic.json")
    if os.path.exists(tp):
        try:
            # measured DRAM bytes per call of the scan (one ncu --set full capture), scaled to this launch's calls
            traffic = float(json.load(open(tp))["dram_bytes_per_call"]) * L * S
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "scan_pairs_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": scan, "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_share_of_step": scan / step_ms if step_ms > 0 else None}

    # ---- e2e through the C-ABI with HOST buffers (rank-local; max over ranks) -------------------
    Lb = min(L, args.e2e_block)
    nblk = (L + Lb - 1) // Lb
    host_gt = ctx.pinned_empty((Lb, S, 3), np.int16)
    host_gt[...] = ctx.block_get_gt(0, Lb)                     # untimed: fill the pinned staging block
    # the pinned block holds loci [0, Lb): every streamed block re-sends it with its own allele tables
    blk_tables = [synth.allele_tables(loci, 0, min(Lb, L - b * Lb)) for b in range(nblk)]
    h2d = d2h = 0

    # two contexts (own stream + device buffers each) ping-pong over the blocks: while one block's kernels and result
    # copies run, the next block's host->device copy is already in flight on the other stream
    ctxs = [ctx, _lib.Context(local_rank)]

    def e2e_step():
        nonlocal h2d, d2h
        h2d = d2h = 0

        def finish(c):
            nonlocal d2h
            c.check(c.lib.trt_harmonize(c.h))
            st_ = c.locus_stats(False, None, 0.01, pinned=True)
            d2h += sum(v.nbytes for v in st_.values())
            return st_

        pending, st = None, None
        for b in range(nblk):
            c = ctxs[b & 1]
            n = blk_tables[b][2].shape[0] - 1
            c.block_begin(n, S, 2, "hipstr")
            c.block_set_gt(host_gt[:n])                       # asynchronous copy from the pinned block
            c.block_set_alleles(*blk_tables[b])
            h2d += host_gt[:n].nbytes + len(blk_tables[b][0]) + sum(a.nbytes for a in blk_tables[b][1:])
            if pending is not None:
                st = finish(pending)
            pending = c
        if pending is not None:
            st = finish(pending)
        return st

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()                                                 # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1000.0 / e2e_steps
    e2e_ms = tdist.max_over_ranks(dist, e2e_ms)
    e2e_value = world * L / (e2e_ms / 1000.0)
    ctx.free_pinned(host_gt)
    ctxs[1].close()

    # ---- the other two tools of the metric on the same resident block (device-timed, results copied to host) ----
    tools = {}
    if not args.statstr_only:
        from trtools_b200 import _lib as L_
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.synth_fill(SEED, rank * L, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=5)   # DP + DFLANKINDEL
        ctx.block_set_alleles(*tables)
        ctx.check(ctx.lib.trt_harmonize(ctx.h))
        rng = np.random.default_rng(SEED + rank)
        traits = np.hstack([rng.standard_normal((S, 1)), rng.standard_normal((S, 10))])
        covars = np.hstack([np.full((S, 1), -1.0), traits])
        covars = (covars - covars.mean(axis=0)) / np.maximum(covars.std(axis=0), 1e-300)
        outcome = covars[:, 1].copy()
        covars[:, 1] = 1.0
        ctx.assoc_set_design(covars, outcome, np.arange(S, dtype=np.int32))

        def timed(fn, steps, warm):
            for _ in range(warm):
                fn()
            scan = []
            barrier()
            ctx.stopwatch_start()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
                scan.append(ctx.last_scan_ms())
            dev = ctx.stopwatch_stop()
            wall = (time.perf_counter() - t0) * 1000.0
            ms_ = max(dev, wall) / steps
            return tdist.max_over_ranks(dist, ms_), float(np.mean(scan))

        def pack_step():
            ctx.check(ctx.lib.trt_pack_length_genotypes(ctx.h))
            ctx.synchronize()
            pack_step.kernel = ctx.last_kernel_ms()

        ms_p, _ = timed(pack_step, 3, 1)
        tools["pack"] = {"value": world * L / (ms_p / 1000.0), "unit": "loci/s", "ms_per_step": ms_p,
                         "workload": "packed int16 [L][S][2] length-genotype tensor from the native GT rows (SURVEY.md 8d K1 without "
                                     "the harmonize kernel); the statistics kernels do not need it (they read the native rows)",
                         "kernel_ms": pack_step.kernel, "algorithmic_bytes_per_call": 10,
                         "roofline_frac": (10.0 * L * S / (pack_step.kernel / 1000.0) / 1e9) / peak}
        ms_a, k_a = timed(lambda: ctx.assoc_ols(20.0, pinned=True), max(2, min(args.steps, 5)), 2)
        tools["associaTR"] = {"value": world * L / (ms_a / 1000.0), "unit": "loci/s", "ms_per_step": ms_a,
                              "workload": "trait ~ TR length + 10 covariates (K=12), non-major cutoff 20; allele-count scan + "
                                          "FP64 moments (thread-per-locus TMA tiles) + mask down-dates + solve, results to host",
                              "kernel_ms": k_a, "algorithmic_bytes_per_call": 6,
                              "note": "one read of the native GT is the algorithmic minimum; the path reads it twice (scan + moments) "
                                      "and the moments kernel is at the FP64 ridge (2*(K+2) flop per 6 B)",
                              "roofline_frac": (6.0 * L * S / (k_a / 1000.0) / 1e9) / peak if k_a > 0 else None}
        cf_specs = [(L_.CF_RATIO_GT, L_.FMT_DFLANKINDEL, 0.15), (L_.CF_MIN, L_.FMT_DP, 20)]
        counts = np.zeros((2, S), np.int64)
        numcalls = np.zeros(S, np.int64)
        totaldp = np.zeros(S)

        def dump_step():
            ctx.call_filters(cf_specs, L_.FMT_DP, counts, numcalls, totaldp, want_mask=False, want_trigger=False, want_gt=False)
            k1 = ctx.last_scan_ms()
            ctx.locus_filters([(L_.LF_HWE, 1e-4)], False, pinned=True)
            dump_step.kernel = k1 + ctx.last_scan_ms()

        ms_d, _ = timed(dump_step, max(2, min(args.steps, 5)), 2)
        tools["dumpSTR"] = {"value": world * L / (ms_d / 1000.0), "unit": "loci/s", "ms_per_step": ms_d,
                            "workload": "call filters min-call-DP 20 + max-call-flank-indel 0.15 (masked GT written), locus filter "
                                        "HWE 1e-4 on the masked genotypes, sample/locus accumulators to host",
                            "kernel_ms": dump_step.kernel, "algorithmic_bytes_per_call": 20,
                            "note": "GT 6 + DP 4 + DFLANKINDEL 4 read, masked GT 6 written; the locus statistics re-read the masked "
                                    "GT (6 more bytes of actual traffic)",
                            "roofline_frac": (20.0 * L * S / (dump_step.kernel / 1000.0) / 1e9) / peak}

    if rank == 0:
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            kind = reference_kind()
            n_sub = cores * (1 if kind == "reference" else 4)
            v, done, wall = cpu_statstr(n_sub, S, cores, kind=kind)
            cpu = {"value": v, "unit": "loci/s", "cores": cores, "kind": kind,
                   "sample": "{} loci x {} samples, statSTR 6 stats (sequence grouping), {} processes, {:.1f} s".format(
                       done, S, cores, wall)}
            if not args.statstr_only:
                cpu["tools"] = cpu_other_tools(cores, S, cores)
        ingest = None
        if world == 1 and not args.no_cpu_baseline:
            ingest = ingest_leg(S)
        out = {
            "metric": METRIC, "value": value, "unit": "loci/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "statSTR all 11 stats (sequence grouping) on synthetic HipSTR, {} loci x {} samples per GPU, "
                                   "GT int16 [L][S][3] generated in HBM".format(L, S),
                       "loci_per_gpu": L, "samples": S, "seed": SEED, "parallelism": "loci sharded x{}".format(world),
                       "l2": "inputs ({:.1f} GB) far larger than L2; no flush needed".format(algo_bytes / 1e9),
                       "e2e": "one pinned {}-locus host block (loci 0..{}) streamed {}x per step through two ping-pong contexts; every copy is a real H2D".format(Lb, Lb - 1, nblk),
                       "device": info["name"], "sm_count": info["sm_count"],
                       "numa_bound_cpus": numa_cpus},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "loci/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches), "clocks": clocks, "tools": tools, "ingest": ingest,
            "timing": {"device_ms_total": ms, "wall_ms_total": wall_ms},
        }
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def ingest_leg(S, n_loci=32):
    """The step before the path, reported beside it (not part of `value` / `e2e`, which start from arrays —
    SURVEY.md 8d): a bounded sample of the workload written as VCF text, read back through the C++ block reader
    (csrc/trt_ingest.cpp) into the arrays the kernels take, and checked against the generator's arrays.  Host work
    only; never allowed to break the bench line."""
    import shutil
    import tempfile
    import time
    tmp = None
    try:
        from trtools_b200 import synth
        from trtools_b200.vcf_ingest import NativeVCF
        loci = synth.make_loci(n_loci, seed=SEED)
        calls = synth.fill_calls(loci, S)
        tmp = tempfile.mkdtemp(prefix="trt_bench_ingest_")
        path = os.path.join(tmp, "sample.vcf")
        synth.write_vcf(path, loci, calls)
        nbytes = os.path.getsize(path)
        best = None
        ok = True
        for _ in range(3):
            t0 = time.time()
            v = NativeVCF(path)
            v._prefetch = ("DP", "DFLANKINDEL", "Q")
            v._native_block_loci = n_loci
            recs = list(v)
            recs[0]._nblk.parse(v._prefetch)       # one pass: GT + the three keys of all records of the run
            gt = recs[0]._nblk.gt
            dt = time.time() - t0
            ok = ok and gt is not None and gt.shape == calls.gt.shape and bool((gt == calls.gt).all()) and \
                bool((recs[0]._nblk.fmt["DFLANKINDEL"][calls.gt[:, :, 1] != -2] ==
                      calls.dflankindel[calls.gt[:, :, 1] != -2]).all())
            v.close()
            best = dt if best is None else min(best, dt)
        return {"value": n_loci / best, "unit": "loci/s", "text_MB_per_s": nbytes / 1e6 / best,
                "host_threads": os.cpu_count(), "arrays_equal_generator": ok,
                "sample": "{} loci x {} samples as HipSTR VCF text ({:.0f} MB), GT+DP+DFLANKINDEL+Q -> int16/int32/float32 "
                          "arrays, best of 3".format(n_loci, S, nbytes / 1e6)}
    except Exception as e:      # pragma: no cover
        return {"error": "{}: {}".format(type(e).__name__, e)}
    finally:
        if tmp:
            shutil.rmtree(tmp, ignore_errors=True)


_REAL_STDOUT = None


def _protect_stdout():
    """stdout must carry exactly ONE JSON line: route everything else that writes to fd 1 (NCCL's version banner,
    library chatter) to stderr and keep a private handle for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--loci", type=int, default=100000)
    ap.add_argument("--samples", type=int, default=50000)
    ap.add_argument("--e2e-block", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin each rank to its GPU's NUMA-local CPUs")
    ap.add_argument("--statstr-only", action="store_true", help="skip the dumpSTR / associaTR measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
