#!/usr/bin/env python3
"""
bench.py — TR loci/sec of the hot path on a synthetic 100k-locus x 50k-sample HipSTR block.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--loci L] [--samples S]
    python bench.py --tool associaTR --loci 1000000 --samples 500000 [--gpus N]      # BASELINE configs[4] (C5)

Default arm: a "step" is one statSTR pass of the hot path (harmonize kernel + GT scan + FP64 epilogue, all 11
statistics, results copied to the host) over the whole synthetic block, which is generated directly in HBM by
``trt_synth_fill`` (30 GB of GT never crosses PCIe).  ``value`` = loci/s with the GT rows resident in HBM; ``e2e`` =
the same pass through the C-ABI with HOST buffers (pinned staging blocks -> H2D -> kernels -> D2H).  ``tools`` carries
the associaTR and dumpSTR passes of the metric on the same block.  ``parity`` compares, in the same run, the GPU rows
with the rows the unmodified reference computes for a sample of the same loci.

``--impl reference`` times the UNMODIFIED reference (baseline/_ref; oracle port only if it is missing) on the box's
host cores: statSTR (the headline ``value``), plus associaTR and dumpSTR under ``tools`` — same config / metric.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement".
"""
import argparse
import json
import math
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
METRIC = "TR loci/sec (statSTR all; associaTR OLS) 100k×50k samp, 1/2/4/8 GPU"
STAT_F64 = ("thresh", "het", "entropy", "mean", "mode", "var", "hwep")       # order of TRT_REGION_STATS
REL_TOL = 1e-6                                                              # north_star: floats within 1e-6 relative


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(L, S, world):
    """The `config` object — identical on the GPU arm and the reference arm."""
    return {
        "workload": "statSTR all 11 statistics (thresh afreq acount nalleles hwep het entropy mean mode var numcalled; "
                    "sequence grouping) on synthetic HipSTR, {} loci x {} samples per GPU; tools: associaTR (trait ~ TR length "
                    "+ 10 PCs, non-major cutoff 20) and dumpSTR (min-call-DP 20, max-call-flank-indel 0.15, min-locus-hwep "
                    "1e-4) on the same block".format(L, S),
        "loci_per_gpu": L, "samples": S, "seed": SEED, "parallelism": "loci sharded x{}".format(world),
        "layout": "cyvcf2 arrays: GT int16 [L][S][3], DP / DFLANKINDEL int32 [L][S]; VCF text parsing excluded on both arms",
        "l2": "inputs ({:.1f} GB of GT per pass) far larger than the 126 MB L2; no flush needed".format(6.0 * L * S / 1e9),
    }


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread in this process (a sample
    every few milliseconds; the timed region of the default run is tens of milliseconds), nvidia-smi as the fallback.
    ``start`` is called before the barrier that opens the timed region, so starting it cannot skew the ranks."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index, interval_s=0.004):
        self.device_index = device_index
        self.interval_s = interval_s
        self.proc = None
        self.lines = []
        self.samples = []          # (t, sm_mhz, reason bits)
        self.sm_max = None
        self.nvml = None
        self.t0 = self.t1 = None
        self._stop = threading.Event()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.device_index])
            except (ValueError, IndexError):
                pass
        return self.device_index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self._physical_index())], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def _poll(self):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((time.perf_counter(), sm, bits))
            except Exception:
                pass
            self._stop.wait(self.interval_s)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=1.0)
            nv = self.nvml
            names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
            inside = [x for x in self.samples if self.t0 is None or (self.t0 <= x[0] <= (self.t1 or x[0]))]
            used = inside if inside else self.samples       # a region shorter than one poll: the samples around it
            reasons = sorted({n for _, _, bits in used for n, m in names if bits & m})
            return {"sm_mhz": float(np.median([x[1] for x in used])) if used else None, "sm_max_mhz": self.sm_max,
                    "reasons": reasons, "samples": len(used), "samples_in_timed_region": len(inside), "source": "nvml"}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/ref_arm.py) on a bounded locus sample, one process per core
# ---------------------------------------------------------------------------------------------------
def cpu_arm(loci_tables, S, cores, tools=True, steps=1, warmup=0):
    """-> dict(statSTR=..., associaTR=..., dumpSTR=...) each {value, unit, cores, kind, sample, rows}.  statSTR runs
    `steps` times (mean loci/s); the other tools once.  One locus per core per run (~1.5 s / 0.6 s / 0.2 s per locus
    at S = 50k on one core)."""
    from oracle import ref_arm
    ref_arm.set_loci(loci_tables)
    kind = ref_arm.reference_kind()
    n_sub = min(cores, loci_tables.n_loci)
    out = {}
    vals, walls, rows = [], [], None
    for i in range(warmup + steps):
        v, done, wall, res = ref_arm.run_pool(ref_arm.statstr_worker, n_sub, S, cores, kind)
        if i >= warmup:
            vals.append(v)
            walls.append(wall)
            rows = [r for part in res for r in part]
    out["statSTR"] = {"value": float(np.mean(vals)), "unit": "loci/s", "cores": cores, "kind": kind,
                      "sample": "{} loci x {} samples per step, {} processes, {:.1f} s per step".format(n_sub, S, cores, float(np.mean(walls))),
                      "rows": rows, "seconds": float(np.sum(walls))}
    if tools:
        tp = ref_arm.save_traits(ref_arm.bench_traits(S, SEED))
        try:
            v, done, wall, res = ref_arm.run_pool(ref_arm.assoc_worker, n_sub, S, cores, kind, extra=(tp,))
        finally:
            os.remove(tp)
        out["associaTR"] = {"value": v, "unit": "loci/s", "cores": cores, "kind": kind,
                            "sample": "{} loci x {} samples, {} processes, {:.1f} s".format(done, S, cores, wall),
                            "rows": [r for part in res for r in part]}
        v, done, wall, res = ref_arm.run_pool(ref_arm.dumpstr_worker, n_sub, S, cores, kind)
        out["dumpSTR"] = {"value": v, "unit": "loci/s", "cores": cores, "kind": kind,
                          "sample": "{} loci x {} samples, {} processes, {:.1f} s".format(done, S, cores, wall),
                          "parts": res}
    return out


def _strip(d):
    return {k: v for k, v in d.items() if k not in ("rows", "parts", "seconds")}


REFERENCE_NOTE = ("the UNMODIFIED reference (trtools.utils.tr_harmonizer.HarmonizeRecord + trtools.statSTR.statSTR / "
                  "trtools.dumpSTR.dumpSTR.ApplyCallFilters+ApplyLocusFilters / trtools.associaTR.associaTR.perform_gwas_helper "
                  "over load_trs; baseline/_ref) on cyvcf2-layout records; loci/s extrapolates linearly (loci are independent)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from trtools_b200 import synth
    cores = os.cpu_count() or 1
    L, S = args.loci, args.samples
    loci = synth.make_loci(L, seed=SEED)
    t0 = time.perf_counter()
    res = cpu_arm(loci, S, cores, tools=not args.statstr_only, steps=max(args.steps, 1), warmup=min(args.warmup, 1))
    total = time.perf_counter() - t0
    st = res["statSTR"]
    out = {
        "impl": "reference", "metric": METRIC, "value": st["value"], "unit": "loci/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * st["seconds"] / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(L, S, max(args.gpus, 1)),
        "cpu_baseline": dict(_strip(st), note=REFERENCE_NOTE if st["kind"] == "reference" else "oracle port of the reference's numpy path"),
        "tools": {k: _strip(v) for k, v in res.items() if k != "statSTR"},
        "e2e": {"value": st["value"], "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": total,
    }
    emit(out)


# ---------------------------------------------------------------------------------------------------
# parity of the GPU rows against the reference rows of the same loci (same run)
# ---------------------------------------------------------------------------------------------------
class Parity:
    def __init__(self):
        self.max_rel = 0.0
        self.int_mismatch = 0
        self.n_float = 0
        self.n_int = 0
        self.worst = None

    def f(self, got, want, what, abs_tol=0.0):
        got, want = float(got), float(want)
        self.n_float += 1
        if math.isnan(got) or math.isnan(want):
            rel = 0.0 if (math.isnan(got) and math.isnan(want)) else float("inf")
        elif got == want:
            rel = 0.0
        elif abs(got - want) <= abs_tol:
            rel = 0.0
        else:
            rel = abs(got - want) / max(abs(got), abs(want))
        if rel > self.max_rel:
            self.max_rel, self.worst = rel, what

    def i(self, got, want, what):
        self.n_int += 1
        if got != want:
            self.int_mismatch += 1
            if self.worst is None or not str(self.worst).startswith("int"):
                self.worst = "int " + what + ": {} vs {}".format(got, want)

    def result(self, n_loci, kind):
        return {"loci": n_loci, "max_rel": self.max_rel, "int_mismatch": self.int_mismatch, "floats_compared": self.n_float,
                "ints_compared": self.n_int, "tolerance": REL_TOL, "ok": bool(self.max_rel <= REL_TOL and self.int_mismatch == 0),
                "worst": self.worst, "against": kind}


def parity_statstr(st, locus_off, rows, kind):
    p = Parity()
    for j, r in enumerate(rows):
        for k in STAT_F64:
            p.f(st[k][j], r[k], "statSTR {} locus {}".format(k, j), abs_tol=1e-300)
        p.i(int(st["nalleles"][j]), r["nalleles"], "nalleles locus %d" % j)
        p.i(int(st["n_called"][j]), r["numcalled"], "numcalled locus %d" % j)
        p.i(st["ac"][locus_off[j]:locus_off[j + 1]].tolist(), r["ac"], "allele counts locus %d" % j)
    return p.result(len(rows), kind)


def parity_assoc(res, pheno_std, K, rows, kind):
    import scipy.stats
    from trtools_b200 import _lib
    reasons = {_lib.AF_NO_CALLED: 'No called samples', _lib.AF_ONE_ALLELE: 'Only one called allele',
               _lib.AF_NCOVARS: 'n covars >= n samples', _lib.AF_NON_MAJOR: 'non-major allele count<20'}
    p = Parity()
    for j, r in enumerate(rows):
        p.i(int(res["n_tested"][j]), r["n_tested"], "n_tested locus %d" % j)
        code = int(res["filter_code"][j])
        p.i("False" if code == _lib.AF_OK else reasons[code], r["filtered"], "locus_filtered locus %d" % j)
        if code != _lib.AF_OK or r["filtered"] != "False":
            continue
        p.f(res["coef"][j] * pheno_std, r["coef"], "assoc coef locus %d" % j)
        p.f(res["se"][j] * pheno_std, r["se"], "assoc se locus %d" % j)
        p.f(res["r2"][j], r["r2"], "assoc r2 locus %d" % j, abs_tol=1e-12)
        # the reference prints p with 3 significant digits; its full-precision value is 2 T_df.sf(|coef/se|)
        want_p = 2.0 * scipy.stats.t.sf(abs(r["coef"] / r["se"]), r["n_tested"] - K)
        p.f(res["p"][j], want_p, "assoc p locus %d" % j, abs_tol=1e-300)
        p.i("{:.2e}".format(res["p"][j]), r["p_text"], "printed p locus %d" % j)
    return p.result(len(rows), kind)


def parity_dumpstr(ctx, loci, S, lf, locus_off, parts, kind):
    """per-locus values from the full-size pass (lf) + per-sample accumulators from a pass over the sampled loci only"""
    from trtools_b200 import _lib, synth
    p = Parity()
    j = 0
    for part in parts:
        for r in part["per_locus"]:
            names = []
            if int(lf["flags"][j]) & 1:
                names.append(part["locus_filter_names"][0])
            if int(lf["flags"][j]) & 0x80000000:
                names.append("NO_CALLS_REMAINING")
            p.i(";".join(names) if names else "PASS", r["filter"], "FILTER locus %d" % j)
            p.i(lf["ac"][locus_off[j]:locus_off[j + 1]].tolist(), r["AC"], "AC/REFAC locus %d" % j)
            p.i(int(lf["n_called"][j]), r["n_called"], "n_called locus %d" % j)
            p.i(int(lf["hrun"][j]), r["HRUN"], "HRUN locus %d" % j)
            p.f(lf["het"][j], r["HET"], "INFO HET locus %d" % j)
            p.f(lf["hwep"][j], r["HWEP"], "INFO HWEP locus %d" % j, abs_tol=1e-300)
            j += 1
    n = j
    # per-sample counters of the sample log over exactly the sampled loci
    ctx.block_begin(n, S, 2, "hipstr")
    ctx.synth_fill(SEED, loci.locus_offset, loci.cum_freq[:n], loci.miss_thresh, loci.half_thresh, with_format=5)
    ctx.block_set_alleles(*synth.allele_tables(loci, 0, n))
    ctx.check(ctx.lib.trt_harmonize(ctx.h))
    specs = [(_lib.CF_RATIO_GT, _lib.FMT_DFLANKINDEL, 0.15), (_lib.CF_MIN, _lib.FMT_DP, 20)]
    counts = np.zeros((2, S), np.int64)
    numcalls = np.zeros(S, np.int64)
    totaldp = np.zeros(S)
    ctx.call_filters(specs, _lib.FMT_DP, counts, numcalls, totaldp, want_mask=False, want_trigger=False, want_gt=False)
    want_nc = sum(part["numcalls"] for part in parts)
    want_dp = sum(part["totaldp"] for part in parts)
    p.i(bool(np.array_equal(numcalls, want_nc)), True, "samplog numcalls")
    p.i(bool(np.array_equal(totaldp, want_dp, equal_nan=True)), True, "samplog totaldp")
    for f, key in enumerate(("HipSTRCallFlankIndels0.15", "HipSTRCallMinDepth20")):
        want = sum(part["counts"][key] for part in parts)
        p.i(bool(np.array_equal(counts[f], want)), True, "samplog " + key)
    return p.result(n, kind)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def _design(S, seed):
    rng = np.random.default_rng(seed)
    traits = np.hstack([rng.standard_normal((S, 1)), rng.standard_normal((S, 10))])
    covars = np.hstack([np.full((S, 1), -1.0), traits])
    pheno_std = float(np.std(covars[:, 1]))
    covars = (covars - covars.mean(axis=0)) / np.maximum(covars.std(axis=0), 1e-300)
    outcome = covars[:, 1].copy()
    covars[:, 1] = 1.0
    return covars, outcome, pheno_std


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from trtools_b200 import _lib, synth, dist as tdist
    numa_cpus = tdist.bind_to_gpu_numa_node(local_rank) if (world > 1 and not args.no_numa_bind) else None
    ctx = _lib.Context(local_rank)
    comm = tdist.init(ctx) if world > 1 else None
    info = ctx.device_info()
    L, S = args.loci, args.samples
    # weak scaling: every rank owns its own L loci (global locus ids rank*L .. rank*L+L-1)
    loci = synth.make_loci(L, seed=SEED, locus_offset=rank * L)
    tables = synth.allele_tables(loci)
    ctx.block_begin(L, S, 2, "hipstr")
    ctx.synth_fill(SEED, rank * L, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=False)
    ctx.block_set_alleles(*tables)
    nA = ctx.nA
    locus_off = ctx.locus_off.copy()

    def barrier():
        if comm is not None:
            comm.barrier()
        ctx.synchronize()

    # ---- NCCL gather of the per-locus result table on rank 0: device buffers, no host bounce ------------------
    # rows gathered per step: the seven float64 statistics + nalleles + n_hom + n_called (10 x 8 B x L) and the allele
    # counts (int32 [nA]) — everything statSTR prints
    gathered = None
    if comm is not None:
        sizes = comm.allgather_i64([10 * 8 * L, 4 * nA])
        if rank == 0:
            gathered = (ctx.pinned_empty((int(sizes[:, 0].sum()),), np.uint8), ctx.pinned_empty((int(sizes[:, 1].sum()),), np.uint8))

    def step():
        ctx.check(ctx.lib.trt_harmonize(ctx.h))
        if comm is None:
            return ctx.locus_stats(False, None, 0.01, pinned=True)      # results land in page-locked host arrays
        ctx.locus_stats(False, None, 0.01, want=())                     # results stay in HBM ...
        comm.gather_region(tdist.REGION_STATS, 0, 10 * 8 * L, sizes[:, 0], 0, None if gathered is None else gathered[0], wait=False)
        comm.gather_region(tdist.REGION_ALLELE_COUNTS, 0, 4 * nA, sizes[:, 1], 0, None if gathered is None else gathered[1], wait=False)
        return None                                                     # ... and travel to rank 0 under the next step

    sampler = ClockSampler(local_rank)
    sampler.start()                                     # before the warm-up: nothing but the barrier precedes the clock
    for _ in range(max(args.warmup, 0)):
        step()
    scan_ms = []
    barrier()
    sampler.mark_begin()
    launches0 = ctx.launch_count()
    t_wall0 = time.perf_counter()
    ctx.stopwatch_start()
    st = None
    for _ in range(args.steps):
        st = step()
        scan_ms.append(ctx.last_scan_ms())
    if comm is not None:
        comm.wait()                                     # the last gather and its host copy belong to the timed region
    ms = ctx.stopwatch_stop()
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1000.0
    sampler.mark_end()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    # the device stopwatch only spans this rank's stream; use the larger of device and wall time, then the max over ranks
    step_ms = max(ms, wall_ms) / args.steps
    per_rank_us = None if comm is None else comm.allgather_i64([int(ms / args.steps * 1000), int(wall_ms / args.steps * 1000)]).tolist()
    step_ms = tdist.max_over_ranks(comm, step_ms)
    value = world * L / (step_ms / 1000.0)
    if comm is not None:
        # rank 0 now holds every rank's table; unpack its own share for the parity check below
        st = ctx.locus_stats(False, None, 0.01, pinned=True)
        if rank == 0:
            own = np.frombuffer(gathered[0][:10 * 8 * L], dtype=np.float64).reshape(10, L)
            assert np.array_equal(own[1], st["het"][0], equal_nan=True) and np.array_equal(own[6], st["hwep"][0], equal_nan=True)
            last = np.frombuffer(gathered[0][-10 * 8 * L:], dtype=np.float64).reshape(10, L)
            assert np.isfinite(last[3]).any()           # the last rank's rows arrived
    st = {k: v[0].copy() for k, v in st.items()}

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
                # the measured peak is a device-to-device COPY (reads + writes); a pure read stream can exceed it
                "frac_of_nominal_7700_GBs": achieved / 7700.0,
                "kernel_share_of_step": scan / step_ms if step_ms > 0 else None}

    # ---- e2e through the C-ABI with HOST buffers (rank-local; max over ranks) -------------------
    # headline: the packed transfer form the library's own block reader parses VCF text into (2 B/call across PCIe);
    # beside it the same pass from cyvcf2-layout int16 arrays (6 B/call), the form a cyvcf2 caller holds
    def regen():        # the e2e legs stream small blocks through this context: restore the full synthetic block first
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.synth_fill(SEED, rank * L, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=False)
        ctx.block_set_alleles(*tables)

    e2e = run_e2e(args, ctx, comm, loci, L, S, local_rank, world, barrier, packed=True, nibble=True)
    regen()
    e2e["packed_2_bytes"] = run_e2e(args, ctx, comm, loci, L, S, local_rank, world, barrier, packed=True)
    regen()
    e2e["cyvcf2_layout"] = run_e2e(args, ctx, comm, loci, L, S, local_rank, world, barrier, packed=False)

    # ---- the other two tools of the metric on the same resident block (device-timed, results copied to host) ----
    tools, assoc_res, lf_res, pheno_std = {}, None, None, 1.0
    if not args.statstr_only:
        L_ = _lib
        ctx.block_begin(L, S, 2, "hipstr")
        ctx.synth_fill(SEED, rank * L, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=5)   # DP + DFLANKINDEL
        ctx.block_set_alleles(*tables)
        ctx.check(ctx.lib.trt_harmonize(ctx.h))
        covars, outcome, pheno_std = _design(S, SEED)
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
            if comm is not None:            # microseconds per step of every rank (device, wall): which rank sets the max, and why
                timed.per_rank = comm.allgather_i64([int(dev / steps * 1000), int(wall / steps * 1000)]).tolist()
            return tdist.max_over_ranks(comm, ms_), float(np.mean(scan))

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

        def assoc_step():
            assoc_step.res = ctx.assoc_ols(20.0, pinned=True)

        timed.per_rank = None
        ms_a, k_a = timed(assoc_step, max(2, min(args.steps, 5)), 2)
        assoc_per_rank = timed.per_rank
        assoc_res = {k: v.copy() for k, v in assoc_step.res.items()}
        tools["associaTR"] = {"value": world * L / (ms_a / 1000.0), "unit": "loci/s", "ms_per_step": ms_a,
                              "workload": "trait ~ TR length + 10 covariates (K=12), non-major cutoff 20; allele counts (GT scan) + "
                                          "exact integer cross moments on the tensor cores (u8 x s8 mma.sync, 7 base-256 digits per "
                                          "design column) + mask down-dates + FP64 solve, results to host",
                              "kernel_ms": k_a, "algorithmic_bytes_per_call": 6, "per_rank_us_device_wall": assoc_per_rank,
                              "roofline_frac": (6.0 * L * S / (k_a / 1000.0) / 1e9) / peak if k_a > 0 else None}
        cf_specs = [(L_.CF_RATIO_GT, L_.FMT_DFLANKINDEL, 0.15), (L_.CF_MIN, L_.FMT_DP, 20)]
        counts = np.zeros((2, S), np.int64)
        numcalls = np.zeros(S, np.int64)
        totaldp = np.zeros(S)

        def dump_step():
            ctx.call_filters(cf_specs, L_.FMT_DP, counts, numcalls, totaldp, want_mask=False, want_trigger=False, want_gt=False)
            k1 = ctx.last_scan_ms()
            dump_step.res = ctx.locus_filters([(L_.LF_HWE, 1e-4)], False, pinned=True)
            dump_step.kernel = k1 + ctx.last_scan_ms()

        ms_d, _ = timed(dump_step, max(2, min(args.steps, 5)), 2)
        lf_res = {k: v.copy() for k, v in dump_step.res.items()}
        tools["dumpSTR"] = {"value": world * L / (ms_d / 1000.0), "unit": "loci/s", "ms_per_step": ms_d,
                            "workload": "call filters min-call-DP 20 + max-call-flank-indel 0.15 (masked GT written), locus filter "
                                        "HWE 1e-4 on the masked genotypes, sample/locus accumulators to host",
                            "kernel_ms": dump_step.kernel, "algorithmic_bytes_per_call": 20,
                            "note": "GT 6 + DP 4 + DFLANKINDEL 4 read, masked GT 6 written",
                            "roofline_frac": (20.0 * L * S / (dump_step.kernel / 1000.0) / 1e9) / peak}

    if rank == 0:
        cores = os.cpu_count() or 1
        cpu, parity, ingest = None, None, None
        if world == 1 and not args.no_cpu_baseline:
            res = cpu_arm(loci, S, cores, tools=not args.statstr_only)
            kind = res["statSTR"]["kind"]
            cpu = dict(_strip(res["statSTR"]), note=REFERENCE_NOTE if kind == "reference" else "oracle port")
            parity = parity_statstr(st, locus_off, res["statSTR"]["rows"], kind)
            parity["tools"] = {}
            if not args.statstr_only:
                cpu["tools"] = {k: _strip(v) for k, v in res.items() if k != "statSTR"}
                parity["tools"]["associaTR"] = parity_assoc(assoc_res, pheno_std, 12, res["associaTR"]["rows"], kind)
                parity["tools"]["dumpSTR"] = parity_dumpstr(ctx, loci, S, lf_res, locus_off, res["dumpSTR"]["parts"], kind)
                parity["ok"] = bool(parity["ok"] and all(t["ok"] for t in parity["tools"].values()))
                parity["max_rel"] = max([parity["max_rel"]] + [t["max_rel"] for t in parity["tools"].values()])
                parity["int_mismatch"] += sum(t["int_mismatch"] for t in parity["tools"].values())
            ingest = ingest_leg(S)
        cfg = workload_config(L, S, world)
        out = {
            "metric": METRIC, "value": value, "unit": "loci/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "tools": tools, "ingest": ingest,
            "device": {"name": info["name"], "sm_count": info["sm_count"], "numa_bound_cpus": numa_cpus},
            "gather": None if comm is None else "per-locus rows (10 x 8 B x L + int32 allele counts) of every rank gathered on rank 0 "
                      "with ncclSend/ncclRecv from device buffers on the context stream (trt_dist_gather_region), host copy on a side stream",
            "timing": {"device_ms_total": ms, "wall_ms_total": wall_ms, "per_rank_us_per_step_device_wall": per_rank_us},
        }
        emit(out)
    if comm is not None:
        comm.barrier()
        comm.close()


def run_e2e(args, ctx, comm, loci, L, S, local_rank, world, barrier, packed=True, nibble=False):
    """The statSTR pass through the C-ABI with HOST buffers: pinned host blocks of the workload's genotypes ->
    trt_block_set_gt_packed / trt_block_set_gt (H2D) -> kernels -> D2H of the statistics.  The host blocks hold DISTINCT
    loci of the workload (copied out of the device-generated block before the clock starts), as many as fit the pinned
    budget; when the budget is smaller than the workload the resident blocks are streamed round-robin (every copy is a
    real H2D).  ``packed``: uint8 [L][S][2] blocks (what trt_vcf_block_parse_packed emits) instead of int16 [L][S][3];
    ``nibble``: uint8 [L][S] blocks (trt_vcf_block_parse_nibble; every locus of the workload has <= 14 alleles)."""
    from trtools_b200 import _lib, synth, dist as tdist
    # the same pinned bytes per block in every form: 4096 loci of cyvcf2 layout = 8192 two-byte = 16384 nibble loci (the
    # per-block fixed cost — table uploads, launches, result copies — is amortised over the loci of the block)
    Lb = min(L, args.e2e_block * (4 if nibble else (2 if packed else 1)))
    nblk = (L + Lb - 1) // Lb
    blk_bytes = Lb * S * (2 if packed else 6)       # (nibble blocks are half of this)
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    # several ranks share one host: keep the sum of the pinned blocks well inside its memory
    budget = min(args.e2e_host_gb * (1 << 30) / (1 if world == 1 else 4), 0.2 * avail / max(world, 1))
    n_host = int(max(1, min(nblk, budget // blk_bytes)))
    host_blocks = []
    for b in range(n_host):
        n = min(Lb, L - b * Lb)
        # like the block reader, a block whose loci all have <= 14 alleles travels as nibbles, any other as two bytes
        if nibble and int(np.max(loci.n_alleles[b * Lb:b * Lb + n])) <= 14:
            hb = ctx.pinned_empty((Lb, S), np.uint8)
            ctx.check(ctx.lib.trt_block_get_gt_nibble(ctx.h, b * Lb, n, hb.ctypes.data, None))
        elif packed:
            hb = ctx.pinned_empty((Lb, S, 2), np.uint8)
            ctx.check(ctx.lib.trt_block_get_gt_packed(ctx.h, b * Lb, n, hb.ctypes.data, None))
        else:
            hb = ctx.pinned_empty((Lb, S, 3), np.int16)
            ctx.check(ctx.lib.trt_block_get_gt(ctx.h, b * Lb, n, hb.ctypes.data))   # untimed: distinct loci b*Lb .. b*Lb+n-1
        host_blocks.append(hb)
    blk_tables = [synth.allele_tables(loci, (b % n_host) * Lb, min((b % n_host) * Lb + min(Lb, L - b * Lb), L)) for b in range(nblk)]
    counters = {"h2d": 0, "d2h": 0}
    # two contexts (own stream + device buffers each) ping-pong over the blocks: while one block's kernels and result
    # copies run, the next block's host->device copy is already in flight on the other stream
    ctxs = [ctx, _lib.Context(local_rank)]

    def e2e_step():
        counters["h2d"] = counters["d2h"] = 0

        def finish(c):
            c.check(c.lib.trt_harmonize(c.h))
            st_ = c.locus_stats(False, None, 0.01, pinned=True)
            counters["d2h"] += sum(v.nbytes for v in st_.values())

        pending = None
        for b in range(nblk):
            c = ctxs[b & 1]
            n = blk_tables[b][2].shape[0] - 1
            hb = host_blocks[b % n_host]
            c.block_begin(n, S, 2, "hipstr")
            # allele tables first: that call returns once its (small, pageable) tables are on the device, and must not
            # wait behind the block's genotype copy
            c.block_set_alleles(*blk_tables[b])
            if hb.ndim == 2:
                c.block_set_gt_nibble(hb[:n])                 # asynchronous copy from the pinned block + expansion kernel
            elif packed:
                c.block_set_gt_packed(hb[:n])
            else:
                c.block_set_gt(hb[:n])
            counters["h2d"] += hb[:n].nbytes + len(blk_tables[b][0]) + sum(a.nbytes for a in blk_tables[b][1:])
            if pending is not None:
                finish(pending)
            pending = c
        if pending is not None:
            finish(pending)

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()                                                 # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1000.0 / e2e_steps
    e2e_ms = tdist.max_over_ranks(comm, e2e_ms)
    host_bytes = sum(hb.nbytes for hb in host_blocks)
    n_nibble = sum(1 for hb in host_blocks if hb.ndim == 2)
    for hb in host_blocks:
        ctx.free_pinned(hb)
    ctxs[1].close()
    return {"value": world * L / (e2e_ms / 1000.0), "unit": "loci/s", "h2d_bytes_per_step": int(counters["h2d"]),
            "d2h_bytes_per_step": int(counters["d2h"]), "ms_per_step": e2e_ms,
            "transfer_form": ("nibble: uint8 [L][S], two 4-bit allele codes per call (trt_block_set_gt_nibble; what "
                              "trt_vcf_block_parse_nibble emits from VCF text when no locus has more than 14 alleles), expanded on "
                              "the device" if nibble else
                              "packed: uint8 [L][S][2] allele codes (trt_block_set_gt_packed; what trt_vcf_block_parse_packed "
                              "emits from VCF text), expanded on the device" if packed else
                              "cyvcf2 layout: int16 [L][S][3] (trt_block_set_gt)"),
            "host_blocks": "{} pinned blocks of {} loci = {} distinct loci ({:.1f} GB) of the workload resident in host memory, "
                           "streamed {} blocks per step through two ping-pong contexts".format(
                               n_host, Lb, min(n_host * Lb, L), host_bytes / 1e9, nblk) +
                           (" ({} of the {} resident blocks as nibbles, the others as two bytes per call)".format(n_nibble, n_host)
                            if nibble else "")}


# ---------------------------------------------------------------------------------------------------
# BASELINE configs[4] (C5): associaTR at biobank scale, loci sharded over the ranks, streamed in device-generated blocks
# ---------------------------------------------------------------------------------------------------
def run_assoc_stream(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from trtools_b200 import _lib, synth, dist as tdist
    ctx = _lib.Context(local_rank)
    comm = tdist.init(ctx) if world > 1 else None
    info = ctx.device_info()
    L, S, B = args.loci, args.samples, args.block_loci
    lo, hi = tdist.locus_shard(L, rank, world)
    # allele tables of ONE block (python-side generator: 0.13 ms per locus) reused by every block; the genotypes of
    # block b are hashed from the GLOBAL locus ids, so every locus of the run has its own calls
    tab = synth.make_loci(min(B, max(hi - lo, 1)), seed=SEED)
    covars, outcome, pheno_std = _design(S, SEED)
    ctx.assoc_set_design(covars, outcome, np.arange(S, dtype=np.int32))
    blocks = [(b0, min(b0 + B, hi)) for b0 in range(lo, hi, B)]
    row_bytes = 5 * 8          # p, coef, se, r2, std_g (f64) per locus: the fixed-width summary row
    # every rank knows every rank's block sizes (contiguous shards, fixed block length): no size exchange per block
    shards = [tdist.locus_shard(L, r, world) for r in range(world)]
    nblk_max = max((h_ - l_ + B - 1) // B for l_, h_ in shards)

    def counts_of(i):
        return np.array([max(0, min(B, h_ - (l_ + i * B))) for l_, h_ in shards], dtype=np.int64)

    recv = [ctx.pinned_empty((int(max(B * world * row_bytes, 16)),), np.uint8) for _ in range(2)] if rank == 0 else [None, None]
    table = np.full((5, L), np.nan) if rank == 0 else None

    def unpack(i, buf):
        counts = counts_of(i)
        off = 0
        for r in range(world):
            c = int(counts[r])
            if c:
                g0 = shards[r][0] + i * B
                table[:, g0:g0 + c] = np.frombuffer(buf[off:off + c * row_bytes], dtype=np.float64).reshape(5, c)
            off += c * row_bytes

    tables_of = {}

    def gen_block(i):
        b0, b1 = blocks[i]
        n = b1 - b0
        if n not in tables_of:
            tables_of[n] = synth.allele_tables(tab, 0, n)
        ctx.block_begin(n, S, 2, "hipstr")
        ctx.synth_fill(SEED, b0, tab.cum_freq[:n], tab.miss_thresh, tab.half_thresh, with_format=False)
        ctx.block_set_alleles(*tables_of[n])
        ctx.synchronize()
        return n

    def run(timed):
        """Per block: [generate] -> harmonize + associaTR kernels -> NCCL gather of the rows (device buffers) whose copy
        to the host runs on the side stream under the NEXT block's generation and kernels."""
        hot_ms = gen_ms = 0.0
        scan_ms = []
        pending = None                                  # (block index, host buffer) of the gather still in flight
        for i in range(nblk_max):
            n = 0
            if i < len(blocks):
                t0 = time.perf_counter()
                n = gen_block(i)
                gen_ms += (time.perf_counter() - t0) * 1000.0
                ctx.stopwatch_start()
                ctx.check(ctx.lib.trt_harmonize(ctx.h))
                res = ctx.assoc_ols(20.0, pinned=True) if comm is None else ctx.assoc_ols(20.0, want=())
                scan_ms.append(ctx.last_scan_ms())
            else:
                ctx.stopwatch_start()
            if comm is not None:
                if pending is not None:                 # the previous block's rows have long arrived: bank them
                    comm.wait()
                    if rank == 0 and timed:
                        unpack(*pending)
                comm.gather_region(tdist.REGION_ASSOC, 0, n * row_bytes, counts_of(i) * row_bytes, 0, recv[i & 1], wait=False)
                pending = (i, recv[i & 1])
            elif timed and n:
                b0, b1 = blocks[i]
                for k, key in enumerate(("p", "coef", "se", "r2", "std_g")):
                    table[k, b0:b1] = res[key]
            hot_ms += ctx.stopwatch_stop()
        if comm is not None and pending is not None:
            t0 = time.perf_counter()
            comm.wait()
            hot_ms += (time.perf_counter() - t0) * 1000.0
            if rank == 0 and timed:
                unpack(*pending)
        return hot_ms, gen_ms, scan_ms

    # warm-up: block 0 through the whole per-block path, untimed — first-use allocations of the 1.4 GB mask / moment
    # buffers, kernel module loading, and NCCL's lazy peer-to-peer connection set-up of the first gather
    if args.warmup > 0:
        n0 = 0
        if blocks:
            n0 = gen_block(0)
            ctx.check(ctx.lib.trt_harmonize(ctx.h))
            ctx.assoc_ols(20.0, want=())
        if comm is not None:
            comm.gather_region(tdist.REGION_ASSOC, 0, n0 * row_bytes, counts_of(0) * row_bytes, 0, recv[0], wait=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    if comm is not None:
        comm.barrier()
    sampler.mark_begin()
    launches0 = ctx.launch_count()
    t0 = time.perf_counter()
    hot_ms, gen_ms, scan_ms = run(True)
    wall_s = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    hot = tdist.max_over_ranks(comm, hot_ms)
    kern = tdist.max_over_ranks(comm, float(np.sum(scan_ms)))
    wall = tdist.max_over_ranks(comm, wall_s)
    peak, peak_src = load_peaks()
    if rank == 0:
        n_ok = int(np.isfinite(table[0]).sum())
        own_bytes = 6.0 * (hi - lo) * S
        out = {
            "metric": METRIC, "value": L / (hot / 1000.0), "unit": "loci/s", "n_gpus": world, "steps": 1, "warmup": 0,
            "ms_per_step": hot, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "associaTR (trait ~ TR length + 10 PCs, cutoff 20) on {} loci x {} samples (BASELINE configs[4]) "
                                   "sharded by locus over {} GPU(s), streamed in device-generated blocks of {} loci; per-locus summary "
                                   "rows gathered on rank 0 with NCCL from device buffers".format(L, S, world, B),
                       "loci": L, "samples": S, "block_loci": B, "seed": SEED,
                       "l2": "each block ({:.1f} GB of GT) is far larger than L2".format(6.0 * min(B, hi - lo) * S / 1e9),
                       "timed": "harmonize + associaTR kernels + NCCL gather + host copy of the rows per block (CUDA events on the "
                                "context stream, summed over blocks, max over ranks); block generation (the stand-in for ingest) is "
                                "reported separately as generation_ms"},
            "roofline": {"bound": "hbm", "kernel": "assoc_tile_kernel (+ down-dates, solve)", "achieved": own_bytes / (kern / 1000.0) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": own_bytes / (kern / 1000.0) / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "kernel_ms": kern, "algorithmic_bytes_per_launch": own_bytes},
            "rows_gathered": n_ok, "loci_tested_ok": n_ok, "generation_ms": tdist.max_over_ranks(None, gen_ms),
            "wall_s": wall, "gpu_launches": int(launches), "clocks": clocks,
            "device": {"name": info["name"], "sm_count": info["sm_count"]},
            "e2e": None, "cpu_baseline": None,
        }
        emit(out)
    if comm is not None:
        comm.barrier()
        comm.close()


def ingest_leg(S, n_loci=32):
    """The step before the path, reported beside it (not part of `value` / `e2e`, which start from arrays —
    SURVEY.md 8d): a bounded sample of the workload written as VCF text, read back through the C++ block reader
    (csrc/trt_ingest.cpp) into the arrays the kernels take, and checked against the generator's arrays.  Host work
    only; never allowed to break the bench line."""
    import shutil
    import tempfile
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
    ap.add_argument("--tool", default="statSTR", choices=["statSTR", "associaTR"],
                    help="associaTR: the streamed biobank-scale run of BASELINE configs[4]")
    ap.add_argument("--loci", type=int, default=100000)
    ap.add_argument("--samples", type=int, default=50000)
    ap.add_argument("--block-loci", type=int, default=16384, help="--tool associaTR: loci per device-generated block")
    ap.add_argument("--e2e-block", type=int, default=4096)
    ap.add_argument("--e2e-host-gb", type=float, default=32.0, help="pinned host memory holding distinct loci for the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin each rank to its GPU's NUMA-local CPUs")
    ap.add_argument("--statstr-only", action="store_true", help="skip the dumpSTR / associaTR measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.tool == "associaTR":
        run_assoc_stream(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
