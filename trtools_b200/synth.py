"""
Synthetic HipSTR-format TR blocks (SURVEY.md §8d "Synthetic generator").

Two halves:

* **locus tables** (host, numpy ``default_rng``): per locus a motif, a reference
  allele and 1+Poisson(3) alternate alleles at +-k repeat units (about 10 % with a
  partial repeat -> fractional repeat lengths, about 5 % same-length sequence
  variants, about 30 % of loci carrying flanking bases that the harmonizer must
  trim), plus a cumulative allele-frequency table (Dirichlet(0.7)) in 32-bit fixed
  point.
* **per-call values** (GT / DP / DSTUTTER / DFLANKINDEL / Q): a counter-based hash of
  ``(seed, field, locus, sample)`` built from integer operations only, so that this
  numpy implementation and the CUDA twin in ``csrc/trt_synth.cu``
  (``trt_synth_fill``) produce bit-identical arrays in cyvcf2 layout without ever
  moving the data across PCIe (100k x 50k does not fit through the host).

Layouts are exactly what cyvcf2 returns per record, stacked over loci:
GT int16 ``[L][S][3]`` (allele, allele, phased), DP/DSTUTTER/DFLANKINDEL int32
``[L][S]``, Q float32 ``[L][S]``.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

MAX_ALLELES = 16            # cumulative-frequency table width (A <= 16)
INT32_MISSING = -2147483648

# field ids of the per-call hash streams (shared with csrc/trt_synth.cu)
F_GT0, F_GT1, F_MISS, F_DP, F_DP2, F_FLANK, F_STUT, F_Q = range(8)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_C_GOLD = np.uint64(0x9E3779B97F4A7C15)
_C_M1 = np.uint64(0xBF58476D1CE4E5B9)
_C_M2 = np.uint64(0x94D049BB133111EB)
_C_L = np.uint64(0xD1B54A32D192ED03)
_C_S = np.uint64(0x8CB92BA72F3D8DD7)
_C_F = np.uint64(0xDB4F0B9175AE2165)


def _mix(x):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over='ignore'):
        x = (x + _C_GOLD) & _M64
        x = ((x ^ (x >> np.uint64(30))) * _C_M1) & _M64
        x = ((x ^ (x >> np.uint64(27))) * _C_M2) & _M64
        return x ^ (x >> np.uint64(31))


def call_hash(seed: int, fld: int, locus, sample):
    """u64 hash of (seed, field, locus, sample); broadcasting numpy arrays."""
    with np.errstate(over='ignore'):
        locus = np.asarray(locus, dtype=np.uint64)
        sample = np.asarray(sample, dtype=np.uint64)
        k = _mix(np.uint64(seed) ^ (np.uint64(fld) * _C_F))
        k = _mix(k + locus * _C_L)
        return _mix(k + sample * _C_S)


def _popcount64(x):
    x = x.astype(np.uint64)
    m1, m2, m4 = np.uint64(0x5555555555555555), np.uint64(0x3333333333333333), np.uint64(0x0F0F0F0F0F0F0F0F)
    x = x - ((x >> np.uint64(1)) & m1)
    x = (x & m2) + ((x >> np.uint64(2)) & m2)
    x = (x + (x >> np.uint64(4))) & m4
    with np.errstate(over='ignore'):
        return ((x * np.uint64(0x0101010101010101)) & _M64) >> np.uint64(56)


@dataclass
class SynthLoci:
    """Per-locus tables of a synthetic HipSTR block."""
    seed: int
    n_loci: int
    chrom: List[str]
    pos: np.ndarray            # int32 [L]  VCF POS (1-based, includes leading flank)
    start: np.ndarray          # int32 [L]  INFO START
    end: np.ndarray            # int32 [L]  INFO END
    period: np.ndarray         # int32 [L]
    ref: List[str]
    alts: List[List[str]]
    n_alleles: np.ndarray      # int32 [L]
    cum_freq: np.ndarray       # uint32 [L][MAX_ALLELES] cumulative thresholds (last used = 2^32-1)
    miss_thresh: int = int(0.02 * 2 ** 32)       # whole call '.'
    half_thresh: int = int(0.001 * 2 ** 32)      # half call 'a|.'
    locus_offset: int = 0      # global index of locus 0 (for sharding: hash uses global ids)


_BASES = np.array(list("ACGT"))


def make_loci(n_loci: int, seed: int = 20261017, locus_offset: int = 0, max_alleles: int = MAX_ALLELES,
              flank_fraction: float = 0.3) -> SynthLoci:
    rng = np.random.default_rng([seed, locus_offset, 0x7A57])
    period = rng.integers(1, 7, size=n_loci).astype(np.int32)
    nrep = rng.integers(8, 31, size=n_loci)
    n_alt = np.minimum(1 + rng.poisson(3.0, size=n_loci), max_alleles - 1)
    pos = np.zeros(n_loci, dtype=np.int32)
    start = np.zeros(n_loci, dtype=np.int32)
    end = np.zeros(n_loci, dtype=np.int32)
    refs, alts_all = [], []
    cum = np.zeros((n_loci, MAX_ALLELES), dtype=np.uint32)
    cursor = 10000
    for i in range(n_loci):
        p = int(period[i])
        motif = "".join(_BASES[rng.integers(0, 4, size=p)])
        core = motif * int(nrep[i])
        lead = trail = ""
        if rng.random() < flank_fraction:
            lead = "".join(_BASES[rng.integers(0, 4, size=int(rng.integers(0, 4)))])
            trail = "".join(_BASES[rng.integers(0, 4, size=int(rng.integers(0, 4)))])
        alts = []
        seen = {core}
        tries = 0
        while len(alts) < int(n_alt[i]) and tries < 200:
            tries += 1
            r = rng.random()
            k = int(rng.integers(1, 6)) * (1 if rng.random() < 0.5 else -1)
            reps = max(1, int(nrep[i]) + k)
            a = motif * reps
            if r < 0.10 and p > 1:                       # partial repeat -> fractional length
                a = a + motif[:int(rng.integers(1, p))]
            elif r < 0.15:                               # same-length sequence variant of the ref
                j = int(rng.integers(0, len(core)))
                sub = _BASES[(np.where(_BASES == core[j])[0][0] + int(rng.integers(1, 4))) % 4]
                a = core[:j] + sub + core[j + 1:]
            if a in seen:
                continue
            seen.add(a)
            alts.append(a)
        n_alt[i] = len(alts)
        cursor += int(rng.integers(200, 2000))
        pos[i] = cursor
        start[i] = cursor + len(lead)
        end[i] = start[i] + len(core) - 1
        refs.append(lead + core + trail)
        alts_all.append([lead + a + trail for a in alts])
        A = len(alts) + 1
        w = rng.dirichlet(np.full(A, 0.7))
        c = np.minimum(np.floor(np.cumsum(w) * 2.0 ** 32), 2.0 ** 32 - 1).astype(np.uint64)
        c[A - 1] = 2 ** 32 - 1
        cum[i, :A] = c.astype(np.uint32)
        cum[i, A:] = np.uint32(2 ** 32 - 1)
    return SynthLoci(seed=seed, n_loci=n_loci, chrom=["1"] * n_loci, pos=pos, start=start, end=end,
                     period=period, ref=refs, alts=alts_all, n_alleles=(n_alt + 1).astype(np.int32),
                     cum_freq=cum, locus_offset=locus_offset)


@dataclass
class SynthCalls:
    gt: np.ndarray            # int16 [L][S][3]
    dp: np.ndarray            # int32 [L][S]
    dstutter: np.ndarray      # int32 [L][S]
    dflankindel: np.ndarray   # int32 [L][S]
    q: np.ndarray             # float32 [L][S]


def fill_calls(loci: SynthLoci, n_samples: int, locus_slice: Optional[slice] = None) -> SynthCalls:
    """numpy twin of ``trt_synth_fill`` (csrc/trt_synth.cu): per-call arrays for the loci."""
    sl = locus_slice or slice(0, loci.n_loci)
    lidx = np.arange(loci.n_loci)[sl]
    L = len(lidx)
    gl = (lidx + loci.locus_offset).astype(np.uint64)[:, None]
    s = np.arange(n_samples, dtype=np.uint64)[None, :]
    seed = loci.seed
    u0 = (call_hash(seed, F_GT0, gl, s) >> np.uint64(32)).astype(np.uint32)
    u1 = (call_hash(seed, F_GT1, gl, s) >> np.uint64(32)).astype(np.uint32)
    um = (call_hash(seed, F_MISS, gl, s) >> np.uint64(32)).astype(np.uint32)
    cum = loci.cum_freq[lidx]                                       # [L][16]
    # allele = number of thresholds strictly below u  (first a with u <= cum[a])
    a0 = np.sum(u0[:, :, None] > cum[:, None, :], axis=2).astype(np.int16)
    a1 = np.sum(u1[:, :, None] > cum[:, None, :], axis=2).astype(np.int16)
    gt = np.empty((L, n_samples, 3), dtype=np.int16)
    gt[:, :, 0] = a0
    gt[:, :, 1] = a1
    gt[:, :, 2] = 1
    missing = um < np.uint32(loci.miss_thresh)
    half = (~missing) & (um < np.uint32(loci.miss_thresh + loci.half_thresh))
    gt[:, :, 0][missing] = -1
    gt[:, :, 1][missing] = -2
    gt[:, :, 2][missing] = 0
    gt[:, :, 1][half] = -1
    h_dp = call_hash(seed, F_DP, gl, s)
    h_dp2 = call_hash(seed, F_DP2, gl, s)
    dp = _popcount64(h_dp).astype(np.int64) + (h_dp2 % np.uint64(33)).astype(np.int64) - 16
    dp = np.maximum(dp, 0).astype(np.int32)
    hf = call_hash(seed, F_FLANK, gl, s)
    # Binomial(64, 1/32): AND of five 64-bit words derived from one hash by re-mixing
    f = hf
    acc_f = hf
    for _ in range(4):
        f = _mix(f)
        acc_f = acc_f & f
    dfl = np.minimum(_popcount64(acc_f).astype(np.int32), dp)
    hs = call_hash(seed, F_STUT, gl, s)
    g = hs
    acc_s = hs
    for _ in range(3):
        g = _mix(g)
        acc_s = acc_s & g
    dst = np.minimum(_popcount64(acc_s).astype(np.int32), dp)
    hq = call_hash(seed, F_Q, gl, s)
    qa = (hq >> np.uint64(40))                                      # 24 bits
    qb = (hq >> np.uint64(16)) & np.uint64(0xFFFFFF)                 # 24 bits
    k = np.uint64(2 ** 24 - 1) - ((qa * qb) >> np.uint64(28))
    q = (k.astype(np.float32) / np.float32(16777216.0)).astype(np.float32)
    dp[missing] = INT32_MISSING
    dfl[missing] = INT32_MISSING
    dst[missing] = INT32_MISSING
    q[missing] = np.nan
    return SynthCalls(gt=gt, dp=dp, dstutter=dst, dflankindel=dfl, q=q)


def make_traits(loci: SynthLoci, calls_gt_locus0: np.ndarray, n_samples: int, n_covars: int = 10,
                seed: int = 20261017) -> np.ndarray:
    """float64 [S, 1 + n_covars]: trait = 0.05 * (allele-index sum at locus 0) + N(0,1), then PCs."""
    rng = np.random.default_rng([seed, 0x7124175])
    g0 = np.clip(calls_gt_locus0[:, :2].astype(float), 0, None).sum(axis=1)
    trait = 0.05 * g0 + rng.standard_normal(n_samples)
    pcs = rng.standard_normal((n_samples, n_covars))
    return np.hstack([trait[:, None], pcs])


def vcf_header(sample_names: List[str]) -> str:
    """A HipSTR-style header for writing synthetic blocks as VCF text (small configs only)."""
    lines = [
        "##fileformat=VCFv4.1",
        '##FILTER=<ID=PASS,Description="All filters passed">',
        "##command=HipSTR-synthetic --trtools-b200",
        "##contig=<ID=1,length=249250621>",
        '##INFO=<ID=START,Number=1,Type=Integer,Description="Inclusive start coodinate for the repetitive portion of the reference allele">',
        '##INFO=<ID=END,Number=1,Type=Integer,Description="Inclusive end coordinate for the repetitive portion of the reference allele">',
        '##INFO=<ID=PERIOD,Number=1,Type=Integer,Description="Length of STR motif">',
        '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
        '##FORMAT=<ID=Q,Number=1,Type=Float,Description="Posterior probability of unphased genotype">',
        '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Number of valid reads used for sample\'s genotype">',
        '##FORMAT=<ID=DSTUTTER,Number=1,Type=Integer,Description="Number of reads with a stutter indel in the STR region">',
        '##FORMAT=<ID=DFLANKINDEL,Number=1,Type=Integer,Description="Number of reads with an indel in the regions flanking the STR">',
        "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(sample_names),
    ]
    return "\n".join(lines) + "\n"


def write_vcf(path: str, loci: SynthLoci, calls: SynthCalls, sample_names: Optional[List[str]] = None):
    """Write a synthetic block as HipSTR-format VCF text (per-call formatting vectorised with np.char)."""
    S = calls.gt.shape[1]
    names = sample_names or ["S%06d" % i for i in range(S)]
    with open(path, "w") as f:
        f.write(vcf_header(names))
        for i in range(calls.gt.shape[0]):
            info = "START={};END={};PERIOD={}".format(loci.start[i], loci.end[i], loci.period[i])
            cols = [loci.chrom[i], str(loci.pos[i]), "STR_%d" % (i + loci.locus_offset), loci.ref[i],
                    ",".join(loci.alts[i]) if loci.alts[i] else ".", ".", ".", info, "GT:Q:DP:DSTUTTER:DFLANKINDEL"]
            gt = calls.gt[i]
            a0 = np.where(gt[:, 0] < 0, ".", gt[:, 0].astype(str))
            a1 = np.where(gt[:, 1] < 0, ".", gt[:, 1].astype(str))
            body = np.char.add(np.char.add(np.char.add(a0, np.where(gt[:, 2] != 0, "|", "/")), a1), ":")
            q = np.array(["{:.8g}".format(x) for x in calls.q[i]])
            for arr in (q, calls.dp[i].astype(str), calls.dstutter[i].astype(str)):
                body = np.char.add(np.char.add(body, arr), ":")
            body = np.char.add(body, calls.dflankindel[i].astype(str))
            body = np.where((gt[:, 0] == -1) & (gt[:, 1] == -2), ".", body)       # no call: a lone '.'
            f.write("\t".join(cols) + "\t" + "\t".join(body.tolist()) + "\n")


def allele_tables(loci: SynthLoci, lo: int = 0, hi: Optional[int] = None):
    """Arrays for ``trt_block_set_alleles`` covering loci [lo, hi): (seqs, allele_off, locus_off,
    pos, start, end, period)."""
    hi = loci.n_loci if hi is None else hi
    parts = []
    lens = []
    counts = []
    for i in range(lo, hi):
        alleles = [loci.ref[i]] + loci.alts[i]
        counts.append(len(alleles))
        for a in alleles:
            parts.append(a)
            lens.append(len(a))
    seqs = "".join(parts).encode("ascii")
    allele_off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=allele_off[1:])
    locus_off = np.zeros(hi - lo + 1, dtype=np.int32)
    np.cumsum(counts, out=locus_off[1:])
    return (seqs, allele_off, locus_off, loci.pos[lo:hi].copy(), loci.start[lo:hi].copy(),
            loci.end[lo:hi].copy(), loci.period[lo:hi].copy())
