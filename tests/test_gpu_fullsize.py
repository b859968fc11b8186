"""
Full-size (BASELINE.json: 100 000 loci x 50 000 samples) property tests of the GT scan, on a device-generated block:
idempotence (a cold first pass equals warm passes bit for bit), a per-locus checksum tying the allele counts to the
call counters, and exact agreement with a numpy recount of sampled rows copied back from HBM.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

L_FULL, S_FULL, SEED = 100000, 50000, 20261017


@pytest.fixture(scope="module")
def full_block():
    from trtools_b200 import _lib, synth
    ctx = _lib.Context(0)                      # a fresh context: its first pass is a cold launch
    info = ctx.device_info()
    if info["free_mem_bytes"] < 40e9:
        pytest.skip("needs 40 GB of free HBM")
    loci = synth.make_loci(L_FULL, seed=SEED)
    ctx.block_begin(L_FULL, S_FULL, 2, "hipstr")
    ctx.synth_fill(SEED, 0, loci.cum_freq, loci.miss_thresh, loci.half_thresh, with_format=False)
    ctx.block_set_alleles(*synth.allele_tables(loci))
    ctx.check(ctx.lib.trt_harmonize(ctx.h))
    keys = ("ac", "n_called", "n_called_nonstrict", "n_hom", "n_padded")
    passes = []
    for _ in range(3):
        st = ctx.locus_stats(False, None, 0.01)
        passes.append({k: st[k][0].copy() for k in keys})
    yield ctx, loci, passes
    ctx.close()


def test_scan_passes_are_bit_identical(full_block):
    _, _, passes = full_block
    for k in passes[0]:
        for i in (1, 2):
            diff = np.nonzero(passes[0][k] != passes[i][k])[0]
            assert diff.size == 0, "{}: pass 0 differs from pass {} at {} entries (first {})".format(k, i, diff.size, diff[:5])


def test_allele_counts_match_call_counters(full_block):
    """sum_a ac[a] = 2 (n_full - n_pad) + n_pad + (n_nonstrict - n_full): every called haplotype is counted once."""
    ctx, _, passes = full_block
    p = passes[0]
    per_locus = np.add.reduceat(p["ac"].astype(np.int64), ctx.locus_off[:-1].astype(np.int64))
    want = p["n_called"] - p["n_padded"] + p["n_called_nonstrict"]
    bad = np.nonzero(per_locus != want)[0]
    assert bad.size == 0, "checksum fails at loci {}".format(bad[:10])
    assert int(p["n_called"].sum()) > 0.9 * L_FULL * S_FULL


def test_sampled_rows_match_numpy_recount(full_block):
    ctx, loci, passes = full_block
    p = passes[0]
    off = ctx.locus_off
    rng = np.random.default_rng(3)
    for l in sorted(set(rng.integers(0, L_FULL, size=48).tolist() + [0, L_FULL - 1, int(np.argmax(loci.n_alleles))])):
        gt = ctx.block_get_gt(l, 1)[0][:, :2].astype(np.int64)
        A = int(loci.n_alleles[l])
        ac = np.bincount(gt[gt >= 0], minlength=A)
        assert np.array_equal(ac, p["ac"][off[l]:off[l + 1]]), l
        nocall = (gt == -1).any(axis=1)
        assert int((~nocall).sum()) == int(p["n_called"][l]), l
        assert int((gt >= 0).any(axis=1).sum()) == int(p["n_called_nonstrict"][l]), l
        assert int(((gt == -2).any(axis=1) & ~nocall).sum()) == int(p["n_padded"][l]), l
