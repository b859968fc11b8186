"""
GPU tests of the multi-GPU plumbing through the library's own NCCL entry points (trt_dist_*, csrc/trt_dist.cu):
device-buffer gathers of the result regions, the counter all-reduces, and the CLIs producing identical files on one
and on two GPUs.  World 1 runs on any GPU box; the world-2 cases need two devices (``gpurun --gpus 2``).
"""
import filecmp
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    from trtools_b200 import _lib
    return _lib.device_count()


def _torchrun(n, script_args, timeout=900):
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + script_args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, OMP_NUM_THREADS="1"))


def test_trt_dist_world1_single_process():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29900 + os.getpid() % 90))
    res = subprocess.run([sys.executable, os.path.join(REPO, "tests", "dist_gpu_worker.py")], capture_output=True, text=True,
                         timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "RANK0 OK" in res.stdout


def test_trt_dist_world2_nccl():
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    res = _torchrun(2, [os.path.join(REPO, "tests", "dist_gpu_worker.py")])
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "RANK0 OK" in res.stdout and "RANK1 OK" in res.stdout


CLI = r'''
import argparse, os, sys
sys.path.insert(0, {repo!r})
tool, out, data = sys.argv[1], sys.argv[2], sys.argv[3]
if tool == "statSTR":
    from trtools_b200 import statSTR
    ns = argparse.Namespace(vcf=os.path.join(data, "many_samples.vcf.gz"), out=out, vcftype="auto", samples=None, sample_prefixes=None,
                            region=None, precision=4, nalleles_thresh=0.01, plot_afreq=False, use_length=False, only_passing=False,
                            block_size=100)
    for s in ("thresh", "afreq", "acount", "nalleles", "hwep", "het", "entropy", "mean", "mode", "var", "numcalled"):
        setattr(ns, s, True)
    sys.exit(statSTR.main(ns))
if tool == "associaTR":
    from trtools_b200 import associaTR
    ns = argparse.Namespace(outfile=out, tr_vcf=os.path.join(data, "many_samples_biallelic.vcf.gz"), phenotype_name="test_pheno",
                            traits=[os.path.join(data, "traits_0.npy"), os.path.join(data, "traits_1.npy")], vcftype=None, same_samples=True,
                            sample_list=None, region=None, non_major_cutoff=5, beagle_dosages=False, plotting_phenotype=None,
                            paired_genotype_plot=False, plot_phenotype_residuals=False, plotting_ci_alphas=[],
                            imputed_ukb_strs_paper_period_check=False, block_size=7)
    associaTR.main(ns)
    sys.exit(0)
if tool == "dumpSTR":
    from trtools_b200 import dumpSTR
    sys.path.insert(0, os.path.join({repo!r}, "tests"))
    from test_gpu_dumpstr import dump_args
    ns = dump_args(vcf=os.path.join(data, "many_samples.vcf.gz"), out=out, vcftype="auto", min_locus_hwep=0.01, min_locus_callrate=0.9,
                   hipstr_min_call_DP=10, hipstr_max_call_flank_indel=0.15, hipstr_min_call_Q=0.9, use_length=True)
    ns.block_size = 64
    sys.exit(dumpSTR.main(ns))
'''


@pytest.mark.parametrize("tool,outs", [("statSTR", [".tab"]), ("associaTR", [""]), ("dumpSTR", [".vcf", ".samplog.tab", ".loclog.tab"])])
def test_cli_output_identical_on_one_and_two_gpus(tool, outs, tmp_path, data_dir):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "cli.py"
    script.write_text(CLI.format(repo=REPO))
    one, two = str(tmp_path / "one"), str(tmp_path / "two")
    r1 = subprocess.run([sys.executable, str(script), tool, one, data_dir], capture_output=True, text=True, timeout=900)
    assert r1.returncode == 0, r1.stdout[-2000:] + r1.stderr[-2000:]
    r2 = _torchrun(2, [str(script), tool, two, data_dir])
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    for suffix in outs:
        a, b = one + suffix, two + suffix
        assert os.path.getsize(a) > 100
        if suffix == ".vcf":
            la = [x for x in open(a) if not x.startswith("##command")]
            lb = [x for x in open(b) if not x.startswith("##command")]
            assert la == lb
        else:
            assert filecmp.cmp(a, b, shallow=False), suffix
