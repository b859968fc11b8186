"""
Build libtrtools_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m trtools_b200.build [--force]

The library travels to the GPU box with the repo snapshot (git-ignored, not gpurun-ignored).
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtrtools_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-fmad=false",
]


def sources():
    """CUDA kernels + the host-side ingest (trt_ingest.cpp: zlib, threads; nvcc hands it to g++)."""
    return sorted(glob.glob(os.path.join(CSRC, "*.cu"))) + sorted(glob.glob(os.path.join(CSRC, "*.cpp")))


def _find_nccl():
    """Prefer the NCCL that ships with torch (2.28.9) so the .so and torch.distributed agree."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            root = list(spec.submodule_search_locations)[0]
            inc, lib = os.path.join(root, "include"), os.path.join(root, "lib")
            if os.path.exists(os.path.join(inc, "nccl.h")) and glob.glob(os.path.join(lib, "libnccl.so*")):
                return inc, lib, os.path.basename(sorted(glob.glob(os.path.join(lib, "libnccl.so*")))[0])
    except Exception:
        pass
    return None, None, None


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(HERE), "include", "trtools_b200.h"), os.path.abspath(__file__)]


def build(force=False, verbose=False):
    """One object per translation unit under build/obj (compiled in parallel, rebuilt only when the unit, a header or
    this script changed), then one link into the in-tree shared object."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    inc, libdir, soname = _find_nccl()
    extra = os.environ.get("TRT_EXTRA_NVCC", "").split()
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + extra
    if verbose:
        cflags += ["-Xptxas", "-v"]
    if inc:
        cflags += ["-I", inc]
    objdir = os.path.join(os.path.dirname(HERE), "build", "obj" + ("_" + "_".join(extra).replace("/", "_") if extra else ""))
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        return job, subprocess.run([nvcc] + cflags + ["-c", src, "-o", obj], capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        results = list(pool.map(compile_one, jobs))
    failed = [(j, r) for j, r in results if r.returncode != 0]
    for (src, obj), r in failed:
        sys.stderr.write(r.stdout + r.stderr)
        if os.path.exists(obj):
            os.remove(obj)
    if failed:
        raise RuntimeError("nvcc failed building " + ", ".join(os.path.basename(j[0]) for j, _ in failed))
    if verbose:
        for _, r in results:
            sys.stderr.write(r.stdout + r.stderr)
    out = os.environ.get("TRT_BUILD_OUT", LIB)
    tmp = out + ".tmp.%d" % os.getpid()          # link next to the target, then rename: never a half-written .so
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    if inc:
        cmd += ["-L", libdir, "-l:" + soname, "-Xlinker", "-rpath=" + libdir]
    else:
        cmd += ["-lnccl"]
    cmd += ["-lz", "-lpthread", "-o", tmp] + [os.path.join(objdir, os.path.basename(s) + ".o") for s in sources()]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed linking libtrtools_b200.so")
    os.replace(tmp, out)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
