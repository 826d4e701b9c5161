"""Build liblctgan_sm100.so (all CUDA kernels + the C ABI) for sm_100a, in-tree.

    python lct-gan_b200/csrc/build.py [--force]

nvcc cross-compiles without a GPU.  The .so lands in lct-gan_b200/lctgan/ (git-ignored, shipped
to the GPU box by gpurun).  Objects are cached in lct-gan_b200/csrc/build/ keyed on mtime.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(HERE), "lctgan")
OUT = os.path.join(PKG, "liblctgan_sm100.so")
BUILD = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _stale(obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    jobs = []
    objs = []
    for src in sources():
        s = os.path.join(HERE, src)
        o = os.path.join(BUILD, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        cmd = [NVCC] + FLAGS + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        if verbose:
            print("[lctgan build] compiled", os.path.basename(s), flush=True)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(cc, jobs))
    if jobs or force or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("[lctgan build] linked", OUT, flush=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
