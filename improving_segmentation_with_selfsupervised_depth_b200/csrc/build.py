"""In-tree nvcc build of the C-ABI library (libsegsde_b200.so) for sm_100a.

`python -m improving_segmentation_with_selfsupervised_depth_b200.csrc.build [--force]`
nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the repo
snapshot to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(PKG, "libsegsde_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC"]


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "segsde_b200.h"))
    hdr_m = max(os.path.getmtime(h) for h in headers)
    jobs = []
    for s in sources():
        src = os.path.join(HERE, s)
        obj = os.path.join(OBJ_DIR, s[:-3] + ".o")
        if force or _newer(src, obj) or os.path.getmtime(obj) < hdr_m:
            jobs.append((src, obj))

    def cc(job):
        src, obj = job
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, r in ex.map(cc, jobs):
                if verbose and r.stderr:
                    print(r.stderr, file=sys.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
    objs = [os.path.join(OBJ_DIR, s[:-3] + ".o") for s in sources()]
    if force or jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
