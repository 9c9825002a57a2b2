"""Build libcellscopes_cuda.so in-tree with nvcc for sm_100a (B200).

    python cellscopes.jl_b200/build.py [--force] [--verbose]

One object per .cu (compiled in parallel), linked into cellscopes.jl_b200/lib/.  nvcc cross-compiles
without a GPU, so this runs on the CPU build box; the .so then travels to the GPU box with the tree.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libcellscopes_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
    "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libcellscopes_cuda.so cannot be built (there is no CPU fallback)")


def _sources():
    # .cu through nvcc; .cpp (host-only code with CPU intrinsics) through nvcc's host compiler pass-through
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "cellscopes_cuda.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(nvcc, src, obj, verbose):
    if src.endswith(".cpp"):
        cmd = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-fvisibility=default", "-I", os.path.join(ROOT, "include"),
               "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
    else:
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hm = _headers_mtime()
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(
            os.path.getmtime(os.path.join(CSRC, src)), hm)
        if stale:
            jobs.append((src, obj))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, rc, log in ex.map(lambda j: _compile(nvcc, j[0], j[1], verbose), jobs):
                if verbose or rc != 0:
                    sys.stderr.write(f"--- {src}\n{log}\n")
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
               "-Xlinker", "--exclude-libs,ALL", "-lz"]      # zlib: gzip'ed MatrixMarket files (csc_read_mtx)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
