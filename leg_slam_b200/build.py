#!/usr/bin/env python
"""Build liblgs.so (the C-ABI hot path, include/lgs.h) in-tree for sm_100a.

    python -m leg_slam_b200.build [--force] [--verbose]

Plain nvcc, one object per .cu (compiled in parallel), linked into
leg_slam_b200/liblgs.so.  No torch, no CMake: the library depends on the CUDA runtime only.
The .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "liblgs.so")
SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "render_fwd.cu", "render_fwd_tc.cu", "render_bwd.cu", "render_bwd_tc.cu",
           "preprocess_bwd.cu", "adam.cu", "dp_adam.cu", "query.cu", "query_tc.cu", "loss.cu", "densify.cu", "ingest.cu", "ply.cu", "bench_kernels.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "ptx.cuh"), os.path.join(CSRC, "adam_math.cuh"), os.path.join(CSRC, "tc.cuh"),
           os.path.join(ROOT, "include", "lgs.h")]


def nvcc_path():
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    return os.path.join(cuda, "bin", "nvcc")


def flags():
    # no --use_fast_math: expf / division / sqrt must round like the reference build
    return ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
            "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    jobs, objs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + HEADERS):
            cmd = [nvcc_path()] + flags() + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        if verbose:
            print("[lgs build]", " ".join(cmd), flush=True)
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0 or verbose:
            sys.stdout.write(r.stdout)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([nvcc_path(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                           "-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
