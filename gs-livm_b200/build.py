"""Build libgsr_b200.so (hand-written sm_100a kernels + C ABI + C++ operator shim) in-tree with nvcc.

The library has no torch / Python dependency: it is the drop-in for GS-LIVM's `src/cuda_rasterizer`
(see include/gsr.h, include/cuda_rasterizer/rasterizer.h).

    python gs-livm_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgsr_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CU_SOURCES = ["preprocess.cu", "binning.cu", "tile_sort.cu", "block_lists.cu", "blend_fwd_sparse.cu", "blend_bwd_sparse.cu", "geom_bwd.cu", "window_allreduce.cu", "window_factored.cu", "train_ops.cu", "vgpr.cu", "gsr_api.cu"]
CXX_SOURCES = ["rasterizer_shim.cpp"]
HEADERS = [os.path.join(CSRC, "gsr_common.cuh"), os.path.join(CSRC, "sh_basis.cuh"), os.path.join(ROOT, "include", "gsr.h"),
           os.path.join(ROOT, "include", "cuda_rasterizer", "rasterizer.h"),
           os.path.join(ROOT, "include", "cuda_rasterizer", "config.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, force: bool, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    if force or _stale(obj, [path] + HEADERS):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
        if verbose:
            sys.stderr.write(res.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = CU_SOURCES + CXX_SOURCES
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-cudart", "static", "-o", LIB] + objs + [
            "-Xlinker", "--exclude-libs=ALL", "-Xlinker", "-Bsymbolic"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
