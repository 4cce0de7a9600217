"""In-tree build of libmassiv_b200.so (nvcc, sm_100a only).

    python -m massiv_b200.build [--force]

Objects go to build/ (git-ignored), the shared library next to this file so that it travels to
the GPU box with the repo snapshot.  No torch involved: the library is plain CUDA runtime + C ABI.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# tuning variants: MB_NVCC_EXTRA="-DMB_T2_WX64=2" MB_VARIANT=wx2 builds libmassiv_b200_wx2.so beside the default
VARIANT = os.environ.get("MB_VARIANT", "")
OBJ = os.path.join(ROOT, "build", "obj" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "libmassiv_b200" + ("_" + VARIANT if VARIANT else "") + ".so")

SOURCES = ["mb_plan.cpp", "mb_api.cu", "k_generic.cu", "k_tile2d.cu", "k_vol3d.cu", "k_life.cu", "mb_shard.cu",
           "k_selftest.cu", "k_sep2d.cu", "k_tile2d_known.cu", "k_vol3d_tb2.cu", "k_linescan.cu", "k_direct2d.cu", "k_post.cu", "k_smallwin.cu", "k_converge.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: belt and braces — the kernels already use explicit *_rn intrinsics, which nvcc
# never contracts; explicit fma() calls (exact-division helper) are unaffected by the flag.
# (tuning variants are built without -lineinfo, MB_NO_LINEINFO=1: the libraries are a third smaller and the GPU box's
# snapshot has a size limit)
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17"] + ([] if os.environ.get("MB_NO_LINEINFO") else ["-lineinfo"]) + ["-fmad=false",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-Xptxas", "-v", "--expt-relaxed-constexpr",
         "-I", os.path.join(ROOT, "include")] + os.environ.get("MB_NVCC_EXTRA", "").split()


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "massiv_b200.h")]


def _compile(src: str, force: bool) -> str:
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    hdrs = [p for p in _deps() if p.endswith((".h", ".cuh"))]
    newest = max(os.path.getmtime(p) for p in [path] + hdrs)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj
    cmd = [NVCC] + FLAGS + (["-x", "cu"] if src.endswith(".cpp") else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, os.path.splitext(src)[0] + ".ptxas.log")
    with open(log, "w") as f:
        f.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    newest_src = max(os.path.getmtime(p) for p in _deps())
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest_src:
        return LIB
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
