"""Builds libdssmd_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m dssmd_b200.build [--force]

The shared library lands next to this file (dssmd_b200/libdssmd_b200.so), is
git-ignored, and is what dssmd_b200._lib loads through ctypes.  nvcc
cross-compiles without a GPU, so this also runs in the CPU-only build container.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OBJ_DIR = os.path.join(HERE, "_obj")
LIB_PATH = os.path.join(HERE, "libdssmd_b200.so")

SOURCES = ["abi.cu", "corr_fwd.cu", "corr_fwd_umma.cu", "corr_fwd_umma2.cu", "corr_bwd.cu", "corr_bwd_reg.cu", "corr_bwd_fused.cu", "corr_bwd_wide.cu", "corr_fwd_reg.cu", "corr_fwd_mma16.cu", "corr_bwd_mma16.cu", "corr_fwd_mma32.cu", "corr_bwd_mma32.cu", "warp.cu", "warp_bwd.cu", "warp_lean.cu", "warp_bwd_band.cu", "warp_bwd_band2.cu", "photometric.cu", "haze.cu", "asm_loss.cu", "epipolar_attn.cu", "epipolar_attn_mma.cu", "epipolar_attn_umma.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "warp_common.cuh"), os.path.join(CSRC, "warp_lean.cuh"), os.path.join(CSRC, "tma.cuh"), os.path.join(CSRC, "corr_reg.cuh"), os.path.join(INCLUDE, "dssmd_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_extension(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link the C-ABI shared library."""
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    deps_common = HEADERS + [os.path.abspath(__file__)]

    def compile_one(src: str) -> str:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        if force or _stale(obj, [path] + deps_common):
            cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))

    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build_extension(force="--force" in sys.argv, verbose=True)
    print("built", p)
