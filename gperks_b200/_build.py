"""Builds gperks_b200/lib/libgpx.so (hand-written sm_100a kernels + C ABI) with nvcc, in-tree."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libgpx.so"
SOURCES = ["gpx_api.cu", "gpx_kmat.cu", "gpx_chol.cu", "gpx_predict.cu", "gpx_hm.cu", "gpx_tf32.cu", "gpx_i8.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; gperks_b200 needs the CUDA toolkit to build libgpx.so")
    return cand


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in SOURCES]
    deps = srcs + [CSRC / "gpx_common.cuh", CSRC / "gpx_tma.cuh", CSRC / "gpx_gemm_tma.cuh", CSRC / "gpx_tc.cuh", PKG.parent / "include" / "gpx.h"]
    if LIB.exists() and not force:
        return LIB
    LIBDIR.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(LIB), *map(str, srcs)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    for d in deps:
        assert d.exists(), d
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
