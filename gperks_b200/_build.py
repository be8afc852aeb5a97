"""Builds gperks_b200/lib/libgpx.so (hand-written sm_100a kernels + C ABI) with nvcc, in-tree."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libgpx.so"
STAMP = LIBDIR / "libgpx.so.sha256"
SOURCES = ["gpx_api.cu", "gpx_kmat.cu", "gpx_chol.cu", "gpx_predict.cu", "gpx_hm.cu", "gpx_gsa.cu", "gpx_tf32.cu", "gpx_i8.cu"]
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


def source_digest(deps) -> str:
    """SHA-256 over the contents of every file the library is built from (and the nvcc flags)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in deps:
        h.update(d.name.encode())
        h.update(d.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Returns the path of libgpx.so, compiling it when it is missing or was built from other sources: the digest of
    every .cu/.cuh/gpx.h it was compiled from is kept beside it (lib/libgpx.so.sha256) and compared on every load, so an
    edited or freshly pulled source tree is never paired with a stale library.  When the sources are unreadable or nvcc is
    absent (a deployed wheel), an existing library is used as it is; ``_native.lib()`` still checks its ABI version."""
    srcs = [CSRC / s for s in SOURCES]
    deps = srcs + [CSRC / "gpx_common.cuh", CSRC / "gpx_math.cuh", CSRC / "gpx_tma.cuh", CSRC / "gpx_gemm_tma.cuh", CSRC / "gpx_tc.cuh", PKG.parent / "include" / "gpx.h"]
    try:
        digest = source_digest(deps)
    except OSError:
        digest = None
    if LIB.exists() and not force:
        built_from = STAMP.read_text().strip() if STAMP.exists() else None
        if digest is None or built_from == digest:
            return LIB
        try:
            nvcc_path()
        except RuntimeError:
            raise RuntimeError(f"{LIB} was built from other sources than the ones in {CSRC} and nvcc is not available to rebuild it")
    LIBDIR.mkdir(parents=True, exist_ok=True)
    tmp = LIBDIR / f"libgpx.{os.getpid()}.tmp.so"          # built aside, renamed atomically: concurrent ranks never load half a file
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(tmp), *map(str, srcs)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB)
    if verbose:
        print(res.stderr)
    if digest is not None:
        STAMP.write_text(digest + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
