"""CPU sweep of the branch-free FP64 exp / sqrt that the pairwise-kernel builders use (gperks_b200/csrc/gpx_math.cuh),
compiled for the host, against numpy/libm.  The device build differs only in the square-root seed (MUFU.RSQ64H instead of a
float reciprocal square root) and in where the coefficients live; tests/test_gpu_parity.py checks the device build."""
import ctypes
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    out = tmp_path_factory.mktemp("mathshim") / "libmathshim.so"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", str(out),
                    str(ROOT / "tests" / "math_host_shim.cu")], check=True, capture_output=True)
    lib = ctypes.CDLL(str(out))
    for f in (lib.gpx_host_exp_neg, lib.gpx_host_sqrt_pos):
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
        f.restype = None
    return lib


def _call(f, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    f(x.ctypes.data, out.ctypes.data, x.size)
    return out


def test_exp_neg_within_one_ulp_of_libm(shim):
    rng = np.random.default_rng(0)
    y = np.concatenate([rng.uniform(0, 700, 200_000), rng.uniform(0, 2, 100_000), 10.0 ** rng.uniform(-18, 0, 50_000),
                        np.array([0.0, 1e-300, np.log(2) / 2, np.log(2), 699.999, 700.0])])
    got = _call(shim.gpx_host_exp_neg, y)
    ref = np.exp(-y)
    rel = np.abs(got - ref) / ref
    assert rel.max() <= 2.3e-16, rel.max()          # <= 1 ulp of a correctly rounded result (libm itself is within 1 ulp)
    assert got[y == 0.0][0] == 1.0


def test_exp_neg_clamps_instead_of_wrapping_the_exponent(shim):
    got = _call(shim.gpx_host_exp_neg, np.array([700.0, 750.0, 1e4, 1e300, 712.3456789]))
    # the clamp works on the upper word: anything above 700 becomes 700 + (low word) * 2^-43 <= 700.0005
    assert np.all(got <= np.exp(-700.0) * (1 + 1e-15)) and np.all(got >= np.exp(-700.0005)) and 0 < got[0] < 1e-303


def test_sqrt_pos_is_correctly_rounded(shim):
    rng = np.random.default_rng(1)
    x = np.concatenate([10.0 ** rng.uniform(-30, 12, 300_000), rng.uniform(0.5, 2.0, 100_000),
                        np.array([1e-30, 1.0, 4.0, 2.0, 1e-15 ** 2])])
    got = _call(shim.gpx_host_sqrt_pos, x)
    assert np.array_equal(got, np.sqrt(x))
