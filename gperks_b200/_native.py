"""ctypes binding of libgpx.so (include/gpx.h).  Thin: raw device pointers in, error codes out.

There is deliberately NO fallback: if the library cannot be loaded, or no CUDA device is present
when a compute entry point is called, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _build

_lib: Optional[C.CDLL] = None

KIND_RBF, KIND_MATERN12, KIND_MATERN32, KIND_MATERN52 = 0, 1, 2, 3
TILE = 128
ABI_VERSION = 203          # gpx_version() of the library these ctypes signatures were written against

_P = C.c_void_p
_SIGS = {
    "gpx_version": (C.c_int, []),
    "gpx_padded": (C.c_int, [C.c_int]),
    "gpx_max_dim": (C.c_int, []),
    "gpx_kmat_sym": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, _P]),
    "gpx_kmat_cross": (C.c_int, [_P, C.c_long, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, _P, C.c_int, _P,
                                 C.c_int, _P]),
    "gpx_potrf": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "gpx_trtri": (C.c_int, [_P, _P, _P, _P, C.c_int, _P]),
    "gpx_lauum": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "gpx_solve_alpha": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P]),
    "gpx_mll_value": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "gpx_mll_grad": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, _P, _P, _P, _P]),
    "gpx_trmm_sumsq": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P]),
    "gpx_predict_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gpx_predict_meanvar": (C.c_int, [_P, C.c_long, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P,
                                      _P, C.c_int, C.c_int, _P, _P, C.c_double, C.c_double, C.c_double, _P, _P, _P,
                                      C.c_size_t, C.c_int, _P]),
    "gpx_trmm_store": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P]),
    "gpx_syrk_sub": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "gpx_split_tf32": (C.c_int, [_P, C.c_long, _P, _P, C.c_int, _P]),
    "gpx_tf32_row_tiles": (C.c_int, [C.c_int]),
    "gpx_trmm_sumsq_tf32": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P]),
    "gpx_predict_tf32_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gpx_predict_meanvar_tf32": (C.c_int, [_P, C.c_long, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, _P, C.c_int, _P,
                                           _P, _P, C.c_int, C.c_int, _P, _P, C.c_double, C.c_double, C.c_double, _P,
                                           _P, _P, C.c_size_t, C.c_int, C.c_int, _P]),
    "gpx_slice_i8": (C.c_int, [_P, C.c_int, C.c_int, C.c_long, _P, C.c_long, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "gpx_i8_row_tiles": (C.c_int, [C.c_int]),
    "gpx_i8_max_n": (C.c_int, [C.c_int]),
    "gpx_i8_safe_n": (C.c_int, [C.c_int]),
    "gpx_i8_digit_bound": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "gpx_i8_clusters": (C.c_int, []),
    "gpx_kmat_cross_i8": (C.c_int, [_P, C.c_long, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, _P, C.c_long, C.c_int, C.c_int,
                                    C.c_int, _P, C.c_int, _P]),
    "gpx_trmm_sumsq_i8": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "gpx_potrf_i8_workspace_bytes": (C.c_size_t, [C.c_int]),
    "gpx_potrf_i8": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "gpx_predict_covar_i8": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, _P, C.c_int, _P, C.c_size_t, _P]),
    "gpx_trtri_i8_workspace_bytes": (C.c_size_t, [C.c_int]),
    "gpx_trtri_i8": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "gpx_lauum_i8_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gpx_lauum_i8": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "gpx_predict_i8_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gpx_predict_i8_planes_offset": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "gpx_predict_meanvar_i8": (C.c_int, [_P, C.c_long, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, _P, C.c_int, C.c_int,
                                         C.c_int, _P, _P, _P, C.c_int, C.c_int, _P, _P, C.c_double, C.c_double, C.c_double,
                                         _P, _P, _P, C.c_size_t, C.c_int, _P, _P]),
    "gpx_implausibility": (C.c_int, [_P, _P, C.c_long, C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    "gpx_wave_i8_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int]),
    "gpx_wave_i8": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_long, _P, _P, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, C.c_int, _P, _P]),
    "gpx_predict_mean": (C.c_int, [_P, C.c_long, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P,
                                   C.c_double, C.c_double, _P, _P]),
    "gpx_saltelli_chunk_rows": (C.c_int, []),
    "gpx_saltelli_partial": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_long, C.c_int, C.c_long, C.c_long, C.c_int, C.c_ulonglong, _P, _P]),
    "gpx_saltelli_reduce": (C.c_int, [_P, C.c_long, C.c_int, C.c_int, _P, _P]),
    "gpx_compact_below_workspace_bytes": (C.c_size_t, [C.c_long]),
    "gpx_compact_below": (C.c_int, [_P, C.c_long, C.c_double, _P, _P, _P, C.c_size_t, _P]),
    "gpx_peak_dmma": (C.c_double, [C.c_int, C.c_int, _P, _P]),
    "gpx_peak_i8mma": (C.c_double, [C.c_int, C.c_int, _P]),
    "gpx_peak_dfma": (C.c_double, [C.c_int, C.c_int, _P, _P]),
}
EXPORTS = tuple(_SIGS)


class I8EmulatorDesc(C.Structure):
    """gpx_i8_emulator of include/gpx.h: a host struct of device pointers describing one factorised emulator."""
    _fields_ = [("X", _P), ("theta", _P), ("xa", _P), ("xr", _P), ("alpha", _P), ("Ls", _P), ("inv_scale", _P),
                ("feat_idx", _P), ("weights", _P), ("bias", _P),
                ("y_mu", C.c_double), ("y_sigma", C.c_double), ("min_var", C.c_double),
                ("N", C.c_int), ("Np", C.c_int), ("kind", C.c_int), ("nslices", C.c_int), ("bits", C.c_int),
                ("n_feat", C.c_int), ("degree", C.c_int), ("reserved", C.c_int)]


def lib() -> C.CDLL:
    """Loads (building if absent) libgpx.so.  Raises if that is impossible -- no fallback."""
    global _lib
    if _lib is None:
        path = _build.build()
        try:
            handle = C.CDLL(str(path))
        except OSError as e:  # pragma: no cover
            raise RuntimeError(f"gperks_b200: cannot load {path}: {e}") from e
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.gpx_version() != ABI_VERSION:      # a library built from other sources than these signatures: never call into it
            raise RuntimeError(f"gperks_b200: {path} reports ABI {handle.gpx_version()}, this package binds ABI {ABI_VERSION}; "
                               f"rebuild with `python -m gperks_b200._build`")
        _lib = handle
    return _lib


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "gperks_b200 computes on an NVIDIA B200 (sm_100a) only: no CUDA device is available and there "
            "is no CPU fallback."
        )


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "gpx expects contiguous CUDA tensors"
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"gpx: {what} failed with code {rc}")


def padded(n: int) -> int:
    return (max(int(n), 1) + TILE - 1) // TILE * TILE
