"""ctypes binding of the C ABI declared in include/cdgpu.h.

`Lib(path, prefix)` binds every entry point of the header under the given symbol
prefix.  The product binds libcdgpu.so with prefix "cdgpu" (see `load_product`);
the test-suite binds the CPU oracle (oracle/libcdref.so, prefix "cdref") through
the very same class so both sides are driven by identical host code.  Nothing in
this package ever loads the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)

OK, EDIM, EARG, ECUDA, ENOMEM, ENCCL, ENODEV, ECAP = range(8)
LOSS_LS, LOSS_WLS, LOSS_SQRT, LOSS_QUAD = range(4)
INIT_SCREENING, INIT_STD, INIT_WARMSTART = range(3)
KERNEL_GAUSSIAN, KERNEL_EPANECHNIKOV = range(2)


class Options(C.Structure):
    """cdgpu_options == CDOptions (src/utils.jl:7-20) + seed."""

    _fields_ = [
        ("maxIter", C.c_int64),
        ("optTol", C.c_double),
        ("randomize", C.c_int32),
        ("warmStart", C.c_int32),
        ("numSteps", C.c_int64),
        ("seed", C.c_uint64),
    ]


class IterOptions(C.Structure):
    """cdgpu_iter_options == IterLassoOptions (src/utils.jl:24-39)."""

    _fields_ = [
        ("maxIter", C.c_int64),
        ("optTol", C.c_double),
        ("initProcedure", C.c_int32),
        ("_pad", C.c_int32),
        ("sinit", C.c_int64),
        ("sigma_init", C.c_double),
        ("optionsCD", Options),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("passes", C.c_int64),
        ("full_passes", C.c_int64),
        ("visits", C.c_int64),
        ("accepted", C.c_int64),
        ("maxH", C.c_double),
        ("converged", C.c_int32),
        ("outer_iters", C.c_int32),
        ("sigma", C.c_double),
        ("device_ms", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_STATS_NAMES = tuple(k for k, _ in Stats._fields_)
_STATS_DTYPE = np.dtype([("passes", "<i8"), ("full_passes", "<i8"), ("visits", "<i8"), ("accepted", "<i8"), ("maxH", "<f8"),
                         ("converged", "<i4"), ("outer_iters", "<i4"), ("sigma", "<f8"), ("device_ms", "<f8")])
assert _STATS_DTYPE.itemsize == C.sizeof(Stats)


def stats_dicts(arr, count):
    """The first `count` entries of a ctypes array of Stats as dicts (one pass through numpy instead of a getattr per
    field: a 100-lambda path has 900 of them)."""
    if count <= 0:
        return []
    rows = np.frombuffer(arr, dtype=_STATS_DTYPE, count=count).tolist()
    return [dict(zip(_STATS_NAMES, r)) for r in rows]


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch (CDGPU_EDIM)."""


class ArgumentError(ValueError):
    """Julia's ArgumentError (CDGPU_EARG)."""


class CdgpuError(RuntimeError):
    """Julia's ErrorException for CUDA / NCCL / memory / capacity failures."""


_SIGS = {
    "version": (C.c_int, []),
    "last_error": (C.c_char_p, []),
    "device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "default_options": (None, [C.POINTER(Options)]),
    "default_iter_options": (None, [C.POINTER(IterOptions)]),
    "naive_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                               C.c_void_p, C.c_void_p, C.c_int]),
    "naive_create_dev": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_int]),
    "quad_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]),
    "quad_create_dev": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int]),
    "gram_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                              C.c_int]),
    "gram_create_dev": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                  C.c_int]),
    "destroy": (C.c_int, [C.c_void_p]),
    "dims": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p, C.POINTER(C.c_int)]),
    "gram_ms": (C.c_int, [C.c_void_p, c_double_p]),
    "quad_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "solve": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.POINTER(Options), C.c_void_p, C.c_void_p, c_int64_p,
                        C.POINTER(Stats)]),
    "path": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(Options), C.c_int64, C.c_int64,
                       C.c_void_p, C.c_void_p, C.c_void_p, c_int64_p, C.c_void_p]),
    "scaled_solve": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.POINTER(IterOptions), C.c_void_p, C.c_void_p,
                               c_int64_p, c_double_p, C.POINTER(Stats)]),
    "state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "stdx": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "lambda_max": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p]),
    "vc_solve": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                           C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(Options),
                           C.c_int, C.c_void_p, C.c_void_p]),
    "vc_solve_refit": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(Options),
                                 C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vc_solve_chain": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(Options),
                                 C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "refit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "vc_lvocv": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64,
                           C.c_int, C.c_double, C.POINTER(Options), C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
}

# only libcdgpu.so has these (the oracle is single-process, single-device)
_SIGS_GPU_ONLY = {
    "launch_count": (C.c_int, [c_int64_p]),
    "comm_unique_id": (C.c_int, [C.c_void_p]),
    "comm_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "comm_destroy": (C.c_int, [C.c_void_p]),
    "gram_create_lazy": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                   C.c_int]),
    "gram_create_lazy_dev": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                       C.c_int]),
    "lazy_stats": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p, c_int64_p, c_double_p]),
    "lazy_form_columns": (C.c_int, [C.c_void_p, c_int64_p]),
    "sweep_ms": (C.c_int, [C.c_void_p, c_double_p]),
    "vc_solve_csc": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                               C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(Options),
                               C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "synth_normal": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int]),
    "gram_create_sharded": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_int]),
}

HEADER_SYMBOLS = tuple(_SIGS) + tuple(_SIGS_GPU_ONLY)


def ptr(a):
    """Host pointer of a numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def f64(a, order="F"):
    """Float64, Julia-layout (column-major) view or copy of `a`."""
    return np.require(a, dtype=np.float64, requirements=["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS",
                                                         "ALIGNED"])


class Lib:
    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise CdgpuError(f"{path} is not built; there is no fallback")
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path, mode=C.RTLD_LOCAL)
        sigs = dict(_SIGS)
        if prefix == "cdgpu":
            sigs.update(_SIGS_GPU_ONLY)
        for name, (res, args) in sigs.items():
            fn = getattr(self.dll, f"{prefix}_{name}")
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def check(self, rc: int):
        if rc == OK:
            return
        msg = (self.last_error() or b"").decode(errors="replace")
        if rc == EDIM:
            raise DimensionMismatch(msg)
        if rc == EARG:
            raise ArgumentError(msg)
        raise CdgpuError(f"[{rc}] {msg}")

    def default_opts(self) -> Options:
        o = Options()
        self.default_options(C.byref(o))
        return o

    def default_iter_opts(self) -> IterOptions:
        o = IterOptions()
        self.default_iter_options(C.byref(o))
        return o


_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_SO = os.path.join(os.path.dirname(_HERE), "csrc", "libcdgpu.so")
_product = None


def load_product() -> Lib:
    """The B200 library.  Raises (never falls back) when it is not built."""
    global _product
    if _product is None:
        _product = Lib(PRODUCT_SO, "cdgpu")
    return _product
