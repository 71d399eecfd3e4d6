"""CPU-side checks of the product library: it loads, exports every symbol of include/cdgpu.h,
and every compute entry point fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cdgpu
from cdgpu import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "cdgpu.h")).read()
    return sorted(set(re.findall(r"\b(cdgpu_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported():
    assert os.path.exists(cdgpu.PRODUCT_SO), "libcdgpu.so is not built (run __graft_entry__.build())"
    dll = C.CDLL(cdgpu.PRODUCT_SO)
    names = header_functions()
    assert len(names) >= 26
    for name in names:
        assert hasattr(dll, name), f"{name} declared in include/cdgpu.h but not exported"
    # the python binding covers the same set
    assert sorted("cdgpu_" + s for s in _ffi.HEADER_SYMBOLS) == names


def test_oracle_exports_same_abi():
    dll = C.CDLL(os.path.join(ROOT, "oracle", "libcdref.so"))
    for name in header_functions():
        ref = name.replace("cdgpu_", "cdref_", 1)
        if any(s in name for s in ("comm_", "sharded", "launch_count", "lazy", "sweep_ms", "synth_", "_csc")):
            continue  # single-process oracle; device-side diagnostics, the lazy covariance form and the device-side CSC
            # compaction of locpolyl1's result have no CPU counterpart
        assert hasattr(dll, ref), ref


def test_struct_layouts_match_header():
    assert C.sizeof(_ffi.Options) == 40 and C.sizeof(_ffi.IterOptions) == 80 and C.sizeof(_ffi.Stats) == 64


def test_version_and_defaults():
    lib = cdgpu.load_product()
    assert lib.version() == 100
    o = lib.default_opts()
    assert (o.maxIter, o.optTol, o.randomize, o.warmStart, o.numSteps) == (2000, 1e-7, 1, 1, 50)  # utils.jl:14-20
    io = lib.default_iter_opts()
    assert (io.maxIter, io.optTol, io.initProcedure, io.sinit, io.sigma_init) == (20, 1e-2, 0, 5, 1.0)


def _no_gpu():
    lib = cdgpu.load_product()
    n = C.c_int()
    return lib.device_count(C.byref(n)) != 0 or n.value == 0


@pytest.mark.skipif(not _no_gpu(), reason="a CUDA device is present")
def test_no_cpu_fallback_without_device():
    be = cdgpu.default()
    X = np.asfortranarray(np.random.default_rng(0).standard_normal((20, 5)))
    y = X[:, 0].copy()
    with pytest.raises(cdgpu.CdgpuError, match="no CPU fallback|no CUDA device"):
        be.CDLeastSquaresLoss(y, X)
    with pytest.raises(cdgpu.CdgpuError):
        be.CDQuadraticLoss(X.T @ X, X.T @ y)
    with pytest.raises(cdgpu.CdgpuError):
        be.locpolyl1(X, np.linspace(0, 1, 20), y, np.array([0.5]), 1, cdgpu.GaussianKernel(0.2), 0.1)


def test_argument_errors_precede_device_use():
    # dimension / argument checks mirror the reference and do not need a device
    be = cdgpu.default()
    X = np.asfortranarray(np.ones((20, 5)))
    with pytest.raises(cdgpu.DimensionMismatch):
        be.CDLeastSquaresLoss(np.ones(19), X)
    with pytest.raises(cdgpu.ArgumentError):
        be.CDQuadraticLoss(np.ones((5, 4)), np.ones(4))
