"""cdgpu — B200-native coordinate-descent solver behind the CoordinateDescent.jl interface.

The compute lives in csrc/ (hand-written sm_100a CUDA behind the C ABI of
include/cdgpu.h); this package is only the host-side mirror of the reference's
names.  Importing it never loads the CPU oracle and never falls back to the CPU.
"""
from ._ffi import (ArgumentError, CdgpuError, DimensionMismatch, Lib, load_product, HEADER_SYMBOLS, PRODUCT_SO)
from .api import (Backend, CDOptions, IterLassoOptions, SparseIterate, ProxL1, LassoSolution, LassoPath,
                  GaussianKernel, EpanechnikovKernel, CDLeastSquaresLoss, CDWeightedLSLoss, CDSqrtLassoLoss,
                  CDQuadraticLoss, default)
