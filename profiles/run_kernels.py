#!/usr/bin/env python
"""One launch of each hot kernel at its BASELINE.json size, for `ncu --set full` (see profiles/README.md):
C2 Gram (gram_syrk_kernel) + covariance path (cov_path_kernel), C3 sqrt-lasso (naive_path_kernel),
C4 varying-coefficient lasso (gram_syrk_kernel<GEMM> + vc_cov_kernel); `c1`: the dense active-set case (C1 at lambda = 0.01:
the team chain engine inside naive_path_kernel)."""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))
import cdgpu  # noqa: E402
from cdgpu import CDOptions, GaussianKernel, ProxL1, SparseIterate  # noqa: E402

which = sys.argv[1:] or ["c2", "c3", "c4"]
be = cdgpu.default()
lib = be.lib
g = torch.Generator(device="cuda")
g.manual_seed(1)


def randn_cols(p, n):
    X = torch.empty((p, n), device="cuda", dtype=torch.float64)
    for j0 in range(0, p, 2000):
        X[j0:j0 + 2000].normal_(generator=g)
    return X


if "c2" in which:
    n, p, s = 10000, 20000, 50
    Xd = randn_cols(p, n)
    yd = Xd[:s].T @ torch.randn(s, device="cuda", dtype=torch.float64, generator=g) + torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
    torch.cuda.synchronize()
    f = cdgpu.CDQuadraticLoss.__new__(cdgpu.CDQuadraticLoss)
    cdgpu.api._Loss.__init__(f, lib)
    f.p = p
    lib.check(lib.gram_create_dev(C.byref(f._h), C.c_void_p(Xd.data_ptr()), n, p, n, C.c_void_p(yd.data_ptr()), 0))
    om = f.stdX()
    lmax = be.findLambdaMax(f, om)
    lams = np.exp(np.linspace(np.log(lmax), np.log(0.05 * lmax), 100))
    path = be.LassoPath(None, None, lams, CDOptions(randomize=False), standardizeX=om, loss=f)
    print("c2: gram %.2f ms, path %.2f ms, visits %d" % (f.gram_ms, path.stats[0]["device_ms"], sum(s["visits"] for s in path.stats)))
    f.close()
    del Xd, yd
if "tall" in which:  # tall sqrt-lasso: the row-distributed sweep of tall_sweep.cu
    n, p, s = 1000000, 500, 10
    Xd = randn_cols(p, n)
    yd = Xd[:s].T @ (1.0 + torch.rand(s, device="cuda", dtype=torch.float64, generator=g)) + torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
    torch.cuda.synchronize()
    f = cdgpu.CDSqrtLassoLoss.__new__(cdgpu.CDSqrtLassoLoss)
    cdgpu.api._Loss.__init__(f, lib)
    f.n, f.p = n, p
    lib.check(lib.naive_create_dev(C.byref(f._h), cdgpu._ffi.LOSS_SQRT, C.c_void_p(Xd.data_ptr()), n, p, n, C.c_void_p(yd.data_ptr()), None, 0))
    om = f.stdX()
    x = SparseIterate(p)
    be.coordinateDescent_(x, f, ProxL1(1.1 * math.sqrt(2 * math.log(p)), om), CDOptions(randomize=False))
    print("tall:", f.last_stats)
    f.close()
    del Xd, yd
if "c3" in which:
    n, p, s = 5000, 50000, 20
    Xd = randn_cols(p, n)
    yd = Xd[:s].T @ torch.randn(s, device="cuda", dtype=torch.float64, generator=g) + torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
    torch.cuda.synchronize()
    f = cdgpu.CDSqrtLassoLoss.__new__(cdgpu.CDSqrtLassoLoss)
    cdgpu.api._Loss.__init__(f, lib)
    f.n, f.p = n, p
    lib.check(lib.naive_create_dev(C.byref(f._h), cdgpu._ffi.LOSS_SQRT, C.c_void_p(Xd.data_ptr()), n, p, n, C.c_void_p(yd.data_ptr()), None, 0))
    x = SparseIterate(p)
    be.coordinateDescent_(x, f, ProxL1(1.1 * math.sqrt(2 * math.log(p))), CDOptions(randomize=False))
    print("c3:", f.last_stats)
    f.close()
    del Xd, yd
if "c1" in which:  # C1 at lambda = 0.01: 828 active entries, 1317 passes -> the team chain engine inside naive_path_kernel
    rng = np.random.default_rng(123)
    n, p, s = 1000, 5000, 10
    X = rng.standard_normal((p, n)).T
    y = X[:, :s] @ (rng.standard_normal(s) * (1.0 + rng.random(s))) + rng.standard_normal(n)
    f = be.CDLeastSquaresLoss(y, X)
    x = SparseIterate(p)
    be.coordinateDescent_(x, f, ProxL1(0.01), CDOptions(maxIter=2000, optTol=1e-7, randomize=False))
    print("c1:", f.last_stats)
    f.close()
if "c4" in which:
    n, p, degree, m = 500, 50, 2, 4096
    rng = np.random.default_rng(125)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    Z = rng.random(n)
    cj = rng.choice([2, 4, 6, 8], size=p)
    Y = np.array([np.sin(cj * Z[i])[:2] @ X[i, :2] for i in range(n)]) + 0.1 * rng.standard_normal(n)
    be.locpolyl1(X, Z, Y, np.linspace(0.01, 0.99, m), degree, GaussianKernel(0.2), 0.01, False, CDOptions(randomize=False))
    print("c4: device %.2f ms" % be.last_vc_stats[0]["device_ms"])
