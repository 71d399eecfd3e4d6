#!/usr/bin/env python
"""Dense active-set chain probe: C1 data (n=1000, p=5000) at lambda=0.01 (nnz ~ 830, ~1300 active passes),
solved in covariance form (cluster kernel) and in naive form (cooperative kernel).  Run with
CDGPU_PROFILE=1 to get the per-phase cycle counters of both kernels."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))
import cdgpu  # noqa: E402
from cdgpu import CDOptions, ProxL1, SparseIterate  # noqa: E402

rng = np.random.default_rng(123)
n, p, s = 1000, 5000, 10
X = rng.standard_normal((p, n)).T
beta = rng.standard_normal(s) * (1.0 + rng.random(s))
y = X[:, :s] @ beta + rng.standard_normal(n)
lam = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
# CDGPU_PROBE_LIB=coordinatedescent.jl_b200/csrc/prof/libcdgpu.so: the diagnostics build (make -C csrc prof) with the
# chain engines' own cycle counters
be = cdgpu.Backend(cdgpu.Lib(os.path.abspath(os.environ["CDGPU_PROBE_LIB"]), "cdgpu")) if os.environ.get("CDGPU_PROBE_LIB") else cdgpu.default()
forms = os.environ.get("CDGPU_PROBE_FORMS", "cov,naive").split(",")
opt = CDOptions(maxIter=2000, optTol=1e-7, randomize=False)
for form in forms:
    f = be.CDQuadraticLoss_from_data(X, y) if form == "cov" else be.CDLeastSquaresLoss(y, X)
    for rep in range(2):
        x = SparseIterate(p)
        be.coordinateDescent_(x, f, ProxL1(lam), opt)
        st = f.last_stats
        print(form, "rep", rep, "nnz", x.nnz, "passes", st["passes"], "visits", st["visits"], "accepted", st["accepted"],
              "device_ms %.3f" % st["device_ms"], flush=True)
    f.close()
