#!/usr/bin/env python
"""C4 (4096 grid points, n=500, p=50, degree 2) with the warm-start chain cut into runs of `chain` grid points
(cdgpu_vc_solve_chain): device time, total passes, and the distance to chain = 1 at the same tolerance."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cdgpu  # noqa: E402
from cdgpu import CDOptions, GaussianKernel  # noqa: E402
from test_baseline_configs import c4_data  # noqa: E402

X, Z, Y = c4_data()
m = int(os.environ.get("C4_M", 4096))
zgrid = np.linspace(0.01, 0.99, m)
be = cdgpu.default()
for tol in (1e-7, 1e-9):
    o = CDOptions(maxIter=200000, optTol=tol, randomize=False)
    base = None
    for chain in (1, 2, 3, 4, 8, 16):
        for rep in range(2):
            out, _ = be.locpolyl1(X, Z, Y, zgrid, 2, GaussianKernel(0.2), 0.01, False, o, chain=chain)
        st = be.last_vc_stats
        if base is None:
            base = out
        print("optTol %.0e chain %2d: device %.2f ms, passes %d, visits %.1f M, max |diff to chain 1| / max|beta| = %.2e" % (
            tol, chain, st[0]["device_ms"], sum(s["passes"] for s in st), sum(s["visits"] for s in st) * 1e-6,
            np.max(np.abs(out - base)) / np.max(np.abs(base))), flush=True)
