#!/usr/bin/env python
"""C4 per-problem statistics of the varying-coefficient path (passes, visits, accepted steps)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))
import cdgpu
from cdgpu import CDOptions, GaussianKernel
rng = np.random.default_rng(7)
n, p, m = 500, 50, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X = np.asfortranarray(rng.standard_normal((n, p)))
z = rng.random(n)
y = sum(X[:, j] * np.sin((2 + 2 * j) * z) for j in range(2)) + 0.1 * rng.standard_normal(n)
zg = np.linspace(0.01, 0.99, m)
be = cdgpu.default()
for rep in range(2):
    t0 = time.perf_counter()
    out, _ = be.locpolyl1(X, z, y, zg, 2, GaussianKernel(0.2), 0.01, options=CDOptions(randomize=False))
    wall = time.perf_counter() - t0
    st = be.last_vc_stats
    tot = {k: sum(s[k] for s in st) for k in ("passes", "full_passes", "visits", "accepted")}
    print("rep", rep, "wall %.1f ms device %.1f ms" % (1e3 * wall, st[0]["device_ms"]), {k: v / m for k, v in tot.items()},
          "max passes", max(s["passes"] for s in st), "min", min(s["passes"] for s in st),
          "nnz/problem %.1f" % (np.count_nonzero(out) / m), flush=True)
