#!/usr/bin/env python
"""Measurements of the BASELINE.json configs that are not the headline bench line:
C1 (lasso n=1000 p=5000), C3 (sqrt-/scaled-lasso n=5000 p=50000, naive form), C4 (4096 local
varying-coefficient problems).  Prints one JSON object per config.  GPU timings are the library's
CUDA-event device_ms; the CPU columns (only through bench.py's cpu_baseline leg) are the oracle port
(-O3, 1 thread, literal reference loops).
Usage: python benchmarks/other_configs.py [c1] [c3] [c4] [lvocv]            (GPU columns only)
       python bench.py --other-configs c1,c3,c4,lvocv [--other-cpu]         (adds the CPU-port columns)
"""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))
import cdgpu  # noqa: E402
from cdgpu import CDOptions, IterLassoOptions, ProxL1, SparseIterate, GaussianKernel  # noqa: E402

HBM = 6535.4
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    pass


def problem(n, p, s, seed, noise=1.0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((p, n)).T
    beta = rng.standard_normal(s) * (1.0 + rng.random(s))
    y = X[:, :s] @ beta + noise * rng.standard_normal(n)
    return X, np.ascontiguousarray(y)


def timed_solve(be, f, lam, om, opt, reps=1):
    best = None
    for _ in range(reps):
        x = SparseIterate(f.p)
        t0 = time.perf_counter()
        be.coordinateDescent_(x, f, ProxL1(lam, om), opt)
        wall = time.perf_counter() - t0
        st = dict(f.last_stats, wall_ms=1e3 * wall, nnz=x.nnz)
        if best is None or st["device_ms"] < best["device_ms"]:
            best = st
    return best, x


def report(name, n, st, cpu=None, extra=None):
    gbs = 8 * n * st["visits"] / (st["device_ms"] * 1e-3) / 1e9
    out = {"config": name, "visits": st["visits"], "accepted": st["accepted"], "passes": st["passes"],
           "full_passes": st["full_passes"], "nnz": st["nnz"], "converged": st["converged"],
           "gpu_device_ms": st["device_ms"], "gpu_wall_ms": st["wall_ms"],
           "gpu_visits_per_s": st["visits"] / (st["device_ms"] * 1e-3),
           "algorithmic_GBps(8n B/visit)": gbs, "hbm_frac_of_measured": gbs / HBM}
    if cpu:
        out["cpu_port_1thread_ms"] = cpu["device_ms"]
        out["cpu_visits_per_s"] = cpu["visits"] / (cpu["device_ms"] * 1e-3)
        out["speedup_device"] = cpu["device_ms"] / st["device_ms"]
        out["same_support"] = cpu.get("same_support")
        out["max_rel_diff"] = cpu.get("max_rel_diff")
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


def compare(xg, xr):
    a, b = xg.toarray(), xr.toarray()
    return bool(np.array_equal(a != 0, b != 0)), float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main(which=None, ref=None):
    """`ref`: a Backend bound to the CPU port, handed in by bench.py's cpu_baseline leg (`bench.py --other-configs ...
    --other-cpu`); this script itself never loads anything from oracle/ — run directly it measures the GPU only."""
    which = which or [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c3", "c4", "lvocv"]
    cpu_on = ref is not None
    gpu = cdgpu.default()
    opt = CDOptions(randomize=False)  # defaults: optTol 1e-7, maxIter 2000

    if "c1" in which:
        n, p = 1000, 5000
        X, y = problem(n, p, 10, 123)
        f = gpu.CDLeastSquaresLoss(y, X)
        fr = ref.CDLeastSquaresLoss(y, X) if cpu_on else None
        for lam in (math.sqrt(2 * math.log(p) / n), 0.05, 0.01):
            st, xg = timed_solve(gpu, f, lam, None, opt, reps=3)
            cpu = None
            if cpu_on:
                cpu, xr = timed_solve(ref, fr, lam, None, opt)
                cpu["same_support"], cpu["max_rel_diff"] = compare(xg, xr)
            report(f"C1 lasso n={n} p={p} lambda={lam:.4f} (naive LS, X {8 * n * p / 1e6:.0f} MB: L2 resident)", n, st, cpu)
        f.close()

    if "c3" in which:
        n, p = 5000, 50000
        X, y = problem(n, p, 20, 124)
        t0 = time.perf_counter()
        f = gpu.CDSqrtLassoLoss(y, X)
        h2d = time.perf_counter() - t0
        lam = 1.1 * math.sqrt(2 * math.log(p))
        st, xg = timed_solve(gpu, f, lam, None, opt, reps=2)
        cpu = None
        if cpu_on:
            fr = ref.CDSqrtLassoLoss(y, X)
            cpu, xr = timed_solve(ref, fr, lam, None, opt)
            cpu["same_support"], cpu["max_rel_diff"] = compare(xg, xr)
        report(f"C3 sqrt-lasso n={n} p={p} lambda={lam:.3f} (naive form, X {8 * n * p / 1e9:.1f} GB in HBM)", n, st, cpu,
               {"create_incl_h2d_ms": 1e3 * h2d})
        f.close()
        # scaled lasso, sigma loop on the device
        lam = math.sqrt(2 * math.log(p) / n)
        io = IterLassoOptions(initProcedure="InitStd", σinit=1.0, optionsCD=opt)
        fl = gpu.CDLeastSquaresLoss(y, X)
        om = fl.stdX()
        fl.close()
        x = SparseIterate(p)
        t0 = time.perf_counter()
        sol = gpu.scaledLasso_(x, X, y, lam, om, io)
        wall = time.perf_counter() - t0
        st = dict(sol.stats, wall_ms=1e3 * wall, nnz=x.nnz)
        cpu = None
        if cpu_on:
            xr = SparseIterate(p)
            solr = ref.scaledLasso_(xr, X, y, lam, om, io)
            cpu = dict(solr.stats)
            cpu["same_support"], cpu["max_rel_diff"] = compare(x, xr)
        report(f"C3 scaled lasso n={n} p={p} lambda={lam:.4f} (naive LS, sigma loop on device; wall includes H2D of X)", n,
               st, cpu, {"sigma": sol.σ, "outer_iters": st["outer_iters"]})

    if "c4" in which:
        n, p, degree, m = 500, 50, 2, 4096
        rng = np.random.default_rng(125)
        X = np.asfortranarray(rng.standard_normal((n, p)))
        Z = rng.random(n)
        cj = rng.choice([2, 4, 6, 8], size=p)
        Y = np.array([np.sin(cj * Z[i])[:2] @ X[i, :2] for i in range(n)]) + 0.1 * rng.standard_normal(n)
        zgrid = np.linspace(0.01, 0.99, m)
        k = GaussianKernel(0.2)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            out, _ = gpu.locpolyl1(X, Z, Y, zgrid, degree, k, 0.01, False, opt)
            wall = time.perf_counter() - t0
            dev = gpu.last_vc_stats[0]["device_ms"]
            if best is None or dev < best[0]:
                best = (dev, wall, gpu.last_vc_stats)
        dev, wall, stats = best
        visits = sum(s["visits"] for s in stats)
        res = {"config": f"C4 locpolyl1 {m} grid points, n={n} p={p} degree={degree} (ep={p * (degree + 1)}), Gaussian h=0.2, lambda0=0.01",
               "gpu_device_ms": dev, "gpu_wall_ms": 1e3 * wall, "problems_per_s_device": m / (dev * 1e-3),
               "problems_per_s_wall": m / wall, "visits": visits, "gpu_visits_per_s": visits / (dev * 1e-3),
               "all_converged": all(s["converged"] for s in stats),
               "algorithmic_GBps(8n(p+2)/problem + 8n/visit)": (8 * n * (p + 2) * m + 8 * n * visits) / (dev * 1e-3) / 1e9}
        if cpu_on:
            ms = 128  # bounded sample of the grid, same chain as the reference
            sub = np.ascontiguousarray(zgrid[:: m // ms])
            t0 = time.perf_counter()
            outr, _ = ref.locpolyl1(X, Z, Y, sub, degree, k, 0.01, False, opt)
            cw = time.perf_counter() - t0
            res["cpu_port_1thread_problems_per_s"] = len(sub) / cw
            res["cpu_sample"] = f"{len(sub)} of {m} grid points (every {m // ms}th), reference's warm-start chain"
            res["speedup_wall"] = (m / wall) / (len(sub) / cw)
            g = out[:, :: m // ms]
            res["same_support"] = bool(np.array_equal(g != 0, outr != 0))
            res["max_abs_diff"] = float(np.max(np.abs(g - outr)))
        print(json.dumps(res), flush=True)

    if "lvocv" in which:
        # lvocv_locpolyl1: n * numH leave-one-out scaled-lasso local problems (the largest natural batch of the package)
        from cdgpu import GaussianKernel as GK
        n, p, degree = 500, 50, 1
        rng = np.random.default_rng(126)
        X = np.asfortranarray(rng.standard_normal((n, p)))
        Z = rng.random(n)
        Y = X[:, 0] * np.sin(4 * Z) + X[:, 1] * np.sin(8 * Z) + 0.1 * rng.standard_normal(n)
        hs = np.exp(np.linspace(np.log(0.01), np.log(0.5), 8))
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            mse = gpu.lvocv_locpolyl1(X, Z, Y, degree, hs, GK, 0.3, opt)
            wall = time.perf_counter() - t0
            dev = gpu.last_vc_stats[0]["device_ms"]
            if best is None or dev < best[0]:
                best = (dev, wall, mse, gpu.last_vc_stats)
        dev, wall, mse, stats = best
        res = {"config": f"lvocv_locpolyl1 n={n} p={p} degree={degree}, {hs.size} bandwidths: {n * hs.size} leave-one-out scaled-lasso problems",
               "gpu_device_ms": dev, "gpu_wall_ms": 1e3 * wall, "problems_per_s_device": n * hs.size / (dev * 1e-3),
               "best_bandwidth": float(hs[int(np.argmin(mse))]), "mean_outer_iters": float(np.mean([s["outer_iters"] for s in stats]))}
        if cpu_on:
            t0 = time.perf_counter()
            mr = ref.lvocv_locpolyl1(X, Z, Y, degree, hs[:1], GK, 0.3, opt)
            cw = time.perf_counter() - t0
            res["cpu_port_1thread_problems_per_s"] = n / cw
            res["cpu_sample"] = "first bandwidth only (500 problems), reference's warm-start chain"
            res["mse_rel_diff_first_bandwidth"] = float(abs(mr[0] - mse[0]) / mr[0])
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
