#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 coordinate-descent path (BASELINE.json configs[1], "C2").

One STEP = one pass of the hot path over one synthetic problem (n = 10000, p = 20000, 50 true non-zeros):
    (X, y)  ->  covariance form of the lasso: diag(A), b = -X'y/n, columns of A = X'X/n
            ->  omega = _stdX!(X), lambda_max, 100 log-spaced lambdas down to 0.05*lambda_max
            ->  warm-started weighted-L1 lasso path by active-set coordinate descent (cluster kernel)
metric  = coordinate updates / s  = descendCoordinate! visits of the whole path / step time
value   = inputs (X, y) already resident in HBM, device pointers through the C ABI
e2e     = the same step through the C ABI with HOST buffers (pinned): H2D of X, y and D2H of the CSC path inside the
          timed region; `e2e_pageable` = the same from ordinary (pageable) numpy arrays, what a Julia Matrix is.
--gram lazy  (default) the covariance handle forms diag(A), b and only the columns of A that the path touches
             (cdgpu_gram_create_lazy: skinny FP64 tensor-core GEMMs on demand);
--gram eager forms the whole p x p Gram first (cdgpu_gram_create: FP64 tensor-core SYRK).  The line of the other
             variant is measured too and reported under `eager` / `lazy`.
N > 1   = one process per GPU.  The warm-started path of ONE problem is a sequential chain and does not shard:
          the headline at N > 1 is N independent replicas (another seed per rank: a CV fold / bootstrap replicate),
          value = sum of visits / max time ("replicas only").  The workloads that DO shard are timed
          in the same run and reported under `sharded`: C5 (tall Gram, rows sharded, one ncclAllReduce; strong
          scaling at fixed n) and C4 (4096 local problems dealt over the ranks; strong scaling).
--impl reference = the reference's CPU implementation of the same step on the host cores: Gram by OpenBLAS (what
          Julia's X'X/n calls) with all threads + the C port of the reference's single-threaded CD loop
          (oracle/libcdref_fast.so).  Each step is a PROPORTIONAL sample of the C2 step: the first L lambdas of the
          path at the full width p and the same fraction of the Gram's columns, so sample visits / sample time estimates
          whole-step visits / whole-step time (the reference's cost per visit does not depend on lambda, the Gram's cost
          is linear in the number of columns formed).
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))

C2 = dict(n=10000, p=20000, s=50, nlambda=100, ratio=0.05, optTol=1e-7, maxIter=2000)
# descendCoordinate! visits of the whole C2 path (seed 123): identical on the device and in the CPU port (asserted by the
# `parity` leg of every N=1 run of this file); the reference arm uses it to size its proportional sample
C2_VISITS = 3990575
# dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full captures under profiles/
NCU_TRAFFIC = {
    "gram_syrk_kernel": (30.228562e9 + 3.203871e9, "profiles/r1d_kernels_ncu.txt"),
    "naive_path_kernel_c3": (6.542697e9 + 0.010799e9, "profiles/r5_naive_path_ncu.txt"),
    "tall_sqrt_kernel": (9.621257e9 + 0.005187e9, "profiles/r8_tall_sqrt_ncu.txt (same n, p; profiles/run_kernels.py tall)"),
}
try:
    NCU_TRAFFIC.update({k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).items()})
except (OSError, ValueError):
    pass


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gram", default="lazy", choices=["lazy", "eager"])
    ap.add_argument("--n", type=int, default=C2["n"])
    ap.add_argument("--p", type=int, default=C2["p"])
    ap.add_argument("--nlambda", type=int, default=C2["nlambda"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-lambdas", type=int, default=0, help="cpu_baseline leg: only the first L lambdas (0 = the whole step)")
    ap.add_argument("--ref-lambdas", type=int, default=8, help="--impl reference: lambdas per sample step")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C3 / C4 side measurements")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the C5 / C4 sharded workloads")
    ap.add_argument("--c5-rows", type=int, default=8_000_000)
    ap.add_argument("--other-configs", default="", help="comma list of c1,c3,c4,lvocv: run benchmarks/other_configs.py instead")
    ap.add_argument("--other-cpu", action="store_true", help="with --other-configs: add the CPU-port columns (cpu_baseline leg)")
    return ap.parse_args()


def make_problem(n, p, s, seed):
    """X ~ N(0,1) (n x p, column-major), y = X[:, :s] beta + N(0,1).  numpy PCG64, fixed seed."""
    rng = np.random.default_rng(seed)
    Xt = rng.standard_normal((p, n))  # C-order (p, n) == F-order (n, p)
    X = Xt.T
    beta = rng.standard_normal(s) * (1.0 + rng.random(s))  # benchmark/cd_bench.jl:14
    y = X[:, :s] @ beta + rng.standard_normal(n)
    return X, np.ascontiguousarray(y)


def lambda_grid(lmax, ratio, m):
    return np.exp(np.linspace(np.log(lmax), np.log(ratio * lmax), m))


def host_threads(k):
    """numpy/OpenBLAS on k threads even when the launcher exported OMP_NUM_THREADS=1 (torchrun does)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=k)
    except Exception:  # noqa: BLE001
        import contextlib
        return contextlib.nullcontext()


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML (a polling
    `nvidia-smi -lms` process was seen to stall allocations and launches by tens of ms)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.stop_flag, self.t = index, [], set(), False, None
        self.mx = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


def run_step(be, make_handle, cfg, opts, keep_path=False):
    """One step against an already chosen input location."""
    f = make_handle()
    om = f.stdX()
    lmax = be.findLambdaMax(f, om)
    lams = lambda_grid(lmax, cfg["ratio"], cfg["nlambda"])
    path = be.LassoPath(None, None, lams, opts, standardizeX=om, loss=f)
    out = dict(visits=sum(s["visits"] for s in path.stats), accepted=sum(s["accepted"] for s in path.stats),
               gram_ms=f.gram_ms, cd_ms=path.stats[0]["device_ms"], nnz_last=path.βpath[-1].nnz,
               passes=sum(s["passes"] for s in path.stats), full_passes=sum(s["full_passes"] for s in path.stats),
               converged=all(s["converged"] for s in path.stats),
               d2h=16 * sum(x.nnz for x in path.βpath) + 8 * (len(lams) + 1) + 8 * f.p + 8)
    out.update(f.sweep_stats())
    if keep_path:
        out["path"], out["lams"], out["omega"] = path, lams, om
    f.close()
    return out


# ---------------------------------------------------------------------------------------------------- side configs
def c3_metrics(be, local, hbm, reps=3):
    """C3 (configs[2]): sqrt-lasso n=5000 p=50000, naive form, X (2 GB) resident in HBM — the HBM-bound sweep kernel.
    Timed region: `reps` solves after one warm-up, CUDA events on the library stream (device_ms), mean reported."""
    import torch

    import cdgpu
    from cdgpu import CDOptions, ProxL1, SparseIterate
    lib = be.lib
    n, p, s = 5000, 50000, 20
    g = torch.Generator(device="cuda")
    g.manual_seed(124)
    Xd = torch.empty((p, n), device="cuda", dtype=torch.float64)  # (p, n) C-order == (n, p) column-major
    for j0 in range(0, p, 5000):
        Xd[j0:j0 + 5000].normal_(generator=g)
    beta = torch.randn(s, device="cuda", dtype=torch.float64, generator=g) * (1.0 + torch.rand(s, device="cuda", dtype=torch.float64, generator=g))
    yd = Xd[:s].T @ beta + torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
    torch.cuda.synchronize()
    f = cdgpu.CDSqrtLassoLoss.__new__(cdgpu.CDSqrtLassoLoss)
    cdgpu.api._Loss.__init__(f, lib)
    f.n, f.p = n, p
    lib.check(lib.naive_create_dev(C.byref(f._h), cdgpu._ffi.LOSS_SQRT, C.c_void_p(Xd.data_ptr()), n, p, n,
                                   C.c_void_p(yd.data_ptr()), None, local))
    lam = 1.1 * math.sqrt(2 * math.log(p))
    runs = []
    for i in range(reps + 1):
        x = SparseIterate(p)
        be.coordinateDescent_(x, f, ProxL1(lam), CDOptions(randomize=False))
        if i:
            runs.append(dict(f.last_stats, nnz=x.nnz))
    f.close()
    del Xd, yd
    ms = float(np.mean([r["device_ms"] for r in runs]))
    st = runs[-1]
    gbs = 8 * n * st["visits"] / (ms * 1e-3) / 1e9
    tr = NCU_TRAFFIC.get("naive_path_kernel_c3", (None, None))
    return {"workload": f"C3 sqrt-lasso n={n} p={p} lambda={lam:.3f}, naive form, X 2.0 GB resident in HBM, one launch per solve",
            "kernel": "naive_path_kernel", "device_ms": ms, "device_ms_runs": [r["device_ms"] for r in runs],
            "visits": st["visits"], "passes": st["passes"], "full_passes": st["full_passes"], "nnz": st["nnz"],
            "converged": bool(st["converged"]), "visits_per_sec": st["visits"] / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                         "bytes_per_visit": 8 * n, "bytes_per_launch": 8 * n * st["visits"], "launch_ms": ms,
                         "traffic": tr[0], "traffic_source": tr[1]}}


def tall_sqrt_metrics(be, local, hbm, reps=3):
    """Tall sqrt-lasso (n = 10^6, p = 500: X 4 GB resident in HBM): CDSqrtLassoLoss beyond one CTA's shared memory, the
    row-distributed sweep of csrc/tall_sweep.cu.  Same timing as c3_metrics."""
    import torch

    import cdgpu
    from cdgpu import CDOptions, ProxL1, SparseIterate
    lib = be.lib
    n, p, s = 1000000, 500, 10
    g = torch.Generator(device="cuda")
    g.manual_seed(125)
    Xd = torch.empty((p, n), device="cuda", dtype=torch.float64)
    for j0 in range(0, p, 50):
        Xd[j0:j0 + 50].normal_(generator=g)
    beta = 1.0 + torch.rand(s, device="cuda", dtype=torch.float64, generator=g)
    yd = Xd[:s].T @ beta + torch.randn(n, device="cuda", dtype=torch.float64, generator=g)
    torch.cuda.synchronize()
    f = cdgpu.CDSqrtLassoLoss.__new__(cdgpu.CDSqrtLassoLoss)
    cdgpu.api._Loss.__init__(f, lib)
    f.n, f.p = n, p
    lib.check(lib.naive_create_dev(C.byref(f._h), cdgpu._ffi.LOSS_SQRT, C.c_void_p(Xd.data_ptr()), n, p, n,
                                   C.c_void_p(yd.data_ptr()), None, local))
    om = f.stdX()
    lam = 1.1 * math.sqrt(2 * math.log(p))
    runs = []
    for i in range(reps + 1):
        x = SparseIterate(p)
        be.coordinateDescent_(x, f, ProxL1(lam, om), CDOptions(randomize=False))
        if i:
            runs.append(dict(f.last_stats, nnz=x.nnz))
    f.close()
    del Xd, yd
    ms = float(np.mean([r["device_ms"] for r in runs]))
    st = runs[-1]
    gbs = 8 * n * st["visits"] / (ms * 1e-3) / 1e9
    tr = NCU_TRAFFIC.get("tall_sqrt_kernel", (None, None))
    return {"workload": f"tall sqrt-lasso n={n} p={p} lambda={lam:.3f} (omega = stdX), residual form with the rows dealt over the "
                        "SMs, X 4.0 GB resident in HBM, one launch per solve",
            "kernel": "tall_sqrt_kernel", "device_ms": ms, "device_ms_runs": [r["device_ms"] for r in runs],
            "visits": st["visits"], "passes": st["passes"], "full_passes": st["full_passes"], "nnz": st["nnz"],
            "converged": bool(st["converged"]), "visits_per_sec": st["visits"] / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                         "bytes_per_visit": 8 * n, "bytes_per_launch": 8 * n * st["visits"], "launch_ms": ms,
                         "traffic": tr[0], "traffic_source": tr[1]}}


def c1_metrics(be, ref, reps=3):
    """C1 (configs[0]): cd_bench.jl-style lasso n=1000 p=5000, 10 true non-zeros, single lambda — the reference's own
    CPU-runnable case.  Three lambdas (10 / ~230 / ~830 non-zeros).  Two device forms: the reference's residual form
    (CDLeastSquaresLoss -> naive_path_kernel) and the covariance form on the same data (lazy covariance handle ->
    cov_path_kernel, cold start from zero like `lasso`).  CPU column: the C port of the reference loop, one thread."""
    from cdgpu import CDOptions, ProxL1, SparseIterate
    n, p = 1000, 5000
    X, y = make_problem(n, p, 10, 123)
    opt = CDOptions(randomize=False)
    out = []
    fn = be.CDLeastSquaresLoss(y, X)
    fr = ref.CDLeastSquaresLoss(y, X) if ref is not None else None
    for lam in (math.sqrt(2 * math.log(p) / n), 0.05, 0.01):
        row = {"lambda": lam}
        runs = []
        for i in range(reps + 1):
            x = SparseIterate(p)
            be.coordinateDescent_(x, fn, ProxL1(lam), opt)
            if i:
                runs.append(fn.last_stats["device_ms"])
        st, bn = fn.last_stats, x.toarray()
        row.update({"naive_device_ms": float(np.mean(runs)), "visits": st["visits"], "accepted": st["accepted"], "passes": st["passes"],
                    "nnz": int(np.count_nonzero(bn)), "converged": bool(st["converged"])})
        runs = []
        for i in range(reps + 1):
            t0 = time.perf_counter()
            fc = be.CDQuadraticLoss_from_data(X, y, lazy=True)  # includes H2D of X (40 MB) and the diag / b pass
            x = SparseIterate(p)
            be.coordinateDescent_(x, fc, ProxL1(lam), opt)
            wall = time.perf_counter() - t0
            stc, ls = fc.last_stats, fc.sweep_stats()
            fc.close()
            if i:
                runs.append((stc["device_ms"], 1e3 * wall))
        bc = x.toarray()
        row.update({"cov_lazy_solve_device_ms": float(np.mean([r[0] for r in runs])),
                    "cov_lazy_wall_ms_incl_create_and_h2d": float(np.mean([r[1] for r in runs])),
                    "cov_lazy_columns": ls["columns"], "cov_lazy_pauses": ls["pauses"],
                    "cov_vs_naive_same_support": bool(np.array_equal(bc != 0, bn != 0)),
                    "cov_vs_naive_max_rel_diff": float(np.max(np.abs(bc - bn)) / max(np.max(np.abs(bn)), 1e-300))})
        if fr is not None:
            xr = SparseIterate(p)
            t0 = time.perf_counter()
            ref.coordinateDescent_(xr, fr, ProxL1(lam), opt)
            row["cpu_port_1thread_ms"] = 1e3 * (time.perf_counter() - t0)
            br = xr.toarray()
            row["naive_vs_cpu_same_support"] = bool(np.array_equal(bn != 0, br != 0))
            row["naive_vs_cpu_max_rel_diff"] = float(np.max(np.abs(bn - br)) / max(np.max(np.abs(br)), 1e-300))
            row["naive_vs_cpu_same_trace"] = bool(fn.last_stats["visits"] == fr.last_stats["visits"] and fn.last_stats["passes"] == fr.last_stats["passes"])
        out.append(row)
    fn.close()
    return {"workload": f"C1 lasso n={n} p={p}, 10 true non-zeros, single lambda, optTol 1e-7, ordered (X 40 MB: L2 resident)", "rows": out}


def c4_data():
    n, p = 500, 50
    rng = np.random.default_rng(125)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    Z = rng.random(n)
    cj = rng.choice([2, 4, 6, 8], size=p)
    Y = np.array([np.sin(cj * Z[i])[:2] @ X[i, :2] for i in range(n)]) + 0.1 * rng.standard_normal(n)
    return X, Z, Y


C4_OPTTOL = 1e-9  # batch (cold starts) and the reference's warm-start chain agree to 1e-6 at this tolerance (tests/test_baseline_configs.py)
C4_CHAIN = 2      # grid points per warm-started run (cdgpu_vc_solve_chain): 2048 runs still fill the GPU, every second point starts warm


def c4_metrics(be, hbm, reps=3):
    """C4 (configs[3]): 4096 kernel-weighted local problems of the varying-coefficient lasso."""
    from cdgpu import CDOptions, GaussianKernel
    n, p, degree, m = 500, 50, 2, 4096
    X, Z, Y = c4_data()
    zgrid = np.linspace(0.01, 0.99, m)
    opt = CDOptions(randomize=False, optTol=C4_OPTTOL)
    runs, cold = [], []
    for i in range(reps + 1):  # the batch with every grid point started from zero, for comparison
        out_cold, _ = be.locpolyl1(X, Z, Y, zgrid, degree, GaussianKernel(0.2), 0.01, False, opt)
        if i:
            cold.append(be.last_vc_stats[0]["device_ms"])
    for i in range(reps + 1):
        t0 = time.perf_counter()
        out_chain, _ = be.locpolyl1(X, Z, Y, zgrid, degree, GaussianKernel(0.2), 0.01, False, opt, chain=C4_CHAIN)
        wall = time.perf_counter() - t0
        if i:
            runs.append((be.last_vc_stats[0]["device_ms"], wall, be.last_vc_stats))
    dev = float(np.mean([r[0] for r in runs]))
    wall = float(np.mean([r[1] for r in runs]))
    stats = runs[-1][2]
    visits = int(sum(s["visits"] for s in stats))
    ep = p * (degree + 1)
    alg = 8 * n * (p + 2) * m + 8 * n * visits  # 8n(p+2) per local problem + 8n per visit
    tr = NCU_TRAFFIC.get("vc_cov_kernel_c4", (None, None))
    return {"workload": f"C4 locpolyl1: {m} grid points, n={n} p={p} degree={degree} (ep={ep}), Gaussian h=0.2, lambda0=0.01, optTol={C4_OPTTOL}, "
                        f"warm-start chain cut into runs of {C4_CHAIN} grid points",
            "kernels": "vc_build_z/v + gram_syrk_kernel<GEMM> (all local Grams as one DMMA GEMM) + vc_cov_kernel",
            "device_ms": dev, "device_ms_runs": [r[0] for r in runs], "wall_ms_incl_h2d_d2h": 1e3 * wall,
            "cold_batch_device_ms": float(np.mean(cold)), "cold_batch_device_ms_runs": cold,
            "max_rel_diff_chain_vs_cold_batch": float(np.max(np.abs(out_chain - out_cold)) / np.max(np.abs(out_cold))),
            "problems_per_sec": m / (dev * 1e-3), "problems_per_sec_e2e": m / wall, "visits": visits,
            "all_converged": all(s["converged"] for s in stats),
            "roofline": {"bound": "hbm", "achieved": alg / (dev * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": alg / (dev * 1e-3) / 1e9 / hbm, "bytes_per_launch": alg, "launch_ms": dev,
                         "traffic": tr[0], "traffic_source": tr[1],
                         "note": "latency-bound chains (one warp or CTA per local problem); bytes = 8n(p+2) per problem + 8n per visit"}}


# ---------------------------------------------------------------------------------------------------- sharded workloads
def sharded_metrics(be, args, rank, world, local, hbm):
    """The workloads that shard over GPUs, timed at this N: max over ranks of device / wall time."""
    import torch
    import torch.distributed as dist

    import cdgpu
    from cdgpu import CDOptions, GaussianKernel
    from cdgpu.distributed import Comm, gram_sharded, locpolyl1_sharded, shard_range
    lib = be.lib
    out = {}

    def mx(*vals):
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def keyed_rows(rows, p, row0, s, seed):
        """rows [row0, row0+rows) of the keyed design (p columns) and response, on this rank's GPU, column-major."""
        Xl = torch.empty((p, rows), device="cuda", dtype=torch.float64)
        lib.check(lib.synth_normal(C.c_void_p(Xl.data_ptr()), rows, p, rows, row0, 0, seed, local))
        noise = torch.empty(rows, device="cuda", dtype=torch.float64)
        lib.check(lib.synth_normal(C.c_void_p(noise.data_ptr()), rows, 1, rows, row0, p, seed, local))
        beta = torch.from_numpy(np.random.default_rng(seed).standard_normal(s) * 1.5).cuda()
        return Xl, Xl[:s].T @ beta + noise

    def make_gram(comm, Xl, yl, nl, n, p):
        if world > 1:
            return gram_sharded(be, comm, Xl.data_ptr(), nl, n, p, nl, yl.data_ptr())
        f = cdgpu.CDQuadraticLoss.__new__(cdgpu.CDQuadraticLoss)
        cdgpu.api._Loss.__init__(f, lib)
        f.p = p
        lib.check(lib.gram_create_dev(C.byref(f._h), C.c_void_p(Xl.data_ptr()), nl, p, nl, C.c_void_p(yl.data_ptr()), local))
        return f

    comm = Comm(be) if world > 1 else None
    # ---- self-check (driver-visible multi-GPU correctness): the row-sharded Gram equals the one-GPU Gram of the same keyed rows
    if world > 1:
        nchk, pchk = 262144, 512
        lo, hi = shard_range(nchk, rank, world)
        Xl, yl = keyed_rows(hi - lo, pchk, lo, 8, 991)
        fs = make_gram(comm, Xl, yl, hi - lo, nchk, pchk)
        As, bs = fs.get()
        fs.close()
        del Xl, yl
        rel = 0.0
        if rank == 0:
            Xf, yf = keyed_rows(nchk, pchk, 0, 8, 991)
            f1 = cdgpu.CDQuadraticLoss.__new__(cdgpu.CDQuadraticLoss)
            cdgpu.api._Loss.__init__(f1, lib)
            f1.p = pchk
            lib.check(lib.gram_create_dev(C.byref(f1._h), C.c_void_p(Xf.data_ptr()), nchk, pchk, nchk, C.c_void_p(yf.data_ptr()), local))
            A1, b1 = f1.get()
            f1.close()
            del Xf, yf
            rel = max(float(np.max(np.abs(As - A1)) / np.max(np.abs(A1))), float(np.max(np.abs(bs - b1)) / np.max(np.abs(b1))))
            out["gram_selfcheck"] = {"what": f"row-sharded Gram over {world} ranks (ncclAllReduce) vs the one-GPU Gram of the same keyed {nchk} x {pchk} rows",
                                     "max_rel_diff": rel, "symmetric": bool(np.array_equal(As, As.T)),
                                     "ok": bool(rel <= 1e-12 and np.array_equal(As, As.T))}
    # ---- C5: tall lasso, rows sharded, strong scaling at fixed n
    free_b, _ = torch.cuda.mem_get_info()
    n, p, s = args.c5_rows, 2000, 20
    while 8 * (n // world + 1) * p > 0.80 * free_b and n > 1_000_000:
        n //= 2
    lo, hi = shard_range(n, rank, world)
    nl = hi - lo
    Xl, yl = keyed_rows(nl, p, lo, s, 555)
    torch.cuda.synchronize()
    # optTol 1e-10 for this path: the Gram differs in the last bits with the number of ranks (summation order of the
    # allreduce), and the supports of the 100 solutions should not depend on that
    opts = CDOptions(randomize=False, optTol=1e-10, maxIter=20000)
    recs = []
    for rep in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f = make_gram(comm, Xl, yl, nl, n, p)
        t1 = time.perf_counter()
        om = f.stdX()
        lmax = be.findLambdaMax(f, om)
        lams = lambda_grid(lmax, 0.05, 100)
        path = be.LassoPath(None, None, lams, opts, standardizeX=om, loss=f)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        gms = f.gram_ms
        f.close()
        if rep:  # the first repetition is the warm-up (NCCL channels, pool growth)
            recs.append((gms, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2 - t0), path))
    gram_dev, gram_wall, path_wall, step_wall = mx(*[float(np.mean([r[i] for r in recs])) for i in range(4)])
    path = recs[-1][4]
    supp = []
    for bcol in path.βpath:  # supports, ignoring coefficients below 1e-6 of the column's largest (not yet decided at optTol)
        v = bcol.toarray()
        supp.append(tuple(int(k) for k in np.flatnonzero(np.abs(v) > 1e-6 * max(np.max(np.abs(v)), 1e-300))))
    import hashlib
    sig = hashlib.sha256(repr(supp).encode()).hexdigest()[:16]
    flops = n * p * (p + 1) + 2 * n * p
    visits = int(sum(st["visits"] for st in path.stats))
    out["c5_tall_gram"] = {
        "workload": f"C5 tall lasso n={n} p={p}: rows sharded over {world} GPU(s), local FP64 DMMA SYRK + ONE ncclAllReduce(sum, f64) of A|b "
                    f"({8 * (p * p + p) / 1e6:.0f} MB), then the 100-lambda covariance path replicated on every rank; keyed data identical for every N",
        "scaling": "strong", "rows_per_gpu": nl, "gram_allreduce_device_ms": gram_dev, "gram_wall_ms": gram_wall,
        "path_wall_ms": path_wall, "step_wall_ms": step_wall, "steps_timed": len(recs),
        "gram_tflops_aggregate": flops / (gram_dev * 1e-3) / 1e12, "visits": visits,
        "steps_per_sec": 1e3 / step_wall, "visits_per_sec": visits / (step_wall * 1e-3),
        "nnz_last": path.βpath[-1].nnz, "support_signature": sig,
        "support_signature_note": "sha256 of the 100 support sets (|beta| > 1e-6 max|beta|, path at optTol 1e-10): the same string at every N",
        "nccl_ranks_in_data_plane_collective": world}
    del Xl, yl
    # ---- C4: 4096 independent local problems dealt round-robin over the ranks (no data-path collective, final all_gather)
    Xc, Zc, Yc = c4_data()
    m = 4096
    zgrid = np.linspace(0.01, 0.99, m)
    opt = CDOptions(randomize=False, optTol=C4_OPTTOL)
    recs = []
    for rep in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if world > 1:
            full = locpolyl1_sharded(be, Xc, Zc, Yc, zgrid, 2, GaussianKernel(0.2), 0.01, opt, chain=C4_CHAIN)
        else:
            full, _ = be.locpolyl1(Xc, Zc, Yc, zgrid, 2, GaussianKernel(0.2), 0.01, False, opt, chain=C4_CHAIN)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if rep:
            recs.append((be.last_vc_stats[0]["device_ms"], 1e3 * wall))
    dev, wall = mx(float(np.mean([r[0] for r in recs])), float(np.mean([r[1] for r in recs])))
    out["c4_vc_lasso"] = {"workload": f"C4 locpolyl1 {m} grid points (n=500 p=50 degree=2) in warm-started runs of {C4_CHAIN}, the runs dealt round-robin over {world} GPU(s); optTol={C4_OPTTOL}",
                          "scaling": "strong", "device_ms": dev, "wall_ms_incl_copies_and_gather": wall,
                          "problems_per_sec": m / (dev * 1e-3), "problems_per_sec_e2e": m / (wall * 1e-3),
                          "coef_checksum": float(np.sum(np.abs(full))), "nnz": int(np.count_nonzero(full)),
                          "limiter": "the hardest single grid point is one sequential chain (passes x active entries dependent steps): "
                                     "device time cannot fall below it however many GPUs share the other problems"}
    if comm:
        comm.close()
    return out


# ---------------------------------------------------------------------------------------------------- reference arm
def reference_setup(cfg):
    import cdgpu
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    ref = cdgpu.Backend(cdgpu.Lib(os.path.join(ROOT, "oracle", "libcdref_fast.so"), "cdref"))
    return ref


def cpu_gram(X, y, cores, ncols=None):
    """A[:, :ncols] = X'X[:, :ncols]/n, b = -X'y/n with numpy/OpenBLAS on `cores` threads (ncols = None: all of A, made
    exactly symmetric); returns (A, b, seconds).  Cost is linear in ncols (flops and output bytes alike)."""
    n, p = X.shape
    with host_threads(cores):
        t0 = time.perf_counter()
        A = X.T @ (X if ncols is None else X[:, :ncols])
        A /= n
        b = -(X.T @ y) / n
        dt = time.perf_counter() - t0
    if ncols is None and not np.array_equal(A, A.T):
        A = (A + A.T) * 0.5
    return np.asfortranarray(A), b, dt


def reference_main(args, cfg, base, workload):
    """CPU arm.  Proportional sample per step: CD over the first L lambdas (fraction f of the path's visits) + the Gram
    of the first f*n rows.  At N > 1 the job is N replicas: N Gram samples back to back (each on all threads) and N CD
    prefixes side by side (one core each) — what the host can do for the same N-replica job."""
    from cdgpu import CDOptions
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    N = max(1, args.gpus)
    ref = reference_setup(cfg)
    n, p = cfg["n"], cfg["p"]
    X, y = make_problem(n, p, cfg["s"], seed=123)
    opts = CDOptions(maxIter=cfg["maxIter"], optTol=cfg["optTol"], randomize=False, warmStart=True)
    A, b, gram_full_s = cpu_gram(X, y, cores)  # setup: the matrix the CD sample runs on (timed once, reported)
    L = max(1, min(args.ref_lambdas, cfg["nlambda"]))
    default_cfg = (n, p, cfg["nlambda"], cfg["s"]) == (C2["n"], C2["p"], C2["nlambda"], C2["s"])

    def cd_prefix(out, i, nl):
        f = ref.CDQuadraticLoss(A, b)
        om = f.stdX()
        lmax = ref.findLambdaMax(f, om)
        lams = lambda_grid(lmax, cfg["ratio"], cfg["nlambda"])[:nl]
        path = ref.LassoPath(None, None, lams, opts, standardizeX=om, loss=f)
        f.close()
        out[i] = sum(s["visits"] for s in path.stats)

    v_total = C2_VISITS
    if not default_cfg:  # unknown path length: run it once (slow, exact)
        tmp = [0]
        cd_prefix(tmp, 0, cfg["nlambda"])
        v_total = tmp[0]

    def step():
        t0 = time.perf_counter()
        vis = [0] * N
        th = [threading.Thread(target=cd_prefix, args=(vis, i, L)) for i in range(N)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        t1 = time.perf_counter()
        frac = vis[0] / v_total
        cols = max(16, int(round(frac * p)))
        tg = 0.0
        for _ in range(N):
            tg += cpu_gram(X, y, cores, ncols=cols)[2]
        t2 = time.perf_counter()
        return sum(vis), t1 - t0, tg, t2 - t0, frac, cols

    for _ in range(min(args.warmup, 1)):
        step()
    tot_v, tc, tg, tt = 0, 0.0, 0.0, 0.0
    for _ in range(args.steps):
        v, a_, g_, w_, frac, rows = step()
        tot_v += v
        tc += a_
        tg += g_
        tt += w_
    val = tot_v / tt
    K = args.steps
    sample = (f"proportional sample of the C2 step per timed step: CD over the first {L} of {cfg['nlambda']} lambdas at the full width "
              f"p={p} on the full Gram ({frac:.4f} of the path's {v_total} visits) + the first {rows} of {p} columns of the Gram (all {n} rows); "
              f"Gram numpy/OpenBLAS on {cores} threads ({tg / K:.2f} s/step; the full Gram took {gram_full_s:.1f} s in setup), CD = C port "
              f"of the reference's single-threaded loop incl. its unconditional O(p) axpy per visit ({tc / K:.2f} s/step)"
              + (f"; N={N} replicas: {N} Gram samples back to back, {N} CD prefixes on {N} cores side by side" if N > 1 else "")
              + f"; estimated whole step on this host: {tt / K / frac:.1f} s")
    out = dict(base, impl="reference", value=val, ms_per_step=1e3 * tt / K,
               config={"workload": workload, "optTol": cfg["optTol"], "randomize": False, "replicas": N},
               cpu_baseline={"value": val, "unit": "visits/s", "cores": cores, "kind": "port", "sample": sample},
               e2e={"value": val, "unit": "visits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------- main
def main():
    args = parse()
    cfg = dict(C2, n=args.n, p=args.p, nlambda=args.nlambda)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"C2: weighted-L1 lasso lambda-path, covariance form, n={cfg['n']} p={cfg['p']}, "
                f"{cfg['nlambda']} lambdas (lambda_max -> {cfg['ratio']} lambda_max) warm-started, covariance formed in the step")
    base = {"metric": "coordinate_updates_per_sec", "unit": "visits/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic"}

    if args.other_configs:
        import cdgpu
        sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
        import other_configs
        ref = None
        if args.other_cpu:  # cpu_baseline leg: the only place outside tests/ and smoke() that executes oracle/
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
            ref = cdgpu.Backend(cdgpu.Lib(os.path.join(ROOT, "oracle", "libcdref_fast.so"), "cdref"))
        other_configs.main(args.other_configs.split(","), ref)
        return

    if args.impl == "reference":
        reference_main(args, cfg, base, workload)
        return

    import torch
    import torch.distributed as dist

    import cdgpu
    from cdgpu import CDOptions
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; libcdgpu has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = cdgpu.Backend(cdgpu.load_product(), device=local)
    lib = be.lib
    opts = CDOptions(maxIter=cfg["maxIter"], optTol=cfg["optTol"], randomize=False, warmStart=True)
    n, p = cfg["n"], cfg["p"]
    X, y = make_problem(n, p, cfg["s"], seed=123 + rank)
    Xpg = np.ascontiguousarray(X.T)  # (p, n) C-order == (n, p) F-order, ordinary pageable memory
    # pinned host copies for the e2e leg, resident device copies for the value leg
    Xp = torch.from_numpy(Xpg).pin_memory()
    yp = torch.from_numpy(y).pin_memory()
    Xd, yd = Xp.cuda(non_blocking=False), yp.cuda(non_blocking=False)
    torch.cuda.synchronize()
    hb0, hb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hb0.record()
    Xd.copy_(Xp, non_blocking=True)
    hb1.record()
    torch.cuda.synchronize()
    h2d_gbs = 8 * n * p / (hb0.elapsed_time(hb1) * 1e-3) / 1e9

    def maker(kind, where):
        dev = where == "dev"
        fn = {("lazy", True): lib.gram_create_lazy_dev, ("lazy", False): lib.gram_create_lazy,
              ("eager", True): lib.gram_create_dev, ("eager", False): lib.gram_create}[(kind, dev)]
        if where == "dev":
            px, py = Xd.data_ptr(), yd.data_ptr()
        elif where == "pinned":
            px, py = Xp.data_ptr(), yp.data_ptr()
        else:
            px, py = Xpg.ctypes.data, y.ctypes.data

        def make():
            f = cdgpu.CDQuadraticLoss.__new__(cdgpu.CDQuadraticLoss)
            cdgpu.api._Loss.__init__(f, lib)
            f.p = p
            lib.check(fn(C.byref(f._h), C.c_void_p(px), n, p, n, C.c_void_p(py), local))
            return f
        return make

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches():
        c = C.c_int64()
        lib.launch_count(C.byref(c))
        return c.value

    def timed(make_handle, steps, warmup, sampler=None):
        for _ in range(warmup):
            run_step(be, make_handle, cfg, opts)
        barrier()
        if sampler:
            sampler.start()
        l0 = launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        res = [run_step(be, make_handle, cfg, opts) for _ in range(steps)]
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        secs = max(wall, e0.elapsed_time(e1) * 1e-3)
        vis = sum(r["visits"] for r in res)
        if world > 1:
            t = torch.tensor([secs], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
            t = torch.tensor([vis], device="cuda", dtype=torch.float64)
            dist.all_reduce(t)
            vis = float(t.item())
        return dict(res=res, secs=secs, visits=vis, launches=launches() - l0, clocks=clocks,
                    value=vis / secs, ms_per_step=1e3 * secs / steps)

    K, W = args.steps, args.warmup
    head, other = args.gram, ("eager" if args.gram == "lazy" else "lazy")
    sampler = ClockSampler(local)
    if os.environ.get("CDGPU_NO_SAMPLER"):
        sampler = None
    main_leg = timed(maker(head, "dev"), K, W, sampler)
    e2e_leg = timed(maker(head, "pinned"), K, W)
    pg_leg = timed(maker(head, "pageable"), K, min(W, 1))
    Ko = max(1, min(K, 5))
    other_leg = timed(maker(other, "dev"), Ko, 1)
    other_e2e = timed(maker(other, "pinned"), Ko, 1)
    clocks = main_leg["clocks"] or {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler disabled"]}

    sharded = None
    if world > 1 and not args.no_sharded:
        peaks0 = {}
        try:
            peaks0 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        del Xd, yd, Xp, yp
        torch.cuda.empty_cache()
        try:
            sharded = sharded_metrics(be, args, rank, world, local, peaks0.get("hbm_gbs", 6650.0))
        except Exception as e:  # noqa: BLE001  (side measurements must never cost the headline line)
            sharded = {"error": repr(e)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    # FP64 tensor ceiling: MEASURED_PEAKS.json has no FP64 entry -> cuBLAS Dgemm measured here, outside the timed region
    a64 = torch.randn(8192, 8192, device="cuda", dtype=torch.float64)
    b64 = torch.randn(8192, 8192, device="cuda", dtype=torch.float64)
    best = 1e9
    for i in range(6):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        torch.mm(a64, b64)
        s1.record()
        torch.cuda.synchronize()
        if i:
            best = min(best, s0.elapsed_time(s1))
    dgemm_tf = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    del a64, b64
    dg_src = "cuBLAS Dgemm 8192^3 FP64 measured in this run (MEASURED_PEAKS.json has no FP64 entry)"

    def mean(leg, key):
        return float(np.mean([r[key] for r in leg["res"]]))

    def leg_summary(leg, legE, kind):
        r0 = leg["res"][-1]
        d = {"gram": kind, "value": leg["value"], "ms_per_step": leg["ms_per_step"], "steps": len(leg["res"]),
             "e2e_value": legE["value"], "e2e_ms_per_step": legE["ms_per_step"],
             "create_device_ms": mean(leg, "gram_ms"), "path_device_ms": mean(leg, "cd_ms"),
             "sweep_kernel_ms": mean(leg, "sweep_ms"), "visits": r0["visits"], "accepted": r0["accepted"],
             "passes": r0["passes"], "full_passes": r0["full_passes"], "nnz_at_last_lambda": r0["nnz_last"],
             "all_converged": bool(r0["converged"]), "gpu_launches": leg["launches"]}
        if kind == "lazy":
            d.update({"columns_formed": r0["columns"], "column_batches": r0["batches"], "kernel_pauses": r0["pauses"],
                      "column_form_ms": mean(leg, "form_ms")})
        return d

    lazy_leg, eager_leg = (main_leg, other_leg) if head == "lazy" else (other_leg, main_leg)
    lazy_e2e, eager_e2e = (e2e_leg, other_e2e) if head == "lazy" else (other_e2e, e2e_leg)
    r0 = main_leg["res"][-1]
    sweep_ms = mean(main_leg, "sweep_ms")
    sweep_bytes = 24 * r0["visits"] + 8 * p * r0["accepted"]  # 24 B/visit + 8p B/accepted step
    tr_sweep = NCU_TRAFFIC.get("cov_path_kernel", (None, None))
    roof_sweep = {"kernel": "cov_path_kernel (cluster CD sweep: whole lambda path)", "bound": "hbm",
                  "achieved": sweep_bytes / (sweep_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                  "frac": sweep_bytes / (sweep_ms * 1e-3) / 1e9 / hbm, "traffic": tr_sweep[0], "traffic_source": tr_sweep[1],
                  "bytes_per_launch": sweep_bytes, "launch_ms": sweep_ms, "peak_source": hbm_src,
                  "note": "a sequential Gauss-Seidel chain on ONE 16-CTA cluster: bounded by the latency of dependent steps, not by bandwidth; "
                          "algorithmic bytes = 24 B/visit + 8p B/accepted step (SURVEY.md §8(d))"}
    gram_flops = n * p * (p + 1) + 2 * n * p  # SYRK (lower triangle) + X'y, algorithmic
    eg_ms = mean(eager_leg, "gram_ms")
    tr_gram = NCU_TRAFFIC.get("gram_syrk_kernel", (None, None))
    roof_gram = {"kernel": "gram_syrk_kernel (FP64 DMMA SYRK, eager covariance form)", "bound": "tensor",
                 "achieved": gram_flops / (eg_ms * 1e-3) / 1e12, "peak": dgemm_tf, "unit": "TFLOP/s",
                 "frac": gram_flops / (eg_ms * 1e-3) / 1e12 / dgemm_tf,
                 "traffic": tr_gram[0] if (n, p) == (C2["n"], C2["p"]) else None, "traffic_source": tr_gram[1],
                 "flops_per_launch": gram_flops, "launch_ms": eg_ms, "peak_source": dg_src}
    lz = lazy_leg["res"][-1]
    lz_form = mean(lazy_leg, "form_ms")
    roof_cols = None
    if lz.get("form_columns"):
        # the BLOCKING batches only (the sweep kernel waits for them; background batches are not timed by the library):
        # 2 n p flops per column formed, widths of 32 columns and up
        fl = 2.0 * n * p * lz["form_columns"]
        roof_cols = {"kernel": "gram_syrk_kernel<GEMM> (FP64 DMMA, columns of A = X'X/n on demand: blocking batches of 32+ columns, row-split)",
                     "bound": "tensor", "achieved": fl / (lz_form * 1e-3) / 1e12, "peak": dgemm_tf, "unit": "TFLOP/s",
                     "frac": fl / (lz_form * 1e-3) / 1e12 / dgemm_tf, "flops_per_solve": fl, "ms_per_solve": lz_form,
                     "columns_timed": lz["form_columns"], "columns_total": lz["columns"],
                     "note": "includes the column gather and the slab reduction; a 32-column batch fills a quarter of the 128-wide DMMA tile",
                     "peak_source": dg_src}
    dominant = roof_sweep if head == "lazy" else roof_gram
    out = dict(base, value=main_leg["value"], ms_per_step=main_leg["ms_per_step"],
               config={"workload": workload, "optTol": cfg["optTol"], "randomize": False, "replicas": world, "gram": head,
                       "l2": "inputs exceed L2: X %.1f GB per replica (eager G %.1f GB)" % (8 * n * p / 1e9, 8 * p * p / 1e9)},
               clocks=clocks,
               e2e={"value": e2e_leg["value"], "unit": "visits/s", "ms_per_step": e2e_leg["ms_per_step"], "steps": K, "warmup": W,
                    "host_memory": "pinned", "h2d_bytes_per_step": 8 * n * p + 8 * n + 8 * p * 2 + 8 * cfg["nlambda"],
                    "h2d_gbs_pinned_measured": h2d_gbs, "d2h_bytes_per_step": int(e2e_leg["res"][-1]["d2h"])},
               e2e_pageable={"value": pg_leg["value"], "unit": "visits/s", "ms_per_step": pg_leg["ms_per_step"], "steps": K,
                             "host_memory": "pageable numpy arrays (what a Julia Matrix{Float64} is), staged by the library"},
               gpu_launches=int(main_leg["launches"]),
               roofline=dominant, roofline_sweep=roof_sweep, roofline_gram=roof_gram, roofline_lazy_columns=roof_cols,
               lazy=leg_summary(lazy_leg, lazy_e2e, "lazy"), eager=leg_summary(eager_leg, eager_e2e, "eager"),
               breakdown={"create_device_ms": mean(main_leg, "gram_ms"), "cd_path_ms": mean(main_leg, "cd_ms"),
                          "sweep_kernel_ms": sweep_ms, "visits_per_step": r0["visits"], "accepted_per_step": r0["accepted"],
                          "passes": r0["passes"], "full_passes": r0["full_passes"], "nnz_at_last_lambda": r0["nnz_last"],
                          "all_converged": bool(r0["converged"]), "cd_only_visits_per_sec": r0["visits"] / (mean(main_leg, "cd_ms") * 1e-3)})
    if args.gpus == 1 and not args.no_cpu_baseline:
        out.update(cpu_leg(args, cfg, be, maker(head, "dev"), opts))
    if world == 1:
        del Xd, yd, Xp, yp
        torch.cuda.empty_cache()
        if not args.no_sharded:  # the N = 1 points of the sharded workloads' scaling curves
            try:
                sharded = sharded_metrics(be, args, rank, world, local, hbm)
            except Exception as e:  # noqa: BLE001
                sharded = {"error": repr(e)}
    if sharded is not None:
        out["sharded"] = sharded
    if args.gpus == 1 and not args.no_secondary:
        sec = {}
        ref1 = None
        if not args.no_cpu_baseline:
            try:
                ref1 = reference_setup(cfg)
            except Exception:  # noqa: BLE001
                ref1 = None
        for name, fn in (("c1_lasso", lambda: c1_metrics(be, ref1)), ("c3_sqrt_lasso", lambda: c3_metrics(be, local, hbm)),
                         ("c4_vc_lasso", lambda: c4_metrics(be, hbm)),
                         ("tall_sqrt_lasso", lambda: tall_sqrt_metrics(be, local, hbm))):
            try:
                sec[name] = fn()
            except Exception as e:  # noqa: BLE001  (side measurements must never cost the headline line)
                sec[name] = {"error": repr(e)}
        out["secondary"] = sec
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_leg(args, cfg, be, make_handle, opts):
    """cpu_baseline (the reference's algorithm on the host cores, ONE whole C2 step: Gram on all threads + the single-threaded CD
    path) and `parity`: the device path against that CPU path, column by column."""
    ref = reference_setup(cfg)
    n, p = cfg["n"], cfg["p"]
    cores = os.cpu_count()
    X, y = make_problem(n, p, cfg["s"], seed=123)
    g = run_step(be, make_handle, cfg, opts, keep_path=True)
    A, b, tg = cpu_gram(X, y, cores)
    L = args.cpu_lambdas if args.cpu_lambdas > 0 else cfg["nlambda"]
    t0 = time.perf_counter()
    f = ref.CDQuadraticLoss(A, b)
    om = f.stdX()
    lmax = ref.findLambdaMax(f, om)
    lams = lambda_grid(lmax, cfg["ratio"], cfg["nlambda"])[:L]
    path = ref.LassoPath(None, None, lams, opts, standardizeX=om, loss=f)
    f.close()
    tc = time.perf_counter() - t0
    v = sum(s["visits"] for s in path.stats)
    gp = g["path"]
    same_support, max_rel, visits_equal, obj_diff = True, 0.0, True, 0.0
    for i in range(L):
        bg, br = gp.βpath[i].toarray(), path.βpath[i].toarray()
        same_support &= bool(np.array_equal(bg != 0, br != 0))
        if np.any(br != 0):
            max_rel = max(max_rel, float(np.max(np.abs(bg - br)) / np.max(np.abs(br))))
        visits_equal &= gp.stats[i]["visits"] == path.stats[i]["visits"] and gp.stats[i]["passes"] == path.stats[i]["passes"]
        if i == L - 1:
            def obj(beta):
                return 0.5 * beta @ (A @ beta) + b @ beta + lams[i] * np.sum(om * np.abs(beta))
            obj_diff = float(abs(obj(bg) - obj(br)))
    whole = L == cfg["nlambda"]
    frac = v / max(1, g["visits"]) if visits_equal else float("nan")
    est_step = tg + tc / frac if not whole else tg + tc
    sample = (f"ONE whole C2 step on the host: " if whole else f"first {L} of {cfg['nlambda']} lambdas (CD time scaled by visits, {frac:.3f} of the path): ")
    sample += (f"Gram numpy/OpenBLAS {cores} threads {tg:.2f} s + C port of the reference CD loop (1 thread, literal always-axpy, "
               f"8p bytes per visit) {tc:.2f} s for {v} visits")
    return {"cpu_baseline": {"value": g["visits"] / est_step, "unit": "visits/s", "cores": cores, "kind": "port", "sample": sample,
                             "step_seconds": est_step, "gram_seconds": tg, "cd_seconds": tc},
            "parity": {"against": "CPU port of the reference (oracle/cdref.c) on the same inputs, same optTol, eager numpy Gram",
                       "lambdas_compared": L, "identical_supports": bool(same_support), "max_rel_coef_diff": max_rel,
                       "objective_abs_diff_last_lambda": obj_diff, "identical_passes_and_visits": bool(visits_equal),
                       "ok": bool(same_support and max_rel <= 1e-6 and obj_diff <= 1e-8)}}


if __name__ == "__main__":
    main()
