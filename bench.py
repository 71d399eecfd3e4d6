#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 coordinate-descent path (BASELINE.json configs[1], "C2").

One STEP = one pass of the hot path over one synthetic problem:
    (X, y)  ->  covariance form  A = X'X/n, b = -X'y/n  (FP64 tensor-core SYRK)
            ->  omega = _stdX!(X), lambda_max, 100 log-spaced lambdas down to 0.05*lambda_max
            ->  warm-started weighted-L1 lasso path by active-set coordinate descent (cluster kernel)
metric  = coordinate updates / s  = descendCoordinate! visits of the whole path / step time
value   = inputs (X, y) already resident in HBM, device pointers through the C ABI
e2e     = the same step through the C ABI with HOST buffers (pinned): H2D of X,y and D2H of the
          CSC path inside the timed region
N > 1   = one process per GPU, every rank solves its own replica (another seed: a CV fold /
          bootstrap replicate); a warm-started path is a sequential chain, so there is no
          data-path collective ("replicas only"); value = sum of visits / max time.
--impl reference = the reference's CPU implementation of the same step on the host cores: Gram by
          OpenBLAS (what Julia's X'X/n calls) with all threads + the C port of the reference's
          single-threaded CD loop (oracle/libcdref_fast.so), on a bounded sample (fewer columns).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))

C2 = dict(n=10000, p=20000, s=50, nlambda=100, ratio=0.05, optTol=1e-7, maxIter=2000)
REF_SAMPLE_P = 4000  # columns of the reference arm's bounded sample
NCU_GRAM_DRAM_BYTES = 30.228562e9 + 3.203871e9  # profiles/r1d_kernels_ncu.txt: gram_syrk_kernel at C2, per launch
NCU_C3_DRAM_BYTES = 6.059584e9 + 0.009213e9     # profiles/r1d_kernels_ncu.txt: naive_path_kernel at C3, per launch


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=C2["n"])
    ap.add_argument("--p", type=int, default=C2["p"])
    ap.add_argument("--nlambda", type=int, default=C2["nlambda"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-p", type=int, default=REF_SAMPLE_P)
    ap.add_argument("--no-secondary", action="store_true", help="skip the C3 / C4 side measurements")
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
            time.sleep(0.2)  # NVML queries take driver locks: a tighter loop was seen to stall cudaMallocAsync/launches

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


def run_step(lib, be, make_handle, cfg, opts):
    """One step against an already chosen input location; returns (visits, accepted, gram_ms, cd_ms, nnz_last, bytes_out)."""
    t0 = time.perf_counter()
    f = make_handle()
    t1 = time.perf_counter()
    om = f.stdX()
    lmax = be.findLambdaMax(f, om)
    lams = lambda_grid(lmax, cfg["ratio"], cfg["nlambda"])
    t2 = time.perf_counter()
    path = be.LassoPath(None, None, lams, opts, standardizeX=om, loss=f)
    t3 = time.perf_counter()
    gram_ms = f.gram_ms
    if os.environ.get("CDGPU_BENCH_VERBOSE") and os.environ.get("CDGPU_BENCH_LAZY"):
        print("lazy:", f.lazy_stats(), file=sys.stderr)
    f.close()
    t4 = time.perf_counter()
    if os.environ.get("CDGPU_BENCH_VERBOSE"):
        print("step: create %.1f ms (gram kernel %.1f) stdx+lmax %.1f path %.1f (device %.1f) close %.1f" % (
            1e3 * (t1 - t0), gram_ms, 1e3 * (t2 - t1), 1e3 * (t3 - t2), path.stats[0]["device_ms"], 1e3 * (t4 - t3)),
            file=sys.stderr)
    visits = sum(s["visits"] for s in path.stats)
    accepted = sum(s["accepted"] for s in path.stats)
    cd_ms = path.stats[0]["device_ms"]
    nnz_tot = sum(x.nnz for x in path.βpath)
    return dict(visits=visits, accepted=accepted, gram_ms=gram_ms, cd_ms=cd_ms, nnz_last=path.βpath[-1].nnz,
                passes=sum(s["passes"] for s in path.stats), full_passes=sum(s["full_passes"] for s in path.stats),
                converged=all(s["converged"] for s in path.stats), d2h=16 * nnz_tot + 8 * (len(lams) + 1) + 8 * f.p + 8)


def secondary_metrics(be, local, hbm):
    """The other two numbers BASELINE.json's metric names, outside the timed region of the headline:
    C3 (configs[2]): sqrt-lasso n=5000 p=50000 in naive form — the HBM-bound sweep kernel, X resident in HBM;
    C4 (configs[3]): 4096 kernel-weighted local problems of the varying-coefficient lasso, problems/s."""
    import math

    import torch

    import cdgpu
    from cdgpu import CDOptions, GaussianKernel, ProxL1, SparseIterate
    lib, out = be.lib, {}
    # ---- C3
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
    best = None
    for _ in range(3):
        x = SparseIterate(p)
        be.coordinateDescent_(x, f, ProxL1(lam), CDOptions(randomize=False))
        st = dict(f.last_stats, nnz=x.nnz)
        if best is None or st["device_ms"] < best["device_ms"]:
            best = st
    f.close()
    del Xd, yd
    gbs = 8 * n * best["visits"] / (best["device_ms"] * 1e-3) / 1e9
    out["c3_sqrt_lasso"] = {"workload": f"sqrt-lasso n={n} p={p} lambda={lam:.3f}, naive form, X 2.0 GB resident in HBM, one launch",
                            "kernel": "naive_path_kernel", "device_ms": best["device_ms"], "visits": best["visits"],
                            "passes": best["passes"], "full_passes": best["full_passes"], "nnz": best["nnz"],
                            "converged": bool(best["converged"]), "visits_per_sec": best["visits"] / (best["device_ms"] * 1e-3),
                            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                         "bytes_per_visit": 8 * n, "traffic": NCU_C3_DRAM_BYTES,
                                         "traffic_source": "profiles/r1d_kernels_ncu.txt (ncu --set full, one launch)"}}
    # ---- C4
    n, p, degree, m = 500, 50, 2, 4096
    rng = np.random.default_rng(125)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    Z = rng.random(n)
    cj = rng.choice([2, 4, 6, 8], size=p)
    Y = np.array([np.sin(cj * Z[i])[:2] @ X[i, :2] for i in range(n)]) + 0.1 * rng.standard_normal(n)
    zgrid = np.linspace(0.01, 0.99, m)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        be.locpolyl1(X, Z, Y, zgrid, degree, GaussianKernel(0.2), 0.01, False, CDOptions(randomize=False))
        wall = time.perf_counter() - t0
        dev = be.last_vc_stats[0]["device_ms"]
        if best is None or dev < best[0]:
            best = (dev, wall, be.last_vc_stats)
    dev, wall, stats = best
    out["c4_vc_lasso"] = {"workload": f"locpolyl1: {m} grid points, n={n} p={p} degree={degree} (ep={p * (degree + 1)}), Gaussian h=0.2, lambda0=0.01",
                          "kernels": "vc_build_z/v + gram_syrk_kernel<GEMM> (all local Grams as one DMMA GEMM) + vc_cov_kernel",
                          "device_ms": dev, "wall_ms_incl_h2d_d2h": 1e3 * wall, "problems_per_sec": m / (dev * 1e-3),
                          "problems_per_sec_e2e": m / wall, "visits": int(sum(s["visits"] for s in stats)),
                          "all_converged": all(s["converged"] for s in stats)}
    return out


def reference_arm(args, cfg):
    """CPU: OpenBLAS Gram (all threads) + C port of the reference CD loop (1 thread) on a bounded sample."""
    import cdgpu
    from cdgpu import CDOptions
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    ref = cdgpu.Backend(cdgpu.Lib(os.path.join(ROOT, "oracle", "libcdref_fast.so"), "cdref"))
    ps = min(args.cpu_sample_p, cfg["p"])
    X, y = make_problem(cfg["n"], ps, cfg["s"], seed=123)
    n = cfg["n"]
    opts = CDOptions(maxIter=cfg["maxIter"], optTol=cfg["optTol"], randomize=False, warmStart=True)
    cores = os.cpu_count()

    def step():
        t0 = time.perf_counter()
        A = X.T @ X
        A /= n
        A = np.asfortranarray((A + A.T) * 0.5) if not np.array_equal(A, A.T) else np.asfortranarray(A)
        b = -(X.T @ y) / n
        t1 = time.perf_counter()
        f = ref.CDQuadraticLoss(A, b)
        om = f.stdX()
        lmax = ref.findLambdaMax(f, om)
        lams = lambda_grid(lmax, cfg["ratio"], cfg["nlambda"])
        path = ref.LassoPath(None, None, lams, opts, standardizeX=om, loss=f)
        f.close()
        t2 = time.perf_counter()
        return sum(s["visits"] for s in path.stats), t1 - t0, t2 - t1

    return step, cores, ps


def main():
    args = parse()
    cfg = dict(C2, n=args.n, p=args.p, nlambda=args.nlambda)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"C2: weighted-L1 lasso lambda-path, covariance form, n={cfg['n']} p={cfg['p']}, "
                f"{cfg['nlambda']} lambdas (lambda_max -> {cfg['ratio']} lambda_max) warm-started, Gram formed in the step")
    base = {"metric": "coordinate_updates_per_sec", "unit": "visits/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic"}

    # ------------------------------------------------------------------ the other BASELINE configs (one JSON line each)
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

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        step, cores, ps = reference_arm(args, cfg)
        for _ in range(min(args.warmup, 1)):
            step()
        tot_v, tg, tc = 0, 0.0, 0.0
        for _ in range(args.steps):
            v, a, b = step()
            tot_v += v
            tg += a
            tc += b
        secs = tg + tc
        val = tot_v / secs
        sample = (f"same generator (seed 123), first {ps} of {cfg['p']} columns, n={cfg['n']}, {cfg['nlambda']} lambdas; "
                  f"Gram by numpy/OpenBLAS on {cores} threads ({tg / args.steps:.2f} s/step) + C port of the reference's "
                  f"single-threaded CD loop incl. its unconditional O(p) axpy per visit ({tc / args.steps:.2f} s/step). "
                  f"The reference moves 8p bytes per visit, so its rate at the full p={cfg['p']} is ~{cfg['p'] // ps}x lower.")
        out = dict(base, impl="reference", value=val, ms_per_step=1e3 * secs / args.steps,
                   config={"workload": workload, "sample_p": ps, "optTol": cfg["optTol"], "randomize": False},
                   cpu_baseline={"value": val, "unit": "visits/s", "cores": cores, "kind": "port", "sample": sample},
                   e2e={"value": val, "unit": "visits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(out))
        return

    # ------------------------------------------------------------------ our arm
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
    # pinned host copies for the e2e leg, resident device copies for the value leg
    Xp = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory()  # (p, n) C-order == (n, p) F-order
    yp = torch.from_numpy(y).pin_memory()
    Xd, yd = Xp.cuda(non_blocking=False), yp.cuda(non_blocking=False)
    torch.cuda.synchronize()
    # the box's pinned H2D rate, for reading the e2e number (outside the timed region)
    hb0, hb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hb0.record()
    Xd.copy_(Xp, non_blocking=True)
    hb1.record()
    torch.cuda.synchronize()
    h2d_gbs = 8 * n * p / (hb0.elapsed_time(hb1) * 1e-3) / 1e9

    def handle_dev():
        f = cdgpu.CDQuadraticLoss.__new__(cdgpu.CDQuadraticLoss)
        cdgpu.api._Loss.__init__(f, lib)
        f.p = p
        mk = lib.gram_create_lazy_dev if os.environ.get("CDGPU_BENCH_LAZY") else lib.gram_create_dev
        lib.check(mk(C.byref(f._h), C.c_void_p(Xd.data_ptr()), n, p, n, C.c_void_p(yd.data_ptr()), local))
        return f

    def handle_host():
        f = cdgpu.CDQuadraticLoss.__new__(cdgpu.CDQuadraticLoss)
        cdgpu.api._Loss.__init__(f, lib)
        f.p = p
        mk = lib.gram_create_lazy if os.environ.get("CDGPU_BENCH_LAZY") else lib.gram_create
        lib.check(mk(C.byref(f._h), C.c_void_p(Xp.data_ptr()), n, p, n, C.c_void_p(yp.data_ptr()), local))
        return f

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches():
        c = C.c_int64()
        lib.launch_count(C.byref(c))
        return c.value

    def timed(make_handle, steps, warmup):
        for _ in range(warmup):
            run_step(lib, be, make_handle, cfg, opts)
        barrier()
        l0 = launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        res = [run_step(lib, be, make_handle, cfg, opts) for _ in range(steps)]
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        dev = e0.elapsed_time(e1) * 1e-3
        secs = max(wall, dev)
        if world > 1:
            t = torch.tensor([secs], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        return res, secs, launches() - l0

    sampler = ClockSampler(local)
    if not os.environ.get("CDGPU_NO_SAMPLER"):
        sampler.start()
    else:
        sampler.nv, sampler.err = None, "disabled" 
    res, secs, nlaunch = timed(handle_dev, args.steps, args.warmup)
    clocks = sampler.stop()
    res_e, secs_e, _ = timed(handle_host, max(1, min(args.steps, 3)), 1)

    visits = sum(r["visits"] for r in res)
    visits_e = sum(r["visits"] for r in res_e)
    if world > 1:
        t = torch.tensor([visits, visits_e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        visits, visits_e = float(t[0].item()), float(t[1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    r0 = res[-1]
    gram_ms = float(np.mean([r["gram_ms"] for r in res]))
    cd_ms = float(np.mean([r["cd_ms"] for r in res]))
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
    gram_flops = n * p * (p + 1) + 2 * n * p  # SYRK (lower triangle) + X'y, algorithmic
    gram_tf = gram_flops / (gram_ms * 1e-3) / 1e12
    sweep_bytes = 24 * r0["visits"] + 8 * p * r0["accepted"]  # 24 B/visit + 8p B/accepted step
    out = dict(base, value=visits / secs, ms_per_step=1e3 * secs / args.steps,
               config={"workload": workload, "optTol": cfg["optTol"], "randomize": False, "replicas": world,
                       "l2": "inputs exceed L2: X %.1f GB, G %.1f GB per replica" % (8 * n * p / 1e9, 8 * p * p / 1e9)},
               clocks=clocks,
               e2e={"value": visits_e / secs_e, "unit": "visits/s", "ms_per_step": 1e3 * secs_e / len(res_e),
                    "h2d_bytes_per_step": 8 * n * p + 8 * n + 8 * p * 2 + 8 * cfg["nlambda"],
                    "h2d_gbs_pinned_measured": h2d_gbs,
                    "d2h_bytes_per_step": int(res_e[-1]["d2h"])},
               gpu_launches=int(nlaunch),
               roofline={"kernel": "gram_syrk_kernel (FP64 DMMA SYRK)", "bound": "tensor", "achieved": gram_tf,
                         "peak": dgemm_tf, "unit": "TFLOP/s", "frac": gram_tf / dgemm_tf,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch at the default size, from the
                         # ncu --set full capture in profiles/r1d_kernels_ncu.txt (X is 1.6 GB, G 3.2 GB)
                         "traffic": NCU_GRAM_DRAM_BYTES if (n, p) == (C2["n"], C2["p"]) else None,
                         "traffic_source": "profiles/r1d_kernels_ncu.txt (ncu --set full, one launch)",
                         "flops_per_launch": gram_flops, "launch_ms": gram_ms,
                         "peak_source": "cuBLAS Dgemm 8192^3 FP64 measured in this run (MEASURED_PEAKS.json has no FP64 entry)"},
               roofline_sweep={"kernel": "cov_path_kernel (cluster CD sweep, whole path)", "bound": "hbm",
                               "achieved": sweep_bytes / (cd_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                               "frac": sweep_bytes / (cd_ms * 1e-3) / 1e9 / hbm, "traffic": None,
                               "bytes_per_launch": sweep_bytes, "launch_ms": cd_ms, "peak_source": hbm_src,
                               "note": "latency-bound sequential chain: one cluster, 24 B/visit + 8p B/accepted step"},
               breakdown={"gram_ms": gram_ms, "cd_path_ms": cd_ms, "visits_per_step": r0["visits"],
                          "accepted_per_step": r0["accepted"], "passes": r0["passes"], "full_passes": r0["full_passes"],
                          "nnz_at_last_lambda": r0["nnz_last"], "all_converged": bool(r0["converged"]),
                          "cd_only_visits_per_sec": r0["visits"] / (cd_ms * 1e-3)})
    if args.gpus == 1 and not args.no_cpu_baseline:
        step, cores, ps = reference_arm(args, cfg)
        v, tg, tc = step()
        out["cpu_baseline"] = {"value": v / (tg + tc), "unit": "visits/s", "cores": cores, "kind": "port",
                               "sample": f"same generator, first {ps} of {p} columns, n={n}, {cfg['nlambda']} lambdas; "
                                         f"Gram numpy/OpenBLAS {cores} threads {tg:.2f} s + C port of the reference CD loop "
                                         f"(1 thread, literal always-axpy) {tc:.2f} s"}
    if args.gpus == 1 and not args.no_secondary:
        try:
            out["secondary"] = secondary_metrics(be, local, hbm)
        except Exception as e:  # noqa: BLE001  (side measurements must never cost the headline line)
            out["secondary"] = {"error": repr(e)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
