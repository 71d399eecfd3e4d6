#!/usr/bin/env python
"""C5 (BASELINE.json configs[4]): tall lasso, n >> p = 2000 — row-sharded FP64 Gram X'X (DMMA SYRK per rank)
+ ONE ncclAllReduce over NVLink, then the covariance-form CD path (replicated on every rank).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
        benchmarks/c5_tall_gram.py [--n-total 10000000] [--p 2000] [--nlambda 100]

X never exists on the host: every rank draws its own n_total/N rows on the device (torch.randn with a
per-rank seed, generated in column-major chunks).  Strong scaling: total work is fixed as N grows.
Prints one JSON line from rank 0.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))
import cdgpu  # noqa: E402
from cdgpu import CDOptions  # noqa: E402
from cdgpu.distributed import Comm, gram_sharded, shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-total", type=int, default=10_000_000)
    ap.add_argument("--p", type=int, default=2000)
    ap.add_argument("--s", type=int, default=20)
    ap.add_argument("--nlambda", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = cdgpu.Backend(cdgpu.load_product(), device=local)
    n, p = args.n_total, args.p
    lo, hi = shard_range(n, rank, world)
    nl = hi - lo
    g = torch.Generator(device="cuda")
    g.manual_seed(1000 + rank)
    Xl = torch.empty((p, nl), device="cuda", dtype=torch.float64)  # (p, n_local) C-order == (n_local, p) F-order
    for j0 in range(0, p, 100):
        Xl[j0:j0 + 100].normal_(generator=g)
    g2 = torch.Generator(device="cuda")
    g2.manual_seed(7)
    beta = torch.randn(args.s, device="cuda", dtype=torch.float64, generator=g2)
    yl = Xl[: args.s].T @ beta + torch.randn(nl, device="cuda", dtype=torch.float64, generator=g)
    torch.cuda.synchronize()
    comm = Comm(be) if world > 1 else None
    opts = CDOptions(randomize=False)
    res = []
    for rep in range(args.reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if world > 1:
            f = gram_sharded(be, comm, Xl.data_ptr(), nl, n, p, nl, yl.data_ptr())
        else:
            f = cdgpu.CDQuadraticLoss.__new__(cdgpu.CDQuadraticLoss)
            cdgpu.api._Loss.__init__(f, be.lib)
            f.p = p
            be.lib.check(be.lib.gram_create_dev(C.byref(f._h), C.c_void_p(Xl.data_ptr()), nl, p, nl, C.c_void_p(yl.data_ptr()), local))
        t1 = time.perf_counter()
        om = f.stdX()
        lmax = be.findLambdaMax(f, om)
        lams = np.exp(np.linspace(np.log(lmax), np.log(0.05 * lmax), args.nlambda))
        path = be.LassoPath(None, None, lams, opts, standardizeX=om, loss=f)
        t2 = time.perf_counter()
        gms = f.gram_ms
        f.close()
        if rep:  # first repetition is the warm-up (NCCL channel setup, pool growth)
            res.append((t1 - t0, t2 - t1, gms, path.stats[0]["device_ms"], sum(s["visits"] for s in path.stats), path.βpath[-1].nnz))
    tg = torch.tensor([min(r[0] for r in res), min(r[1] for r in res), min(r[2] for r in res)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    if rank == 0:
        gram_wall, path_wall, gram_dev = (float(v) for v in tg.tolist())
        flops = n * p * (p + 1) + 2 * n * p
        print(json.dumps({"config": f"C5 tall lasso n={n} p={p}: row-sharded DMMA Gram + ncclAllReduce, then {args.nlambda}-lambda cov path",
                          "n_gpus": world, "scaling": "strong", "rows_per_gpu": nl,
                          "gram_device_ms(max over ranks, incl. allreduce)": gram_dev, "gram_wall_ms": 1e3 * gram_wall,
                          "gram_TFLOPs_aggregate(n p (p+1) flops)": flops / (gram_dev * 1e-3) / 1e12,
                          "allreduce_payload_MB": 8 * (p * p + p) / 1e6, "path_wall_ms": 1e3 * path_wall,
                          "path_device_ms": res[-1][3], "visits": res[-1][4], "nnz_last": res[-1][5],
                          "total_wall_ms": 1e3 * (gram_wall + path_wall)}))
    if comm:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
