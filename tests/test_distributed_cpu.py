"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: shard ranges and the gather of
independently solved grid-point blocks.  The compute behind it here is the CPU oracle; the same
code path runs libcdgpu on a GPU box (tests/test_multi_gpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from cdgpu.distributed import shard_range
    for m in (0, 1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(m, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == m
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))
    import torch.distributed as dist

    import cdgpu
    from cdgpu import CDOptions, GaussianKernel
    from cdgpu.distributed import locpolyl1_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = cdgpu.Backend(cdgpu.Lib(os.path.join(ROOT, "oracle", "libcdref.so"), "cdref"))
    rng = np.random.default_rng(3)
    n, p, degree = 150, 5, 1
    X = np.asfortranarray(rng.standard_normal((n, p)))
    Z = rng.random(n)
    Y = np.sin(4 * Z) * X[:, 0] + 0.1 * rng.standard_normal(n)
    zgrid = np.linspace(0.1, 0.9, 9)
    o = CDOptions(randomize=False, optTol=1e-12, maxIter=20000)
    full = locpolyl1_sharded(ref, X, Z, Y, zgrid, degree, GaussianKernel(0.2), 0.02, o)
    whole, _ = ref.locpolyl1(X, Z, Y, zgrid, degree, GaussianKernel(0.2), 0.02, False, o)
    # both sharding modes of the grid, and the leave-one-out problems of lvocv_locpolyl1 (one all_reduce of the sums)
    blocks = locpolyl1_sharded(ref, X, Z, Y, zgrid, degree, GaussianKernel(0.2), 0.02, o, interleave=False)
    from cdgpu.distributed import lvocv_locpolyl1_sharded
    hs = np.array([0.1, 0.3])
    oc = CDOptions(randomize=False, warmStart=False, optTol=1e-10, maxIter=20000)
    mse_sharded = lvocv_locpolyl1_sharded(ref, X[:60], Z[:60], Y[:60], degree, hs, GaussianKernel, 0.3, oc)
    mse_whole = ref.lvocv_locpolyl1(X[:60], Z[:60], Y[:60], degree, hs, GaussianKernel, 0.3, oc)
    # runs of two warm-started grid points (cdgpu_vc_solve_chain) dealt over the ranks: the same runs as in one process,
    # so the result is identical even at a loose tolerance (9 grid points: the last run has one)
    ol = CDOptions(randomize=False, optTol=1e-4, maxIter=20000)
    chained = locpolyl1_sharded(ref, X, Z, Y, zgrid, degree, GaussianKernel(0.2), 0.02, ol, chain=2)
    chained_whole, _ = ref.locpolyl1(X, Z, Y, zgrid, degree, GaussianKernel(0.2), 0.02, False, ol, chain=2)
    assert np.array_equal(chained, chained_whole)
    err = max(float(np.max(np.abs(full - whole))), float(np.max(np.abs(blocks - whole))),
              float(np.max(np.abs(mse_sharded - mse_whole) / mse_whole)))
    q.put((rank, err, int(np.count_nonzero(whole))))
    dist.destroy_process_group()


def test_locpolyl1_sharded_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    for _, err, nnz in res:
        assert err < 1e-9 and nnz > 5  # shards (each started from zero) == the chained single-process run
