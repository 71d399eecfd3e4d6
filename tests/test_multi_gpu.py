"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box): run with
   gpurun --gpus 2 -- python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
       --master-port 29517 tests/test_multi_gpu.py
The row-sharded Gram (DMMA SYRK per rank + ncclAllReduce inside libcdgpu) must equal the single-GPU
Gram up to summation order, be exactly symmetric, and drive the same path on every rank."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _main():
    import torch
    import torch.distributed as dist

    import cdgpu
    from cdgpu import CDOptions, GaussianKernel
    from cdgpu.distributed import Comm, gram_sharded, locpolyl1_sharded, shard_range
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = cdgpu.Backend(cdgpu.load_product(), device=local)
    rng = np.random.default_rng(7)
    n, p, s = 4096 + 37, 700, 12
    X = rng.standard_normal((p, n)).T  # F-order (n, p)
    y = X[:, :s] @ rng.standard_normal(s) + rng.standard_normal(n)
    lo, hi = shard_range(n, rank, world)
    Xl = torch.from_numpy(np.ascontiguousarray(X[lo:hi].T)).cuda()  # (p, n_local) C-order == F-order shard
    yl = torch.from_numpy(np.ascontiguousarray(y[lo:hi])).cuda()
    comm = Comm(be)
    f = gram_sharded(be, comm, Xl.data_ptr(), hi - lo, n, p, hi - lo, yl.data_ptr())
    A, b = f.get()
    assert np.array_equal(A, A.T)
    assert np.allclose(A, X.T @ X / n, rtol=1e-12, atol=1e-13) and np.allclose(b, -X.T @ y / n, rtol=1e-12, atol=1e-13)
    f1 = be.CDQuadraticLoss_from_data(np.asfortranarray(X), y)
    om = f1.stdX()
    lams = np.exp(np.linspace(np.log(be.findLambdaMax(f1, om)), np.log(0.1 * be.findLambdaMax(f1, om)), 20))
    o = CDOptions(randomize=False, optTol=1e-12, maxIter=20000)
    pa = be.LassoPath(None, None, lams, o, standardizeX=om, loss=f)
    pb = be.LassoPath(None, None, lams, o, standardizeX=om, loss=f1)
    for xa, xb in zip(pa.βpath, pb.βpath):
        a_, b_ = xa.toarray(), xb.toarray()
        assert np.array_equal(a_ != 0, b_ != 0) and np.max(np.abs(a_ - b_)) <= 1e-6 * max(np.max(np.abs(b_)), 1e-300)
    # grid-point sharding with the NCCL gather
    Z = rng.random(300)
    Xs = np.asfortranarray(rng.standard_normal((300, 8)))
    Y = np.sin(4 * Z) * Xs[:, 0] + 0.1 * rng.standard_normal(300)
    zgrid = np.linspace(0.1, 0.9, 33)
    full = locpolyl1_sharded(be, Xs, Z, Y, zgrid, 1, GaussianKernel(0.2), 0.02, o)
    whole, _ = be.locpolyl1(Xs, Z, Y, zgrid, 1, GaussianKernel(0.2), 0.02, False, o)
    assert np.array_equal(full, whole)
    # the warm-start chain in runs of two grid points (cdgpu_vc_solve_chain): the runs are the units dealt over the ranks,
    # so the sharded result is the one-GPU result bit for bit even at a loose tolerance (33 grid points: the last run has one)
    ol = CDOptions(randomize=False, optTol=1e-4, maxIter=20000)
    full2 = locpolyl1_sharded(be, Xs, Z, Y, zgrid, 1, GaussianKernel(0.2), 0.02, ol, chain=2)
    whole2, _ = be.locpolyl1(Xs, Z, Y, zgrid, 1, GaussianKernel(0.2), 0.02, False, ol, chain=2)
    assert np.array_equal(full2, whole2)
    comm.close()
    dist.barrier()
    if rank == 0:
        print(f"multi-gpu ok: world={world} sharded gram {f.gram_ms:.2f} ms vs single {f1.gram_ms:.2f} ms")
    dist.destroy_process_group()


@pytest.mark.gpu
def test_multi_gpu_needs_torchrun():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs; run under torchrun (see module docstring)")
    import subprocess
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", __file__], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


if __name__ == "__main__":
    _main()
