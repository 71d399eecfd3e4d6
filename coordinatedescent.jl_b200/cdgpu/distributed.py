"""One-process-per-GPU helpers (torch.distributed is plumbing only).

* independent units (grid points of locpolyl1, replicas, folds) are sharded with NO data-path
  collective; the only communication is the final gather of the results;
* the row-sharded Gram has the one real exchange step: ncclAllReduce inside libcdgpu
  (csrc/nccl_comm.cu); the 128-byte ncclUniqueId is broadcast here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .api import Backend, CDQuadraticLoss, _Loss
from ._ffi import f64, ptr


def shard_range(m: int, rank: int, world: int):
    """Contiguous block of [0, m) owned by `rank`: sizes differ by at most one."""
    base, extra = divmod(m, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def locpolyl1_sharded(be: Backend, X, z, y, zgrid, degree, kernel, λ0, options=None, group=None, interleave=True, chain=1):
    """locpolyl1 with the grid points split over the ranks of `group`; every rank returns the full
    ep x m matrix (all_gather of the owned columns).  No data-path collective: the local problems are
    independent.  `interleave=True` deals the grid points round-robin (rank r owns zgrid[r::world]):
    neighbouring grid points cost about the same number of passes, so the ranks finish together;
    `interleave=False` gives every rank one contiguous block ([m_begin, m_end) of the C ABI).  `chain` > 1: the units
    dealt are runs of `chain` consecutive grid points, each run warm-started point to point (cdgpu_vc_solve_chain), so
    the result does not depend on the number of ranks."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    zgrid = f64(zgrid)
    m = zgrid.size
    chain = max(1, int(chain))

    def owned(r):  # grid points of rank r when runs of `chain` points are dealt round-robin
        runs = np.arange(r, -(-m // chain), world)
        idx = (runs[:, None] * chain + np.arange(chain)[None, :]).ravel()
        return idx[idx < m]

    if interleave:
        mine_idx = owned(rank)
        out, _ = be.locpolyl1(X, z, y, np.ascontiguousarray(zgrid[mine_idx]), degree, kernel, λ0, False, options,
                              chain=chain if chain > 1 else None)  # (None: the library's default)
        cols = out
    else:
        if chain > 1:
            raise ValueError("chain > 1 needs interleave=True (runs are the units dealt)")
        lo, hi = shard_range(m, rank, world)
        out, _ = be.locpolyl1(X, z, y, zgrid, degree, kernel, λ0, False, options, shard=(lo, hi))
        cols = out[:, lo:hi]
    ep = out.shape[0]
    # gather variable-sized column blocks: pad to the largest block
    width = max(owned(r).size for r in range(world)) if interleave else -(-m // world)
    mine = torch.zeros(width * ep, dtype=torch.float64)
    mine[: cols.shape[1] * ep] = torch.from_numpy(np.ascontiguousarray(cols.T).ravel())
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = mine.to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    full = np.zeros((ep, m), order="F")
    for r, t in enumerate(parts):
        if interleave:
            idx = owned(r)
            full[:, idx] = t.cpu().numpy()[: idx.size * ep].reshape(idx.size, ep).T
        else:
            a, b = shard_range(m, r, world)
            full[:, a:b] = t.cpu().numpy()[: (b - a) * ep].reshape(b - a, ep).T
    return full


def lvocv_locpolyl1_sharded(be: Backend, X, z, y, degree, hArr, kernelType, λ0, options=None, group=None):
    """lvocv_locpolyl1 with the numH*n leave-one-out problems split over the ranks of `group` (contiguous slices of
    the bandwidth-major problem list; every problem is independent).  The only communication is one all_reduce of the
    numH partial sums; every rank returns the full MSE vector."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    hArr = f64(hArr)
    m = hArr.size * np.asarray(X).shape[0]
    lo, hi = shard_range(m, rank, world)
    part = be.lvocv_locpolyl1(X, z, y, degree, hArr, kernelType, λ0, options, shard=(lo, hi))
    t = torch.from_numpy(np.ascontiguousarray(part))
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, group=group)
    return t.cpu().numpy()


class Comm:
    """NCCL communicator owned by libcdgpu (cdgpu_comm_*), bootstrapped over torch.distributed."""

    def __init__(self, be: Backend, group=None):
        import torch
        import torch.distributed as dist

        self.lib = be.lib
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        idbuf = (C.c_ubyte * 128)()
        if rank == 0:
            self.lib.check(self.lib.comm_unique_id(C.cast(idbuf, C.c_void_p)))
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        idbuf = (C.c_ubyte * 128).from_buffer_copy(raw)
        self._c = C.c_void_p()
        self.lib.check(self.lib.comm_init(C.byref(self._c), C.cast(idbuf, C.c_void_p), rank, world, be.device))
        self.rank, self.world = rank, world

    def close(self):
        if self._c:
            self.lib.comm_destroy(self._c)
            self._c = C.c_void_p()

    __del__ = close


def gram_sharded(be: Backend, comm: Comm, dX_local_ptr: int, n_local: int, n_total: int, p: int, ldx: int,
                 dy_local_ptr: int) -> CDQuadraticLoss:
    """Covariance-form handle from row shards resident on each rank's GPU (device pointers)."""
    f = CDQuadraticLoss.__new__(CDQuadraticLoss)
    _Loss.__init__(f, be.lib)
    f.p = p
    be.lib.check(be.lib.gram_create_sharded(C.byref(f._h), C.c_void_p(dX_local_ptr), n_local, n_total, p, ldx,
                                            C.c_void_p(dy_local_ptr), comm._c, be.device))
    return f
