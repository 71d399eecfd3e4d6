"""Host-side mirror of the CoordinateDescent.jl interface for the CD hot path.

Julia is not available where this is built, so the host side above the C ABI is
Python with the reference's names and argument meaning (`!` becomes a trailing
underscore).  Every call is one C-ABI call into libcdgpu.so; there is no Python
or CPU implementation of the algorithm here.

    reference (Julia)                                   here
    --------------------------------------------------  ---------------------------
    CDOptions(;...)                 utils.jl:7-20       CDOptions(...)
    IterLassoOptions(;...)          utils.jl:24-39      IterLassoOptions(...)
    SparseIterate(p)                ProximalBase        SparseIterate(p)
    ProxL1(l0[, l])                 ProximalBase        ProxL1(l0, l=None)
    CDLeastSquaresLoss(y, X)        cd_diff...:52       be.CDLeastSquaresLoss(y, X)
    CDWeightedLSLoss(y, X, w)       cd_diff...:128      be.CDWeightedLSLoss(y, X, w)
    CDSqrtLassoLoss(y, X)           cd_diff...:211      be.CDSqrtLassoLoss(y, X)
    CDQuadraticLoss(A, b)           cd_diff...:305      be.CDQuadraticLoss(A, b)
    coordinateDescent!(x,f,g,opt)   coordinate_descent.jl:7   be.coordinateDescent_(x,f,g,opt)
    lasso / sqrtLasso / scaledLasso! / LassoPath        lasso.jl:26-260
    locpolyl1 / GaussianKernel / EpanechnikovKernel     varying_coefficient_lasso.jl:3-79

`be` is a `Backend` bound to a loaded library; `default()` is the B200 one.
Python indices are 0-based at this surface; the 1-based SparseIterate triple is
what crosses the ABI.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _ffi
from ._ffi import ArgumentError, CdgpuError, DimensionMismatch, Lib, f64, ptr


# ----------------------------------------------------------------- options --
@dataclass
class CDOptions:
    maxIter: int = 2000
    optTol: float = 1e-7
    randomize: bool = True
    warmStart: bool = True
    numSteps: int = 50
    seed: int = 0  # stands in for Julia's global RNG

    def c(self) -> _ffi.Options:
        return _ffi.Options(int(self.maxIter), float(self.optTol), int(self.randomize), int(bool(self.warmStart)),
                            int(self.numSteps), int(self.seed))


_INIT = {"Screening": _ffi.INIT_SCREENING, "InitStd": _ffi.INIT_STD, "WarmStart": _ffi.INIT_WARMSTART}


@dataclass
class IterLassoOptions:
    maxIter: int = 20
    optTol: float = 1e-2
    initProcedure: str = "Screening"  # :Screening, :InitStd, :WarmStart
    sinit: int = 5
    σinit: float = 1.0
    optionsCD: CDOptions = field(default_factory=CDOptions)

    def c(self) -> _ffi.IterOptions:
        if self.initProcedure not in _INIT:
            raise ArgumentError("Incorrect initialization Symbol")  # lasso.jl:128
        return _ffi.IterOptions(int(self.maxIter), float(self.optTol), _INIT[self.initProcedure], 0, int(self.sinit),
                                float(self.σinit), self.optionsCD.c())


# ----------------------------------------------------------------- iterate --
class SparseIterate:
    """ProximalBase.SparseIterate{Float64}: the (nzval, nzval2ind, nnz) triple.

    `nzval2ind` holds 1-BASED coordinates (it is the array that crosses the ABI);
    `x[k]` and `nonzero()` are 0-based.
    """

    def __init__(self, p_or_vec, _triple=None):
        if _triple is not None:  # (nzval, nzval2ind) of exactly nnz entries: a read-only snapshot (path column)
            self.p = int(p_or_vec)
            self.nzval, self.nzval2ind = _triple
            self._nnz = C.c_int64(len(self.nzval))
        elif np.isscalar(p_or_vec):
            self.p = int(p_or_vec)
            self.nzval = np.zeros(self.p)
            self.nzval2ind = np.zeros(self.p, dtype=np.int64)
            self._nnz = C.c_int64(0)
        else:  # SparseIterate(sprand(p, 0.6)) — stores the non-zeros in index order
            v = np.asarray(p_or_vec, dtype=np.float64).ravel()
            self.p = v.size
            self.nzval = np.zeros(self.p)
            self.nzval2ind = np.zeros(self.p, dtype=np.int64)
            nz = np.flatnonzero(v)
            self.nzval[: nz.size] = v[nz]
            self.nzval2ind[: nz.size] = nz + 1
            self._nnz = C.c_int64(nz.size)

    @property
    def nnz(self) -> int:
        return int(self._nnz.value)

    def __len__(self):
        return self.p

    def __getitem__(self, k: int) -> float:
        hit = np.flatnonzero(self.nzval2ind[: self.nnz] == k + 1)
        return float(self.nzval[hit[0]]) if hit.size else 0.0

    def _grow(self):
        if self.nzval.size < self.p:  # snapshot made by LassoPath: give it ProximalBase's dense capacity
            n = self.nnz
            nz, ind = np.zeros(self.p), np.zeros(self.p, dtype=np.int64)
            nz[:n], ind[:n] = self.nzval[:n], self.nzval2ind[:n]
            self.nzval, self.nzval2ind = nz, ind

    def __setitem__(self, k: int, v: float):
        if not 0 <= k < self.p:
            raise IndexError(k)
        self._grow()
        hit = np.flatnonzero(self.nzval2ind[: self.nnz] == k + 1)
        if hit.size:
            self.nzval[hit[0]] = v
        elif v != 0.0:
            n = self.nnz
            self.nzval[n], self.nzval2ind[n] = v, k + 1
            self._nnz.value = n + 1

    def nonzero(self) -> np.ndarray:
        n = self.nnz
        return np.sort(self.nzval2ind[:n][self.nzval[:n] != 0.0] - 1)

    def toarray(self) -> np.ndarray:
        out = np.zeros(self.p)
        n = self.nnz
        out[self.nzval2ind[:n] - 1] = self.nzval[:n]
        return out

    def copy(self) -> "SparseIterate":
        o = SparseIterate(self.p)
        n = self.nnz
        o.nzval[:n], o.nzval2ind[:n], o._nnz.value = self.nzval[:n], self.nzval2ind[:n], n
        return o

    def __eq__(self, other):
        return isinstance(other, SparseIterate) and self.p == other.p and np.array_equal(self.toarray(),
                                                                                         other.toarray())


@dataclass
class ProxL1:
    """ProximalBase.ProxL1(λ0) / ProxL1(λ0, λ): penalty λ0·Σ λ_k |x_k|."""

    λ0: float
    λ: Optional[np.ndarray] = None

    def __post_init__(self):
        if self.λ is not None:
            self.λ = np.ascontiguousarray(self.λ, dtype=np.float64)


@dataclass
class LassoSolution:  # lasso.jl:7-17
    x: SparseIterate
    residuals: np.ndarray
    penalty: ProxL1
    σ: Optional[float]


@dataclass
class LassoPath:  # lasso.jl:203-206
    λpath: np.ndarray
    βpath: list
    stats: list = field(default_factory=list)


@dataclass
class GaussianKernel:  # varying_coefficient_lasso.jl:5-7,17
    h: float
    kind = _ffi.KERNEL_GAUSSIAN


@dataclass
class EpanechnikovKernel:  # varying_coefficient_lasso.jl:9-11,18-21
    h: float
    kind = _ffi.KERNEL_EPANECHNIKOV


def _std(r: np.ndarray) -> float:
    """Statistics.std (corrected)."""
    return float(np.std(r, ddof=1))


# ------------------------------------------------------------------ losses --
class _Loss:
    """A CoordinateDifferentiableFunction living behind a library handle."""

    kind = -1

    def __init__(self, lib: Lib):
        self.lib = lib
        self._h = C.c_void_p()
        self._keep = []  # host arrays the handle may alias (oracle) — keep alive
        self.last_stats = None

    def __del__(self):
        try:
            if self._h:
                self.lib.destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    close = __del__

    @property
    def numCoordinates(self) -> int:
        p = C.c_int64()
        self.lib.check(self.lib.dims(self._h, None, C.byref(p), None))
        return p.value

    def _state(self, n) -> np.ndarray:
        out = np.empty(n)
        self.lib.check(self.lib.state(self._h, ptr(out)))
        return out


class _NaiveLoss(_Loss):
    def __init__(self, lib, y, X, w=None, device=0):
        super().__init__(lib)
        X, y = f64(X), f64(y)
        if X.ndim != 2 or y.ndim != 1:
            raise DimensionMismatch("X must be a matrix and y a vector")
        if y.shape[0] != X.shape[0] or (w is not None and np.shape(w) != y.shape):
            raise DimensionMismatch()  # cd_differentiable_function.jl:53,129,212
        w = None if w is None else f64(w)
        self.n, self.p = X.shape
        self._keep = [X, y, w]
        lib.check(lib.naive_create(C.byref(self._h), self.kind, ptr(X), self.n, self.p, self.n, ptr(y), ptr(w),
                                   device))

    @property
    def r(self) -> np.ndarray:
        return self._state(self.n)

    def stdX(self, w=None) -> np.ndarray:
        out = np.empty(self.p)
        w = None if w is None else f64(w)
        self.lib.check(self.lib.stdx(self._h, ptr(w), ptr(out)))
        return out


class CDLeastSquaresLoss(_NaiveLoss):
    kind = _ffi.LOSS_LS


class CDWeightedLSLoss(_NaiveLoss):
    kind = _ffi.LOSS_WLS

    def __init__(self, lib, y, X, w, device=0):
        super().__init__(lib, y, X, w, device)


class CDSqrtLassoLoss(_NaiveLoss):
    kind = _ffi.LOSS_SQRT


class CDQuadraticLoss(_Loss):
    kind = _ffi.LOSS_QUAD

    def __init__(self, lib, A, b, device=0):
        super().__init__(lib)
        A, b = f64(A), f64(b)
        if A.ndim != 2 or A.shape[0] != A.shape[1] or b.ndim != 1 or b.shape[0] != A.shape[1]:
            raise ArgumentError("A must be square and length(b) == size(A, 2)")  # :306
        self.p = A.shape[0]
        self._keep = [A, b]
        lib.check(lib.quad_create(C.byref(self._h), ptr(A), self.p, self.p, ptr(b), device))

    @classmethod
    def from_data(cls, lib, X, y, device=0, lazy=False):
        """A = X'X/n, b = -X'y/n formed by the library (FP64 tensor-core SYRK).  lazy=True: only diag(A) and b up
        front, columns of A on demand (cdgpu_gram_create_lazy)."""
        self = cls.__new__(cls)
        _Loss.__init__(self, lib)
        X, y = f64(X), f64(y)
        if X.ndim != 2 or y.shape != (X.shape[0],):
            raise DimensionMismatch()
        self.p = X.shape[1]
        create = lib.gram_create_lazy if lazy else lib.gram_create
        lib.check(create(C.byref(self._h), ptr(X), X.shape[0], self.p, X.shape[0], ptr(y), device))
        return self

    def sweep_stats(self):
        """device ms inside the sweep kernel during the last solve / path, plus lazy_stats()."""
        ms = C.c_double()
        self.lib.check(self.lib.sweep_ms(self._h, C.byref(ms)))
        return dict(self.lazy_stats(), sweep_ms=ms.value)

    def lazy_stats(self):
        cols, nb, npause, ms = C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
        self.lib.check(self.lib.lazy_stats(self._h, C.byref(cols), C.byref(nb), C.byref(npause), C.byref(ms)))
        fc = C.c_int64()
        self.lib.check(self.lib.lazy_form_columns(self._h, C.byref(fc)))
        return {"columns": cols.value, "batches": nb.value, "pauses": npause.value, "form_ms": ms.value,
                "form_columns": fc.value}

    def stdX(self) -> np.ndarray:
        """sqrt(diag(A)) == _stdX!(X) when A = X'X/n (utils.jl:127-138)."""
        out = np.empty(self.p)
        self.lib.check(self.lib.stdx(self._h, None, ptr(out)))
        return out

    @property
    def Ax(self) -> np.ndarray:
        return self._state(self.p)

    @property
    def gram_ms(self) -> float:
        ms = C.c_double()
        self.lib.check(self.lib.gram_ms(self._h, C.byref(ms)))
        return ms.value

    def get(self):
        A, b = np.empty((self.p, self.p), order="F"), np.empty(self.p)
        self.lib.check(self.lib.quad_get(self._h, ptr(A), ptr(b)))
        return A, b


# ----------------------------------------------------------------- backend --
class Backend:
    """The reference's function set bound to one loaded library."""

    def __init__(self, lib: Lib, device: int = 0):
        self.lib, self.device = lib, device

    # constructors
    def CDLeastSquaresLoss(self, y, X):
        return CDLeastSquaresLoss(self.lib, y, X, device=self.device)

    def CDWeightedLSLoss(self, y, X, w):
        return CDWeightedLSLoss(self.lib, y, X, w, device=self.device)

    def CDSqrtLassoLoss(self, y, X):
        return CDSqrtLassoLoss(self.lib, y, X, device=self.device)

    def CDQuadraticLoss(self, A, b):
        return CDQuadraticLoss(self.lib, A, b, device=self.device)

    def CDQuadraticLoss_from_data(self, X, y, lazy=False):
        return CDQuadraticLoss.from_data(self.lib, X, y, device=self.device, lazy=lazy)

    # coordinateDescent!(x, f, g, options)          coordinate_descent.jl:7-39
    def coordinateDescent_(self, x: SparseIterate, f: _Loss, g: ProxL1, options: CDOptions = None):
        options = options or CDOptions()
        p = f.numCoordinates
        if x.p != p:
            raise DimensionMismatch()  # :13
        if g.λ is not None and g.λ.shape != (p,):
            raise DimensionMismatch()  # :15
        o, st = options.c(), _ffi.Stats()
        x._grow()
        self.lib.check(self.lib.solve(f._h, float(g.λ0), ptr(g.λ), C.byref(o), ptr(x.nzval), ptr(x.nzval2ind),
                                      C.byref(x._nnz), C.byref(st)))
        f.last_stats = st.as_dict()
        return x

    # lasso(X, y, λ[, ω], options)                   lasso.jl:26-53
    def lasso(self, X, y, λ, ω=None, options: CDOptions = None):
        if isinstance(ω, CDOptions):
            ω, options = None, ω
        X = f64(X)
        x = SparseIterate(X.shape[1])
        f = self.CDLeastSquaresLoss(y, X)
        g = ProxL1(float(λ), ω)
        self.coordinateDescent_(x, f, g, options)
        r = f.r
        sol = LassoSolution(x, r, g, _std(r))
        sol.stats = f.last_stats
        f.close()
        return sol

    # sqrtLasso(X, y, λ[, ω], options; standardizeX)  lasso.jl:62-98
    def sqrtLasso(self, X, y, λ, ω=None, options: CDOptions = None, standardizeX=True):
        if isinstance(ω, CDOptions):
            ω, options = None, ω
        X = f64(X)
        x = SparseIterate(X.shape[1])
        f = self.CDSqrtLassoLoss(y, X)
        if ω is None and standardizeX:
            # lasso.jl:72-75 is dead code on Julia >= 1.0 (Array{T}(p)); its evident
            # intent is ω = _stdX!(X)
            ω = f.stdX()
        g = ProxL1(float(λ), ω)
        self.coordinateDescent_(x, f, g, options)
        r = f.r
        sol = LassoSolution(x, r, g, _std(r))
        sol.stats = f.last_stats
        f.close()
        return sol

    # scaledLasso!(x, X, y, λ, ω, options)           lasso.jl:107-144
    def scaledLasso_(self, x: SparseIterate, X, y, λ, ω, options: IterLassoOptions = None):
        options = options or IterLassoOptions()
        X = f64(X)
        ω = f64(ω)
        if x.p != X.shape[1] or ω.shape != (X.shape[1],):
            raise DimensionMismatch()
        o = options.c()
        x._grow()
        f = self.CDLeastSquaresLoss(y, X)
        st, sig = _ffi.Stats(), C.c_double()
        self.lib.check(self.lib.scaled_solve(f._h, float(λ), ptr(ω), C.byref(o), ptr(x.nzval), ptr(x.nzval2ind),
                                             C.byref(x._nnz), C.byref(sig), C.byref(st)))
        sol = LassoSolution(x, f.r, ProxL1(float(λ) * st.sigma, ω), sig.value)
        sol.stats = st.as_dict()
        f.close()
        return sol

    # LassoPath(X, Y, λpath, options; max_hat_s, standardizeX)   lasso.jl:229-260
    def LassoPath(self, X, Y, λpath, options: CDOptions = None, max_hat_s=math.inf, standardizeX=True, loss=None):
        """`loss` (an existing handle, e.g. a covariance-form CDQuadraticLoss) replaces
        the reference's CDLeastSquaresLoss(Y, X); then X may be None and ω must be
        passed through `standardizeX` as an array."""
        options = options or CDOptions()
        own = loss is None
        f = self.CDLeastSquaresLoss(Y, f64(X)) if own else loss
        p = f.numCoordinates
        if isinstance(standardizeX, np.ndarray):
            stdX = f64(standardizeX)
        elif standardizeX:
            stdX = f.stdX()  # _stdX!(stdX, X)  :239
        else:
            stdX = np.ones(p)  # :241
        lam = np.ascontiguousarray(λpath, dtype=np.float64)
        m = lam.size
        cap = int(min(m * p, max(1 << 22, 4 * p)))
        colptr, rowval = np.zeros(m + 1, dtype=np.int64), np.empty(cap, dtype=np.int64)
        nzval, done = np.empty(cap), C.c_int64()
        stats = (_ffi.Stats * m)()
        mhs = -1 if math.isinf(max_hat_s) else int(max_hat_s)
        o = options.c()
        self.lib.check(self.lib.path(f._h, ptr(lam), m, ptr(stdX), C.byref(o), mhs, cap, ptr(colptr), ptr(rowval),
                                     ptr(nzval), C.byref(done), C.cast(stats, C.c_void_p)))
        nd = done.value
        tot = int(colptr[nd])
        vals, inds, cp = nzval[:tot].copy(), rowval[:tot].copy(), colptr[: nd + 1].tolist()  # columns are views of one copy
        βpath = [SparseIterate(p, _triple=(vals[cp[i]:cp[i + 1]], inds[cp[i]:cp[i + 1]])) for i in range(nd)]
        if own:
            f.close()
        return LassoPath(lam[: done.value].copy(), βpath, _ffi.stats_dicts(stats, done.value))

    # refitLassoPath(path, X, Y)                        lasso.jl:208-225
    def refitLassoPath(self, path: LassoPath, X, Y, loss=None):
        """Least-squares refit on each distinct support of the path (`X[:, S] \\ Y`), through the library
        (`cdgpu_refit`: normal equations on the support + Cholesky on the device).  Keys are 0-based index tuples
        in increasing order.  `loss`: an existing handle on (X, Y) to reuse."""
        own = loss is None
        f = self.CDLeastSquaresLoss(Y, f64(X)) if own else loss
        out = {}
        for β in path.βpath:
            S = tuple(sorted(int(k) for k in β.nonzero()))
            if S in out:
                continue
            coef = np.zeros(len(S))
            if S:
                idx = np.asarray(S, dtype=np.int64) + 1
                self.lib.check(self.lib.refit(f._h, ptr(idx), len(S), ptr(coef)))
            out[S] = coef
        if own:
            f.close()
        return out

    # locpolyl1(X, z, y, zgrid, degree, kernel, λ0, refit, options)  varying_coefficient_lasso.jl:30-79
    def locpolyl1(self, X, z, y, zgrid, degree, kernel, λ0, refit=False, options: CDOptions = None, shard=None, chain=None,
                  sparse=False):
        """`sparse=True`: the results come back as the reference's SparseMatrixCSC (`spzeros(ep, m)`, :46-47), compacted on
        the device (`cdgpu_vc_solve_csc`), as `(colptr, rowval, nzval)` triples with 0-based offsets and 1-based rows;
        the second triple is None without refit.  `chain`: grid points per warm-started run.  None: the library's default (device: every grid point from zero,
        all of them concurrently; CPU oracle: the reference's chain over the grid points it is given); 1: every grid
        point from zero on either library; k: runs of k consecutive grid points, each point from its predecessor's
        iterate; `len(zgrid)`: the reference's loop exactly (varying_coefficient_lasso.jl:56,68), as one sequential
        chain."""
        options = options or CDOptions()
        X, z, y, zgrid = f64(X), f64(z), f64(y), f64(zgrid)
        n, p = X.shape
        if z.shape != (n,) or y.shape != (n,):
            raise DimensionMismatch()
        m, ep = zgrid.size, p * (degree + 1)
        lo, hi = shard if shard is not None else (0, m)
        out = np.zeros((ep, m), order="F")
        stats = (_ffi.Stats * m)()
        o = options.c()
        if sparse:
            cap = max(1, min(ep * (hi - lo), 64 * (hi - lo)))
            while True:
                cp, rv, nz = np.zeros(m + 1, dtype=np.int64), np.zeros(cap, dtype=np.int64), np.zeros(cap)
                cpR, rvR, nzR = (np.zeros(m + 1, dtype=np.int64), np.zeros(cap, dtype=np.int64), np.zeros(cap)) if refit \
                    else (None, None, None)
                rc = self.lib.vc_solve_csc(ptr(X), n, p, n, ptr(z), ptr(y), ptr(zgrid), m, lo, hi, int(degree), kernel.kind,
                                           float(kernel.h), float(λ0), C.byref(o), 1 if chain is None else int(chain),
                                           self.device, cap, ptr(cp), ptr(rv), ptr(nz), ptr(cpR), ptr(rvR), ptr(nzR),
                                           C.cast(stats, C.c_void_p))
                need = max(int(cp[m]), int(cpR[m]) if refit else 0)
                if rc == _ffi.ECAP and need > cap:  # the library reports the entries it needs: one retry
                    cap = need
                    continue
                self.lib.check(rc)
                break
            self.last_vc_stats = [stats[i].as_dict() for i in range(lo, hi)]
            tri = (cp, rv[:cp[m]].copy(), nz[:cp[m]].copy())
            return tri, ((cpR, rvR[:cpR[m]].copy(), nzR[:cpR[m]].copy()) if refit else None)
        if chain is not None:
            outR = np.zeros((ep, m), order="F") if refit else None
            self.lib.check(self.lib.vc_solve_chain(ptr(X), n, p, n, ptr(z), ptr(y), ptr(zgrid), m, lo, hi, int(degree),
                                                   kernel.kind, float(kernel.h), float(λ0), C.byref(o), int(chain),
                                                   self.device, ptr(out), ptr(outR), C.cast(stats, C.c_void_p)))
            self.last_vc_stats = [stats[i].as_dict() for i in range(lo, hi)]
            return out, outR
        if not refit:
            self.lib.check(self.lib.vc_solve(ptr(X), n, p, n, ptr(z), ptr(y), ptr(zgrid), m, lo, hi, int(degree),
                                             kernel.kind, float(kernel.h), float(λ0), C.byref(o), self.device, ptr(out),
                                             C.cast(stats, C.c_void_p)))
            self.last_vc_stats = [stats[i].as_dict() for i in range(lo, hi)]
            return out, None
        # refit on the selected groups (varying_coefficient_lasso.jl:71-76, get_nonzero_coordinates! :488-512) inside the
        # same library call: the normal equations come from the moment blocks already on the device
        outR = np.zeros((ep, m), order="F")
        self.lib.check(self.lib.vc_solve_refit(ptr(X), n, p, n, ptr(z), ptr(y), ptr(zgrid), m, lo, hi, int(degree),
                                               kernel.kind, float(kernel.h), float(λ0), C.byref(o), self.device, ptr(out),
                                               ptr(outR), C.cast(stats, C.c_void_p)))
        self.last_vc_stats = [stats[i].as_dict() for i in range(lo, hi)]
        return out, outR

    # lvocv_locpolyl1(X, z, y, degree, hArr, kernelType, λ0, options)   varying_coefficient_lasso.jl:81-137
    def lvocv_locpolyl1(self, X, z, y, degree, hArr, kernelType, λ0, options: CDOptions = None, shard=None):
        """Leave-one-out MSE per bandwidth.  `kernelType` is GaussianKernel or EpanechnikovKernel (the class, as in the
        reference's `createKernel(kernelType, h)`).  All numH*n local problems run as one batch; `shard=(q0, q1)`
        restricts the call to a slice of the (bandwidth-major) problem list and returns the partial sums."""
        options = options or CDOptions()
        X, z, y, hArr = f64(X), f64(z), f64(y), f64(hArr)
        n, p = X.shape
        if z.shape != (n,) or y.shape != (n,):
            raise DimensionMismatch()
        numH = hArr.size
        m = numH * n
        lo, hi = shard if shard is not None else (0, m)
        sq = np.zeros(m)
        stats = (_ffi.Stats * m)()
        o = options.c()
        self.lib.check(self.lib.vc_lvocv(ptr(X), n, p, n, ptr(z), ptr(y), int(degree), ptr(hArr), numH, kernelType(1.0).kind,
                                         float(λ0), C.byref(o), lo, hi, self.device, ptr(sq), C.cast(stats, C.c_void_p)))
        self.last_vc_stats = [stats[i].as_dict() for i in range(lo, hi)]
        self.last_lvocv_sqerr = sq
        return sq.reshape(numH, n).sum(axis=1)  # MSE[indH] += (Yh - y[i])^2, :131

    def findLambdaMax(self, f: _Loss, ω=None) -> float:  # coordinate_descent.jl:118-149 at x = 0
        out = C.c_double()
        ω = None if ω is None else f64(ω)
        self.lib.check(self.lib.lambda_max(f._h, ptr(ω), C.byref(out)))
        return out.value


_default = None


def default() -> Backend:
    """Backend on libcdgpu.so; raises when the CUDA library is not built."""
    global _default
    if _default is None:
        _default = Backend(_ffi.load_product())
    return _default
