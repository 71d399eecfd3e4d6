"""Pins the oracle against the reference's known-answer tests."""
import ctypes as C
import math

import numpy as np
import pytest

import cdgpu
from cdgpu import CDOptions, ProxL1, SparseIterate


def test_small_proxl1_kat(ref):
    # test/coordinate_descent.jl:13-25: A = I2, b = -[1, 1.5], λ = 1.2  =>  x = [0, 0.3]
    Y = np.array([1.0, 1.5])
    f = ref.CDQuadraticLoss(np.eye(2), -Y)
    x = SparseIterate(2)
    ref.coordinateDescent_(x, f, ProxL1(1.2), CDOptions(maxIter=100, optTol=1e-8, warmStart=True, randomize=False))
    assert np.allclose(x.toarray(), [0.0, 0.3], rtol=1e-8, atol=0)


def test_sparse_iterate_insertion_order(ref):
    # test/atom_iterator.jl:13-28: x[2]=1; x[1]=2  => sparse pass visits [2, 1]
    lib = ref.lib.dll
    keys = np.array([2, 1], dtype=np.int64)
    vals = np.array([1.0, 2.0])
    nzval, ind, nnz = np.zeros(5), np.zeros(5, dtype=np.int64), C.c_int64()
    rc = lib.cdref_sparse_iterate_replay(C.c_int64(5), keys.ctypes, vals.ctypes, C.c_int64(2), 0, nzval.ctypes,
                                         ind.ctypes, C.byref(nnz))
    assert rc == 0 and nnz.value == 2 and list(ind[:2]) == [2, 1]
    visited = np.zeros(5, dtype=np.int64)
    lib.cdref_iterator_collect(C.c_int64(5), ind.ctypes, C.c_int64(2), 1, 0, C.c_uint64(0), C.c_uint64(0),
                               visited.ctypes, None)
    assert list(visited) == [1, 2, 3, 4, 5]  # full pass
    lib.cdref_iterator_collect(C.c_int64(5), ind.ctypes, C.c_int64(2), 0, 0, C.c_uint64(0), C.c_uint64(0),
                               visited.ctypes, None)
    assert list(visited[:2]) == [2, 1]  # pass over non-zeros


@pytest.mark.parametrize("mode", [1, 2])
def test_random_iterator_is_a_permutation(ref, mode):
    # test/atom_iterator.jl:50-67
    lib = ref.lib.dll
    p = 50
    rng = np.random.default_rng(1)
    nz = rng.permutation(p)[:10].astype(np.int64) + 1
    visited, order = np.zeros(p, dtype=np.int64), np.zeros(p, dtype=np.int64)
    lib.cdref_iterator_collect(C.c_int64(p), nz.ctypes, C.c_int64(10), 1, mode, C.c_uint64(7), C.c_uint64(3),
                               visited.ctypes, order.ctypes)
    assert sorted(visited) == list(range(1, p + 1)) and list(visited) == list(order)
    assert list(visited) != list(range(1, p + 1))
    lib.cdref_iterator_collect(C.c_int64(p), nz.ctypes, C.c_int64(10), 0, mode, C.c_uint64(7), C.c_uint64(3),
                               visited.ctypes, order.ctypes)
    assert list(visited[:10]) == [nz[order[i] - 1] for i in range(10)]
    assert sorted(order[:10]) == list(range(1, 11))


def test_dropzeros_and_explicit_zero(ref):
    lib = ref.lib.dll
    keys = np.array([3, 1, 4, 1, 5], dtype=np.int64)
    vals = np.array([1.0, 2.0, 3.0, 0.0, 0.0])  # x[1] becomes an explicit zero; x[5]=0 is not stored
    nzval, ind, nnz = np.zeros(6), np.zeros(6, dtype=np.int64), C.c_int64()
    lib.cdref_sparse_iterate_replay(C.c_int64(6), keys.ctypes, vals.ctypes, C.c_int64(5), 0, nzval.ctypes, ind.ctypes,
                                    C.byref(nnz))
    assert nnz.value == 3 and list(ind[:3]) == [3, 1, 4] and list(nzval[:3]) == [1.0, 0.0, 3.0]
    lib.cdref_sparse_iterate_replay(C.c_int64(6), keys.ctypes, vals.ctypes, C.c_int64(5), 1, nzval.ctypes, ind.ctypes,
                                    C.byref(nnz))
    assert nnz.value == 2 and list(ind[:2]) == [3, 4] and list(nzval[:2]) == [1.0, 3.0]


def test_kernels_kat(ref):
    # test/varying_coefficient_lasso.jl:16-21
    ev = ref.lib.dll.cdref_kernel_evaluate
    ev.restype = C.c_double
    ev.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    assert ev(0, 1.0, 0.3, 0.4) == pytest.approx(math.exp(-0.01), rel=1e-15)
    # src/varying_coefficient_lasso.jl:18-21
    assert ev(1, 0.5, 0.3, 0.4) == pytest.approx(0.75 * (1 - 0.04) / 0.5, rel=1e-15)
    assert ev(1, 0.1, 0.3, 0.4) == 0.0


def test_expand_X_kat(ref):
    # test/varying_coefficient_lasso.jl:43-66
    ex = ref.lib.dll.cdref_expand_X
    ex.restype = None
    X = np.asfortranarray(np.arange(1.0, 7.0).reshape(3, 2).T)  # reshape(collect(1.:6.), 2, 3)
    z = np.array([0.2, 0.4])
    z0 = 0.3
    for degree, Q in [(0, np.ones((2, 1))), (1, np.array([[1, -0.1], [1, 0.1]])),
                      (2, np.array([[1, -0.1, 0.01], [1, 0.1, 0.01]]))]:
        tX = np.zeros((2, 3 * (degree + 1)), order="F")
        ex(tX.ctypes, X.ctypes, C.c_int64(2), C.c_int64(3), C.c_int64(2), z.ctypes, C.c_double(z0), degree)
        want = np.stack([np.kron(X[i, :], Q[i, :]) for i in range(2)])
        assert np.allclose(tX, want, rtol=1e-13, atol=1e-15)


def test_shrink(ref):
    s = ref.lib.dll.cdref_shrink
    s.restype = C.c_double
    s.argtypes = [C.c_double, C.c_double]
    assert s(1.5, 1.2) == pytest.approx(0.3) and s(-1.5, 1.2) == pytest.approx(-0.3) and s(1.0, 1.2) == 0.0
