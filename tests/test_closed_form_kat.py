"""Oracle-INDEPENDENT known answers for every loss on the hot path (VERDICT r1, next #1c).

None of the expected values below comes from oracle/cdref.c or from the CUDA library: they are closed forms of
the reference's objectives (cited per test) or optimality certificates (KKT systems) evaluated with numpy.  The
same checks run against the CPU oracle (`-m "not gpu"`, pins the oracle) and against libcdgpu.so (`-m gpu`).

  CDLeastSquaresLoss   ||y - X b||^2/(2n) + l0 sum w_k|b_k|     cd_differentiable_function.jl:43-111
  CDWeightedLSLoss     sum_i w_i r_i^2/(2n) + l0 sum w_k|b_k|   :118-194
  CDSqrtLassoLoss      ||y - X b||_2 + l0 sum w_k|b_k|          :202-291 (threshold :277, closed form :277-283)
  CDQuadraticLoss      b'Ab/2 + c'b + l0 sum w_k|b_k|           :299-348
  scaledLasso!         fixed point sigma = sqrt(||r||^2/n)       lasso.jl:107-144
"""
import numpy as np
import pytest

from cdgpu import CDOptions, IterLassoOptions, ProxL1, SparseIterate

TIGHT = CDOptions(maxIter=200000, optTol=1e-14, randomize=False)


@pytest.fixture(params=["ref", pytest.param("gpu", marks=pytest.mark.gpu)])
def be(request):
    return request.getfixturevalue(request.param)


def soft(v, c):
    return np.sign(v) * np.maximum(np.abs(v) - c, 0.0)


def orthonormal_design(n, p, rng, w=None):
    """Columns orthogonal in the (w-weighted) inner product, with distinct non-unit norms."""
    M = rng.standard_normal((n, p))
    sw = np.ones(n) if w is None else np.sqrt(w)
    Q, _ = np.linalg.qr(M * sw[:, None])
    return np.asfortranarray(Q / sw[:, None] * rng.uniform(0.5, 3.0, p)[None, :] * np.sqrt(n))


def test_ls_orthogonal_design_is_soft_threshold(be):
    # X'X diagonal  =>  b_k = S(X_k'y / a_k, n*l0*w_k / a_k), a_k = ||X_k||^2   (:96-103 has no coupling left)
    rng = np.random.default_rng(11)
    n, p = 64, 12
    X = orthonormal_design(n, p, rng)
    y = X @ rng.standard_normal(p) / n + 0.3 * rng.standard_normal(n)
    om = rng.uniform(0.5, 2.0, p)
    lam = 0.02
    a = (X ** 2).sum(0)
    want = soft(X.T @ y / a, n * lam * om / a)
    assert 0 < np.count_nonzero(want) < p
    got = be.lasso(X, y, lam, om, TIGHT).x.toarray()
    assert np.array_equal(got != 0, want != 0)
    assert np.allclose(got, want, rtol=1e-10, atol=1e-13)
    got1 = be.lasso(X, y, lam, TIGHT).x.toarray()  # ProxL1{T,Nothing}
    assert np.allclose(got1, soft(X.T @ y / a, n * lam / a), rtol=1e-10, atol=1e-13)


def test_wls_orthogonal_design_is_soft_threshold(be):
    # sum_i w_i X_ij X_ik = 0 (j != k)  =>  b_k = S(sum w X_k y / a_k, n*l0*w_k / a_k), a_k = sum w X_k^2  (:179-186)
    rng = np.random.default_rng(12)
    n, p = 80, 9
    w = rng.uniform(0.1, 2.0, n)
    X = orthonormal_design(n, p, rng, w)
    y = X @ rng.standard_normal(p) / n + 0.3 * rng.standard_normal(n)
    om = rng.uniform(0.5, 2.0, p)
    lam = 0.01
    a = (w[:, None] * X ** 2).sum(0)
    want = soft((w * y) @ X / a, n * lam * om / a)
    assert 0 < np.count_nonzero(want) < p
    f = be.CDWeightedLSLoss(y, X, w)
    x = SparseIterate(p)
    be.coordinateDescent_(x, f, ProxL1(lam, om), TIGHT)
    got = x.toarray()
    assert np.array_equal(got != 0, want != 0)
    assert np.allclose(got, want, rtol=1e-10, atol=1e-13)


def test_quad_diagonal_and_coupled_closed_forms(be):
    # diagonal A: b_k = S(-c_k/A_kk, l0*w_k/A_kk) (:330-337)
    rng = np.random.default_rng(13)
    p = 10
    d = rng.uniform(0.5, 4.0, p)
    c = rng.standard_normal(p)
    om = rng.uniform(0.5, 2.0, p)
    lam = 0.4
    want = soft(-c / d, lam * om / d)
    assert 0 < np.count_nonzero(want) < p
    x = SparseIterate(p)
    be.coordinateDescent_(x, be.CDQuadraticLoss(np.diag(d), c), ProxL1(lam, om), TIGHT)
    assert np.array_equal(x.toarray() != 0, want != 0) and np.allclose(x.toarray(), want, rtol=1e-12, atol=1e-15)
    # coupled 3x3 with a known sign pattern: on the support S with signs s, b_S = A_SS^{-1}(-c_S - l0*w_S*s) and the
    # inactive coordinate satisfies |A_kS b_S + c_k| <= l0*w_k  (the KKT system of the objective)
    A = np.array([[2.0, 0.6, 0.3], [0.6, 1.5, -0.2], [0.3, -0.2, 1.0]])
    om = np.array([1.0, 2.0, 1.5])
    lam = 0.5
    c = np.array([-3.0, 2.5, -0.6])
    S, sg = [0, 1], np.array([1.0, -1.0])
    bS = np.linalg.solve(A[np.ix_(S, S)], -c[S] - lam * om[S] * sg)
    assert np.all(np.sign(bS) == sg) and abs(A[2, S] @ bS + c[2]) < lam * om[2]
    x = SparseIterate(3)
    be.coordinateDescent_(x, be.CDQuadraticLoss(A, c), ProxL1(lam, om), TIGHT)
    assert np.allclose(x.toarray(), [bS[0], bS[1], 0.0], rtol=1e-11, atol=0) and x.toarray()[2] == 0.0


def test_sqrt_lasso_single_column_closed_form(be):
    # p = 1: minimise ||y - x b||_2 + l|b|.  With s = x'y, xx = x'x, yy = y'y: b = 0 iff |s| <= l*sqrt(yy); otherwise
    # the stationarity condition (s - xx b)/sqrt(yy - 2 s b + xx b^2) = l sign(b) solves to
    # b = (s -+ l*sqrt((yy - s^2/xx)/(1 - l^2/xx)))/xx  — the formula of cd_differentiable_function.jl:277-283.
    rng = np.random.default_rng(14)
    n = 40
    x = rng.standard_normal(n)
    for sign in (1.0, -1.0):
        y = sign * 1.7 * x + 0.5 * rng.standard_normal(n)
        s, xx, yy = x @ y, x @ x, y @ y
        for lam, active in ((0.5, True), (2.0, True), (0.99 * abs(s) / np.sqrt(yy), True), (1.01 * abs(s) / np.sqrt(yy), False)):
            want = 0.0
            if active:
                want = (s - np.sign(s) * lam * np.sqrt((yy - s * s / xx) / (1.0 - lam * lam / xx))) / xx
                # independent check of the closed form itself: stationarity of the objective at `want`
                r = y - x * want
                assert abs(x @ r / np.linalg.norm(r) - lam * np.sign(want)) < 1e-10
            sol = be.sqrtLasso(np.asfortranarray(x[:, None]), y, lam, TIGHT, standardizeX=False)
            got = sol.x.toarray()[0]
            assert (got != 0.0) == active
            assert got == pytest.approx(want, rel=1e-10, abs=0)


def kkt_l1(grad, beta, thr, tol):
    """-grad_k = thr_k*sign(b_k) on the support, |grad_k| <= thr_k off it."""
    on = beta != 0
    assert np.all(np.abs(grad[on] + thr[on] * np.sign(beta[on])) <= tol * np.maximum(thr[on], 1.0))
    assert np.all(np.abs(grad[~on]) <= thr[~on] * (1 + tol))


@pytest.mark.parametrize("kind", ["ls", "wls", "quad", "sqrt"])
def test_kkt_certificate_general_design(be, kind):
    # convex objective + KKT system satisfied to 1e-9  =>  the minimiser, whatever solver produced it
    rng = np.random.default_rng(15)
    n, p, s = 120, 60, 6
    X = np.asfortranarray(rng.standard_normal((n, p)))
    y = X[:, :s] @ rng.uniform(1.0, 2.0, s) + 0.5 * rng.standard_normal(n)
    om = rng.uniform(0.7, 1.5, p)
    w = rng.uniform(0.2, 1.8, n)
    x = SparseIterate(p)
    if kind == "ls":
        lam = 0.15
        be.coordinateDescent_(x, be.CDLeastSquaresLoss(y, X), ProxL1(lam, om), TIGHT)
        grad = -X.T @ (y - X @ x.toarray()) / n
    elif kind == "wls":
        lam = 0.15
        be.coordinateDescent_(x, be.CDWeightedLSLoss(y, X, w), ProxL1(lam, om), TIGHT)
        grad = -X.T @ (w * (y - X @ x.toarray())) / n
    elif kind == "quad":
        lam = 0.15
        A, c = X.T @ X / n, -X.T @ y / n
        A = (A + A.T) / 2
        be.coordinateDescent_(x, be.CDQuadraticLoss(A, c), ProxL1(lam, om), TIGHT)
        grad = A @ x.toarray() + c
    else:
        lam = 3.0
        be.coordinateDescent_(x, be.CDSqrtLassoLoss(y, X), ProxL1(lam, om), TIGHT)
        r = y - X @ x.toarray()
        grad = -X.T @ r / np.linalg.norm(r)
    b = x.toarray()
    assert 0 < np.count_nonzero(b) < p
    kkt_l1(grad, b, lam * om, 1e-9)


def test_scaled_lasso_fixed_point(be):
    # lasso.jl:131-141 iterated to a fixed point: beta solves the lasso at l*sigma (KKT) AND sigma = sqrt(||r||^2/n)
    rng = np.random.default_rng(16)
    n, p, s = 150, 80, 5
    X = np.asfortranarray(rng.standard_normal((n, p)))
    y = X[:, :s] @ rng.uniform(1.0, 2.0, s) + 0.7 * rng.standard_normal(n)
    om = np.sqrt((X ** 2).sum(0) / n)
    lam = np.sqrt(2 * np.log(p) / n)
    x = SparseIterate(p)
    sol = be.scaledLasso_(x, X, y, lam, om, IterLassoOptions(maxIter=200, optTol=1e-12, initProcedure="InitStd", σinit=1.0,
                                                              optionsCD=TIGHT))
    b = x.toarray()
    r = y - X @ b
    sig = np.sqrt(r @ r / n)
    assert abs(sol.stats["sigma"] - sig) / sig < 1e-10
    kkt_l1(-X.T @ r / n, b, lam * sig * om, 1e-8)
    assert sol.σ == pytest.approx(np.std(r, ddof=1), rel=1e-12)
