"""Pins the oracle against the reference's solver-independent properties
(test/lasso.jl, test/coordinate_descent.jl) and against scikit-learn."""
import numpy as np
import pytest

from cdgpu import CDOptions, IterLassoOptions, ProxL1, SparseIterate
from helpers import gauss_problem, sprand_iterate


def test_lasso_zero_above_lambda_max(ref):
    # test/lasso.jl:23-34
    rng = np.random.default_rng(1)
    n, p = 100, 10
    X = np.asfortranarray(rng.standard_normal((n, p)))
    Y = X @ np.ones(p) + 0.1 * rng.standard_normal(n)
    lam = np.max(np.abs(X.T @ Y / n)) + 0.1
    out = ref.lasso(X, Y, lam)
    assert out.x == SparseIterate(p) and out.x.nnz == 0


def test_lasso_weighted_equals_covariance_form_and_kkt(ref):
    # test/lasso.jl:36-56
    rng = np.random.default_rng(2)
    n, p, s = 100, 10, 5
    X = np.asfortranarray(rng.standard_normal((n, p)))
    Y = X[:, :s] @ np.ones(s) + 0.1 * rng.standard_normal(n)
    lam = np.full(p, 0.3)
    beta = ref.lasso(X, Y, 1.0, lam, CDOptions(optTol=1e-12))
    f = ref.CDQuadraticLoss(X.T @ X / n, -X.T @ Y / n)
    x1 = SparseIterate(p)
    ref.coordinateDescent_(x1, f, ProxL1(1.0, lam), CDOptions(optTol=1e-12))
    assert np.allclose(beta.x.toarray(), x1.toarray(), atol=1e-5)
    kkt = np.max(np.abs(X.T @ (Y - X @ beta.x.toarray()) / n))
    assert abs(kkt - 0.3) / 0.3 < 1e-5


def test_lasso_scalar_equals_ones(ref):
    # test/lasso.jl:58-72
    X, Y, _ = gauss_problem(500, 500, 50, seed=3)
    x1 = ref.lasso(X, Y, 0.1)
    x2 = ref.lasso(X, Y, 0.1, np.ones(500))
    assert np.allclose(x1.x.toarray(), x2.x.toarray(), atol=1e-5)


def test_cd_lasso_ls_equals_quadratic(ref):
    # test/lasso.jl:76-101
    X, Y, _ = gauss_problem(200, 50, 10, seed=4, noise=0.1)
    n = 200
    g = ProxL1(0.2)
    f1 = ref.CDQuadraticLoss(X.T @ X / n, -X.T @ Y / n)
    f2 = ref.CDLeastSquaresLoss(Y, X)
    x1, x2 = SparseIterate(50), SparseIterate(50)
    ref.coordinateDescent_(x1, f1, g, CDOptions(optTol=1e-12))
    ref.coordinateDescent_(x2, f2, g, CDOptions(optTol=1e-12))
    assert np.max(np.abs(x1.toarray() - x2.toarray())) < 1e-5
    for x in (x1, x2):
        assert abs(np.max(np.abs(X.T @ (Y - X @ x.toarray()) / n)) - 0.2) / 0.2 < 1e-5


@pytest.mark.parametrize("weighted,tol", [(False, 1e-12), (True, 1e-8)])
def test_warm_cold_ordered_random_agree(ref, weighted, tol):
    # test/coordinate_descent.jl:29-63 (ProxL1) and :65-99 (AProxL1)
    rng = np.random.default_rng(5)
    n, p, s = 500, 50, 10 if weighted else 5
    X, Y, _ = gauss_problem(n, p, s, seed=6)
    g = ProxL1(0.01, rng.random(p)) if weighted else ProxL1(0.02)
    f = ref.CDLeastSquaresLoss(Y, X)
    sols = []
    for warm in (True, False):
        for rand in (0, 1, 2):
            x = SparseIterate(sprand_iterate(p, 0.6, rng))
            ref.coordinateDescent_(x, f, g, CDOptions(maxIter=5000, optTol=tol, warmStart=warm, randomize=rand,
                                                      seed=11))
            assert f.last_stats["converged"] == 1
            sols.append(x.toarray())
    for s_ in sols[1:]:
        assert np.allclose(s_, sols[0], atol=1e-5)


def test_sqrt_lasso_kkt(ref):
    # test/lasso.jl:106-125
    X, Y, _ = gauss_problem(100, 50, 5, seed=7)
    lam = 2.8
    f = ref.CDSqrtLassoLoss(Y, X)
    x1 = SparseIterate(50)
    ref.coordinateDescent_(x1, f, ProxL1(lam), CDOptions(maxIter=5000, optTol=1e-8))
    r = Y - X @ x1.toarray()
    assert max(0, np.max(np.abs(X.T @ r / np.linalg.norm(r))) - lam) / lam < 1e-3
    assert np.allclose(f.r, r, atol=1e-9)


def test_sqrt_lasso_interfaces(ref):
    # test/lasso.jl:127-181 (p reduced from 500 to 200 to keep the CPU suite short)
    rng = np.random.default_rng(8)
    n, p, s = 500, 200, 20
    X, Y, _ = gauss_problem(n, p, s, seed=9)
    lam = 1.5
    f = ref.CDSqrtLassoLoss(Y, X)
    sols = []
    for warm in (True, False):
        for rand in (0, 1):
            o = CDOptions(maxIter=5000, optTol=1e-10, warmStart=warm, randomize=rand, seed=5)
            x = SparseIterate(sprand_iterate(p, 0.6, rng))
            ref.coordinateDescent_(x, f, ProxL1(lam), o)
            sols.append(x.toarray())
            sols.append(ref.sqrtLasso(X, Y, lam, o, standardizeX=False).x.toarray())
            sols.append(ref.sqrtLasso(X, Y, lam, np.ones(p), o).x.toarray())
    for s_ in sols[1:]:
        assert np.allclose(s_, sols[0], atol=1e-4)


def test_scaled_lasso_kkt_and_init_independence(ref):
    # test/lasso.jl:186-216
    n, p, s = 1000, 500, 50
    X, Y, _ = gauss_problem(n, p, s, seed=10)
    lam = 0.13
    cd = CDOptions(maxIter=5000, optTol=1e-8)
    opt1 = IterLassoOptions(maxIter=100, optTol=1e-8, optionsCD=cd)
    opt2 = IterLassoOptions(maxIter=100, optTol=1e-8, initProcedure="InitStd", σinit=2.0, optionsCD=cd)
    x1, x2 = SparseIterate(p), SparseIterate(p)
    sol1 = ref.scaledLasso_(x1, X, Y, lam, np.ones(p), opt1)
    sol2 = ref.scaledLasso_(x2, X, Y, lam, np.ones(p), opt2)
    for x, sol in ((x1, sol1), (x2, sol2)):
        kkt = np.max(np.abs(X.T @ (Y - X @ x.toarray()) / n))
        # exact property: KKT holds at λ·σ of the last ProxL1 (lasso.jl:131-140) ...
        assert abs(kkt - lam * sol.stats["sigma"]) / (sol.stats["sigma"] * lam) < 1e-6
        # ... and the reference's assertion against the returned std(r) (n-1, centred) holds up to O(1/n)
        assert max(kkt - lam * sol.σ, 0.0) / (sol.σ * lam) < 5e-3
    assert np.allclose(x1.toarray(), x2.toarray(), atol=1e-4)


@pytest.mark.parametrize("standardize", [False, True])
def test_lasso_path_equals_pointwise(ref, standardize):
    # test/lasso.jl:220-288
    n, p, s = 1000, 500, 50
    X, Y, _ = gauss_problem(n, p, s, seed=12)
    opt = CDOptions(maxIter=5000, optTol=1e-8)
    lam1, lam2 = 0.3, 0.1
    if standardize:
        load = np.sqrt((X ** 2).sum(0) / n)  # _stdX!
        f = ref.CDLeastSquaresLoss(Y, X)
        assert np.allclose(f.stdX(), load, rtol=1e-13)
        x1, x2 = ref.lasso(X, Y, lam1, load, opt), ref.lasso(X, Y, lam2, load, opt)
    else:
        x1, x2 = ref.lasso(X, Y, lam1, opt), ref.lasso(X, Y, lam2, opt)
    path = ref.LassoPath(X, Y, [lam1, lam2], opt, standardizeX=standardize)
    assert np.allclose(path.βpath[0].toarray(), x1.x.toarray(), atol=1e-5)
    assert np.allclose(path.βpath[1].toarray(), x2.x.toarray(), atol=1e-5)
    # max_hat_s stops the path early (lasso.jl:253-256)
    short = ref.LassoPath(X, Y, [lam1, lam2], opt, standardizeX=standardize, max_hat_s=1)
    assert len(short.βpath) == 1 and short.λpath.tolist() == [lam1]


def test_against_sklearn(ref):
    sk = pytest.importorskip("sklearn.linear_model")
    X, Y, _ = gauss_problem(300, 120, 8, seed=13)
    for lam in (0.3, 0.1, 0.03):
        m = sk.Lasso(alpha=lam, fit_intercept=False, tol=1e-14, max_iter=100000).fit(X, Y)
        out = ref.lasso(X, Y, lam, CDOptions(optTol=1e-12, randomize=False, maxIter=20000))
        b = out.x.toarray()
        assert np.array_equal(b != 0, m.coef_ != 0)
        assert np.max(np.abs(b - m.coef_)) < 1e-8
    # weighted L1 == column rescale
    om = 0.5 + np.random.default_rng(0).random(120)
    m = sk.Lasso(alpha=0.1, fit_intercept=False, tol=1e-14, max_iter=100000).fit(X / om, Y)
    out = ref.lasso(X, Y, 0.1, om, CDOptions(optTol=1e-12, randomize=False, maxIter=20000))
    assert np.max(np.abs(out.x.toarray() - m.coef_ / om)) < 1e-8


def test_wls_equals_sqrtw_scaled_ls(ref):
    # the disabled reference test test/coordinate_descent.jl:101-161 states this oracle for CDWeightedLSLoss
    rng = np.random.default_rng(14)
    X, Y, _ = gauss_problem(300, 40, 6, seed=15)
    w = rng.random(300) + 0.1
    g = ProxL1(0.05, 0.5 + rng.random(40))
    o = CDOptions(maxIter=5000, optTol=1e-12, randomize=False)
    x1, x2 = SparseIterate(40), SparseIterate(40)
    ref.coordinateDescent_(x1, ref.CDWeightedLSLoss(Y, X, w), g, o)
    sw = np.sqrt(w)
    ref.coordinateDescent_(x2, ref.CDLeastSquaresLoss(sw * Y, np.asfortranarray(sw[:, None] * X)), g, o)
    assert np.allclose(x1.toarray(), x2.toarray(), atol=1e-9)


def test_errors(ref):
    import cdgpu
    X, Y, _ = gauss_problem(20, 5, 2, seed=16)
    with pytest.raises(cdgpu.DimensionMismatch):
        ref.CDLeastSquaresLoss(Y[:-1], X)
    with pytest.raises(cdgpu.ArgumentError):
        A = X.T @ X
        A[0, 1] += 1e-3
        ref.CDQuadraticLoss(A, np.zeros(5))
    f = ref.CDLeastSquaresLoss(Y, X)
    with pytest.raises(cdgpu.DimensionMismatch):
        ref.coordinateDescent_(SparseIterate(4), f, ProxL1(0.1))
    with pytest.raises(cdgpu.DimensionMismatch):
        ref.coordinateDescent_(SparseIterate(5), f, ProxL1(0.1, np.ones(4)))
    with pytest.raises(cdgpu.ArgumentError):
        ref.scaledLasso_(SparseIterate(5), X, Y, 0.1, np.ones(5), IterLassoOptions(initProcedure="Nope"))


def test_locpolyl1_matches_alt_formulation(ref):
    # intended oracle of the disabled test test/varying_coefficient_lasso.jl:121-142 and
    # benchmark/locpoly_bench.jl:72-120: per grid point, lasso on sqrt(w)-scaled expanded data
    from cdgpu import GaussianKernel
    rng = np.random.default_rng(17)
    n, p, degree = 200, 6, 1
    X = np.asfortranarray(rng.standard_normal((n, p)))
    Z = rng.random(n)
    c = rng.choice([2, 4, 6, 8], size=p)
    Y = np.array([np.sin(c * Z[i])[:2] @ X[i, :2] for i in range(n)]) + 0.1 * rng.standard_normal(n)
    zgrid = np.array([0.2, 0.5, 0.8])
    k = GaussianKernel(0.2)
    o = CDOptions(randomize=False, optTol=1e-12, maxIter=20000)
    out, _ = ref.locpolyl1(X, Z, Y, zgrid, degree, k, 0.05, False, o)
    ep = p * (degree + 1)
    for gi, z0 in enumerate(zgrid):
        w = np.exp(-(Z - z0) ** 2 / k.h) / k.h
        eX = np.zeros((n, ep), order="F")
        for j in range(p):
            for l in range(degree + 1):
                eX[:, j * (degree + 1) + l] = X[:, j] * (Z - z0) ** l
        sd = np.sqrt((w[:, None] * eX ** 2).sum(0) / n)
        sw = np.sqrt(w)
        alt = ref.lasso(np.asfortranarray(sw[:, None] * eX), sw * Y, 0.05, sd, o)
        assert np.allclose(out[:, gi], alt.x.toarray(), atol=1e-8)
    assert np.count_nonzero(out) > 0


def test_lvocv_oracle_prefers_the_true_bandwidth_and_chain_cut_agree(ref):
    """lvocv_locpolyl1 (varying_coefficient_lasso.jl:81-137) in the oracle: the warm-start chain across observations
    (reference) and independent problems (what the device batches) give the same leave-one-out MSE, and an absurdly
    wide bandwidth loses against a reasonable one on a strongly varying coefficient."""
    import cdgpu
    from cdgpu import CDOptions, GaussianKernel
    rng = np.random.default_rng(96)
    n, p = 80, 5
    X = np.asfortranarray(rng.standard_normal((n, p)))
    Z = rng.random(n)
    Y = X[:, 0] * np.sin(6 * Z) + 0.1 * rng.standard_normal(n)
    h = np.array([0.02, 5.0])
    o = dict(randomize=False, maxIter=20000, optTol=1e-10)
    chain = ref.lvocv_locpolyl1(X, Z, Y, 1, h, GaussianKernel, 0.2, CDOptions(warmStart=True, **o))
    cut = ref.lvocv_locpolyl1(X, Z, Y, 1, h, GaussianKernel, 0.2, CDOptions(warmStart=False, **o))
    assert np.allclose(chain, cut, rtol=1e-3)
    assert cut[0] < 0.5 * cut[1]


def test_locpolyl1_chain_runs_in_the_oracle(ref):
    """cdref_vc_solve_chain: the reference's warm-start chain (varying_coefficient_lasso.jl:56,68) cut into runs of k
    grid points.  k = m is the plain oracle call; a run equals the plain call on its own grid points; warm starts change
    the pass counts of a run's later points but, at a tight tolerance, not the solution."""
    from cdgpu import CDOptions, GaussianKernel
    rng = np.random.default_rng(97)
    n, p, m = 120, 6, 7
    X = np.asfortranarray(rng.standard_normal((n, p)))
    Z = rng.random(n)
    Y = X[:, 0] * np.sin(4 * Z) + X[:, 1] * Z + 0.1 * rng.standard_normal(n)
    zgrid = np.linspace(0.15, 0.85, m)
    loose = CDOptions(randomize=False, maxIter=20000, optTol=1e-4)
    whole, _ = ref.locpolyl1(X, Z, Y, zgrid, 1, GaussianKernel(0.25), 0.02, False, loose)
    sw = ref.last_vc_stats
    chained, _ = ref.locpolyl1(X, Z, Y, zgrid, 1, GaussianKernel(0.25), 0.02, False, loose, chain=m)
    assert np.array_equal(whole, chained) and [s["passes"] for s in sw] == [s["passes"] for s in ref.last_vc_stats]
    runs, _ = ref.locpolyl1(X, Z, Y, zgrid, 1, GaussianKernel(0.25), 0.02, False, loose, chain=3)
    sr = ref.last_vc_stats
    for b0 in range(0, m, 3):
        own, _ = ref.locpolyl1(X, Z, Y, zgrid[b0:b0 + 3], 1, GaussianKernel(0.25), 0.02, False, loose)
        assert np.array_equal(runs[:, b0:b0 + 3], own)
        assert [s["passes"] for s in sr[b0:b0 + 3]] == [s["passes"] for s in ref.last_vc_stats]
    tight = CDOptions(randomize=False, maxIter=200000, optTol=1e-12)
    a, _ = ref.locpolyl1(X, Z, Y, zgrid, 1, GaussianKernel(0.25), 0.02, False, tight, chain=1)
    b, _ = ref.locpolyl1(X, Z, Y, zgrid, 1, GaussianKernel(0.25), 0.02, False, tight, chain=m)
    assert np.array_equal(a != 0, b != 0) and np.max(np.abs(a - b)) <= 1e-8 * np.max(np.abs(b))
    with pytest.raises(ValueError):
        ref.locpolyl1(X, Z, Y, zgrid, 1, GaussianKernel(0.25), 0.02, False, loose, chain=0)


def test_oracle_refits_match_numpy(ref):
    """refitLassoPath (lasso.jl:208-225, test/lasso.jl:236-241) and locpolyl1(refit=true)
    (varying_coefficient_lasso.jl:71-76) in the oracle against numpy's own least squares."""
    from cdgpu import CDOptions, GaussianKernel
    X, y, _ = gauss_problem(200, 60, 6, seed=51)
    o = CDOptions(randomize=False, maxIter=20000, optTol=1e-12)
    path = ref.LassoPath(X, y, [0.3, 0.1, 0.03], o, standardizeX=False)
    rf = ref.refitLassoPath(path, X, y)
    assert len(rf) >= 2
    for S, coef in rf.items():
        assert np.allclose(coef, np.linalg.lstsq(X[:, list(S)], y, rcond=None)[0], rtol=1e-9, atol=1e-12)
    rng = np.random.default_rng(52)
    Xs = np.asfortranarray(rng.standard_normal((150, 5)))
    Z = rng.random(150)
    Y = np.sin(4 * Z) * Xs[:, 0] + 0.1 * rng.standard_normal(150)
    zg = np.array([0.25, 0.5, 0.75])
    out, outR = ref.locpolyl1(Xs, Z, Y, zg, 2, GaussianKernel(0.2), 0.05, True, o)
    for g in range(3):
        grp = np.flatnonzero(np.any(out[:, g].reshape(5, 3) != 0, axis=1))
        S = (grp[:, None] * 3 + np.arange(3)).ravel()
        d = Z - zg[g]
        w = np.exp(-d ** 2 / 0.2) / 0.2
        E = np.stack([Xs[:, k // 3] * d ** (k % 3) for k in S], axis=1)
        t = E.T * w
        assert np.allclose(outR[S, g], np.linalg.solve(t @ E, t @ Y), rtol=1e-9, atol=1e-12)
        assert np.count_nonzero(outR[:, g]) == S.size
