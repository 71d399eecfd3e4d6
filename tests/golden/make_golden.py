#!/usr/bin/env python
"""Generates the golden fixtures in this directory (run from the repo root: python tests/golden/make_golden.py).

The reference (Julia + ProximalBase) cannot run in the build image and its test-suite stores NO
golden vectors (every random test draws from Julia's RNG), so the fixtures come from two sources,
recorded per case in the `source` field:

  * "sklearn"  — scikit-learn's Lasso(fit_intercept=False, tol=1e-14), an independent
                 implementation of exactly CDLeastSquaresLoss + ProxL1 (lasso.jl:26-53: objective
                 ||y - X b||^2/(2n) + lambda * sum_k omega_k |b_k|; weighted via column rescaling).
                 These pin the ORACLE as well as the CUDA path.
  * "kat"      — the reference's own known-answer test (test/coordinate_descent.jl:13-25).
  * "oracle"   — oracle/libcdref.so outputs at optTol = 1e-13 for the losses sklearn does not
                 implement (sqrt-lasso, scaled lasso, weighted LS, covariance path, locpolyl1).
                 These are regression fixtures: they pin the CUDA path and guard the oracle
                 against drift, but are not independent of the oracle.
Inputs are stored with the outputs, so the fixtures do not depend on numpy's RNG stream.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))
import cdgpu  # noqa: E402
from cdgpu import CDOptions, IterLassoOptions, ProxL1, SparseIterate, GaussianKernel, EpanechnikovKernel  # noqa: E402


def problem(n, p, s, seed, noise=0.5):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    beta = np.zeros(p)
    beta[:s] = rng.standard_normal(s) * (1.0 + rng.random(s))
    y = X[:, :s] @ beta[:s] + noise * rng.standard_normal(n)
    return X, y


def main():
    from sklearn.linear_model import Lasso

    ref = cdgpu.Backend(cdgpu.Lib(os.path.join(ROOT, "oracle", "libcdref.so"), "cdref"))
    tight = CDOptions(maxIter=200000, optTol=1e-13, randomize=False)
    out = {}

    # ---- sklearn: LS + (weighted) L1 at three lambdas
    X, y = problem(60, 90, 6, 11)
    n, p = X.shape
    om = np.sqrt((X ** 2).sum(0) / n) * (1.0 + 0.5 * np.random.default_rng(5).random(p))
    lams = np.array([0.4, 0.15, 0.05])
    B_plain, B_w = [], []
    for lam in lams:
        m = Lasso(alpha=lam, fit_intercept=False, tol=1e-14, max_iter=1000000).fit(X, y)
        B_plain.append(m.coef_.copy())
        m = Lasso(alpha=lam, fit_intercept=False, tol=1e-14, max_iter=1000000).fit(X / om, y)
        B_w.append(m.coef_ / om)
    out["sk_X"], out["sk_y"], out["sk_omega"], out["sk_lambdas"] = X, y, om, lams
    out["sk_beta_plain"], out["sk_beta_weighted"] = np.array(B_plain), np.array(B_w)

    # ---- reference KAT
    out["kat_A"], out["kat_b"], out["kat_lambda"], out["kat_x"] = np.eye(2), -np.array([1.0, 1.5]), 1.2, np.array([0.0, 0.3])

    # ---- oracle: sqrt-lasso, scaled lasso, WLS, covariance path, locpolyl1
    X, y = problem(80, 120, 5, 21, noise=1.0)
    n, p = X.shape
    om = np.sqrt((X ** 2).sum(0) / n)
    out["or_X"], out["or_y"], out["or_omega"] = X, y, om
    lam_sqrt = 1.1 * np.sqrt(2 * np.log(p))
    out["sqrt_lambda"] = lam_sqrt
    out["sqrt_beta"] = ref.sqrtLasso(X, y, lam_sqrt, om, tight).x.toarray()
    lam_sc = np.sqrt(2 * np.log(p) / n)
    x = SparseIterate(p)
    sol = ref.scaledLasso_(x, X, y, lam_sc, om, IterLassoOptions(initProcedure="InitStd", σinit=1.0, optionsCD=tight))
    out["scaled_lambda"], out["scaled_beta"], out["scaled_sigma"] = lam_sc, sol.x.toarray(), sol.σ
    w = 0.25 + np.random.default_rng(3).random(n)
    f = ref.CDWeightedLSLoss(y, X, w)
    x = SparseIterate(p)
    ref.coordinateDescent_(x, f, ProxL1(0.12, om), tight)
    out["wls_w"], out["wls_lambda"], out["wls_beta"] = w, 0.12, x.toarray()
    f.close()
    A = np.asfortranarray(X.T @ X / n)
    A = (A + A.T) / 2
    b = -(X.T @ y) / n
    f = ref.CDQuadraticLoss(A, b)
    lmax = ref.findLambdaMax(f, om)
    lp = np.exp(np.linspace(np.log(lmax), np.log(0.05 * lmax), 12))
    path = ref.LassoPath(None, None, lp, tight, standardizeX=om, loss=f)
    out["cov_A"], out["cov_b"], out["cov_lambdas"] = A, b, lp
    out["cov_betas"] = np.array([x.toarray() for x in path.βpath])
    f.close()

    rng = np.random.default_rng(31)
    n, p = 120, 6
    Xv = np.asfortranarray(rng.standard_normal((n, p)))
    z = rng.random(n)
    yv = Xv[:, 0] * np.sin(2 * z) + Xv[:, 1] * np.sin(4 * z) + 0.1 * rng.standard_normal(n)
    zg = np.linspace(0.1, 0.9, 5)
    out["vc_X"], out["vc_z"], out["vc_y"], out["vc_zgrid"] = Xv, z, yv, zg
    out["vc_gauss_d1"] = ref.locpolyl1(Xv, z, yv, zg, 1, GaussianKernel(0.2), 0.02, options=tight)[0]
    out["vc_epan_d2"] = ref.locpolyl1(Xv, z, yv, zg, 2, EpanechnikovKernel(0.4), 0.02, options=tight)[0]
    np.savez_compressed(os.path.join(HERE, "cd_golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "cd_golden_v1.npz"), {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
