"""Shared test helpers: synthetic data (numpy PCG64, seeded), objectives, KKT residuals."""
import numpy as np


def gauss_problem(n, p, s, seed, noise=1.0, beta_scale=1.0):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    beta = np.zeros(p)
    beta[:s] = rng.standard_normal(s) * beta_scale
    y = X[:, :s] @ beta[:s] + noise * rng.standard_normal(n)
    return X, y, beta


def sprand_iterate(p, density, rng):
    v = np.zeros(p)
    mask = rng.random(p) < density
    v[mask] = rng.random(mask.sum())
    return v


def lasso_objective(X, y, beta, lam, omega=None):
    n = X.shape[0]
    om = np.ones(X.shape[1]) if omega is None else omega
    r = y - X @ beta
    return r @ r / (2 * n) + lam * np.sum(om * np.abs(beta))


def quad_objective(A, b, beta, lam, omega=None):
    om = np.ones(A.shape[0]) if omega is None else omega
    return 0.5 * beta @ A @ beta + b @ beta + lam * np.sum(om * np.abs(beta))


def sqrt_objective(X, y, beta, lam, omega=None):
    om = np.ones(X.shape[1]) if omega is None else omega
    return np.linalg.norm(y - X @ beta) + lam * np.sum(om * np.abs(beta))


def rel_err(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def assert_parity(beta_gpu, beta_ref, obj_gpu=None, obj_ref=None, rtol=1e-6, otol=1e-8):
    """north_star tolerance: identical supports, coefficients within 1e-6 relative,
    objective within 1e-8."""
    assert np.array_equal(beta_gpu != 0, beta_ref != 0), "support sets differ"
    assert rel_err(beta_gpu, beta_ref) <= rtol, f"coefficients differ: {rel_err(beta_gpu, beta_ref):.3e}"
    if obj_gpu is not None:
        assert abs(obj_gpu - obj_ref) <= otol * max(1.0, abs(obj_ref)), f"objective differs {obj_gpu} vs {obj_ref}"
