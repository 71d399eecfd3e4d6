"""Golden fixtures (tests/golden/cd_golden_v1.npz, generator: tests/golden/make_golden.py).

`-m "not gpu"`: the ORACLE against the fixtures — the scikit-learn and KAT cases are independent of
it (they pin it), the "oracle" cases guard it against drift.  `-m gpu`: the CUDA path through the C
ABI against the same fixtures, at the north-star tolerance (identical supports, 1e-6 relative)."""
import os

import numpy as np
import pytest

import cdgpu
from cdgpu import CDOptions, IterLassoOptions, ProxL1, SparseIterate, GaussianKernel, EpanechnikovKernel
from helpers import assert_parity

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cd_golden_v1.npz"))
TIGHT = CDOptions(maxIter=200000, optTol=1e-13, randomize=False)


def _close(got, want, rtol):
    assert np.array_equal(got != 0, want != 0), "support sets differ"
    assert np.max(np.abs(got - want)) <= rtol * max(np.max(np.abs(want)), 1e-300)


def check_all(be, rtol):
    X, y, om = np.asfortranarray(G["sk_X"]), G["sk_y"], G["sk_omega"]
    for i, lam in enumerate(G["sk_lambdas"]):  # independent implementation (scikit-learn)
        _close(be.lasso(X, y, lam, None, TIGHT).x.toarray(), G["sk_beta_plain"][i], rtol)
        _close(be.lasso(X, y, lam, om, TIGHT).x.toarray(), G["sk_beta_weighted"][i], rtol)
    f = be.CDQuadraticLoss(np.asfortranarray(G["kat_A"]), G["kat_b"])  # test/coordinate_descent.jl:13-25
    x = SparseIterate(2)
    be.coordinateDescent_(x, f, ProxL1(float(G["kat_lambda"])), CDOptions(maxIter=100, optTol=1e-8, randomize=False))
    assert np.allclose(x.toarray(), G["kat_x"], rtol=1e-8, atol=0)
    f.close()
    X, y, om = np.asfortranarray(G["or_X"]), G["or_y"], G["or_omega"]
    p = X.shape[1]
    _close(be.sqrtLasso(X, y, float(G["sqrt_lambda"]), om, TIGHT).x.toarray(), G["sqrt_beta"], rtol)
    sol = be.scaledLasso_(SparseIterate(p), X, y, float(G["scaled_lambda"]), om,
                          IterLassoOptions(initProcedure="InitStd", σinit=1.0, optionsCD=TIGHT))
    _close(sol.x.toarray(), G["scaled_beta"], rtol)
    assert abs(sol.σ - float(G["scaled_sigma"])) <= 1e-8 * float(G["scaled_sigma"])
    f = be.CDWeightedLSLoss(y, X, G["wls_w"])
    x = SparseIterate(p)
    be.coordinateDescent_(x, f, ProxL1(float(G["wls_lambda"]), om), TIGHT)
    _close(x.toarray(), G["wls_beta"], rtol)
    f.close()
    f = be.CDQuadraticLoss(np.asfortranarray(G["cov_A"]), G["cov_b"])
    path = be.LassoPath(None, None, G["cov_lambdas"], TIGHT, standardizeX=om, loss=f)
    assert len(path.βpath) == len(G["cov_lambdas"])
    for x, want in zip(path.βpath, G["cov_betas"]):
        _close(x.toarray(), want, rtol)
    f.close()
    Xv, z, yv, zg = np.asfortranarray(G["vc_X"]), G["vc_z"], G["vc_y"], G["vc_zgrid"]
    _close(be.locpolyl1(Xv, z, yv, zg, 1, GaussianKernel(0.2), 0.02, options=TIGHT)[0], G["vc_gauss_d1"], rtol)
    _close(be.locpolyl1(Xv, z, yv, zg, 2, EpanechnikovKernel(0.4), 0.02, options=TIGHT)[0], G["vc_epan_d2"], rtol)


def test_oracle_matches_golden(ref):
    check_all(ref, 1e-8)


@pytest.mark.gpu
def test_cuda_path_matches_golden(gpu):
    check_all(gpu, 1e-6)
