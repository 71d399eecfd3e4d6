"""Moment (covariance) form of the varying-coefficient path against the residual-form kernels on mid-size problems:
per-problem pass counts and the largest coefficient difference.  Diagnostic, not a test."""
import os, sys, subprocess, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "coordinatedescent.jl_b200"))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import cdgpu
    from cdgpu import CDOptions, GaussianKernel
    n, p, dgr, m = [int(v) for v in sys.argv[2:6]]
    rng = np.random.default_rng(7)
    X = np.asfortranarray(rng.standard_normal((n, p)))
    z = rng.random(n)
    y = sum(X[:, j] * np.sin((2 + 2 * j) * z) for j in range(2)) + 0.1 * rng.standard_normal(n)
    zg = np.linspace(0.01, 0.99, m)
    be = cdgpu.default()
    out, _ = be.locpolyl1(X, z, y, zg, dgr, GaussianKernel(0.2), float(sys.argv[6]), options=CDOptions(randomize=False, optTol=1e-10, maxIter=20000))
    st = be.last_vc_stats
    np.save(sys.argv[7], out)
    print(json.dumps({"passes": [s["passes"] for s in st][:8], "conv": sum(s["converged"] for s in st), "nnz": int(np.count_nonzero(out))}))
else:
    for cfg in (["200", "20", "2", "64", "0.01"], ["500", "50", "2", "64", "0.01"], ["500", "50", "2", "64", "0.002"]):
        res = {}
        for name, env in (("naive", {"CDGPU_VC_FORM": "naive"}), ("cov", {})):
            f = f"/tmp/vc_{name}.npy"
            r = subprocess.run([sys.executable, __file__, "child", *cfg, f], env=dict(os.environ, **env), capture_output=True, text=True)
            print(cfg, name, r.stdout.strip()[-300:], r.stderr.strip()[-300:])
            res[name] = np.load(f)
        for name in ("cov",):
            d = np.abs(res[name] - res["naive"]).max(axis=0)
            print(cfg, name, "max abs diff vs naive %.3e" % d.max(), "problems off by >1e-6:", int((d > 1e-6).sum()),
                  "support equal:", bool(np.array_equal(res[name] != 0, res["naive"] != 0)))
