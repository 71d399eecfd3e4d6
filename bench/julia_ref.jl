# bench/julia_ref.jl — times the TRUE reference (mlakolar/CoordinateDescent.jl, Julia) on the host cores, for a box that
# has Julia >= 1.5 with CoordinateDescent.jl + ProximalBase v0.3.0 installed (the build container has neither, so this
# script has never been executed there; bench.py's CPU arm uses the C port oracle/cdref.c instead — BASELINE.md §2).
#
#   julia --project=<env with CoordinateDescent> bench/julia_ref.jl [c1] [c2] [c3] [c4]      (default: all)
#
# One JSON line per configuration of BASELINE.json, same shapes, tolerances and lambda grids as bench.py /
# benchmarks/other_configs.py / tests/test_baseline_configs.py.  Julia's RNG differs from numpy's, so the DATA are a
# different draw of the same distribution: compare rates and wall times, not coefficients (parity is pinned by the
# oracle, tests/).  Threads: the reference is single-threaded; X'X/n goes to OpenBLAS with BLAS.get_num_threads().
using CoordinateDescent, ProximalBase, LinearAlgebra, Random, Statistics, Printf

json(d) = "{" * join(["\"$k\": " * (v isa AbstractString ? "\"$v\"" : string(v)) for (k, v) in d], ", ") * "}"

function problem(n, p, s, seed; noise=1.0)
    Random.seed!(seed)
    X = randn(n, p)
    β = randn(s) .* (1 .+ rand(s))                    # benchmark/cd_bench.jl:14
    y = X[:, 1:s] * β + noise * randn(n)
    X, y
end

# the reference keeps no counters: count descendCoordinate! visits by wrapping the pass structure is not possible
# without touching the package, so rates are reported per solve / per path, not per visit
function c1()
    n, p = 1000, 5000
    X, y = problem(n, p, 10, 123)
    opt = CDOptions(; randomize=false)
    for λ in (sqrt(2 * log(p) / n), 0.05, 0.01)
        lasso(X, y, λ, opt)                            # compile
        t = @elapsed sol = lasso(X, y, λ, opt)
        println(json(["config" => "C1 lasso n=$n p=$p lambda=$(round(λ, digits=4))", "julia_ms" => 1e3t, "nnz" => nnz(sol.x),
                      "threads" => 1]))
    end
end

function c2(; n=10000, p=20000, s=50, m=100, ratio=0.05)
    X, y = problem(n, p, s, 123)
    tg = @elapsed begin
        A = X' * X / n
        b = -X' * y / n
    end
    A = (A + A') / 2
    ω = sqrt.(diag(A))
    λmax = maximum(abs.(b) ./ ω)
    λs = exp.(range(log(λmax), log(ratio * λmax), length=m))
    f = CDQuadraticLoss(A, b)
    opt = CDOptions(; randomize=false, warmStart=true)
    x = SparseIterate(p)
    tc = @elapsed for λ in λs
        coordinateDescent!(x, f, ProxL1(λ, ω), opt)
    end
    println(json(["config" => "C2 covariance path n=$n p=$p $m lambdas", "gram_s" => tg, "blas_threads" => BLAS.get_num_threads(),
                  "cd_path_s" => tc, "step_s" => tg + tc, "nnz_last" => nnz(x)]))
end

function c3()
    n, p = 5000, 50000
    X, y = problem(n, p, 20, 124)
    opt = CDOptions(; randomize=false)
    λ = 1.1 * sqrt(2 * log(p))
    t = @elapsed sol = sqrtLasso(X, y, λ, opt; standardizeX=false)
    println(json(["config" => "C3 sqrt-lasso n=$n p=$p", "julia_ms" => 1e3t, "nnz" => nnz(sol.x)]))
    ω = sqrt.(vec(sum(abs2, X, dims=1)) ./ n)
    x = SparseIterate(p)
    io = IterLassoOptions(; initProcedure=:InitStd, σinit=1.0, optionsCD=opt)
    t = @elapsed sol = scaledLasso!(x, X, y, sqrt(2 * log(p) / n), ω, io)
    println(json(["config" => "C3 scaled lasso n=$n p=$p", "julia_ms" => 1e3t, "nnz" => nnz(x), "sigma" => sol.σ]))
end

function c4(; m=4096)
    n, p = 500, 50
    Random.seed!(125)
    X = randn(n, p); Z = rand(n)
    c = rand([2, 4, 6, 8], p)
    Y = [sin(c[1] * Z[i]) * X[i, 1] + sin(c[2] * Z[i]) * X[i, 2] for i in 1:n] + 0.1 * randn(n)
    zgrid = collect(range(0.01, 0.99, length=m))
    opt = CDOptions(; randomize=false, optTol=1e-9)
    locpolyl1(X, Z, Y, zgrid[1:8], 2, GaussianKernel(0.2), 0.01, false, opt)   # compile
    t = @elapsed locpolyl1(X, Z, Y, zgrid, 2, GaussianKernel(0.2), 0.01, false, opt)
    println(json(["config" => "C4 locpolyl1 $m grid points n=$n p=$p degree=2", "julia_s" => t, "problems_per_s" => m / t]))
end

which = isempty(ARGS) ? ["c1", "c2", "c3", "c4"] : ARGS
"c1" in which && c1()
"c2" in which && c2()
"c3" in which && c3()
"c4" in which && c4()
