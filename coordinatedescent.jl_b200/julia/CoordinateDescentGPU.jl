# CoordinateDescentGPU.jl — Julia host side of libcdgpu.so: the reference's names, `ccall` underneath.
#
# REVIEWED, NOT EXECUTED: Julia and ProximalBase are not available where this repository is built
# (SURVEY.md §8c); the executable contract of the C ABI is the Python/ctypes harness in
# ../cdgpu/ and tests/.  This file is the binding a CoordinateDescent.jl maintainer would add:
# it keeps the package's types and exported functions and overrides the hot path
# (coordinateDescent!, lasso, sqrtLasso, scaledLasso!, LassoPath, locpolyl1) with one ccall each.
#
# Two deliberate differences from the reference's loss objects:
#  * COPY, NOT ALIAS.  The reference's losses alias the caller's arrays (locpolyl1 mutates `expandX` and `w` in place
#    between solves, varying_coefficient_lasso.jl:54,63-65).  A device handle snapshots X / y / w / A / b at
#    construction: after mutating them call `update!(f)` (or build a new loss).  The front-ends of this module never
#    rely on aliasing.
#  * The iterate is exchanged through ProximalBase's PUBLIC interface only (`fill!`, `setindex!`, `nnz`, the
#    `nzval`/`nzval2ind` fields the reference itself reads in atom_iterator.jl:25,74): no assumption about the
#    mutability of `SparseIterate` or about constructors the reference does not use.
#
#   using CoordinateDescentGPU            # instead of `using CoordinateDescent`
#   f = CDQuadraticLoss(X'X/n, -X'y/n)    # or CDQuadraticLoss(X, y, Val(:data)) / Val(:lazy): covariance form on the GPU
#   coordinateDescent!(x, f, ProxL1(λ, ω), CDOptions(; optTol=1e-8))
#
module CoordinateDescentGPU

using ProximalBase            # SparseIterate, ProxL1 (the iterate and penalty types stay ProximalBase's)
using SparseArrays
using Statistics
using LinearAlgebra

export lasso, sqrtLasso, scaledLasso!, LassoPath, LassoSolution,
       IterLassoOptions, CDOptions,
       CoordinateDifferentiableFunction,
       CDLeastSquaresLoss, CDWeightedLSLoss, CDQuadraticLoss, CDSqrtLassoLoss,
       coordinateDescent!,
       GaussianKernel, SmoothingKernel, EpanechnikovKernel, evaluate, createKernel, locpolyl1,
       refitLassoPath, lvocv_locpolyl1, update!

const libcdgpu = get(ENV, "LIBCDGPU", joinpath(@__DIR__, "..", "csrc", "libcdgpu.so"))

# ---------------------------------------------------------------- status codes (include/cdgpu.h)
const CDGPU_OK, CDGPU_EDIM, CDGPU_EARG = Cint(0), Cint(1), Cint(2)

function check(rc::Cint)
    rc == CDGPU_OK && return nothing
    msg = unsafe_string(ccall((:cdgpu_last_error, libcdgpu), Cstring, ()))
    rc == CDGPU_EDIM && throw(DimensionMismatch(msg))        # coordinate_descent.jl:13,15
    rc == CDGPU_EARG && throw(ArgumentError(msg))            # cd_differentiable_function.jl:306
    error("libcdgpu [$rc]: $msg")                            # CUDA / NCCL / memory: ErrorException
end

# ---------------------------------------------------------------- options (src/utils.jl:7-39)
struct CDOptions
    maxIter::Int64
    optTol::Float64
    randomize::Bool
    warmStart::Bool
    numSteps::Int64
    seed::UInt64      # stands in for Julia's global RNG (atom_iterator.jl:60)
end
CDOptions(; maxIter::Int64=2000, optTol::Float64=1e-7, randomize::Bool=true, warmStart::Bool=true,
          numSteps::Int=50, seed::Integer=rand(UInt64)) =
    CDOptions(maxIter, optTol, randomize, warmStart, numSteps, UInt64(seed))

struct cdgpu_options       # layout of `struct cdgpu_options`
    maxIter::Int64; optTol::Float64; randomize::Int32; warmStart::Int32; numSteps::Int64; seed::UInt64
end
c_opts(o::CDOptions) = cdgpu_options(o.maxIter, o.optTol, o.randomize, o.warmStart, o.numSteps, o.seed)

struct IterLassoOptions
    maxIter::Int64
    optTol::Float64
    initProcedure::Symbol     # :Screening, :InitStd, :WarmStart
    sinit::Int64
    σinit::Float64
    optionsCD::CDOptions
end
IterLassoOptions(; maxIter::Int64=20, optTol::Float64=1e-2, initProcedure::Symbol=:Screening, sinit::Int64=5,
                 σinit::Float64=1., optionsCD::CDOptions=CDOptions()) =
    IterLassoOptions(maxIter, optTol, initProcedure, sinit, σinit, optionsCD)

struct cdgpu_iter_options
    maxIter::Int64; optTol::Float64; initProcedure::Int32; _pad::Int32; sinit::Int64; sigma_init::Float64
    optionsCD::cdgpu_options
end
function c_opts(o::IterLassoOptions)
    init = o.initProcedure == :Screening ? 0 : o.initProcedure == :InitStd ? 1 : o.initProcedure == :WarmStart ? 2 :
           throw(ArgumentError("Incorrect initialization Symbol"))          # lasso.jl:128
    cdgpu_iter_options(o.maxIter, o.optTol, init, 0, o.sinit, o.σinit, c_opts(o.optionsCD))
end

struct cdgpu_stats
    passes::Int64; full_passes::Int64; visits::Int64; accepted::Int64; maxH::Float64
    converged::Int32; outer_iters::Int32; sigma::Float64; device_ms::Float64
end

# ---------------------------------------------------------------- losses = device handles
abstract type CoordinateDifferentiableFunction end

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(p::Ptr{Cvoid})
        h = new(p)
        finalizer(h -> (ccall((:cdgpu_destroy, libcdgpu), Cint, (Ptr{Cvoid},), h.ptr); nothing), h)
        h
    end
end

for (T, kind) in ((:CDLeastSquaresLoss, 0), (:CDSqrtLassoLoss, 2))
    @eval begin
        struct $T{T<:AbstractFloat, S, U} <: CoordinateDifferentiableFunction
            y::S; X::U; r::Vector{T}; h::Handle
        end
        function $T(y::AbstractVector{Float64}, X::StridedMatrix{Float64}; device::Integer=0)
            length(y) == size(X, 1) || throw(DimensionMismatch())            # cd_differentiable_function.jl:53,212
            hp = Ref{Ptr{Cvoid}}(C_NULL)
            GC.@preserve X y check(ccall((:cdgpu_naive_create, libcdgpu), Cint,
                (Ref{Ptr{Cvoid}}, Cint, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Cint),
                hp, $kind, X, size(X, 1), size(X, 2), stride(X, 2), y, C_NULL, device))
            $T{Float64, typeof(y), typeof(X)}(y, X, copy(y), Handle(hp[]))
        end
    end
end

struct CDWeightedLSLoss{T<:AbstractFloat, S, U} <: CoordinateDifferentiableFunction
    y::S; X::U; w::S; r::Vector{T}; h::Handle
end
function CDWeightedLSLoss(y::AbstractVector{Float64}, X::StridedMatrix{Float64}, w::AbstractVector{Float64}; device::Integer=0)
    length(y) == size(X, 1) == length(w) || throw(DimensionMismatch())        # :129
    hp = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve X y w check(ccall((:cdgpu_naive_create, libcdgpu), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Cint),
        hp, 1, X, size(X, 1), size(X, 2), stride(X, 2), y, w, device))
    CDWeightedLSLoss{Float64, typeof(y), typeof(X)}(y, X, w, copy(y), Handle(hp[]))
end

struct CDQuadraticLoss{T<:AbstractFloat, S, U} <: CoordinateDifferentiableFunction
    A::S; b::U; Ax::Vector{T}; h::Handle
end
function CDQuadraticLoss(A::StridedMatrix{Float64}, b::AbstractVector{Float64}; device::Integer=0)
    (size(A, 1) == size(A, 2) && length(b) == size(A, 2)) || throw(ArgumentError("A must be square, length(b) == size(A,2)"))
    hp = Ref{Ptr{Cvoid}}(C_NULL)      # issymmetric(A) is checked on the device -> CDGPU_EARG -> ArgumentError (:306)
    GC.@preserve A b check(ccall((:cdgpu_quad_create, libcdgpu), Cint,
        (Ref{Ptr{Cvoid}}, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Cint), hp, A, size(A, 1), stride(A, 2), b, device))
    CDQuadraticLoss{Float64, typeof(A), typeof(b)}(A, b, zeros(length(b)), Handle(hp[]))
end
"""
    CDQuadraticLoss(X, y, ::Val{:data})

Covariance form straight from the data: `A = X'X/n`, `b = -X'y/n` formed by the FP64 tensor-core
SYRK kernel on the device (what users of the reference write by hand, test/lasso.jl:48,88).
"""
function CDQuadraticLoss(X::StridedMatrix{Float64}, y::AbstractVector{Float64}, ::Val{:data}; device::Integer=0)
    length(y) == size(X, 1) || throw(DimensionMismatch())
    hp = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve X y check(ccall((:cdgpu_gram_create, libcdgpu), Cint,
        (Ref{Ptr{Cvoid}}, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Cint),
        hp, X, size(X, 1), size(X, 2), stride(X, 2), y, device))
    p = size(X, 2)
    CDQuadraticLoss{Float64, Nothing, Nothing}(nothing, nothing, zeros(p), Handle(hp[]))
end

"""
    CDQuadraticLoss(X, y, Val(:lazy))

The same covariance-form loss without forming the p x p matrix: `diag(A)` and `b` up front, columns of `A = X'X/n` on
demand as coordinates become non-zero (cdgpu_gram_create_lazy).  For paths whose supports stay small.
"""
function CDQuadraticLoss(X::StridedMatrix{Float64}, y::AbstractVector{Float64}, ::Val{:lazy}; device::Integer=0)
    length(y) == size(X, 1) || throw(DimensionMismatch())
    hp = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve X y check(ccall((:cdgpu_gram_create_lazy, libcdgpu), Cint,
        (Ref{Ptr{Cvoid}}, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Cint),
        hp, X, size(X, 1), size(X, 2), stride(X, 2), y, device))
    CDQuadraticLoss{Float64, Nothing, Nothing}(nothing, nothing, zeros(size(X, 2)), Handle(hp[]))
end

"""
    update!(f)

Re-upload the arrays a loss was built from after they were mutated in place (device handles copy, they do not alias).
"""
function update!(f::Union{CDLeastSquaresLoss, CDSqrtLassoLoss})
    g = typeof(f).name.wrapper(f.y, f.X)
    f.h.ptr, g.h.ptr = g.h.ptr, f.h.ptr      # the old handle is destroyed by g's finalizer
    f
end
function update!(f::CDWeightedLSLoss)
    g = CDWeightedLSLoss(f.y, f.X, f.w)
    f.h.ptr, g.h.ptr = g.h.ptr, f.h.ptr
    f
end

numCoordinates(f::CDQuadraticLoss) = length(f.Ax)
numCoordinates(f::CoordinateDifferentiableFunction) = size(f.X, 2)
state(f::CDQuadraticLoss) = f.Ax
state(f::CoordinateDifferentiableFunction) = f.r
sync_state!(f) = (s = state(f); GC.@preserve s check(ccall((:cdgpu_state, libcdgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}), f.h.ptr, s)); f)

# ---------------------------------------------------------------- SparseIterate <-> (nzval, nzval2ind, nnz)
# The triple crosses the ABI in caller-owned scratch copies; the iterate itself is only touched through
# `fill!` (coordinate_descent.jl:25 does the same) and `setindex!`, which appends in the order given — the
# insertion order pinned by test/atom_iterator.jl:13-28.  The library returns compacted lists (no explicit zeros).
function export_iterate(x::SparseIterate{Float64})
    p = length(x)
    nzv = zeros(Float64, p); ind = zeros(Int64, p); m = nnz(x)
    @inbounds for i = 1:m
        nzv[i] = x.nzval[i]; ind[i] = x.nzval2ind[i]
    end
    nzv, ind, m
end
function import_iterate!(x::SparseIterate{Float64}, nzv::Vector{Float64}, ind::Vector{Int64}, m::Integer)
    fill!(x, 0.0)
    @inbounds for i = 1:m
        x[ind[i]] = nzv[i]
    end
    x
end

# ---------------------------------------------------------------- coordinateDescent! (coordinate_descent.jl:7-39)
function coordinateDescent!(x::SparseIterate{Float64}, f::CoordinateDifferentiableFunction, g::ProxL1,
                            options::CDOptions=CDOptions())
    ProximalBase.numCoordinates(x) == numCoordinates(f) || throw(DimensionMismatch())
    weighted = !isa(g, ProxL1{typeof(g.λ0), Nothing})
    weighted && (length(g.λ) == numCoordinates(f) || throw(DimensionMismatch()))
    nzv, ind, m = export_iterate(x)
    cnt = Ref{Int64}(m)
    st = Ref{cdgpu_stats}()
    ω = weighted ? convert(Vector{Float64}, g.λ) : Float64[]
    GC.@preserve nzv ind ω check(ccall((:cdgpu_solve, libcdgpu), Cint,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Ref{cdgpu_options}, Ptr{Float64}, Ptr{Int64}, Ref{Int64}, Ref{cdgpu_stats}),
        f.h.ptr, g.λ0, weighted ? pointer(ω) : C_NULL, Ref(c_opts(options)), nzv, ind, cnt, st))
    import_iterate!(x, nzv, ind, cnt[])
    sync_state!(f)          # f.r / f.Ax must reflect the final iterate (LassoSolution aliases f.r, lasso.jl:37)
    x
end

# ---------------------------------------------------------------- front-ends (src/lasso.jl)
struct LassoSolution{T, S}
    x::SparseIterate{T}
    residuals::Vector{T}
    penalty::S
    σ::Union{T, Nothing}
end

function lasso(X::StridedMatrix{Float64}, y::StridedVector{Float64}, λ::Float64, options::CDOptions=CDOptions())
    x = SparseIterate(size(X, 2)); f = CDLeastSquaresLoss(y, X); g = ProxL1(λ)
    coordinateDescent!(x, f, g, options)
    LassoSolution{Float64, typeof(g)}(x, f.r, g, std(f.r))
end
function lasso(X::StridedMatrix{Float64}, y::StridedVector{Float64}, λ::Float64, ω::Array{Float64}, options::CDOptions=CDOptions())
    x = SparseIterate(size(X, 2)); f = CDLeastSquaresLoss(y, X); g = ProxL1(λ, ω)
    coordinateDescent!(x, f, g, options)
    LassoSolution{Float64, typeof(g)}(x, f.r, g, std(f.r))
end

_stdX!(out::Vector{Float64}, f::CoordinateDifferentiableFunction) =
    (GC.@preserve out check(ccall((:cdgpu_stdx, libcdgpu), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), f.h.ptr, C_NULL, out)); out)

function sqrtLasso(X::StridedMatrix{Float64}, y::StridedVector{Float64}, λ::Float64, options::CDOptions=CDOptions();
                   standardizeX=true)
    p = size(X, 2); x = SparseIterate(p); f = CDSqrtLassoLoss(y, X)
    g = standardizeX ? ProxL1(λ, _stdX!(Array{Float64}(undef, p), f)) : ProxL1(λ)    # intent of lasso.jl:72-75
    coordinateDescent!(x, f, g, options)
    LassoSolution{Float64, typeof(g)}(x, f.r, g, std(f.r))
end
function sqrtLasso(X::StridedMatrix{Float64}, y::StridedVector{Float64}, λ::Float64, ω::Array{Float64}, options::CDOptions=CDOptions())
    x = SparseIterate(size(X, 2)); f = CDSqrtLassoLoss(y, X); g = ProxL1(λ, ω)
    coordinateDescent!(x, f, g, options)
    LassoSolution{Float64, typeof(g)}(x, f.r, g, std(f.r))
end

# scaledLasso! (lasso.jl:107-144): the σ loop runs on the device, one ccall
function scaledLasso!(x::SparseIterate{Float64}, X::AbstractMatrix{Float64}, y::AbstractVector{Float64}, λ::Float64,
                      ω::AbstractVector{Float64}, options::IterLassoOptions=IterLassoOptions())
    f = CDLeastSquaresLoss(y, X)
    nzv, ind, m = export_iterate(x)
    cnt = Ref{Int64}(m); σ = Ref{Float64}(0.0); st = Ref{cdgpu_stats}()
    ωv = convert(Vector{Float64}, ω)
    GC.@preserve nzv ind ωv check(ccall((:cdgpu_scaled_solve, libcdgpu), Cint,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Ref{cdgpu_iter_options}, Ptr{Float64}, Ptr{Int64}, Ref{Int64}, Ref{Float64}, Ref{cdgpu_stats}),
        f.h.ptr, λ, ωv, Ref(c_opts(options)), nzv, ind, cnt, σ, st))
    import_iterate!(x, nzv, ind, cnt[]); sync_state!(f)
    g = ProxL1(λ * st[].sigma, ωv)
    LassoSolution{Float64, typeof(g)}(x, f.r, g, σ[])
end

struct LassoPath{T<:AbstractFloat}
    λpath::Vector{T}
    βpath::Vector{SparseIterate{T,1}}
end

# LassoPath (lasso.jl:229-260): the whole warm-started path is ONE ccall; β comes back as CSC
function LassoPath(X::StridedMatrix{Float64}, Y::StridedVector{Float64}, λpath::Vector{Float64}, options=CDOptions();
                   max_hat_s=Inf, standardizeX::Bool=true, loss::Union{Nothing,CoordinateDifferentiableFunction}=nothing)
    f = loss === nothing ? CDLeastSquaresLoss(Y, X) : loss      # pass a CDQuadraticLoss for the covariance-form path
    p = numCoordinates(f)
    stdX = standardizeX ? _stdX!(Array{Float64}(undef, p), f) : ones(p)
    m = length(λpath)
    cap = min(m * p, max(1 << 22, 4p))
    colptr = zeros(Int64, m + 1); rowval = zeros(Int64, cap); nzval = zeros(Float64, cap); done = Ref{Int64}(0)
    stats = Vector{cdgpu_stats}(undef, m)
    GC.@preserve λpath stdX colptr rowval nzval stats check(ccall((:cdgpu_path, libcdgpu), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ref{cdgpu_options}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ref{Int64}, Ptr{cdgpu_stats}),
        f.h.ptr, λpath, m, stdX, Ref(c_opts(options)), isinf(max_hat_s) ? -1 : Int64(max_hat_s), cap, colptr, rowval, nzval, done, stats))
    βpath = Vector{SparseIterate{Float64,1}}(undef, done[])
    for i = 1:done[]
        xi = SparseIterate(Float64, p)                               # the constructor of lasso.jl:244
        rng = colptr[i]+1:colptr[i+1]
        βpath[i] = import_iterate!(xi, nzval[rng], rowval[rng], length(rng))
    end
    done[] < m && resize!(λpath, done[])                          # lasso.jl:253-256
    LassoPath{Float64}(copy(λpath), βpath)
end

# ---------------------------------------------------------------- varying-coefficient lasso
abstract type SmoothingKernel{T} end
struct GaussianKernel{T} <: SmoothingKernel{T}; h::T; end
struct EpanechnikovKernel{T} <: SmoothingKernel{T}; h::T; end
createKernel(::Type{GaussianKernel{T}}, h::T) where {T<:AbstractFloat} = GaussianKernel{T}(h)
createKernel(::Type{EpanechnikovKernel{T}}, h::T) where {T<:AbstractFloat} = EpanechnikovKernel{T}(h)
evaluate(k::GaussianKernel{T}, x::T, y::T) where {T<:AbstractFloat} = exp(-(x-y)^2. / k.h) / k.h
function evaluate(k::EpanechnikovKernel{T}, x::T, y::T) where {T<:AbstractFloat}
    u = (x - y) / k.h
    abs(u) >= 1. ? zero(T) : 0.75 * (1. - u^2.) / k.h
end
kernel_kind(::GaussianKernel) = Cint(0)
kernel_kind(::EpanechnikovKernel) = Cint(1)

# locpolyl1 (varying_coefficient_lasso.jl:30-79): all grid points in one batched call; with refit=true the weighted
# normal equations on the selected groups (:71-76) are formed from the same moment blocks and solved on the device
function locpolyl1(X::Matrix{Float64}, z::Vector{Float64}, y::Vector{Float64}, zgrid::Vector{Float64}, degree::Int64,
                   kernel::SmoothingKernel{Float64}, λ0::Float64, refit::Bool, options::CDOptions=CDOptions(); device::Integer=0,
                   chain::Integer=1, sparse_output::Bool=true)
    # chain: grid points per warm-started run (cdgpu_vc_solve_chain).  1: every grid point from zero, all of them concurrently
    # (the default); length(zgrid): the reference's loop exactly (`beta` carried from grid point to grid point,
    # varying_coefficient_lasso.jl:56,68) as one sequential chain on the device; k: runs of k grid points.
    n, p = size(X); ep = p * (degree + 1); m = length(zgrid)
    if sparse_output
        # SparseMatrixCSC built from the device-side compaction (cdgpu_vc_solve_csc): only colptr and the stored entries
        # cross the ABI.  One retry when the first guess of the capacity is too small (colptr[m+1] then holds the need).
        cap = max(1, min(ep * m, 64 * m))
        while true
            cp = zeros(Int64, m + 1); rv = zeros(Int64, cap); nz = zeros(Float64, cap)
            cpR = zeros(Int64, m + 1); rvR = zeros(Int64, refit ? cap : 0); nzR = zeros(Float64, refit ? cap : 0)
            rc = GC.@preserve X z y zgrid cp rv nz cpR rvR nzR ccall((:cdgpu_vc_solve_csc, libcdgpu), Cint,
                (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Cint, Cint, Float64,
                 Float64, Ref{cdgpu_options}, Int64, Cint, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64},
                 Ptr{Float64}, Ptr{Cvoid}),
                X, n, p, n, z, y, zgrid, m, 0, m, degree, kernel_kind(kernel), kernel.h, λ0, Ref(c_opts(options)), chain, device, cap,
                cp, rv, nz, refit ? pointer(cpR) : Ptr{Int64}(C_NULL), refit ? pointer(rvR) : Ptr{Int64}(C_NULL),
                refit ? pointer(nzR) : Ptr{Float64}(C_NULL), C_NULL)
            need = max(cp[m+1], cpR[m+1])
            if rc == 7 && need > cap   # CDGPU_ECAP
                cap = need
                continue
            end
            check(rc)
            out = SparseMatrixCSC(ep, m, cp .+ 1, rv[1:cp[m+1]], nz[1:cp[m+1]])
            outR = refit ? SparseMatrixCSC(ep, m, cpR .+ 1, rvR[1:cpR[m+1]], nzR[1:cpR[m+1]]) : spzeros(Float64, ep, m)
            return out, outR
        end
    end
    out = zeros(Float64, ep, m)
    if chain != 1
        outR = refit ? zeros(Float64, ep, m) : nothing
        GC.@preserve X z y zgrid out outR check(ccall((:cdgpu_vc_solve_chain, libcdgpu), Cint,
            (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Cint, Cint, Float64, Float64,
             Ref{cdgpu_options}, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
            X, n, p, n, z, y, zgrid, m, 0, m, degree, kernel_kind(kernel), kernel.h, λ0, Ref(c_opts(options)), chain, device, out,
            refit ? pointer(outR) : Ptr{Float64}(C_NULL), C_NULL))
        return sparse(out), (refit ? sparse(outR) : spzeros(Float64, ep, m))
    end
    if !refit
        GC.@preserve X z y zgrid out check(ccall((:cdgpu_vc_solve, libcdgpu), Cint,
            (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Cint, Cint, Float64, Float64,
             Ref{cdgpu_options}, Cint, Ptr{Float64}, Ptr{Cvoid}),
            X, n, p, n, z, y, zgrid, m, 0, m, degree, kernel_kind(kernel), kernel.h, λ0, Ref(c_opts(options)), device, out, C_NULL))
        return sparse(out), spzeros(Float64, ep, m)
    end
    outR = zeros(Float64, ep, m)
    GC.@preserve X z y zgrid out outR check(ccall((:cdgpu_vc_solve_refit, libcdgpu), Cint,
        (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Cint, Cint, Float64, Float64,
         Ref{cdgpu_options}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Cvoid}),
        X, n, p, n, z, y, zgrid, m, 0, m, degree, kernel_kind(kernel), kernel.h, λ0, Ref(c_opts(options)), device, out, outR, C_NULL))
    sparse(out), sparse(outR)
end

# refitLassoPath (lasso.jl:208-225): least squares on each distinct support of the path, on the device
function refitLassoPath(path::LassoPath{Float64}, X::Matrix{Float64}, Y::Vector{Float64}; device::Integer=0)
    f = CDLeastSquaresLoss(Y, X; device=device)
    out = Dict{Vector{Int64},Vector{Float64}}()
    for β in path.βpath
        S = sort(β.nzval2ind[1:β.nnz])
        haskey(out, S) && continue
        coef = zeros(Float64, length(S))
        isempty(S) || GC.@preserve S coef check(ccall((:cdgpu_refit, libcdgpu), Cint,
            (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float64}), f.h.ptr, S, length(S), coef))
        out[S] = coef
    end
    out
end

# lvocv_locpolyl1 (varying_coefficient_lasso.jl:81-137): every (bandwidth, left-out observation) pair is one local
# scaled-lasso problem of a single batched call; MSE[indH] is the sum of the squared prediction errors over i
function lvocv_locpolyl1(X::Matrix{Float64}, z::Vector{Float64}, y::Vector{Float64}, degree::Int64, hArr::Vector{Float64},
                         kernelType::Type{<:SmoothingKernel}, λ0::Float64, options::CDOptions=CDOptions(); device::Integer=0)
    n, p = size(X); numH = length(hArr)
    sqerr = zeros(Float64, n * numH)
    KT = kernelType isa UnionAll ? kernelType{Float64} : kernelType   # GaussianKernel or GaussianKernel{Float64}
    kind = kernel_kind(createKernel(KT, 1.0))
    GC.@preserve X z y hArr sqerr check(ccall((:cdgpu_vc_lvocv, libcdgpu), Cint,
        (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Int64, Cint, Float64,
         Ref{cdgpu_options}, Int64, Int64, Cint, Ptr{Float64}, Ptr{Cvoid}),
        X, n, p, n, z, y, degree, hArr, numH, kind, λ0, Ref(c_opts(options)), 0, n * numH, device, sqerr, C_NULL))
    vec(sum(reshape(sqerr, n, numH), dims=1))
end

end # module
