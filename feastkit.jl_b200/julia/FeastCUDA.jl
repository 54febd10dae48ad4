# FeastCUDA.jl -- the Julia host shim: FeastKit.jl's own method names, bodies replaced by `ccall`s into libfeastcuda.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: `julia` is absent from the build image and from the GPU box.  The file is kept
# mechanical (one table of names -> one generic driver) so that its correctness follows from the C-ABI tests, which drive
# the very same entry points through ctypes (feastkit.jl_b200/feastcuda/, tests/test_gpu_*.py).
#
# Usage inside FeastKit.jl (see INTEGRATION.md):   include("FeastCUDA.jl"); using .FeastCUDA
# file:line citations are relative to FeastKit.jl/src.
module FeastCUDA

using LinearAlgebra, SparseArrays

const libfeastcuda = get(ENV, "FEASTCUDA_LIB", joinpath(@__DIR__, "..", "lib", "libfeastcuda.so"))

# ---- mirrors of include/feastcuda.h -------------------------------------------------------------------------------
const FEASTCUDA_A, FEASTCUDA_B = Cint(0), Cint(1)
const FEASTCUDA_CSC = Cint(1)
const FEASTCUDA_SYM, FEASTCUDA_HERM, FEASTCUDA_GEN = Cint(0), Cint(1), Cint(2)
const SOLVER_DIRECT, SOLVER_BICGSTAB, SOLVER_MSLANCZOS = Cint(0), Cint(1), Cint(2)
const FILTER_REFERENCE, FILTER_TRUE = Cint(0), Cint(1)

struct SolverOpts              # feastcuda_solver_opts (96 bytes, C layout)
    solver::Cint
    tol::Cdouble
    maxiter::Cint
    restart::Cint
    inner_rel::Cdouble
    ritz_guess::Cint
    filter::Cint
    shard::Cint
    check_every::Cint
    q0_real::Cint
    x_real::Cint
    inner_rel0::Cdouble
    maxiter0::Cint
    keep_going::Cint
    adaptive::Cint
    mixed::Cint
    eps_floor::Cdouble
    b_delta::Cdouble           # generalized Lanczos path: accuracy of the Chebyshev solves with B (0 -> 1e-4)
end

# FeastResult{T,VT} (core/feast_types.jl:85-98) -- same fields, arrays trimmed to M
struct FeastResult{T<:Real,VT}
    lambda::Vector{T}
    q::Matrix{VT}
    M::Int
    res::Vector{T}
    info::Int
    epsout::T
    loop::Int
end

mutable struct Handle
    ptr::Ptr{Cvoid}
end

function check(rc::Cint, h::Ptr{Cvoid}=C_NULL)
    rc == 0 && return
    msg = unsafe_string(ccall((:feastcuda_last_error, libfeastcuda), Cstring, (Ptr{Cvoid},), h))
    rc == 1 ? throw(ArgumentError(msg)) : error("libfeastcuda status $rc: $msg")
end

const _handle = Ref{Union{Nothing,Handle}}(nothing)
function handle()
    if _handle[] === nothing
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:feastcuda_create, libfeastcuda), Cint, (Ptr{Ptr{Cvoid}}, Cint), p, parse(Cint, get(ENV, "LOCAL_RANK", "0"))))
        h = Handle(p[])
        finalizer(x -> ccall((:feastcuda_destroy, libfeastcuda), Cint, (Ptr{Cvoid},), x.ptr), h)
        _handle[] = h
    end
    return _handle[].ptr
end

# ---- parameters / contours (core/feast_parameters.jl:7-18,41-386; core/feast_tools.jl:212-371) --------------------
function feastinit!(fpm::Vector{Int})
    length(fpm) >= 64 || throw(ArgumentError("fpm array must have at least 64 elements"))
    check(ccall((:feastcuda_feastinit, libfeastcuda), Cint, (Ptr{Int64},), fpm))
    return fpm
end
feastinit() = feastinit!(zeros(Int, 64))
function feastdefault!(fpm::Vector{Int})
    ccall((:feastcuda_feastdefault, libfeastcuda), Cint, (Ptr{Int64},), fpm) == 0 ||
        throw(ArgumentError("Invalid fpm parameter"))
    return fpm
end
function feast_contour(Emin::Real, Emax::Real, fpm::Vector{Int})
    feastdefault!(fpm)
    ne = fpm[2]
    Zne = Vector{ComplexF64}(undef, ne); Wne = Vector{ComplexF64}(undef, ne)
    check(ccall((:feastcuda_contour, libfeastcuda), Cint, (Cdouble, Cdouble, Ptr{Int64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                Emin, Emax, fpm, Zne, Wne))
    return (Zne=Zne, Wne=Wne)
end
function feast_gcontour(Emid::Number, r::Real, fpm::Vector{Int})
    feastdefault!(fpm)
    ne = fpm[8]
    Zne = Vector{ComplexF64}(undef, ne); Wne = Vector{ComplexF64}(undef, ne)
    z = ComplexF64(Emid)
    check(ccall((:feastcuda_gcontour, libfeastcuda), Cint, (Cdouble, Cdouble, Cdouble, Ptr{Int64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                real(z), imag(z), r, fpm, Zne, Wne))
    return (Zne=Zne, Wne=Wne)
end

# ---- operators: SparseMatrixCSC{Tv,Int} is CSC, 1-based Int64 -> passed as is (fmt = CSC, index_base = 1) ----------
function set_sparse!(h, which::Cint, A::SparseMatrixCSC{Float64,Int}, structure::Cint)
    GC.@preserve A check(ccall((:feastcuda_set_csr_d, libfeastcuda), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Cint, Cint, Cint),
        h, which, size(A, 1), nnz(A), A.colptr, A.rowval, A.nzval, 1, FEASTCUDA_CSC, structure), h)
end
function set_sparse!(h, which::Cint, A::SparseMatrixCSC{ComplexF64,Int}, structure::Cint)
    GC.@preserve A check(ccall((:feastcuda_set_csr_z, libfeastcuda), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}, Cint, Cint, Cint),
        h, which, size(A, 1), nnz(A), A.colptr, A.rowval, A.nzval, 1, FEASTCUDA_CSC, structure), h)
end
function set_dense!(h, which::Cint, A::Matrix{Float64}, structure::Cint)
    GC.@preserve A check(ccall((:feastcuda_set_dense_d, libfeastcuda), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Int64, Cint),
        h, which, size(A, 1), A, stride(A, 2), structure), h)
end
function set_dense!(h, which::Cint, A::Matrix{ComplexF64}, structure::Cint)
    GC.@preserve A check(ccall((:feastcuda_set_dense_z, libfeastcuda), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{ComplexF64}, Int64, Cint),
        h, which, size(A, 1), A, stride(A, 2), structure), h)
end
function set_band!(h, which::Cint, AB::Matrix{T}, k::Int, structure::Cint) where {T<:Union{Float64,ComplexF64}}
    sym = T === Float64 ? :feastcuda_set_band_d : :feastcuda_set_band_z
    GC.@preserve AB check(ccall((sym, libfeastcuda), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{T}, Int64, Cint),
        h, which, size(AB, 2), k, AB, size(AB, 1), structure), h)
end

# ---- the generic interval driver: every Hermitian-family name below lands here -------------------------------------
# replaces _feast_dense_complex_hermitian (dense/feast_dense.jl:78-351), _feast_sparse_hermitian
# (sparse/feast_sparse.jl:246-499) and _feast_banded_complex_hermitian (banded/feast_banded.jl:561-823)
function _solve_interval(setA!, setB!, N::Int, Emin, Emax, M0::Int, fpm::Vector{Int}, ::Type{VT};
                         Zne=nothing, Wne=nothing, solver::Symbol=:direct, solver_tol::Real=0.0, solver_maxiter::Int=500,
                         solver_restart::Int=30, sparse::Bool=false, eps_floor::Float64=0.0, mixed::Int=(sparse ? 2 : 0)) where {VT}
    # mixed: 0 = FP64 Krylov vectors, 1 = FP32 vectors, 2 = follow fpm[42] (the reference's default fpm[42] = 1, core/feast_parameters.jl:316-319,
    # is honoured for sparse problems: real symmetric standard pencils run FP32 Lanczos vectors inside the FP64 refinement loop)
    feastdefault!(fpm)
    # check_feast_srci_input (core/feast_aux.jl:369-399): thrown before any device call, exactly as the reference does
    N > 0 || throw(ArgumentError("Matrix size N must be positive"))
    0 < M0 <= N || throw(ArgumentError("Number of eigenvalues M0 must be between 1 and N"))
    Emin < Emax || throw(ArgumentError("Search interval [Emin, Emax] must be valid"))
    solver_choice = solver == :iterative ? :gmres : solver
    solver_choice in (:direct, :gmres, :bicgstab, :mslanczos) ||
        throw(ArgumentError("Unsupported solver option '$solver'. Use :direct, :gmres, or :iterative."))
    h = handle()
    setA!(h)
    setB! === nothing ? check(ccall((:feastcuda_clear_b, libfeastcuda), Cint, (Ptr{Cvoid},), h), h) : setB!(h)
    if Zne === nothing
        c = feast_contour(Emin, Emax, fpm); Zne, Wne = c.Zne, c.Wne
    end
    real_result = VT <: Real
    eng_solver = sparse ? (solver_choice == :bicgstab ? SOLVER_BICGSTAB : SOLVER_MSLANCZOS) :
                          (solver_choice == :direct ? SOLVER_DIRECT : SOLVER_BICGSTAB)
    opts = Ref(SolverOpts(eng_solver, solver_tol, solver_maxiter, solver_restart == 30 ? 3 : solver_restart,
                          sparse ? 1e-3 : 0.0, sparse ? 1 : 0, FILTER_TRUE, 0, 16, 0, real_result ? 1 : 0, 0.0, 0, 0, sparse ? 1 : 0,
                          Cint(mixed) #= 1: FP32 Lanczos vectors, 2: follow fpm[42] =#, eps_floor, 0.0))
    lambda = zeros(Float64, M0); res = zeros(Float64, M0); X = zeros(VT, N, M0)
    M = Ref{Int64}(0); info = Ref{Int64}(0); loop = Ref{Int64}(0); epsout = Ref{Float64}(0.0)
    GC.@preserve fpm Zne Wne lambda res X begin
        check(ccall((:feastcuda_solve_interval, libfeastcuda), Cint,
            (Ptr{Cvoid}, Cdouble, Cdouble, Int64, Ptr{Int64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Int64, Ptr{Cvoid}, Ref{SolverOpts},
             Ptr{Float64}, Ptr{VT}, Ptr{Float64}, Ref{Int64}, Ref{Int64}, Ref{Float64}, Ref{Int64}),
            h, Emin, Emax, M0, fpm, Zne, Wne, length(Zne), C_NULL, opts, lambda, X, res, M, info, epsout, loop), h)
    end
    m = Int(M[])
    return FeastResult{Float64,VT}(lambda[1:m], X[:, 1:m], m, res[1:m], Int(info[]), epsout[], Int(loop[]))
end

# ---- reference names (one line each; `x` variants add the contour) --------------------------------------------------
# sparse/feast_sparse.jl:1516-1529, 713-731, 759-788, 815-831
feast_scsrev!(A::SparseMatrixCSC{<:Real}, Emin, Emax, M0, fpm; kw...) =
    _solve_interval(h -> set_sparse!(h, FEASTCUDA_A, SparseMatrixCSC{Float64,Int}(A), FEASTCUDA_SYM), nothing, size(A, 1), Emin, Emax, M0, fpm, Float64; sparse=true, kw...)
feast_scsrgv!(A::SparseMatrixCSC{<:Real}, B::SparseMatrixCSC{<:Real}, Emin, Emax, M0, fpm; kw...) =
    _solve_interval(h -> set_sparse!(h, FEASTCUDA_A, SparseMatrixCSC{Float64,Int}(A), FEASTCUDA_SYM),
                    h -> set_sparse!(h, FEASTCUDA_B, SparseMatrixCSC{Float64,Int}(B), FEASTCUDA_SYM), size(A, 1), Emin, Emax, M0, fpm, Float64; sparse=true, kw...)
function feast_hcsrev!(A::SparseMatrixCSC{<:Complex}, Emin, Emax, M0, fpm; kw...)
    ishermitian(A) || throw(ArgumentError("Matrix A must be Hermitian"))
    _solve_interval(h -> set_sparse!(h, FEASTCUDA_A, SparseMatrixCSC{ComplexF64,Int}(A), FEASTCUDA_HERM), nothing, size(A, 1), Emin, Emax, M0, fpm, ComplexF64; sparse=true, kw...)
end
function feast_hcsrgv!(A::SparseMatrixCSC{<:Complex}, B::SparseMatrixCSC{<:Complex}, Emin, Emax, M0, fpm; kw...)
    ishermitian(A) || throw(ArgumentError("Matrix A must be Hermitian"))
    ishermitian(B) || throw(ArgumentError("Matrix B must be Hermitian"))
    _solve_interval(h -> set_sparse!(h, FEASTCUDA_A, SparseMatrixCSC{ComplexF64,Int}(A), FEASTCUDA_HERM),
                    h -> set_sparse!(h, FEASTCUDA_B, SparseMatrixCSC{ComplexF64,Int}(B), FEASTCUDA_HERM), size(A, 1), Emin, Emax, M0, fpm, ComplexF64; sparse=true, kw...)
end
# dense/feast_dense.jl:776-797, 356-370, 390-400, 799-810
feast_syev!(A::Matrix{<:Real}, Emin, Emax, M0, fpm; kw...) =
    _solve_interval(h -> set_dense!(h, FEASTCUDA_A, Matrix{Float64}(A), FEASTCUDA_SYM), nothing, size(A, 1), Emin, Emax, M0, fpm, Float64; kw...)
feast_sygv!(A::Matrix{<:Real}, B::Matrix{<:Real}, Emin, Emax, M0, fpm; kw...) =
    _solve_interval(h -> set_dense!(h, FEASTCUDA_A, Matrix{Float64}(A), FEASTCUDA_SYM), h -> set_dense!(h, FEASTCUDA_B, Matrix{Float64}(B), FEASTCUDA_SYM),
                    size(A, 1), Emin, Emax, M0, fpm, Float64; kw...)
feast_heev!(A::Matrix{<:Complex}, Emin, Emax, M0, fpm; kw...) =
    _solve_interval(h -> set_dense!(h, FEASTCUDA_A, Matrix{ComplexF64}(A), FEASTCUDA_HERM), nothing, size(A, 1), Emin, Emax, M0, fpm, ComplexF64; kw...)
feast_hegv!(A::Matrix{<:Complex}, B::Matrix{<:Complex}, Emin, Emax, M0, fpm; kw...) =
    _solve_interval(h -> set_dense!(h, FEASTCUDA_A, Matrix{ComplexF64}(A), FEASTCUDA_HERM), h -> set_dense!(h, FEASTCUDA_B, Matrix{ComplexF64}(B), FEASTCUDA_HERM),
                    size(A, 1), Emin, Emax, M0, fpm, ComplexF64; kw...)
# banded/feast_banded.jl:1410-1420, 9-186, 326-383, 385-421
feast_sbev!(A::Matrix{<:Real}, kla::Int, Emin, Emax, M0, fpm; kw...) =
    _solve_interval(h -> set_band!(h, FEASTCUDA_A, Matrix{Float64}(A), kla, FEASTCUDA_SYM), nothing, size(A, 2), Emin, Emax, M0, fpm, Float64; kw...)
feast_sbgv!(A::Matrix{<:Real}, B::Matrix{<:Real}, kla::Int, klb::Int, Emin, Emax, M0, fpm; kw...) =
    _solve_interval(h -> set_band!(h, FEASTCUDA_A, Matrix{Float64}(A), kla, FEASTCUDA_SYM), h -> set_band!(h, FEASTCUDA_B, Matrix{Float64}(B), klb, FEASTCUDA_SYM),
                    size(A, 2), Emin, Emax, M0, fpm, Float64; kw...)
feast_hbev!(A::Matrix{<:Complex}, kla::Int, Emin, Emax, M0, fpm; kw...) =
    _solve_interval(h -> set_band!(h, FEASTCUDA_A, Matrix{ComplexF64}(A), kla, FEASTCUDA_HERM), nothing, size(A, 2), Emin, Emax, M0, fpm, ComplexF64; kw...)
feast_hbgv!(A::Matrix{<:Complex}, B::Matrix{<:Complex}, kla::Int, klb::Int, Emin, Emax, M0, fpm; kw...) =
    _solve_interval(h -> set_band!(h, FEASTCUDA_A, Matrix{ComplexF64}(A), kla, FEASTCUDA_HERM), h -> set_band!(h, FEASTCUDA_B, Matrix{ComplexF64}(B), klb, FEASTCUDA_HERM),
                    size(A, 2), Emin, Emax, M0, fpm, ComplexF64; kw...)

# custom-contour `x` variants (sparse/feast_sparse.jl:751-757 etc.): same call with the caller's nodes
for f in (:feast_scsrev, :feast_hcsrev, :feast_syev, :feast_heev)
    @eval $(Symbol(f, "x!"))(A, Emin, Emax, M0, fpm, Zne, Wne; kw...) = $(Symbol(f, "!"))(A, Emin, Emax, M0, fpm; Zne=Zne, Wne=Wne, kw...)
end
for f in (:feast_scsrgv, :feast_hcsrgv, :feast_sygv, :feast_hegv)
    @eval $(Symbol(f, "x!"))(A, B, Emin, Emax, M0, fpm, Zne, Wne; kw...) = $(Symbol(f, "!"))(A, B, Emin, Emax, M0, fpm; Zne=Zne, Wne=Wne, kw...)
end
for f in (:feast_sbev, :feast_hbev)     # banded/feast_banded.jl:1422-1440, 326-383
    @eval $(Symbol(f, "x!"))(A, kla, Emin, Emax, M0, fpm, Zne, Wne; kw...) = $(Symbol(f, "!"))(A, kla, Emin, Emax, M0, fpm; Zne=Zne, Wne=Wne, kw...)
end
for f in (:feast_sbgv, :feast_hbgv)
    @eval $(Symbol(f, "x!"))(A, B, kla, klb, Emin, Emax, M0, fpm, Zne, Wne; kw...) = $(Symbol(f, "!"))(A, B, kla, klb, Emin, Emax, M0, fpm; Zne=Zne, Wne=Wne, kw...)
end

# precision / parallel alias families (interfaces/feast_precision_aliases.jl:10-117,163-423,497-771): pure forwarding.
# With `comm`/`use_threads` the reference picks an MPI/threads backend; here every rank of the job (one process per GPU,
# feastcuda_nccl_init) takes part and the keyword is accepted and ignored.
for (alias, target) in ((:dfeast_scsrev!, :feast_scsrev!), (:dfeast_scsrgv!, :feast_scsrgv!), (:zfeast_hcsrev!, :feast_hcsrev!),
                        (:zfeast_hcsrgv!, :feast_hcsrgv!), (:dfeast_syev!, :feast_syev!), (:dfeast_sygv!, :feast_sygv!),
                        (:zfeast_heev!, :feast_heev!), (:zfeast_hegv!, :feast_hegv!), (:dfeast_sbev!, :feast_sbev!),
                        (:dfeast_sbgv!, :feast_sbgv!), (:zfeast_hbev!, :feast_hbev!), (:zfeast_hbgv!, :feast_hbgv!))
    @eval $alias(args...; comm=nothing, use_threads=nothing, kw...) = $target(args...; kw...)
    @eval $(Symbol("p", alias))(args...; comm=nothing, use_threads=nothing, kw...) = $target(args...; kw...)
end

# Float32 / ComplexF32 names (interfaces/feast_precision_aliases.jl:10-117,163-423): inputs are widened, the engine computes in
# Float64 and stops at the reference's single-precision tolerance max(10^-fpm[3], sqrt(eps(Float32))) (core/feast_parameters.jl:398-405),
# results are narrowed -- FeastResult{Float32,...} like the reference returns.
const _EPS32 = Float64(sqrt(eps(Float32)))
_narrow(r::FeastResult{Float64,Float64}) = FeastResult{Float32,Float32}(Float32.(r.lambda), Float32.(r.q), r.M, Float32.(r.res), r.info, Float32(r.epsout), r.loop)
_narrow(r::FeastResult{Float64,ComplexF64}) = FeastResult{Float32,ComplexF32}(Float32.(r.lambda), ComplexF32.(r.q), r.M, Float32.(r.res), r.info, Float32(r.epsout), r.loop)
for (alias, target) in ((:sfeast_scsrev!, :feast_scsrev!), (:sfeast_scsrgv!, :feast_scsrgv!), (:cfeast_hcsrev!, :feast_hcsrev!),
                        (:cfeast_hcsrgv!, :feast_hcsrgv!), (:sfeast_syev!, :feast_syev!), (:sfeast_sygv!, :feast_sygv!),
                        (:cfeast_heev!, :feast_heev!), (:cfeast_hegv!, :feast_hegv!), (:sfeast_sbev!, :feast_sbev!),
                        (:sfeast_sbgv!, :feast_sbgv!), (:cfeast_hbev!, :feast_hbev!), (:cfeast_hbgv!, :feast_hbgv!))
    mx = target === :feast_scsrev! ? 1 : 0       # single-precision sparse names: FP32 Krylov vectors inside the FP64 refinement loop
    @eval $alias(args...; comm=nothing, use_threads=nothing, kw...) = _narrow($target(args...; eps_floor=_EPS32, mixed=$mx, kw...))
    @eval $(Symbol("p", alias))(args...; comm=nothing, use_threads=nothing, kw...) = _narrow($target(args...; eps_floor=_EPS32, mixed=$mx, kw...))
end

# ---- matrix-free: feast_matvec(A_mul!, B_mul!, N, interval) interfaces/feast_interfaces.jl:465-481 -> feast_sparse_matvec!
# sparse/feast_sparse.jl:1284-1471.  The operator is a DEVICE callback of C type feastcuda_apply_fn
#   (ctx::Ptr{Cvoid}, n::Int64, ncols::Int64, X::Ptr{Float64}, ldx::Int64, Y::Ptr{Float64}, ldy::Int64, stream::Ptr{Cvoid}) -> Cvoid
# that enqueues Y = A*X for row-major device blocks on `stream`: either a symbol of the user's own CUDA library
# (`cglobal((:my_apply, "libmyops.so"))`) or an `@cfunction` whose body launches the user's kernels.  B must be the identity.
function feast_matvec(apply::Ptr{Cvoid}, ctx::Ptr{Cvoid}, N::Int, interval::Tuple{<:Real,<:Real}; M0::Int=10, fpm=nothing, kw...)
    fpm = fpm === nothing ? feastinit() : fpm
    setA! = h -> check(ccall((:feastcuda_set_matfree_d, libfeastcuda), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}), h, N, apply, ctx), h)
    return _solve_interval(setA!, nothing, N, Float64(interval[1]), Float64(interval[2]), M0, fpm, Float64; sparse=true, kw...)
end

# ---- general (non-Hermitian) family: feast_gcsrgv!/gcsrev! sparse/feast_sparse.jl:873-1006,1531-1545;
# feast_gegv!/geev! dense/feast_dense.jl:402-593,812-823; feast_gbgv!/gbev! banded/feast_banded.jl:1548-1600 ----------------
struct FeastGeneralResult{T<:Real}            # core/feast_types.jl:100-118
    lambda::Vector{Complex{T}}
    q::Matrix{Complex{T}}
    M::Int
    res::Vector{T}
    info::Int
    epsout::T
    loop::Int
end

function _solve_contour(setA!, setB!, N::Int, Emid::Number, r::Real, M0::Int, fpm::Vector{Int};
                        Zne=nothing, Wne=nothing, solver::Symbol=:direct, solver_tol::Real=0.0, solver_maxiter::Int=500,
                        solver_restart::Int=30, sparse::Bool=false)
    feastdefault!(fpm)
    # check_feast_grci_input (core/feast_aux.jl:401-425)
    N > 0 || throw(ArgumentError("Matrix size N must be positive"))
    0 < M0 <= N || throw(ArgumentError("Number of eigenvalues M0 must be between 1 and N"))
    r > 0 || throw(ArgumentError("Search radius r must be positive"))
    h = handle()
    setA!(h)
    setB! === nothing ? check(ccall((:feastcuda_clear_b, libfeastcuda), Cint, (Ptr{Cvoid},), h), h) : setB!(h)
    if Zne === nothing
        c = feast_gcontour(Emid, r, fpm); Zne, Wne = c.Zne, c.Wne
    end
    z = ComplexF64(Emid)
    opts = Ref(SolverOpts(sparse ? SOLVER_BICGSTAB : SOLVER_DIRECT, solver_tol, solver_maxiter, solver_restart == 30 ? 3 : solver_restart,
                          0.0, 0, FILTER_REFERENCE, 0, 16, 0, 0, 0.0, 0, 0, 0, Cint(0), 0.0, 0.0))
    lambda = zeros(ComplexF64, M0); res = zeros(Float64, M0); X = zeros(ComplexF64, N, M0)
    M = Ref{Int64}(0); info = Ref{Int64}(0); loop = Ref{Int64}(0); epsout = Ref{Float64}(0.0)
    GC.@preserve fpm Zne Wne lambda res X begin
        check(ccall((:feastcuda_solve_contour, libfeastcuda), Cint,
            (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Int64, Ptr{Int64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Int64, Ptr{Cvoid}, Ref{SolverOpts},
             Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Float64}, Ref{Int64}, Ref{Int64}, Ref{Float64}, Ref{Int64}),
            h, real(z), imag(z), r, M0, fpm, Zne, Wne, length(Zne), C_NULL, opts, lambda, X, res, M, info, epsout, loop), h)
    end
    m = Int(M[])
    return FeastGeneralResult{Float64}(lambda[1:m], X[:, 1:m], m, res[1:m], Int(info[]), epsout[], Int(loop[]))
end

feast_gcsrev!(A::SparseMatrixCSC, Emid, r, M0, fpm; kw...) =
    _solve_contour(h -> set_sparse!(h, FEASTCUDA_A, SparseMatrixCSC{ComplexF64,Int}(A), FEASTCUDA_GEN), nothing, size(A, 1), Emid, r, M0, fpm; sparse=true, kw...)
feast_gcsrgv!(A::SparseMatrixCSC, B::SparseMatrixCSC, Emid, r, M0, fpm; kw...) =
    _solve_contour(h -> set_sparse!(h, FEASTCUDA_A, SparseMatrixCSC{ComplexF64,Int}(A), FEASTCUDA_GEN),
                   h -> set_sparse!(h, FEASTCUDA_B, SparseMatrixCSC{ComplexF64,Int}(B), FEASTCUDA_GEN), size(A, 1), Emid, r, M0, fpm; sparse=true, kw...)
feast_geev!(A::Matrix, Emid, r, M0, fpm; kw...) =
    _solve_contour(h -> set_dense!(h, FEASTCUDA_A, Matrix{ComplexF64}(A), FEASTCUDA_GEN), nothing, size(A, 1), Emid, r, M0, fpm; kw...)
feast_gegv!(A::Matrix, B::Matrix, Emid, r, M0, fpm; kw...) =
    _solve_contour(h -> set_dense!(h, FEASTCUDA_A, Matrix{ComplexF64}(A), FEASTCUDA_GEN), h -> set_dense!(h, FEASTCUDA_B, Matrix{ComplexF64}(B), FEASTCUDA_GEN),
                   size(A, 1), Emid, r, M0, fpm; kw...)
for f in (:feast_gcsrev, :feast_geev)
    @eval $(Symbol(f, "x!"))(A, Emid, r, M0, fpm, Zne, Wne; kw...) = $(Symbol(f, "!"))(A, Emid, r, M0, fpm; Zne=Zne, Wne=Wne, kw...)
end
for f in (:feast_gcsrgv, :feast_gegv)
    @eval $(Symbol(f, "x!"))(A, B, Emid, r, M0, fpm, Zne, Wne; kw...) = $(Symbol(f, "!"))(A, B, Emid, r, M0, fpm; Zne=Zne, Wne=Wne, kw...)
end
function feast_general(A::AbstractMatrix, center::Number, radius::Real; M0::Int=10, fpm=nothing, kw...)   # interfaces/feast_interfaces.jl:274-379
    size(A, 1) == size(A, 2) || throw(ArgumentError("Matrix must be square"))
    fpm = fpm === nothing ? feastinit() : fpm
    M0 = min(M0, size(A, 1))
    return issparse(A) ? feast_gcsrev!(A, center, radius, M0, fpm; kw...) : feast_geev!(Matrix(A), center, radius, M0, fpm; kw...)
end

# general band storage (2k+1) x n, diagonal in row k+1: feast_gbgv!/gbev! banded/feast_banded.jl:1548-1600
feast_gbgv!(A::Matrix, B::Matrix, ka::Int, kb::Int, Emid, r, M0, fpm; kw...) =
    _solve_contour(h -> set_band!(h, FEASTCUDA_A, Matrix{ComplexF64}(A), ka, FEASTCUDA_GEN), h -> set_band!(h, FEASTCUDA_B, Matrix{ComplexF64}(B), kb, FEASTCUDA_GEN),
                   size(A, 2), Emid, r, M0, fpm; kw...)
feast_gbev!(A::Matrix, ka::Int, Emid, r, M0, fpm; kw...) =
    _solve_contour(h -> set_band!(h, FEASTCUDA_A, Matrix{ComplexF64}(A), ka, FEASTCUDA_GEN), nothing, size(A, 2), Emid, r, M0, fpm; kw...)
feast_gbgvx!(A, B, ka, kb, Emid, r, M0, fpm, Zne, Wne; kw...) = feast_gbgv!(A, B, ka, kb, Emid, r, M0, fpm; Zne=Zne, Wne=Wne, kw...)
feast_gbevx!(A, ka, Emid, r, M0, fpm, Zne, Wne; kw...) = feast_gbev!(A, ka, Emid, r, M0, fpm; Zne=Zne, Wne=Wne, kw...)

# ---- complex-symmetric names (A == transpose(A)): dense/feast_dense.jl:1261-1286, sparse/feast_sparse.jl:1038-1095,
# banded/feast_banded.jl:1469-1525.  Validation as in the reference (core/feast_aux.jl:665-668), then the general-contour
# engine (orthonormalised one-sided Rayleigh-Ritz instead of the transpose-bilinear CS-RR projection: same eigenvalues inside).
_check_complex_symmetric(A) = issymmetric(A) || throw(ArgumentError("Matrix must be complex-symmetric (equal to its transpose)."))
function feast_gegv_complex_sym!(A::Matrix{<:Complex}, B::Matrix{<:Complex}, Emid, r, M0, fpm; kw...)
    _check_complex_symmetric(A); _check_complex_symmetric(B)
    return feast_gegv!(A, B, Emid, r, M0, fpm; kw...)
end
function feast_geev_complex_sym!(A::Matrix{<:Complex}, Emid, r, M0, fpm; kw...)
    _check_complex_symmetric(A)
    return feast_geev!(A, Emid, r, M0, fpm; kw...)
end
function feast_scsrgv_complex!(A::SparseMatrixCSC{<:Complex}, B::SparseMatrixCSC{<:Complex}, Emid, r, M0, fpm; kw...)
    _check_complex_symmetric(A); _check_complex_symmetric(B)
    return feast_gcsrgv!(A, B, Emid, r, M0, fpm; kw...)
end
function feast_scsrev_complex!(A::SparseMatrixCSC{<:Complex}, Emid, r, M0, fpm; kw...)
    _check_complex_symmetric(A)
    return feast_gcsrev!(A, Emid, r, M0, fpm; kw...)
end
# symmetric upper band (k+1) x n (diagonal in the last row) -> general band (2k+1) x n, no conjugation
function _symmetric_band_to_general(AB::Matrix{<:Complex}, k::Int)
    n = size(AB, 2)
    GB = zeros(ComplexF64, 2k + 1, n)
    for d in 0:k
        GB[k + 1 - d, d+1:n] .= AB[k + 1 - d, d+1:n]            # A[j-d, j]
        d > 0 && (GB[k + 1 + d, 1:n-d] .= AB[k + 1 - d, d+1:n])  # mirrored entry A[j, j-d]
    end
    return GB
end
feast_sbgv_complex!(A::Matrix{<:Complex}, B::Matrix{<:Complex}, ka::Int, kb::Int, Emid, r, M0, fpm; kw...) =
    feast_gbgv!(_symmetric_band_to_general(A, ka), _symmetric_band_to_general(B, kb), ka, kb, Emid, r, M0, fpm; kw...)
feast_sbev_complex!(A::Matrix{<:Complex}, ka::Int, Emid, r, M0, fpm; kw...) =
    feast_gbev!(_symmetric_band_to_general(A, ka), ka, Emid, r, M0, fpm; kw...)

# ---- polynomial problems: feast_pep! dense/feast_dense.jl:715-772 (first companion linearisation, M0*d columns, leading-block
# eigenvectors), wrappers :946-987, feast_polynomial interfaces/feast_interfaces.jl:448-462
function feast_pep!(A::Vector{<:Matrix}, d::Int, Emid, r, M0::Int, fpm::Vector{Int}; kw...)
    length(A) == d + 1 || throw(ArgumentError("Need d+1 coefficient matrices"))
    N = size(A[1], 1)
    all(size(Ai) == (N, N) for Ai in A) || throw(ArgumentError("All matrices must be same size"))
    DN = d * N
    A_lin = zeros(ComplexF64, DN, DN); B_lin = zeros(ComplexF64, DN, DN)
    for i in 1:d-1
        A_lin[(i-1)*N+1:i*N, i*N+1:(i+1)*N] .= Matrix{ComplexF64}(I, N, N)
        B_lin[(i-1)*N+1:i*N, (i-1)*N+1:i*N] .= Matrix{ComplexF64}(I, N, N)
    end
    for j in 1:d
        A_lin[(d-1)*N+1:d*N, (j-1)*N+1:j*N] .= -A[j]
    end
    B_lin[(d-1)*N+1:d*N, (d-1)*N+1:d*N] .= A[d+1]
    res = feast_gegv!(A_lin, B_lin, Emid, r, M0 * d, fpm; kw...)
    return FeastGeneralResult{Float64}(res.lambda, res.q[1:N, :], res.M, res.res, res.info, res.epsout, res.loop)
end
feast_gepev!(A, d, Emid, r, M0, fpm; kw...) = feast_pep!(A, d, Emid, r, M0, fpm; kw...)
feast_hepev!(A, d, Emid, r, M0, fpm; kw...) = feast_pep!(A, d, Emid, r, M0, fpm; kw...)
feast_sypev!(A::Vector{<:Matrix{<:Real}}, d, Emid, r, M0, fpm; kw...) = feast_pep!([ComplexF64.(Ai) for Ai in A], d, Emid, r, M0, fpm; kw...)
for f in (:feast_gepev, :feast_hepev, :feast_sypev)
    @eval $(Symbol(f, "x!"))(A, d, Emid, r, M0, fpm, Zne, Wne; kw...) = $(Symbol(f, "!"))(A, d, Emid, r, M0, fpm; Zne=Zne, Wne=Wne, kw...)
end
function feast_polynomial(coeffs::Vector{<:AbstractMatrix}, center::Number, radius::Real; M0::Int=10, fpm=nothing, kw...)
    fpm = fpm === nothing ? feastinit() : fpm
    return feast_pep!([Matrix{ComplexF64}(c) for c in coeffs], length(coeffs) - 1, center, radius, M0, fpm; kw...)
end

# feast_banded(A, kla, interval; B, klb, M0, fpm) -- interfaces/feast_interfaces.jl:381-420
function feast_banded(A::Matrix, kla::Int, interval::Tuple; B=nothing, klb::Int=0, M0::Int=10, fpm=nothing, kw...)
    fpm = fpm === nothing ? feastinit() : fpm
    Emin, Emax = interval
    M0 = min(M0, size(A, 2))
    if eltype(A) <: Complex
        return B === nothing ? feast_hbev!(copy(A), kla, Emin, Emax, M0, fpm; kw...) : feast_hbgv!(copy(A), copy(B), kla, klb, Emin, Emax, M0, fpm; kw...)
    end
    return B === nothing ? feast_sbev!(copy(A), kla, Emin, Emax, M0, fpm; kw...) : feast_sbgv!(copy(A), copy(B), kla, klb, Emin, Emax, M0, fpm; kw...)
end

# ---- multi-GPU attach (replaces the MPI communicator of parallel/feast_mpi.jl:9-54): rank 0 creates the id, the
# application broadcasts its 128 bytes (MPI.Bcast!, a file, Distributed), every rank attaches its handle ------------------
function nccl_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:feastcuda_nccl_unique_id, libfeastcuda), Cint, (Ptr{UInt8},), id))
    return id
end
nccl_init!(nranks::Integer, rank::Integer, id::Vector{UInt8}) =
    check(ccall((:feastcuda_nccl_init, libfeastcuda), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), handle(), nranks, rank, id), handle())

# Row sharding (after nccl_init!; real symmetric sparse standard problems): every rank owns a block of rows of A and of every block
# vector; Q0 / X stay global arrays of which a rank reads / writes its own rows (INTEGRATION.md section 5)
row_sharding!(on::Bool=true) =
    check(ccall((:feastcuda_set_row_sharding, libfeastcuda), Cint, (Ptr{Cvoid}, Cint), handle(), on ? 1 : 0), handle())
function row_range()
    r0 = Ref{Int64}(0); nr = Ref{Int64}(0); ng = Ref{Int64}(0)
    check(ccall((:feastcuda_row_range, libfeastcuda), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), handle(), r0, nr, ng), handle())
    return (r0[] + 1):(r0[] + nr[]), ng[]          # this rank's rows (1-based), global order
end

# ---- stage-level wrappers: the block arithmetic inside feast_srci!/feast_hrci!/feast_grci! (kernel/feast_kernel.jl) ------
# Q_proj .+= w .* Y (:143,:519,:766)
accumulate!(Qacc::Matrix{ComplexF64}, w::Number, Y::Matrix{ComplexF64}) = (check(ccall((:feastcuda_accumulate, libfeastcuda), Cint,
    (Ptr{Cvoid}, Cdouble, Cdouble, Int64, Ptr{ComplexF64}, Ptr{ComplexF64}), handle(), real(w), imag(w), size(Y, 2), Y, Qacc), handle()); Qacc)
# Q0' * Y (:147,:522)
function gram(X::Matrix{ComplexF64}, Y::Matrix{ComplexF64})
    C = zeros(ComplexF64, size(X, 2), size(X, 2))
    check(ccall((:feastcuda_gram, libfeastcuda), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                handle(), size(X, 1), size(X, 2), X, Y, C), handle())
    return C
end
# Q_proj * V (:187,:547,:838-845)
function rowtransform(X::Matrix{ComplexF64}, T::Matrix{ComplexF64})
    Y = zeros(ComplexF64, size(X, 1), size(T, 2))
    check(ccall((:feastcuda_rowtransform, libfeastcuda), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                handle(), size(X, 1), size(X, 2), size(T, 2), X, T, Y), handle())
    return Y
end
# eigen(Sq, Aq) (:175,:539,:812)
function eig_general(S::Matrix{ComplexF64}, B::Union{Nothing,Matrix{ComplexF64}}=nothing)
    r = size(S, 1); lambda = zeros(ComplexF64, r); V = zeros(ComplexF64, r, r)
    check(ccall((:feastcuda_eig_general, libfeastcuda), Cint, (Ptr{Cvoid}, Int64, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                handle(), r, S, B === nothing ? C_NULL : pointer(B), lambda, V), handle())
    return lambda, V
end

# high-level feast(A[,B],(Emin,Emax); M0, fpm) -- interfaces/feast_interfaces.jl:143-272 (dispatch only)
# backend keywords (interfaces/feast_interfaces.jl:24-58): validated like the reference; every accepted choice runs on the GPU
# engine (:distributed / :mpi = all ranks attached with nccl_init!; there is no threads backend: one process drives one GPU)
function _select_backend(backend, parallel, strict_backend::Bool, nranks::Int)
    norm(p) = p === true ? :auto : p === false ? :serial : p isa Symbol ? p : throw(ArgumentError("Invalid parallel option: $p"))
    requested = backend !== nothing ? backend : parallel !== nothing ? norm(parallel) : :serial
    backend !== nothing && parallel !== nothing && norm(parallel) != backend &&
        throw(ArgumentError("Conflicting backend requests: backend=$backend and parallel=$(norm(parallel))"))
    requested in (:serial, :auto, :threads, :distributed, :mpi) ||
        throw(ArgumentError("Unknown backend: $requested. Use :serial, :auto, :threads, :distributed, or :mpi"))
    fallback = !strict_backend && (backend === :auto || (backend === nothing && (parallel === true || parallel === :auto)))
    requested in (:serial, :auto) && return requested
    available = requested !== :threads && nranks > 1
    available || fallback || throw(ArgumentError("Backend :$requested is not available"))
    return available ? requested : :serial
end

function feast(A::AbstractMatrix, interval::Tuple; M0::Int=10, fpm=nothing, backend=nothing, parallel=nothing,
               strict_backend::Bool=false, use_threads=nothing, comm=nothing, nranks::Int=1, kw...)
    _select_backend(backend, parallel, strict_backend, nranks)
    size(A, 1) == size(A, 2) || throw(ArgumentError("Matrix must be square"))
    fpm = fpm === nothing ? feastinit() : fpm
    M0 = min(M0, size(A, 1))
    Emin, Emax = interval
    if eltype(A) <: Complex
        ishermitian(A) || throw(ArgumentError("Matrix must be Hermitian; use feast_general"))
        return issparse(A) ? feast_hcsrev!(A, Emin, Emax, M0, fpm; kw...) : feast_heev!(Matrix(A), Emin, Emax, M0, fpm; kw...)
    end
    issymmetric(A) || throw(ArgumentError("Matrix must be symmetric; use feast_general"))
    return issparse(A) ? feast_scsrev!(A, Emin, Emax, M0, fpm; kw...) : feast_syev!(Matrix(A), Emin, Emax, M0, fpm; kw...)
end

export feastinit, feastinit!, feastdefault!, feast_contour, feast_gcontour, feast, FeastResult,
       feast_scsrev!, feast_scsrgv!, feast_hcsrev!, feast_hcsrgv!, feast_syev!, feast_sygv!, feast_heev!, feast_hegv!,
       feast_sbev!, feast_sbgv!, feast_hbev!, feast_hbgv!,
       feast_matvec, FeastGeneralResult, feast_general, feast_gcsrev!, feast_gcsrgv!, feast_geev!, feast_gegv!, feast_gbev!, feast_gbgv!,
       feast_geev_complex_sym!, feast_gegv_complex_sym!, feast_scsrev_complex!, feast_scsrgv_complex!, feast_sbev_complex!,
       feast_sbgv_complex!, feast_pep!, feast_gepev!, feast_hepev!, feast_sypev!, feast_polynomial, feast_banded

end # module
