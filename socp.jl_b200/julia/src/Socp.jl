# Socp.jl -- Julia host side of the B200-native hot path.
#
# Keeps the reference's entry points (BenChung/Socp.jl src/Socp.jl:8-78, src/solver.jl:1-40):
#   POC, SOC, Problem, State, SolverState, solve_socp, the KKTSolver / AbstractScaling seam,
# and adds the batch entry the GPU needs (BatchProblem, solve_socp_batch).  Everything below
# solve_socp's loop runs on the GPU behind the C ABI of include/socp_b200.h, reached with plain
# `ccall`; no CUDA.jl, no kernel generation, no CPU fallback (a missing library is an error).
#
# NOTE: Julia is not installed in the build image, so this file is review-only there; the same
# C ABI is exercised by the Python mirror (socp.jl_b200/socp_b200/api.py) in tests/ and bench.py.
module Socp

using LinearAlgebra
using SparseArrays

export POC, SOC, Problem, BatchProblem, SparseBatchProblem, State, SolverState, B200Solver, B200Scaling,
       solve_socp, solve_socp_batch, compute_scaling, sqr_scaling, setup_iter, solve_kkt, scale!, iscale!,
       vprod, iprod, make_e, max_step, compute_step, deg

const libsocp = get(ENV, "SOCP_B200_LIB", joinpath(@__DIR__, "..", "..", "lib", "libsocp_b200.so"))

# ---------------------------------------------------------------- cones (reference src/Socp.jl:8-18)
abstract type Cone{D} end
struct POC{D} <: Cone{D}
    offs::Int
    POC(offs, dim) = new{dim}(offs)
end
struct SOC{D} <: Cone{D}
    offs::Int
    SOC(offs, dim) = new{dim}(offs)
end
conedim(::Cone{D}) where {D} = D
conekind(::POC) = Int32(0)
conekind(::SOC) = Int32(1)
deg(c::POC{D}) where {D} = D                      # reference src/vectors.jl:165-171
deg(::SOC) = 1
deg(cs::Tuple{Vararg{Cone}}) = sum(deg, cs)

abstract type AbstractScaling end                 # reference src/Socp.jl:77
abstract type KKTSolver{T} end                    # reference src/Socp.jl:78

# ---------------------------------------------------------------- C structs (include/socp_b200.h)
struct CLayout
    n::Int32; p::Int32; k::Int32; ncones::Int32
    cone_kind::Ptr{Int32}; cone_offs::Ptr{Int32}; cone_dim::Ptr{Int32}
end
mutable struct CParams
    max_iter::Int32; path::Int32; tol::Float64; step_damp::Float64; init_eps::Float64
end
function default_params()
    p = CParams(0, 0, 0.0, 0.0, 0.0)
    ccall((:socp_b200_default_params, libsocp), Cvoid, (Ref{CParams},), p)
    return p                                        # 40, auto, 1e-5, 0.99, 1e-10 (src/solver.jl:105,122,146,91)
end

struct SocpError <: Exception
    code::Int
    msg::String
end

# ---------------------------------------------------------------- problems
"""
    BatchProblem(c, A, b, G, h, cones)

`B` independent problems sharing one layout.  `c` is n x B, `b` p x B, `h` k x B, `A` p x n x B,
`G` k x n x B (column-major slices, batch last = the library's layout, no copy); a 2-D `A`/`G`
is shared by the whole batch.  `sing` (reference src/Socp.jl:49-56) is computed on the device.
"""
struct BatchProblem{C<:Tuple{Vararg{Cone}}}
    c::Matrix{Float64}
    A::Array{Float64}
    b::Matrix{Float64}
    G::Array{Float64}
    h::Matrix{Float64}
    cones::C
    n::Int; m::Int; k::Int; B::Int
    function BatchProblem(c::AbstractMatrix, A::AbstractArray, b::AbstractMatrix, G::AbstractArray,
                          h::AbstractMatrix, cones::C) where {C<:Tuple{Vararg{Cone}}}
        n, B = size(c)
        m = size(b, 1)
        k = size(h, 1)
        @assert size(b, 2) == B && size(h, 2) == B
        @assert size(G, 1) == k && size(G, 2) == n          # src/Socp.jl:45-47
        @assert m == 0 || (size(A, 1) == m && size(A, 2) == n)   # src/Socp.jl:43-44
        @assert sum(conedim, cones) == k
        new{C}(Matrix{Float64}(c), Array{Float64}(A), Matrix{Float64}(b), Array{Float64}(G),
               Matrix{Float64}(h), cones, n, m, k, B)
    end
end

"""
    SparseBatchProblem(c, A, b, G, h, cones; Avals, Gvals)

The same batch with `A` and `G` held the way the reference holds them (`SparseMatrixCSC{Float64,Int64}`,
src/Socp.jl:25,29): one sparsity pattern per matrix for the whole batch; `Avals` / `Gvals` (nnz x B) carry the
per-problem values on that pattern (without them the one matrix is shared by every problem).  Nothing is densified on
the host: colptr / rowval / nzval go to the library as they are (`socp_b200_solve_host_csc`).
"""
struct SparseBatchProblem{C<:Tuple{Vararg{Cone}}}
    c::Matrix{Float64}
    A::SparseMatrixCSC{Float64,Int64}
    Avals::Union{Nothing,Matrix{Float64}}
    b::Matrix{Float64}
    G::SparseMatrixCSC{Float64,Int64}
    Gvals::Union{Nothing,Matrix{Float64}}
    h::Matrix{Float64}
    cones::C
    n::Int; m::Int; k::Int; B::Int
    function SparseBatchProblem(c::AbstractMatrix, A::SparseMatrixCSC, b::AbstractMatrix, G::SparseMatrixCSC,
                                h::AbstractMatrix, cones::C; Avals = nothing, Gvals = nothing) where {C<:Tuple{Vararg{Cone}}}
        n, B = size(c)
        m = size(b, 1)
        k = size(h, 1)
        @assert size(b, 2) == B && size(h, 2) == B
        @assert size(G) == (k, n) && size(A) == (m, n)          # src/Socp.jl:43-47
        @assert sum(conedim, cones) == k
        @assert Gvals === nothing || size(Gvals) == (nnz(G), B)
        @assert Avals === nothing || size(Avals) == (nnz(A), B)
        new{C}(Matrix{Float64}(c), SparseMatrixCSC{Float64,Int64}(A), Avals, Matrix{Float64}(b),
               SparseMatrixCSC{Float64,Int64}(G), Gvals, Matrix{Float64}(h), cones, n, m, k, B)
    end
end
const AnyBatchProblem = Union{BatchProblem,SparseBatchProblem}

"Problem(c, A, b, G, h, cones) -- reference src/Socp.jl:40-59.  Sparse A and G stay sparse (a batch of one on the CSC path)."
function Problem(c::AbstractVector, A::SparseMatrixCSC, b::AbstractVector, G::SparseMatrixCSC,
                 h::AbstractVector, cones::Tuple{Vararg{Cone}})
    @assert length(b) == size(A, 1)
    @assert size(A, 2) == length(c) && size(G, 2) == length(c)
    @assert length(h) == size(G, 1)
    SparseBatchProblem(reshape(Vector{Float64}(c), :, 1), A, reshape(Vector{Float64}(b), :, 1), G,
                       reshape(Vector{Float64}(h), :, 1), cones)
end
function Problem(c::AbstractVector, A::AbstractMatrix, b::AbstractVector, G::AbstractMatrix,
                 h::AbstractVector, cones::Tuple{Vararg{Cone}})
    @assert length(b) == size(A, 1)
    @assert size(A, 2) == length(c) && size(G, 2) == length(c)
    @assert length(h) == size(G, 1)
    BatchProblem(reshape(Vector{Float64}(c), :, 1), Matrix{Float64}(A), reshape(Vector{Float64}(b), :, 1),
                 Matrix{Float64}(G), reshape(Vector{Float64}(h), :, 1), cones)
end

mutable struct State                               # reference src/Socp.jl:62-75 (+ what it lacks)
    x::Vector{Float64}; y::Vector{Float64}; z::Vector{Float64}; s::Vector{Float64}
    status::Int32; iters::Int32; pobj::Float64; dobj::Float64
end

# ---------------------------------------------------------------- solver state = device handle
mutable struct B200Scaling <: AbstractScaling      # fields as struct Scaling, src/scalings.jl:1-20
    l::Matrix{Float64}; wbs::Matrix{Float64}; mu::Matrix{Float64}; fail::Vector{Int32}
end
# The solver object only names the back end and the devices to shard over (DenseSolver's role in the plug-in seam,
# src/densesolver.jl:1-39); the device handle is owned by the SolverState built from it, so one B200Solver can seed
# several SolverStates without their finalizers sharing (and leaking) a handle.
mutable struct B200Solver <: KKTSolver{B200Scaling}
    devices::Vector{Int32}
    B200Solver(prob = nothing; devices = Int32[]) = new(Vector{Int32}(devices))
end

mutable struct SolverState{S<:KKTSolver}
    solver::S
    handle::Ptr{Cvoid}
    scaling::B200Scaling
    n::Int; m::Int; k::Int; B::Int; ncones::Int
    function SolverState(pr::AnyBatchProblem, solver::B200Solver)
        kind = Int32[conekind(c) for c in pr.cones]
        offs = Int32[c.offs for c in pr.cones]
        dim = Int32[conedim(c) for c in pr.cones]
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = GC.@preserve kind offs dim begin
            lay = CLayout(pr.n, pr.m, pr.k, length(pr.cones), pointer(kind), pointer(offs), pointer(dim))
            devs = isempty(solver.devices) ? C_NULL : pointer(solver.devices)
            ccall((:socp_b200_create, libsocp), Cint,
                  (Ref{Ptr{Cvoid}}, Ref{CLayout}, Int64, Ptr{Int32}, Int32),
                  h, lay, pr.B, devs, length(solver.devices))
        end
        rc == 0 || throw(SocpError(rc, unsafe_string(ccall((:socp_b200_last_error, libsocp), Cstring, (Ptr{Cvoid},), C_NULL))))
        sc = B200Scaling(zeros(pr.k, pr.B), zeros(pr.k, pr.B), zeros(length(pr.cones), pr.B), zeros(Int32, pr.B))
        ss = new{B200Solver}(solver, h[], sc, pr.n, pr.m, pr.k, pr.B, length(pr.cones))
        finalizer(s -> (s.handle != C_NULL && ccall((:socp_b200_destroy, libsocp), Cint, (Ptr{Cvoid},), s.handle);
                        s.handle = C_NULL), ss)
        return ss
    end
end

check(ss, rc, what) = rc == 0 || throw(SocpError(rc, what * ": " *
    unsafe_string(ccall((:socp_b200_last_error, libsocp), Cstring, (Ptr{Cvoid},), ss.handle))))

function load!(ss::SolverState, pr::BatchProblem)
    flags = Int32((ndims(pr.A) == 2 && pr.m > 0 ? 1 : 0) | (ndims(pr.G) == 2 ? 2 : 0))
    rc = GC.@preserve pr ccall((:socp_b200_set_data, libsocp), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Int32),
        ss.handle, pr.c, pr.m > 0 ? pointer(pr.A) : C_NULL, pr.m > 0 ? pointer(pr.b) : C_NULL,
        pr.G, pr.h, C_NULL, flags)
    check(ss, rc, "socp_b200_set_data")
end

# The reference keeps A and G as SparseMatrixCSC{Float64,Int64} (src/Socp.jl:25,29).  Their three fields are
# handed to the library as they are (1-based): only nnz values per problem cross PCIe and the dense operands
# of the KKT path are assembled on the device.  `Avals` / `Gvals` (nnz x B) give per-problem values on the
# shared pattern; without them the one matrix is shared by the whole batch.
struct CCsc
    nnz::Int64; colptr::Ptr{Int64}; rowval::Ptr{Int64}; nzval::Ptr{Float64}; index_base::Int32
end
function load!(ss::SolverState, c::Matrix{Float64}, A::Union{Nothing,SparseMatrixCSC{Float64,Int64}}, b::Matrix{Float64},
               G::SparseMatrixCSC{Float64,Int64}, h::Matrix{Float64};
               Avals::Union{Nothing,Matrix{Float64}} = nothing, Gvals::Union{Nothing,Matrix{Float64}} = nothing,
               sing::Union{Nothing,Vector{UInt8}} = nothing)
    @assert size(G) == (ss.k, ss.n) && (ss.m == 0 || size(A) == (ss.m, ss.n))       # src/Socp.jl:43-47
    @assert Gvals === nothing || size(Gvals) == (nnz(G), ss.B)
    @assert Avals === nothing || size(Avals) == (nnz(A), ss.B)
    flags = Int32((Avals === nothing && ss.m > 0 ? 1 : 0) | (Gvals === nothing ? 2 : 0))
    gv = Gvals === nothing ? G.nzval : Gvals
    rc = GC.@preserve c A b G h gv Avals sing begin
        gs = Ref(CCsc(nnz(G), pointer(G.colptr), pointer(G.rowval), pointer(gv), 1))
        as = ss.m > 0 ? Ref(CCsc(nnz(A), pointer(A.colptr), pointer(A.rowval),
                                 pointer(Avals === nothing ? A.nzval : Avals), 1)) : Ref(CCsc(0, C_NULL, C_NULL, C_NULL, 1))
        ccall((:socp_b200_set_data_csc, libsocp), Cint,
              (Ptr{Cvoid}, Ptr{Float64}, Ptr{CCsc}, Ptr{Float64}, Ptr{CCsc}, Ptr{Float64}, Ptr{UInt8}, Int32),
              ss.handle, c, ss.m > 0 ? as : C_NULL,
              ss.m > 0 ? pointer(b) : Ptr{Float64}(C_NULL), gs, h,
              sing === nothing ? Ptr{UInt8}(C_NULL) : pointer(sing), flags)
    end
    check(ss, rc, "socp_b200_set_data_csc")
end

"""
    solve_socp_batch(prob, ss; params, sing)

solve_socp (src/solver.jl:40-152) for every problem of the batch on the GPU, as ONE library call from host data
to host results (`socp_b200_solve_host`: upload, solve and download of chunks overlap).  `sing` is the reference's
5th type parameter per problem (src/Socp.jl:49-56: `cholesky(G'G)` fails).  The default `nothing` has the library
run that test on the device, as the reference's constructor does on the host; pass `zeros(UInt8, B)` only when every
`G` is known to have full column rank (skips the test).
"""
function solve_socp_batch(pr::BatchProblem, ss::SolverState; params = default_params(),
                          sing::Union{Nothing,Vector{UInt8}} = nothing)
    x = zeros(pr.n, pr.B); y = zeros(pr.m, pr.B); z = zeros(pr.k, pr.B); s = zeros(pr.k, pr.B)
    status = zeros(Int32, pr.B); iters = zeros(Int32, pr.B); pobj = zeros(pr.B); dobj = zeros(pr.B)
    flags = Int32((ndims(pr.A) == 2 && pr.m > 0 ? 1 : 0) | (ndims(pr.G) == 2 ? 2 : 0))
    rc = GC.@preserve pr sing ccall((:socp_b200_solve_host, libsocp), Cint,
        (Ptr{Cvoid}, Ref{CParams}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{UInt8}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
        ss.handle, params, pr.c, pr.m > 0 ? pointer(pr.A) : C_NULL, pr.m > 0 ? pointer(pr.b) : C_NULL,
        pr.G, pr.h, sing === nothing ? Ptr{UInt8}(C_NULL) : pointer(sing), flags,
        x, pr.m > 0 ? pointer(y) : C_NULL, z, s, status, iters, pobj, dobj)
    check(ss, rc, "socp_b200_solve_host")
    return (x = x, y = y, z = z, s = s, status = status, iters = iters, pobj = pobj, dobj = dobj)
end

# The same one-shot call from the reference's own storage: colptr / rowval / nzval (1-based) as they are.
function solve_socp_batch(pr::SparseBatchProblem, ss::SolverState; params = default_params(),
                          sing::Union{Nothing,Vector{UInt8}} = nothing)
    x = zeros(pr.n, pr.B); y = zeros(pr.m, pr.B); z = zeros(pr.k, pr.B); s = zeros(pr.k, pr.B)
    status = zeros(Int32, pr.B); iters = zeros(Int32, pr.B); pobj = zeros(pr.B); dobj = zeros(pr.B)
    flags = Int32((pr.Avals === nothing && pr.m > 0 ? 1 : 0) | (pr.Gvals === nothing ? 2 : 0))
    gv = pr.Gvals === nothing ? pr.G.nzval : pr.Gvals
    av = pr.Avals === nothing ? pr.A.nzval : pr.Avals
    rc = GC.@preserve pr gv av sing x y z s status iters pobj dobj begin
        gs = Ref(CCsc(nnz(pr.G), pointer(pr.G.colptr), pointer(pr.G.rowval), pointer(gv), 1))
        as = Ref(CCsc(nnz(pr.A), pointer(pr.A.colptr), pointer(pr.A.rowval), pr.m > 0 ? pointer(av) : Ptr{Float64}(C_NULL), 1))
        ccall((:socp_b200_solve_host_csc, libsocp), Cint,
              (Ptr{Cvoid}, Ref{CParams}, Ptr{Float64}, Ptr{CCsc}, Ptr{Float64}, Ptr{CCsc}, Ptr{Float64},
               Ptr{UInt8}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
               Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
              ss.handle, params, pr.c, pr.m > 0 ? as : C_NULL, pr.m > 0 ? pointer(pr.b) : Ptr{Float64}(C_NULL), gs, pr.h,
              sing === nothing ? Ptr{UInt8}(C_NULL) : pointer(sing), flags,
              x, pr.m > 0 ? pointer(y) : Ptr{Float64}(C_NULL), z, s, status, iters, pobj, dobj)
    end
    check(ss, rc, "socp_b200_solve_host_csc")
    return (x = x, y = y, z = z, s = s, status = status, iters = iters, pobj = pobj, dobj = dobj)
end

"solve_socp(prob, ss) -> State, reference src/solver.jl:40-152 (a batch of one)."
function solve_socp(pr::AnyBatchProblem, ss::SolverState; params = default_params())
    @assert pr.B == 1
    r = solve_socp_batch(pr, ss; params = params)
    State(r.x[:, 1], r.y[:, 1], r.z[:, 1], r.s[:, 1], r.status[1], r.iters[1], r.pobj[1], r.dobj[1])
end

# ---------------------------------------------------------------- the plug-in seam, batched (step level)
"compute_scaling(cones, scaling, s, z) -- reference src/scalings.jl:101-110; s, z are k x B."
function compute_scaling(ss::SolverState, s::Matrix{Float64}, z::Matrix{Float64})
    sc = ss.scaling
    rc = ccall((:socp_b200_compute_scaling, libsocp), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
        ss.handle, s, z, sc.l, sc.wbs, sc.mu, sc.fail)
    check(ss, rc, "socp_b200_compute_scaling")
    return sc
end
"setup_iter(solver, prob, state, scaling) -- reference src/densesolver.jl:41-52."
function setup_iter(ss::SolverState)
    fail = zeros(Int32, ss.B)
    check(ss, ccall((:socp_b200_setup_iter, libsocp), Cint, (Ptr{Cvoid}, Ptr{Int32}), ss.handle, fail), "socp_b200_setup_iter")
    return fail
end
"solve_kkt(solver, prob, state, scaling, dx,dy,dz,ds, cx,cy,cz,cs) -- reference src/densesolver.jl:54-90."
function solve_kkt(ss::SolverState, dx, dy, dz, ds, cx, cy, cz, cs)
    pn(a) = isempty(a) ? Ptr{Float64}(C_NULL) : pointer(a)
    rc = GC.@preserve dx dy dz ds cx cy cz cs ccall((:socp_b200_solve_kkt, libsocp), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ss.handle, pn(dx), pn(dy), pn(dz), pn(ds), pn(cx), pn(cy), pn(cz), pn(cs))
    check(ss, rc, "socp_b200_solve_kkt")
end
for (jl, cfn) in ((:scale!, :socp_b200_scale), (:iscale!, :socp_b200_iscale))   # src/scalings.jl:159-173
    @eval function $jl(ss::SolverState, inp::Matrix{Float64}, out::Matrix{Float64})
        check(ss, ccall(($(QuoteNode(cfn)), libsocp), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                        ss.handle, inp, out), $(string(cfn)))
        return out
    end
end
"""
    sqr_scaling(ss) -> (D, u, v)

The SqrScaling form of the scaling computed last (`compute_scaling(cones, ::SqrScaling, s, z)`, reference
src/sqrscalings.jl:177-185; per cone :50-58, :98-128): `W^-2 = Diagonal(D) + u u' - v v'` cone by cone.  Each result is
k x B; the reference's per-cone vectors `us[c]`, `vs[c]` (support on cone c only) are packed into one k-vector.
"""
function sqr_scaling(ss::SolverState)
    D = zeros(ss.k, ss.B); u = zeros(ss.k, ss.B); v = zeros(ss.k, ss.B)
    check(ss, ccall((:socp_b200_sqr_scaling, libsocp), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    ss.handle, D, u, v), "socp_b200_sqr_scaling")
    return D, u, v
end
function vprod(ss::SolverState, u::Matrix{Float64}, v::Matrix{Float64})          # src/vectors.jl:58-81
    out = similar(u)
    check(ss, ccall((:socp_b200_vprod, libsocp), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    ss.handle, u, v, out), "socp_b200_vprod")
    out
end
function iprod(ss::SolverState, lam::Matrix{Float64}, v::Matrix{Float64})        # src/vectors.jl:99-131
    out = similar(v)
    check(ss, ccall((:socp_b200_iprod, libsocp), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    ss.handle, lam, v, out), "socp_b200_iprod")
    out
end
function make_e(ss::SolverState)                                                  # src/vectors.jl:7-24
    out = zeros(ss.k, ss.B)
    check(ss, ccall((:socp_b200_make_e, libsocp), Cint, (Ptr{Cvoid}, Ptr{Float64}), ss.handle, out), "socp_b200_make_e")
    out
end
function max_step(ss::SolverState, x::Matrix{Float64})                            # src/mats.jl:1-28
    out = zeros(ss.B)
    check(ss, ccall((:socp_b200_max_step, libsocp), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ss.handle, x, out), "socp_b200_max_step")
    out
end
function compute_step(ss::SolverState, l::Matrix{Float64}, ds::Matrix{Float64}, dz::Matrix{Float64})   # src/mats.jl:30-40
    out = zeros(ss.B)
    check(ss, ccall((:socp_b200_compute_step, libsocp), Cint,
                    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), ss.handle, l, ds, dz, out), "socp_b200_compute_step")
    out
end

include("moi.jl")

end # module
