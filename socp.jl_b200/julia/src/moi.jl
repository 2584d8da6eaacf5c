# moi.jl -- MathOptInterface entry point kept from the reference (src/moi.jl:59-68, :200-224): an
# `Optimizer` whose `optimize!` assembles (c, A, b, G, h, cones) in the ECOS-style row order
# (zero cone rows -> A/b, nonnegative rows first in G, then each second-order cone) and calls
# solve_socp -- here the B200 path, as a batch of one.  Unlike the committed reference wrapper this
# one passes a cone TUPLE and a SolverState, and it reports TerminationStatus / ObjectiveValue from
# the status and objective words the device returns (the reference has neither).
#
# Only the solve entry and the result getters live here; the allocate-load / copy_to plumbing is
# MathOptInterface version specific and out of scope for the hot path (SURVEY.md section 8(f)-2).
# Loaded only when MathOptInterface is available.
const _HAVE_MOI = Base.find_package("MathOptInterface") !== nothing
if _HAVE_MOI
    import MathOptInterface
    const MOI = MathOptInterface

    mutable struct Optimizer <: MOI.AbstractOptimizer
        c::Vector{Float64}
        IA::Vector{Int}; JA::Vector{Int}; VA::Vector{Float64}; b::Vector{Float64}
        IG::Vector{Int}; JG::Vector{Int}; VG::Vector{Float64}; h::Vector{Float64}
        l::Int                      # nonnegative-orthant rows
        q::Vector{Int}              # second-order cone dimensions
        maxsense::Bool
        objconstant::Float64
        sol::Union{Nothing,State}
        options::Dict{Symbol,Any}
        Optimizer(; kwargs...) = new(Float64[], Int[], Int[], Float64[], Float64[], Int[], Int[], Float64[],
                                     Float64[], 0, Int[], false, 0.0, nothing, Dict{Symbol,Any}(kwargs))
    end
    MOI.get(::Optimizer, ::MOI.SolverName) = "Socp (B200)"
    MOI.is_empty(o::Optimizer) = isempty(o.c)

    function MOI.optimize!(o::Optimizer)
        n = length(o.c)
        A = Matrix(sparse(o.IA, o.JA, o.VA, length(o.b), n))
        G = Matrix(sparse(o.IG, o.JG, o.VG, length(o.h), n))
        cones = Cone[]
        offs = 0
        if o.l > 0
            push!(cones, POC(0, o.l)); offs = o.l
        end
        for q in o.q
            push!(cones, SOC(offs, q)); offs += q
        end
        prob = Problem(o.c, A, o.b, G, o.h, Tuple(cones))
        o.sol = solve_socp(prob, SolverState(prob, B200Solver(prob)))
        return
    end

    function MOI.get(o::Optimizer, ::MOI.TerminationStatus)
        o.sol === nothing && return MOI.OPTIMIZE_NOT_CALLED
        o.sol.status == 0 ? MOI.OPTIMAL : (o.sol.status == 1 ? MOI.ITERATION_LIMIT : MOI.NUMERICAL_ERROR)
    end
    MOI.get(o::Optimizer, ::MOI.ObjectiveValue) = (o.maxsense ? -1 : 1) * o.sol.pobj + o.objconstant
    MOI.get(o::Optimizer, ::MOI.DualObjectiveValue) = (o.maxsense ? -1 : 1) * o.sol.dobj + o.objconstant
    MOI.get(o::Optimizer, ::MOI.VariablePrimal, vi::MOI.VariableIndex) = o.sol.x[vi.value]
    MOI.get(o::Optimizer, ::MOI.ResultCount) = o.sol === nothing ? 0 : 1
end
