# moi.jl -- MathOptInterface wrapper (MOI 0.9.x, the version the reference pins: Manifest.toml:137-141).
#
# Same surface as the reference's wrapper (src/moi.jl:59-272): `Socp.Optimizer`, model loading through the
# allocate-load interface of MOI.Utilities (`copy_to` -> `automatic_copy_to`), `optimize!`, and the result getters --
# with what the committed reference wrapper lacks or gets wrong (SURVEY.md section 0): it passes the cones as a TUPLE
# and a SolverState to solve_socp (the reference passes a Vector and no state, src/moi.jl:215-222), the constraint
# getters do not call undefined helpers (`scalecoef`, src/moi.jl:262,269), and TerminationStatus / PrimalStatus /
# DualStatus / ObjectiveValue / DualObjectiveValue are reported from the status and objective words the device returns.
#
# Canonical form handed to the solver (the ECOS convention the reference follows, src/moi.jl:155-158):
#     min c'x   s.t.   A x = b,   h - G x in K = R^l_+ x Q^{q_1} x ... x Q^{q_N}
# A VectorAffineFunction f(x) = F x + g in a set S is stored as rows (-F | g): Zeros rows go to (A, b), Nonnegatives
# rows to the first l rows of (G, h), every SecondOrderCone to its own block after them.  The index of a constraint is
# its first row inside its group (unique per (function, set) type pair, as MOI requires).
#
# A and G stay sparse all the way: `optimize!` builds SparseMatrixCSC{Float64,Int64} (the reference's own storage,
# src/moi.jl:208-210, src/Socp.jl:25,29) and hands colptr / rowval / nzval to socp_b200_solve_host_csc as they are.
#
# Review-only in the build image (no Julia, no MathOptInterface there); loaded only when MathOptInterface is found.
const _HAVE_MOI = Base.find_package("MathOptInterface") !== nothing
if _HAVE_MOI
    import MathOptInterface
    const MOI = MathOptInterface
    const MOIU = MOI.Utilities

    const AffineVec = MOI.VectorAffineFunction{Float64}
    const ConeSets = Union{MOI.Zeros, MOI.Nonnegatives, MOI.SecondOrderCone}

    # rows of the three groups as they are allocated, and what the getters need afterwards
    mutable struct RowLayout
        neq::Int                        # rows of Zeros constraints so far          -> (A, b)
        nlp::Int                        # rows of Nonnegatives constraints so far   -> G rows 1..nlp
        nsoc::Int                       # rows of SecondOrderCone constraints so far-> G rows nlp+1..
        socdims::Vector{Int}            # dimension of every second-order cone, in allocation order
        eqlen::Dict{Int,Int}            # first row (0-based, inside (A, b)) -> number of rows
        conelen::Dict{Int,Int}          # first row (0-based, inside (G, h)) -> number of rows
        RowLayout() = new(0, 0, 0, Int[], Dict{Int,Int}(), Dict{Int,Int}())
    end

    # triplets collected between copy_to and optimize!
    mutable struct Triplets
        nvar::Int
        c::Vector{Float64}
        objconstant::Float64
        Ai::Vector{Int}; Aj::Vector{Int}; Av::Vector{Float64}; b::Vector{Float64}
        Gi::Vector{Int}; Gj::Vector{Int}; Gv::Vector{Float64}; h::Vector{Float64}
    end

    mutable struct Optimizer <: MOI.AbstractOptimizer
        rows::RowLayout
        maxsense::Bool
        objconstant::Float64
        data::Union{Nothing,Triplets}       # non-nothing between copy_to and optimize!
        sol::Union{Nothing,State}
        silent::Bool
        devices::Vector{Int32}
        params::CParams
        function Optimizer(; devices = Int32[], max_iter = nothing, tol = nothing, kwargs...)
            prm = default_params()
            max_iter === nothing || (prm.max_iter = Int32(max_iter))
            tol === nothing || (prm.tol = Float64(tol))
            new(RowLayout(), false, 0.0, nothing, nothing, false, Vector{Int32}(devices), prm)
        end
    end

    MOI.get(::Optimizer, ::MOI.SolverName) = "Socp (B200)"
    MOI.supports(::Optimizer, ::MOI.Silent) = true
    MOI.set(o::Optimizer, ::MOI.Silent, v::Bool) = (o.silent = v)
    MOI.get(o::Optimizer, ::MOI.Silent) = o.silent
    MOI.is_empty(o::Optimizer) = !o.maxsense && o.data === nothing && o.sol === nothing
    function MOI.empty!(o::Optimizer)
        o.rows = RowLayout()
        o.maxsense = false
        o.objconstant = 0.0
        o.data = nothing
        o.sol = nothing
        return
    end

    MOI.supports(::Optimizer, ::Union{MOI.ObjectiveSense, MOI.ObjectiveFunction{MOI.ScalarAffineFunction{Float64}}}) = true
    MOI.supports_constraint(::Optimizer, ::Type{AffineVec}, ::Type{<:ConeSets}) = true

    # ---- allocate-load (MOI.Utilities, MOI 0.9): the first pass sizes the row groups, the second fills the triplets
    MOIU.supports_allocate_load(::Optimizer, copy_names::Bool) = !copy_names
    MOI.copy_to(dest::Optimizer, src::MOI.ModelLike; kws...) = MOIU.automatic_copy_to(dest, src; kws...)

    function MOIU.allocate_variables(o::Optimizer, nvars::Integer)
        o.rows = RowLayout()
        return MOI.VariableIndex.(1:nvars)
    end
    function MOIU.load_variables(o::Optimizer, nvars::Integer)
        r = o.rows
        o.data = Triplets(nvars, zeros(nvars), 0.0, Int[], Int[], Float64[], zeros(r.neq),
                          Int[], Int[], Float64[], zeros(r.nlp + r.nsoc))
        return
    end

    # first row of a new constraint inside its group
    function take_rows!(r::RowLayout, s::MOI.Zeros)
        first = r.neq; r.neq += MOI.dimension(s); first
    end
    function take_rows!(r::RowLayout, s::MOI.Nonnegatives)
        first = r.nlp; r.nlp += MOI.dimension(s); first
    end
    function take_rows!(r::RowLayout, s::MOI.SecondOrderCone)
        first = r.nsoc; r.nsoc += MOI.dimension(s); push!(r.socdims, MOI.dimension(s)); first
    end
    function MOIU.allocate_constraint(o::Optimizer, ::F, s::S) where {F<:MOI.AbstractFunction, S<:MOI.AbstractSet}
        return MOI.ConstraintIndex{F,S}(take_rows!(o.rows, s))
    end

    # 0-based first row of a constraint inside its matrix: second-order cones follow the l orthant rows of G
    first_row(::RowLayout, ci::MOI.ConstraintIndex{<:MOI.AbstractFunction, MOI.Zeros}) = ci.value
    first_row(::RowLayout, ci::MOI.ConstraintIndex{<:MOI.AbstractFunction, MOI.Nonnegatives}) = ci.value
    first_row(r::RowLayout, ci::MOI.ConstraintIndex{<:MOI.AbstractFunction, MOI.SecondOrderCone}) = r.nlp + ci.value

    function MOIU.load_constraint(o::Optimizer, ci::MOI.ConstraintIndex, f::AffineVec, s::ConeSets)
        fc = MOIU.canonical(f)                    # duplicates merged, zeros dropped
        d = o.data
        r0 = first_row(o.rows, ci)
        nrow = MOI.dimension(s)
        eq = s isa MOI.Zeros
        (eq ? o.rows.eqlen : o.rows.conelen)[r0] = nrow
        rhs = eq ? d.b : d.h
        Is, Js, Vs = eq ? (d.Ai, d.Aj, d.Av) : (d.Gi, d.Gj, d.Gv)
        # f(x) = F x + g in S   <=>   g - (-F) x in S: the matrix rows are -F, the right-hand side is g
        rhs[r0 .+ (1:nrow)] .= fc.constants
        for t in fc.terms
            push!(Is, r0 + t.output_index)
            push!(Js, t.scalar_term.variable_index.value)
            push!(Vs, -t.scalar_term.coefficient)
        end
        return
    end

    MOIU.allocate(o::Optimizer, ::MOI.ObjectiveSense, sense::MOI.OptimizationSense) = (o.maxsense = sense == MOI.MAX_SENSE)
    MOIU.allocate(::Optimizer, ::MOI.ObjectiveFunction, ::MOI.ScalarAffineFunction{Float64}) = nothing
    MOIU.load(::Optimizer, ::MOI.ObjectiveSense, ::MOI.OptimizationSense) = nothing
    function MOIU.load(o::Optimizer, ::MOI.ObjectiveFunction, f::MOI.ScalarAffineFunction{Float64})
        c = zeros(o.data.nvar)
        for t in f.terms
            c[t.variable_index.value] += t.coefficient
        end
        o.objconstant = f.constant
        o.data.objconstant = f.constant
        o.data.c = o.maxsense ? -c : c            # the solver minimises
        return
    end

    # ---- solve: Problem(c, A, b, G, h, cones) + solve_socp(prob, ss), reference src/moi.jl:200-224
    function MOI.optimize!(o::Optimizer)
        d = o.data
        d === nothing && return                   # already solved, nothing new copied (reference :201-204)
        r = o.rows
        A = sparse(d.Ai, d.Aj, d.Av, r.neq, d.nvar)
        G = sparse(d.Gi, d.Gj, d.Gv, r.nlp + r.nsoc, d.nvar)
        cones = Cone[]
        offs = 0
        if r.nlp > 0
            push!(cones, POC(0, r.nlp)); offs = r.nlp
        end
        for q in r.socdims
            push!(cones, SOC(offs, q)); offs += q
        end
        prob = Problem(d.c, A, d.b, G, d.h, Tuple(cones))          # sparse A, G: kept as CSC (Socp.jl)
        ss = SolverState(prob, B200Solver(prob; devices = o.devices))
        o.sol = solve_socp(prob, ss; params = o.params)
        finalize(ss)                                               # release the device handle now
        o.data = nothing
        return
    end

    # ---- results
    function MOI.get(o::Optimizer, ::MOI.TerminationStatus)
        o.sol === nothing && return MOI.OPTIMIZE_NOT_CALLED
        o.sol.status == 0 ? MOI.OPTIMAL : (o.sol.status == 1 ? MOI.ITERATION_LIMIT : MOI.NUMERICAL_ERROR)
    end
    function MOI.get(o::Optimizer, ::MOI.RawStatusString)
        o.sol === nothing && return "optimize! not called"
        ("stop test |rx| + |ry| + z's < tol met", "iteration limit reached", "numerical failure (Cholesky or cone membership)")[o.sol.status + 1]
    end
    MOI.get(o::Optimizer, ::MOI.ResultCount) = o.sol === nothing ? 0 : 1
    MOI.get(o::Optimizer, ::MOI.PrimalStatus) = o.sol === nothing ? MOI.NO_SOLUTION : (o.sol.status == 0 ? MOI.FEASIBLE_POINT : MOI.UNKNOWN_RESULT_STATUS)
    MOI.get(o::Optimizer, ::MOI.DualStatus) = MOI.get(o, MOI.PrimalStatus())
    MOI.get(o::Optimizer, ::MOI.ObjectiveValue) = (o.maxsense ? -1 : 1) * o.sol.pobj + o.objconstant
    MOI.get(o::Optimizer, ::MOI.DualObjectiveValue) = (o.maxsense ? -1 : 1) * o.sol.dobj + o.objconstant
    MOI.get(o::Optimizer, ::MOI.BarrierIterations) = Int(o.sol.iters)

    MOI.get(o::Optimizer, ::MOI.VariablePrimal, vi::MOI.VariableIndex) = o.sol.x[vi.value]
    MOI.get(o::Optimizer, a::MOI.VariablePrimal, vis::Vector{MOI.VariableIndex}) = [MOI.get(o, a, vi) for vi in vis]

    # value of the constraint function F x + g: identically zero on a Zeros constraint at a feasible point, and the
    # slack s = h - G x on a cone constraint (reference src/moi.jl:246-263)
    function MOI.get(o::Optimizer, ::MOI.ConstraintPrimal, ci::MOI.ConstraintIndex{AffineVec, MOI.Zeros})
        return zeros(o.rows.eqlen[first_row(o.rows, ci)])
    end
    function MOI.get(o::Optimizer, ::MOI.ConstraintPrimal, ci::MOI.ConstraintIndex{AffineVec, <:Union{MOI.Nonnegatives, MOI.SecondOrderCone}})
        r0 = first_row(o.rows, ci)
        return o.sol.s[r0 .+ (1:o.rows.conelen[r0])]
    end
    # multipliers: y on the equality rows, z on the cone rows (reference :265-270).  The solver's stationarity
    # c + A'y + G'z = 0 with A = G = -F reads c - F'lambda = 0, MOI's convention for a minimisation; for a maximisation
    # the solver saw -c, which gives c + F'lambda = 0 -- MOI's convention for a maximisation -- so the multipliers are
    # returned as they are in both cases.
    function MOI.get(o::Optimizer, ::MOI.ConstraintDual, ci::MOI.ConstraintIndex{AffineVec, MOI.Zeros})
        r0 = first_row(o.rows, ci)
        return o.sol.y[r0 .+ (1:o.rows.eqlen[r0])]
    end
    function MOI.get(o::Optimizer, ::MOI.ConstraintDual, ci::MOI.ConstraintIndex{AffineVec, <:Union{MOI.Nonnegatives, MOI.SecondOrderCone}})
        r0 = first_row(o.rows, ci)
        return o.sol.z[r0 .+ (1:o.rows.conelen[r0])]
    end
end
