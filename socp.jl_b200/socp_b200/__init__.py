"""socp_b200 -- Python host-side mirror of Socp.jl's solver interface over the
C ABI of libsocp_b200 (include/socp_b200.h).  The Julia host side lives in
../julia/src; Julia is not installed in this image, so this mirror is what the
tests and bench.py drive.  Names follow the reference: POC, SOC, Problem, State,
SolverState, solve_socp, compute_scaling, setup_iter, solve_kkt, scale_, iscale_,
vprod, iprod, make_e, max_step, compute_step (reference src/Socp.jl,
src/solver.jl, src/densesolver.jl, src/scalings.jl, src/vectors.jl, src/mats.jl)."""
from .api import (POC, SOC, Cone, Problem, State, B200Solver, B200Scaling, SolverState, solve_socp,
                  BatchProblem, CscMatrix, BatchSolverState, solve_socp_batch, BatchResult, SocpError,
                  compute_scaling, setup_iter, solve_kkt, scale_, iscale_, iwiw, sqr_scaling, vprod, iprod, make_e,
                  max_step, compute_step, deg, default_params,
                  STATUS_CONVERGED, STATUS_MAXITER, STATUS_NUMERICAL, PATH_AUTO, PATH_TILED, PATH_FUSED)
from . import generators

__all__ = [n for n in dir() if not n.startswith("_")]
