/*
 * socp_b200.h -- C ABI of the B200-native replacement for the per-iteration KKT
 * hot path of BenChung/Socp.jl (dense primal-dual interior-point SOCP solver),
 * for batches of independent dense SOCPs that share one cone layout.
 *
 *     minimize c'x   s.t.  A x = b,   G x + s = h,   s in K
 *     K = product of positive-orthant blocks (POC) and second-order cones (SOC)
 *
 * Everything is Float64.  Matrices are column-major exactly as a Julia
 * Matrix{Float64} (G is k x n with leading dimension k, A is p x n with leading
 * dimension p); the batch index is the slowest one (problem b's G starts at
 * G + b*k*n).  All pointers in this header are HOST pointers owned by the caller
 * for the duration of the call only, unless the function name ends in `_dev`.
 *
 * Return value of every function: 0 = OK, <0 = usage error (the analogue of the
 * reference's @assert's, src/Socp.jl:43-47), >0 = CUDA runtime error code; the
 * message is available from socp_b200_last_error().  Per-problem outcomes are
 * DATA (status[b]), never a non-zero return.  There is no CPU fallback: without
 * a CUDA device every call except _version/_last_error fails with a CUDA error.
 *
 * "reference" below = /root/reference (BenChung/Socp.jl), cited file:line.
 */
#ifndef SOCP_B200_H
#define SOCP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOCP_B200_VERSION 100

/* cone kinds: POC{D} / SOC{D}, reference src/Socp.jl:8-18 */
#define SOCP_CONE_POC 0
#define SOCP_CONE_SOC 1

/* per-problem status words (defined by the oracle; the reference has none, it
 * breaks / runs out of iterations / throws -- src/solver.jl:105,122-124) */
#define SOCP_STATUS_CONVERGED 0  /* stop test src/solver.jl:122 satisfied           */
#define SOCP_STATUS_MAXITER   1  /* max_iter Mehrotra steps taken                   */
#define SOCP_STATUS_NUMERICAL 2  /* where the reference would throw (DomainError in
                                    sqrt, PosDefException in cholesky!)             */

/* error codes (<0) */
#define SOCP_ERR_NULL     -1
#define SOCP_ERR_LAYOUT   -2
#define SOCP_ERR_SIZE     -3
#define SOCP_ERR_STATE    -4   /* call order (e.g. solve before set_data)            */
#define SOCP_ERR_NOMEM    -5

/* set_data flags */
#define SOCP_FLAG_SHARED_A 1   /* one A (p x n) shared by every problem of the batch */
#define SOCP_FLAG_SHARED_G 2   /* one G (k x n) shared by every problem of the batch */

/* which path _solve uses */
#define SOCP_PATH_AUTO   0     /* fused whole-solve kernel when the layout fits in
                                  shared memory, tiled global-memory path otherwise  */
#define SOCP_PATH_TILED  1     /* force the tiled (multi-kernel) path                */
#define SOCP_PATH_FUSED  2     /* force the fused kernel (error if it does not fit)  */

typedef struct socp_handle socp_handle;

/* Replaces the type parameters of Problem{C,n,m,k,sing} (reference
 * src/Socp.jl:20-38) and the cone tuple: one layout per handle, shared by all
 * problems of the batch.  `p` is the reference's `m` (equality rows). Cones must
 * tile 0..k-1 contiguously in order (cone_offs[i] = sum of earlier dims), POC
 * blocks first as the reference assumes (src/scalings.jl:102). */
typedef struct socp_layout {
    int32_t n;                 /* variables                      (Problem.n) */
    int32_t p;                 /* equality rows                  (Problem.m) */
    int32_t k;                 /* cone rows                      (Problem.k) */
    int32_t ncones;
    const int32_t* cone_kind;  /* SOCP_CONE_POC / SOCP_CONE_SOC              */
    const int32_t* cone_offs;  /* 0-based offset = Cone.offs                 */
    const int32_t* cone_dim;   /* D                                           */
} socp_layout;

/* The reference's literals: 40 (src/solver.jl:105), 1e-5 (:122), 0.99 (:146),
 * 1e-10 (:91,:97).  socp_b200_default_params() fills them in. */
typedef struct socp_params {
    int32_t max_iter;
    int32_t path;              /* SOCP_PATH_*                                 */
    double  tol;
    double  step_damp;
    double  init_eps;
} socp_params;

/* device-side wall times (CUDA events) of the last _solve / _solve_dev, ms */
typedef struct socp_timings {
    double h2d_ms;             /* set_data upload                             */
    double solve_ms;           /* initial point + Mehrotra loop on the device */
    double d2h_ms;             /* result download                             */
    int64_t kernel_launches;   /* kernels launched by the last solve          */
    int32_t iterations_max;    /* largest per-problem iteration count         */
    int32_t path_used;         /* SOCP_PATH_TILED or SOCP_PATH_FUSED          */
} socp_timings;

int  socp_b200_version(void);
int  socp_b200_device_count(void);
void socp_b200_default_params(socp_params* out);

/* Handle = SolverState + DenseSolver workspaces of the reference
 * (src/solver.jl:1-38, src/densesolver.jl:1-39) for `batch` problems, resident
 * on the listed devices (contiguous shard per device, no collective).  devices
 * == NULL / ndev == 0 means device 0 (or the current one).  Reusable across
 * set_data/solve calls, like the reference's SolverState (test/runtests.jl:243). */
int  socp_b200_create(socp_handle** out, const socp_layout* layout, int64_t batch,
                      const int32_t* devices, int32_t ndev);
int  socp_b200_destroy(socp_handle* h);
const char* socp_b200_last_error(const socp_handle* h);

/* Problem(c, A, b, G, h, cones), reference src/Socp.jl:40-59.  Copies the data
 * to the device(s).  `sing` (one byte per problem, the 5th type parameter:
 * "cholesky(G'G) failed", src/Socp.jl:49-56) may be NULL, in which case it is
 * computed on the device with the same test. */
int  socp_b200_set_data(socp_handle* h, const double* c, const double* A, const double* b,
                        const double* G, const double* hvec, const uint8_t* sing, int32_t flags);

/* The same constructor for callers that hold A and G the way the reference
 * stores them: SparseMatrixCSC{Float64,Int64} (fields colptr, rowval, nzval;
 * src/Socp.jl:25,29, built in src/moi.jl:208-210).  One sparsity pattern per
 * matrix for the whole batch; nzval is [batch][nnz], or [nnz] with the
 * SOCP_FLAG_SHARED_* bit.  index_base = 1 for Julia's arrays, 0 for C / scipy.
 * Row indices must increase strictly inside every column (the invariant of
 * SparseMatrixCSC); anything else is SOCP_ERR_LAYOUT.  Only nnz values per
 * problem cross PCIe; the dense column-major operands of the KKT path are
 * assembled on the device.  A may be NULL when p == 0. */
typedef struct socp_csc {
    int64_t        nnz;         /* stored entries of one matrix                  */
    const int64_t* colptr;      /* [cols + 1]                                    */
    const int64_t* rowval;      /* [nnz]                                         */
    const double*  nzval;       /* [batch][nnz], or [nnz] when shared            */
    int32_t        index_base;  /* 0 or 1                                        */
} socp_csc;
int  socp_b200_set_data_csc(socp_handle* h, const double* c, const socp_csc* A, const double* b,
                            const socp_csc* G, const double* hvec, const uint8_t* sing, int32_t flags);

/* solve_socp(prob, ss), reference src/solver.jl:40-152, for the whole batch:
 * initial point (:68-104) + Mehrotra loop (:105-151) on the device, results
 * copied back.  Any output pointer may be NULL.  Shapes: x[batch][n],
 * y[batch][p], z[batch][k], s[batch][k], status/iters/pobj/dobj[batch].
 * iters = completed loop bodies; pobj = c'x; dobj = -b'y - h'z. */
int  socp_b200_solve(socp_handle* h, const socp_params* params,
                     double* x, double* y, double* z, double* s,
                     int32_t* status, int32_t* iters, double* pobj, double* dobj);

/* Problem(c, A, b, G, h, cones) followed by solve_socp(prob, ss) -- reference
 * src/Socp.jl:40-59 + src/solver.jl:40-152 -- as ONE call from host data to host
 * results: what a caller of the reference's solve_socp pays end to end.  Same
 * arguments as _set_data followed by _solve.  On the fused path (sing given, no
 * sing problem) the batch is cut into chunks whose upload, solve and download
 * overlap; otherwise it is exactly _set_data + _solve.  Pinned host buffers make
 * the copies asynchronous; pageable ones work but serialise. */
int  socp_b200_solve_host(socp_handle* h, const socp_params* params,
                          const double* c, const double* A, const double* b, const double* G,
                          const double* hvec, const uint8_t* sing, int32_t flags,
                          double* x, double* y, double* z, double* s,
                          int32_t* status, int32_t* iters, double* pobj, double* dobj);

/* The one-shot call for callers that hold A and G as the reference does
 * (SparseMatrixCSC, src/Socp.jl:25,29): Problem(c, A, b, G, h, cones) +
 * solve_socp(prob, ss) with the arguments of _set_data_csc followed by _solve.
 * Pipelined like _solve_host; only the stored values cross PCIe, the dense
 * operands of every chunk are assembled on the device before its solve, and
 * the compressed-G plan of the fused kernel comes straight from the pattern. */
int  socp_b200_solve_host_csc(socp_handle* h, const socp_params* params,
                              const double* c, const socp_csc* A, const double* b, const socp_csc* G,
                              const double* hvec, const uint8_t* sing, int32_t flags,
                              double* x, double* y, double* z, double* s,
                              int32_t* status, int32_t* iters, double* pobj, double* dobj);

/* Same, but results stay on the device (inputs already resident after
 * set_data): the region bench.py times for `value`.  Fetch with _get_results. */
int  socp_b200_solve_dev(socp_handle* h, const socp_params* params);
int  socp_b200_get_results(socp_handle* h, double* x, double* y, double* z, double* s,
                           int32_t* status, int32_t* iters, double* pobj, double* dobj);
int  socp_b200_get_sing(socp_handle* h, uint8_t* sing);
int  socp_b200_timings(const socp_handle* h, socp_timings* out);

/* ---- step-level calls: the reference's plug-in seam, batched ----------------
 * All arrays are [batch][len] host arrays.  They operate on the handle's
 * problem data and on the scaling / factor left by the previous step call. */

/* compute_scaling(cones, scaling, s, z), reference src/scalings.jl:101-110 (per
 * cone :22-99; lambda/wbs/mu as in src/sqrscalings.jl:130-138).  Outputs (any
 * may be NULL): lambda[batch][k] (= scaling.l), wbs[batch][k], mu[batch][ncones]
 * (eta; 0 for POC blocks).  fail[batch] != 0 where the reference would throw. */
int  socp_b200_compute_scaling(socp_handle* h, const double* s, const double* z,
                               double* lambda, double* wbs, double* mu, int32_t* fail);
/* setup_iter(solver, prob, state, scaling), reference src/densesolver.jl:41-52:
 * H = G'W^-2 G (+A'A if sing), Cholesky, A H^-1 A', Cholesky.  fail[batch] != 0
 * on a non-positive pivot (PosDefException). */
int  socp_b200_setup_iter(socp_handle* h, int32_t* fail);
/* solve_kkt(solver, prob, state, scaling, dx,dy,dz,ds, cx,cy,cz,cs), reference
 * src/densesolver.jl:54-90. */
int  socp_b200_solve_kkt(socp_handle* h, const double* dx, const double* dy, const double* dz,
                         const double* ds, double* cx, double* cy, double* cz, double* cs);
/* scale!(cones, scaling, in, out) = W in; iscale! = W^-1 in.  src/scalings.jl:112-173 */
int  socp_b200_scale(socp_handle* h, const double* in, double* out);
int  socp_b200_iscale(socp_handle* h, const double* in, double* out);
/* out = W^-2 in (the dense iWiW gemv of src/densesolver.jl:86 in closed form) */
int  socp_b200_iwiw(socp_handle* h, const double* in, double* out);
/* The SqrScaling form of the scaling computed last (compute_scaling(cones, ::SqrScaling, s, z), reference
 * src/sqrscalings.jl:177-185; per cone :50-58 and :66-139): W^-2 = diag(D) + sum_c (u_c u_c' - v_c v_c').  D, u, v
 * are [batch][k]; the reference keeps one k-vector u_c, v_c per cone with support on that cone only
 * (src/sqrscalings.jl:119-128), here the supports are packed into one k-vector each (zero on orthant rows). */
int  socp_b200_sqr_scaling(socp_handle* h, double* D, double* u, double* v);
/* make_e!, vprod!, iprod!: reference src/vectors.jl:7-24, :58-81, :99-131 */
int  socp_b200_make_e(socp_handle* h, double* out);
int  socp_b200_vprod(socp_handle* h, const double* u, const double* v, double* out);
int  socp_b200_iprod(socp_handle* h, const double* lambda, const double* v, double* out);
/* max_step(cones, x) -> out[batch], reference src/mats.jl:1-28 */
int  socp_b200_max_step(socp_handle* h, const double* x, double* out);
/* compute_step(cones, l, ds, dz) -> out[batch], reference src/mats.jl:30-86 */
int  socp_b200_compute_step(socp_handle* h, const double* lambda, const double* ds,
                            const double* dz, double* out);
/* debug: the reduced KKT matrix H (n x n, column-major, full symmetric) and its
 * lower Cholesky factor L of the last setup_iter (GWiWiG / GWiWiGfact,
 * src/densesolver.jl:9,15) */
int  socp_b200_get_H(socp_handle* h, double* out);
int  socp_b200_get_L(socp_handle* h, double* out);

/* Parity aid (no reference counterpart as a call; it exposes the inputs and outputs of the reference's
 * compute_scaling / setup_iter / solve_kkt, src/densesolver.jl:41-90, from inside the fused whole-solve kernel): problem
 * `index` is solved again on its own and at Mehrotra iteration `iter` (0-based) the kernel writes out s, z [k] (what
 * the scaling is computed from), H [n*n, column-major] = G'W^-2 G (+A'A if sing) before its factorisation, the
 * right-hand side dx [n], dy [p], dz [k], ds [k] and the result cx, cy, cz, cs of the affine (phase 1) or combined
 * (phase 2) solve_kkt.  Null outputs are skipped.  Needs a layout the fused kernel takes (16 < n <= 64). */
int  socp_b200_debug_fused_step(socp_handle* h, int64_t index, int32_t iter, int32_t phase,
                                double* s, double* z, double* H, double* dx, double* dy, double* dz, double* ds,
                                double* cx, double* cy, double* cz, double* cs);

/* Measurement utility (no reference counterpart): mean CUDA-event time, in ms, of
 * the device kernels behind one step-level call over `reps` repetitions on the data
 * resident after _set_data + _compute_scaling.  which: 0 compute_scaling, 1 scale!,
 * 2 iscale!, 3 vprod!, 4 iprod!, 5 compute_step, 6 W^-1 G, 7 SYRK G'W^-2 G,
 * 8 Cholesky, 9 L L' solve (1 rhs), 10 G'v, 11 G v.  Single-device handles only. */
int  socp_b200_profile_step(socp_handle* h, int32_t which, int32_t reps, double* ms_per_call);

#ifdef __cplusplus
}
#endif
#endif /* SOCP_B200_H */
