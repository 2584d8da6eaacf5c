// tiled_kernels.cuh -- per-problem cone / driver kernels of the tiled path.
//
// One CTA per problem; warp w owns cones w, w+nw, ... (cone_ops.cuh).  Vectors
// live in global memory ([batch][len]); a problem whose `active` flag is 0 is
// skipped by every kernel (per-problem early exit of the Mehrotra loop,
// reference src/solver.jl:122-124).
#pragma once
#include "cone_ops.cuh"

namespace socp {

// All device pointers of one device shard.  Passed by value to the kernels.

#define SOCP_VEC(ptr, len) ((ptr) + (int64_t)b * (len))

// ------------------------------------------------------------ step-level kernels
__device__ __forceinline__ void dev_scaling(const ConeLayout& L, int b, const double* s, const double* z,
                                            double* lam, double* wb, double* iwb, double* eta, int* fail) {
    const int lane = threadIdx.x & 31;
    const double* sb = SOCP_VEC(s, L.k);
    const double* zb = SOCP_VEC(z, L.k);
    double* lb = SOCP_VEC(lam, L.k);
    double* wbb = SOCP_VEC(wb, L.k);
    double* iwbb = SOCP_VEC(iwb, L.k);
    double* eb = SOCP_VEC(eta, 4 * L.ncones);
    int f = 0;
    SOCP_FOR_EACH_CONE(L, c, kind, offs, dim) {
        if (kind == KIND_POC) {
            f |= warp_poc_scaling(sb + offs, zb + offs, dim, lane, lb + offs, wbb + offs, iwbb + offs);
            if (lane == 0) eb[c] = 0.0;
        } else {
            f |= warp_soc_scaling(sb + offs, zb + offs, dim, lane, lb + offs, wbb + offs, eb + c, L.ncones);
        }
    }
    if (f && lane == 0) atomicOr(fail + b, 1);
}
__global__ void k_scaling(ConeLayout L, const double* __restrict__ s, const double* __restrict__ z,
                          double* __restrict__ lam, double* __restrict__ wb, double* __restrict__ iwb,
                          double* __restrict__ eta, int* __restrict__ fail, const int* __restrict__ active) {
    const int b = blockIdx.x;
    if (active && !active[b]) return;
    dev_scaling(L, b, s, z, lam, wb, iwb, eta, fail);
}

template <int MODE>
__global__ void k_apply(ConeLayout L, const double* __restrict__ wb, const double* __restrict__ iwb,
                        const double* __restrict__ eta, const double* v, double* out) {
    const int b = blockIdx.x;
    cta_apply<MODE>(L, SOCP_VEC(wb, L.k), SOCP_VEC(iwb, L.k), SOCP_VEC(eta, 4 * L.ncones), SOCP_VEC(v, L.k),
                    SOCP_VEC(out, L.k));
}
__global__ void k_vprod(ConeLayout L, const double* u, const double* v, double* t) {
    const int b = blockIdx.x;
    cta_vprod(L, SOCP_VEC(u, L.k), SOCP_VEC(v, L.k), SOCP_VEC(t, L.k));
}
__global__ void k_iprod(ConeLayout L, const double* lam, const double* v, double* t) {
    const int b = blockIdx.x;
    cta_iprod(L, SOCP_VEC(lam, L.k), SOCP_VEC(v, L.k), SOCP_VEC(t, L.k));
}
__global__ void k_make_e(ConeLayout L, double* out) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31;
    double* o = SOCP_VEC(out, L.k);
    SOCP_FOR_EACH_CONE(L, c, kind, offs, dim) {
        for (int i = lane; i < dim; i += 32) o[offs + i] = e_value(kind, i);
    }
}
__global__ void k_max_step(ConeLayout L, const double* x, double* out) {
    __shared__ double scratch[128];
    const int b = blockIdx.x;
    const double m = block_max(cta_max_step_partial(L, SOCP_VEC(x, L.k)), scratch);
    if (threadIdx.x == 0) out[b] = m;
}
__global__ void k_compute_step(ConeLayout L, const double* lam, const double* ds, const double* dz, double* out) {
    __shared__ double scratch[128];
    const int b = blockIdx.x;
    int f = 0;
    const double* lb = SOCP_VEC(lam, L.k);
    double m = fmax(cta_scmax_partial(L, lb, SOCP_VEC(ds, L.k), &f), cta_scmax_partial(L, lb, SOCP_VEC(dz, L.k), &f));
    m = block_max(m, scratch);
    if (threadIdx.x == 0) out[b] = step_from_t(m);
}

// ------------------------------------------------------------ Gt = W^-1 G (or copy)
// grid (ceil(n/COLS), batch); warps loop over (column, cone) pairs.
// identity != 0: Gt = G (initial point / sing detection, W = I).
__global__ void __launch_bounds__(256)
k_build_gt(ConeLayout L, const double* __restrict__ G, int64_t sG, const double* __restrict__ wb,
           const double* __restrict__ iwb, const double* __restrict__ eta, double* __restrict__ Gt, int ldgt,
           int identity, int cols_per_cta, const int* __restrict__ active, int nbatch) {
    const int b = batch_index();
    if (b >= nbatch) return;
    if (active && !active[b]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double* Gb = G + (int64_t)b * sG;
    double* Gtb = Gt + (int64_t)b * ldgt * L.n;
    const double* wbb = SOCP_VEC(wb, L.k);
    const double* iwbb = SOCP_VEC(iwb, L.k);
    const double* eb = SOCP_VEC(eta, 4 * L.ncones);
    const int c0 = blockIdx.x * cols_per_cta;
    const int c1 = min(L.n, c0 + cols_per_cta);
    const int ntask = (c1 - c0) * L.ncones;
    for (int task = warp; task < ntask; task += nw) {
        const int col = c0 + task / L.ncones, c = task % L.ncones;
        const int kind = L.kind[c], offs = L.offs[c], dim = L.dim[c];
        const double* src = Gb + (int64_t)col * L.k + offs;
        double* dst = Gtb + (int64_t)col * ldgt + offs;
        if (identity) {
            for (int i = lane; i < dim; i += 32) dst[i] = src[i];
        } else if (kind == KIND_POC) {
            warp_poc_apply<APPLY_WINV>(wbb + offs, iwbb + offs, src, dst, dim, lane);
        } else {
            warp_soc_apply<APPLY_WINV>(wbb + offs, eb + c, L.ncones, src, dst, dim, lane);
        }
    }
}

// padded copy of A (p x n, ld p) into Ap (ld ldap), pad rows left zero
__global__ void k_pad_copy(const double* __restrict__ A, int64_t sA, int rows, int cols, double* __restrict__ Ap, int ldap, int nbatch) {
    const int b = batch_index();
    if (b >= nbatch) return;
    const double* Ab = A + (int64_t)b * sA;
    double* Apb = Ap + (int64_t)b * ldap * cols;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < rows * cols; q += gridDim.x * blockDim.x) {
        const int r = q % rows, c = q / rows;
        Apb[(int64_t)c * ldap + r] = Ab[(int64_t)c * rows + r];
    }
}
// HiAt[:, i] = A[i, :]   (n x p, ld n)
__global__ void k_transpose_A(const double* __restrict__ A, int64_t sA, int p, int n, double* __restrict__ out,
                              const int* __restrict__ active, int nbatch) {
    const int b = batch_index();
    if (b >= nbatch) return;
    if (active && !active[b]) return;
    const double* Ab = A + (int64_t)b * sA;
    double* ob = out + (int64_t)b * n * p;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < p * n; q += gridDim.x * blockDim.x) {
        const int c = q % n, i = q / n;
        ob[(int64_t)i * n + c] = Ab[(int64_t)c * p + i];
    }
}
// M (p x p, ld ldm) = A (p x n) * HiAt (n x p); one thread per entry (i fastest)
__global__ void k_small_gemm(const double* __restrict__ A, int64_t sA, int p, int n, const double* __restrict__ B,
                             double* __restrict__ M, int ldm, const int* __restrict__ active, int nbatch) {
    const int b = batch_index();
    if (b >= nbatch) return;
    if (active && !active[b]) return;
    const double* Ab = A + (int64_t)b * sA;
    const double* Bb = B + (int64_t)b * n * p;
    double* Mb = M + (int64_t)b * ldm * p;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < p * p; q += gridDim.x * blockDim.x) {
        const int i = q % p, j = q / p;
        double acc = 0.0;
        for (int c = 0; c < n; ++c) acc = fma(Ab[(int64_t)c * p + i], Bb[(int64_t)j * n + c], acc);
        Mb[(int64_t)j * ldm + i] = acc;
    }
}

// ------------------------------------------------------------ driver kernels
// solve_kkt head, reference src/densesolver.jl:61-66: k0 = lam \ ds; k1 = W k0;
// k2 = dz - k1; u = W^-2 k2 (the vector G' is applied to next).
__device__ __forceinline__ void kkt_head(const Ws& w, int b) {
    const ConeLayout& L = w.L;
    const int lane = threadIdx.x & 31;
    const double* lam = SOCP_VEC(w.lam, L.k);
    const double* wb = SOCP_VEC(w.wb, L.k);
    const double* iwb = SOCP_VEC(w.iwb, L.k);
    const double* eta = SOCP_VEC(w.eta, 4 * L.ncones);
    const int nc = L.ncones;
    const double* ds = SOCP_VEC(w.ds, L.k);
    const double* dz = SOCP_VEC(w.dz, L.k);
    double* k0 = SOCP_VEC(w.k0, L.k);
    double* k2 = SOCP_VEC(w.k2, L.k);
    double* u = SOCP_VEC(w.u, L.k);
    SOCP_FOR_EACH_CONE(L, c, kind, offs, dim) {
        if (kind == KIND_POC) {
            warp_poc_iprod(lam + offs, ds + offs, k0 + offs, dim, lane);
            __syncwarp();
            warp_poc_apply<APPLY_W>(wb + offs, iwb + offs, k0 + offs, k2 + offs, dim, lane);
        } else {
            warp_soc_iprod(lam + offs, ds + offs, k0 + offs, dim, lane);
            __syncwarp();
            warp_soc_apply<APPLY_W>(wb + offs, eta + c, nc, k0 + offs, k2 + offs, dim, lane);
        }
        __syncwarp();
        for (int i = lane; i < dim; i += 32) k2[offs + i] = dz[offs + i] - k2[offs + i];
        __syncwarp();
        if (kind == KIND_POC) warp_poc_apply<APPLY_WINV2>(wb + offs, iwb + offs, k2 + offs, u + offs, dim, lane);
        else warp_soc_apply<APPLY_WINV2>(wb + offs, eta + c, nc, k2 + offs, u + offs, dim, lane);
    }
}
// solve_kkt tail, reference src/densesolver.jl:86-89: on entry u = G cx - k2;
// cz = W^-2 u; cs = W (k0 - W cz).  Writes rz (= cz), rs (= cs); then the
// driver's kt3 = W cz, kt2 = W^-1 cs (src/solver.jl:128-129).  W cz is computed
// once and reused for both.
__device__ __forceinline__ void kkt_tail(const Ws& w, int b) {
    const ConeLayout& L = w.L;
    const int lane = threadIdx.x & 31;
    const double* wb = SOCP_VEC(w.wb, L.k);
    const double* iwb = SOCP_VEC(w.iwb, L.k);
    const double* eta = SOCP_VEC(w.eta, 4 * L.ncones);
    const int nc = L.ncones;
    const double* u = SOCP_VEC(w.u, L.k);
    double* k0 = SOCP_VEC(w.k0, L.k);
    double* rz = SOCP_VEC(w.rz, L.k);
    double* rs = SOCP_VEC(w.rs, L.k);
    double* kt3 = SOCP_VEC(w.kt3, L.k);
    double* kt2 = SOCP_VEC(w.kt2, L.k);
    SOCP_FOR_EACH_CONE(L, c, kind, offs, dim) {
        const bool poc = (kind == KIND_POC);
        if (poc) warp_poc_apply<APPLY_WINV2>(wb + offs, iwb + offs, u + offs, rz + offs, dim, lane);
        else warp_soc_apply<APPLY_WINV2>(wb + offs, eta + c, nc, u + offs, rz + offs, dim, lane);
        __syncwarp();
        if (poc) warp_poc_apply<APPLY_W>(wb + offs, iwb + offs, rz + offs, kt3 + offs, dim, lane);
        else warp_soc_apply<APPLY_W>(wb + offs, eta + c, nc, rz + offs, kt3 + offs, dim, lane);
        __syncwarp();
        for (int i = lane; i < dim; i += 32) k0[offs + i] -= kt3[offs + i];      // :88  (k0 - W cz = W^-1 cs = kt2)
        __syncwarp();
        if (poc) warp_poc_apply<APPLY_W>(wb + offs, iwb + offs, k0 + offs, rs + offs, dim, lane);
        else warp_soc_apply<APPLY_W>(wb + offs, eta + c, nc, k0 + offs, rs + offs, dim, lane);
        __syncwarp();
        // kt2 = W^-1 cs, reference src/solver.jl:129 (applied, not shortcut, to keep the rounding)
        if (poc) warp_poc_apply<APPLY_WINV>(wb + offs, iwb + offs, rs + offs, kt2 + offs, dim, lane);
        else warp_soc_apply<APPLY_WINV>(wb + offs, eta + c, nc, rs + offs, kt2 + offs, dim, lane);
    }
}

// Initial shift, reference src/solver.jl:86-104.  On entry z holds z0 = G x - h.
__device__ __forceinline__ void dev_init_shift(const Ws& w, int b, const LoopParams& P, double* scratch) {
    const ConeLayout& L = w.L;
    const int lane = threadIdx.x & 31;
    double* z = SOCP_VEC(w.z, L.k);
    double* s = SOCP_VEC(w.s, L.k);
    // s temporarily holds -z0 so that max_step(-z0) can use the same routine
    for (int i = threadIdx.x; i < L.k; i += blockDim.x) s[i] = -z[i];
    __syncthreads();
    const double alphp = block_max(cta_max_step_partial(L, s), scratch);     // :88
    const double alphd = block_max(cta_max_step_partial(L, z), scratch);     // :89
    const bool shp = !(fabs(alphp) < P.init_eps), shd = !(fabs(alphd) < P.init_eps);
    SOCP_FOR_EACH_CONE(L, c, kind, offs, dim) {
        for (int i = lane; i < dim; i += 32) {
            const double e = e_value(kind, i);
            const double z0 = z[offs + i];
            s[offs + i] = shp ? (-z0 + (1.0 + alphp) * e) : -z0;             // :91-95
            z[offs + i] = shd ? (z0 + (1.0 + alphd) * e) : z0;               // :97-101
        }
    }
}
__global__ void k_init_shift(Ws w, LoopParams P) {
    __shared__ double scratch[128];
    const int b = blockIdx.x;
    if (!w.active[b]) return;
    dev_init_shift(w, b, P, scratch);
}

// After the residual gemvs (dx,dy,dz hold the NEGATED residuals, src/solver.jl:125):
// stop test (:122-124), ds = -lambda o lambda (:120,:125), solve_kkt head.
__device__ __forceinline__ int dev_pre(const Ws& w, int b, const LoopParams& P, int it, double* scratch) {
    const ConeLayout& L = w.L;
    if (w.fail[b]) {                      // compute_scaling threw
        if (threadIdx.x == 0) { w.status[b] = ST_NUMERICAL; w.active[b] = 0; }
        return 0;
    }
    const double* dx = SOCP_VEC(w.dx, L.n);
    const double* dy = SOCP_VEC(w.dy, L.p);
    const double* z = SOCP_VEC(w.z, L.k);
    const double* s = SOCP_VEC(w.s, L.k);
    const double* lam = SOCP_VEC(w.lam, L.k);
    double* ds = SOCP_VEC(w.ds, L.k);
    double nx = 0.0, ny = 0.0, gap = 0.0, ll = 0.0;
    for (int i = threadIdx.x; i < L.n; i += blockDim.x) nx = fma(dx[i], dx[i], nx);
    for (int i = threadIdx.x; i < L.p; i += blockDim.x) ny = fma(dy[i], dy[i], ny);
    for (int i = threadIdx.x; i < L.k; i += blockDim.x) {
        gap = fma(z[i], s[i], gap);
        ll = fma(lam[i], lam[i], ll);
    }
    block_sum4(nx, ny, gap, ll, scratch);
    const double resid = sqrt(nx) + sqrt(ny) + gap;
    if (threadIdx.x == 0) {
        w.sc[b].resid = resid;
        w.sc[b].gap = gap;
        w.sc[b].ll = ll;
    }
    if (resid < P.tol) {                  // :122-124
        if (threadIdx.x == 0) { w.status[b] = ST_CONVERGED; w.active[b] = 0; }
        return 0;
    }
    if (threadIdx.x == 0 && w.nactive) atomicAdd(w.nactive + it, 1);
    cta_vprod(L, lam, lam, ds);           // :120
    __syncthreads();
    for (int i = threadIdx.x; i < L.k; i += blockDim.x) ds[i] = -ds[i];      // :125
    __syncthreads();
    kkt_head(w, b);
    return 1;
}
__global__ void k_pre(Ws w, LoopParams P, int it) {
    __shared__ double scratch[128];
    const int b = blockIdx.x;
    if (!w.active[b]) return;
    dev_pre(w, b, P, it, scratch);
}

// Between the two solves: finish solve #1, centering parameter (:130-134),
// combined right-hand side (:136-140), head of solve #2.
__device__ __forceinline__ int dev_mid(const Ws& w, int b, const LoopParams& P, double* scratch, int* iscratch) {
    const ConeLayout& L = w.L;
    if (w.fail[b]) {                      // cholesky! threw
        if (threadIdx.x == 0) { w.status[b] = ST_NUMERICAL; w.active[b] = 0; }
        return 0;
    }
    const int lane = threadIdx.x & 31;
    kkt_tail(w, b);
    __syncthreads();
    const double* lam = SOCP_VEC(w.lam, L.k);
    double* kt2 = SOCP_VEC(w.kt2, L.k);
    double* kt3 = SOCP_VEC(w.kt3, L.k);
    double* ds = SOCP_VEC(w.ds, L.k);
    int f = 0;
    double m = fmax(cta_scmax_partial(L, lam, kt3, &f), cta_scmax_partial(L, lam, kt2, &f));
    m = block_max(m, scratch);
    f = block_or(f, iscratch);
    const double t = step_from_t(m);      // :130
    double dot = 0.0;
    for (int i = threadIdx.x; i < L.k; i += blockDim.x) dot = fma(kt2[i], kt3[i], dot);
    dot = block_sum(dot, scratch);
    const double ll = w.sc[b].ll;
    const double rho = 1.0 - t - t * t * dot * fast_rcp(ll);   // :132 (minus: reference quirk)
    const double cl = fmax(0.0, fmin(1.0, rho));
    const double sig = cl * cl * cl;                           // :133
    const double mu = ll / (double)L.deg;                      // :134
    const double scf = 1.0 - sig;                              // :136
    if (threadIdx.x == 0) { w.sc[b].t = t; w.sc[b].sigma = sig; w.sc[b].mu = mu; }
    if (f) {
        if (threadIdx.x == 0) { w.status[b] = ST_NUMERICAL; w.active[b] = 0; }
        return 0;
    }
    // ds += sig*mu*e - kt2 o kt3      :137-139   (kt2 is reused as scratch for the product)
    cta_vprod(L, kt2, kt3, kt2);
    __syncthreads();
    const double sm = sig * mu;
    SOCP_FOR_EACH_CONE(L, c, kind, offs, dim) {
        for (int i = lane; i < dim; i += 32) ds[offs + i] += sm * e_value(kind, i) - kt2[offs + i];
    }
    double* dx = SOCP_VEC(w.dx, L.n);
    double* dy = SOCP_VEC(w.dy, L.p);
    double* dz = SOCP_VEC(w.dz, L.k);
    for (int i = threadIdx.x; i < L.n; i += blockDim.x) dx[i] *= scf;        // :140
    for (int i = threadIdx.x; i < L.p; i += blockDim.x) dy[i] *= scf;
    for (int i = threadIdx.x; i < L.k; i += blockDim.x) dz[i] *= scf;
    __syncthreads();
    kkt_head(w, b);
    return 1;
}
__global__ void k_mid(Ws w, LoopParams P) {
    __shared__ double scratch[128];
    __shared__ int iscratch[32];
    const int b = blockIdx.x;
    if (!w.active[b]) return;
    dev_mid(w, b, P, scratch, iscratch);
}

// After solve #2: tail, step length (:143-146), iterate update (:147-150).
__device__ __forceinline__ int dev_post(const Ws& w, int b, const LoopParams& P, double* scratch, int* iscratch) {
    const ConeLayout& L = w.L;
    kkt_tail(w, b);
    __syncthreads();
    const double* lam = SOCP_VEC(w.lam, L.k);
    const double* kt2 = SOCP_VEC(w.kt2, L.k);
    const double* kt3 = SOCP_VEC(w.kt3, L.k);
    int f = 0;
    double m = fmax(cta_scmax_partial(L, lam, kt3, &f), cta_scmax_partial(L, lam, kt2, &f));
    m = block_max(m, scratch);
    const double step = step_from_t(m) * P.step_damp;          // :145-146
    double* x = SOCP_VEC(w.x, L.n);
    double* y = SOCP_VEC(w.y, L.p);
    double* z = SOCP_VEC(w.z, L.k);
    double* s = SOCP_VEC(w.s, L.k);
    const double* rx = SOCP_VEC(w.rx, L.n);
    const double* ry = SOCP_VEC(w.ry, L.p);
    const double* rz = SOCP_VEC(w.rz, L.k);
    const double* rs = SOCP_VEC(w.rs, L.k);
    // the reference would carry NaN/Inf into the next cholesky! and throw there
    for (int i = threadIdx.x; i < L.n; i += blockDim.x) f |= !isfinite(rx[i]);
    for (int i = threadIdx.x; i < L.p; i += blockDim.x) f |= !isfinite(ry[i]);
    for (int i = threadIdx.x; i < L.k; i += blockDim.x) f |= !isfinite(rz[i]) | !isfinite(rs[i]);
    f |= !isfinite(step);
    f = block_or(f, iscratch);
    if (f) {
        if (threadIdx.x == 0) { w.status[b] = ST_NUMERICAL; w.active[b] = 0; }
        return 0;
    }
    for (int i = threadIdx.x; i < L.n; i += blockDim.x) x[i] = fma(rx[i], step, x[i]);   // :147
    for (int i = threadIdx.x; i < L.p; i += blockDim.x) y[i] = fma(ry[i], step, y[i]);   // :148
    for (int i = threadIdx.x; i < L.k; i += blockDim.x) {
        z[i] = fma(rz[i], step, z[i]);                                                    // :149
        s[i] = fma(rs[i], step, s[i]);                                                    // :150
    }
    if (threadIdx.x == 0) { w.sc[b].step = step; w.iters[b] += 1; }
    return 1;
}
__global__ void k_post(Ws w, LoopParams P) {
    __shared__ double scratch[128];
    __shared__ int iscratch[32];
    const int b = blockIdx.x;
    if (!w.active[b]) return;
    dev_post(w, b, P, scratch, iscratch);
}

// Problems that failed in the initial factorisation / still running at the end.
__global__ void k_finalize(Ws w, int phase) {
    __shared__ double scratch[128];
    const int b = blockIdx.x;
    const ConeLayout& L = w.L;
    if (phase == 0) {          // after the initial-point factorisation
        if (threadIdx.x == 0 && w.fail[b]) { w.status[b] = ST_NUMERICAL; w.active[b] = 0; }
        return;
    }
    if (threadIdx.x == 0 && w.active[b]) { w.status[b] = ST_MAXITER; w.active[b] = 0; }
    const double* c = SOCP_VEC(w.c, L.n);
    const double* bb = SOCP_VEC(w.b, L.p);
    const double* h = SOCP_VEC(w.h, L.k);
    const double* x = SOCP_VEC(w.x, L.n);
    const double* y = SOCP_VEC(w.y, L.p);
    const double* z = SOCP_VEC(w.z, L.k);
    double po = 0.0, d = 0.0;
    for (int i = threadIdx.x; i < L.n; i += blockDim.x) po = fma(c[i], x[i], po);
    for (int i = threadIdx.x; i < L.p; i += blockDim.x) d = fma(-bb[i], y[i], d);
    for (int i = threadIdx.x; i < L.k; i += blockDim.x) d = fma(-h[i], z[i], d);
    po = block_sum(po, scratch);
    d = block_sum(d, scratch);
    if (threadIdx.x == 0) { w.sc[b].pobj = po; w.sc[b].dobj = d; w.pobj[b] = po; w.dobj[b] = d; }
}

// reset of the per-problem state at the start of a solve
__global__ void k_reset(Ws w, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    w.status[b] = ST_RUNNING;
    w.iters[b] = 0;
    w.active[b] = 1;
    w.fail[b] = 0;
}
__global__ void k_fail_to_sing(const int* fail, uint8_t* sing, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) sing[b] = fail[b] ? 1 : 0;
}
__global__ void k_sing_to_active(const uint8_t* sing, int* active, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) active[b] = sing[b] ? 1 : 0;
}

}  // namespace socp
