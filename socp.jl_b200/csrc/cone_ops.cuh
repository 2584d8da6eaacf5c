// cone_ops.cuh -- warp-cooperative Jordan-algebra / Nesterov-Todd primitives.
//
// Every function is called by all 32 lanes of one warp for ONE cone block; the
// vectors may live in global or shared memory (generic pointers, already offset
// to the first entry of the block).  Reductions are xor-butterfly shuffles, so
// every lane holds the same, order-deterministic result.  Formulas follow the
// reference (/root/reference = BenChung/Socp.jl), cited per function; the O(d^2)
// loops of the reference are replaced by their O(d) closed forms (SURVEY.md
// appendix A), which are the same maps.
#pragma once
#include "common.cuh"

namespace socp {

// ---------------------------------------------------------------------------
// compute_scaling                      reference src/scalings.jl:22-30 (POC),
//                                      :32-99 (SOC)
// Outputs: lam (scaling.l), wb (scaling.wbs), *eta (scaling.mu[cind]; 0 for POC).
// Returns nonzero in every lane where the reference would throw a DomainError
// (sqrt of a negative) -- NaN inputs are flagged as well.
// ---------------------------------------------------------------------------
// iwb receives 1/wb (the diagonal of W^-1, sqrt(z/s) in the reference) so that the applies multiply.
__device__ __forceinline__ int warp_poc_scaling(const double* s, const double* z, int d, int lane,
                                                double* lam, double* wb, double* iwb) {
    int fail = 0;
    for (int i = lane; i < d; i += 32) {
        const double si = s[i], zi = z[i];
        const double q = si * fast_rcp(zi), qi = zi * fast_rcp(si), pz = si * zi;
        fail |= !(q >= 0.0) | !(pz >= 0.0);
        wb[i] = fast_sqrt(q);       // W = sqrt(s/z)
        iwb[i] = fast_sqrt(qi);     // W^-1 = sqrt(z/s)                    :26
        lam[i] = fast_sqrt(pz);
    }
    return warp_or(fail);
}

// Per-cone scalars: es[0] = eta (scaling.mu), es[nc] = 1/eta, es[2nc] = 1/eta^2, es[3nc] = 1/(1 + wbar_0).
__device__ __forceinline__ int warp_soc_scaling(const double* s, const double* z, int d, int lane,
                                                double* lam, double* wb, double* es, int nc) {
    const double s0 = s[0], z0 = z[0];
    double ss = 0.0, zz = 0.0;
    for (int i = 1 + lane; i < d; i += 32) {
        const double si = s[i], zi = z[i];
        ss = fma(si, si, ss);
        zz = fma(zi, zi, zz);
    }
    ss = warp_sum(ss);
    zz = warp_sum(zz);
    const double onrms = s0 * s0 - ss;            // :39-45
    const double onrmz = z0 * z0 - zz;
    int fail = !(onrms >= 0.0) | !(onrmz >= 0.0);
    const double is = fast_rsqrt(onrms), iz = fast_rsqrt(onrmz);   // 1/nrms, 1/nrmz      :46-49
    const double nrms = onrms * is, nrmz = onrmz * iz;
    const double sb0 = s0 * is, zb0 = z0 * iz;
    double ns = 0.0;
    for (int i = 1 + lane; i < d; i += 32) ns = fma(z[i] * iz, s[i] * is, ns);
    ns = warp_sum(ns) + zb0 * sb0;                // :53-56
    const double g2 = (1.0 + ns) / 2.0;
    fail |= !(g2 >= 0.0);
    const double rg = fast_rsqrt(g2);
    const double gamma = g2 * rg;                 // :57
    const double ig = 0.5 * rg;                   // 1/(2 gamma)   :64
    const double eta = fast_sqrt(nrms * iz);      // sqrt(nrms/nrmz)  :68
    const double tmv1 = fast_sqrt(nrms * nrmz);   // :91
    const double mult = tmv1 * fast_rcp(zb0 + sb0 + 2.0 * gamma);   // :92
    const double cs = gamma + zb0, cz = gamma + sb0;        // :93-94
    for (int i = 1 + lane; i < d; i += 32) {
        const double sb = s[i] * is, zb = z[i] * iz;
        wb[i] = (sb - zb) * ig;                   // :62,:64
        lam[i] = (sb * cs + zb * cz) * mult;      // :95-97
    }
    if (lane == 0) {
        const double w0 = (sb0 + zb0) * ig;
        wb[0] = w0;                               // :60
        lam[0] = gamma * tmv1;                    // :98
        const double ie = fast_rcp(eta);
        es[0] = eta;                              // :69
        es[nc] = ie;
        es[2 * nc] = ie * ie;
        es[3 * nc] = fast_rcp(1.0 + w0);
    }
    return fail;
}

// ---------------------------------------------------------------------------
// scale! / iscale! / W^-2 apply        reference src/scalings.jl:112-156,
//                                      src/densesolver.jl:86 (iWiW gemv)
// in-place (out == v) is allowed.
// ---------------------------------------------------------------------------
enum ApplyMode { APPLY_W = 0, APPLY_WINV = 1, APPLY_WINV2 = 2 };

template <int MODE>
__device__ __forceinline__ void warp_poc_apply(const double* wb, const double* iwb, const double* v, double* out,
                                               int d, int lane) {
    for (int i = lane; i < d; i += 32) {
        const double vi = v[i];
        double r;
        if (MODE == APPLY_W) r = wb[i] * vi;                       // :112-117
        else if (MODE == APPLY_WINV) r = iwb[i] * vi;              // 1/wb * v  :119-124
        else { const double iw = iwb[i]; r = iw * iw * vi; }       // (iW*iW')[i,i] = iW[i]^2
        out[i] = r;
    }
}

template <int MODE>
__device__ __forceinline__ void warp_soc_apply(const double* wb, const double* es, int nc, const double* v, double* out,
                                               int d, int lane) {
    const double v0 = v[0], w0 = wb[0];
    const double eta = es[0], ie = es[nc], ie2 = es[2 * nc], r1w = es[3 * nc];
    double dl = 0.0;
    for (int i = 1 + lane; i < d; i += 32) dl = fma(wb[i], v[i], dl);
    dl = warp_sum(dl);                                          // :129-132 / :145-148
    if (MODE == APPLY_W) {
        const double cst = v0 + dl * r1w;                       // v0 + del/(1+wb0)  :135
        for (int i = 1 + lane; i < d; i += 32) out[i] = eta * (v[i] + cst * wb[i]);   // :137-139
        if (lane == 0) out[0] = eta * (w0 * v0 + dl);           // :136
    } else if (MODE == APPLY_WINV) {
        const double cst = -v0 + dl * r1w;                      // :151
        for (int i = 1 + lane; i < d; i += 32) out[i] = ie * (v[i] + cst * wb[i]);    // :153-155
        if (lane == 0) out[0] = ie * (w0 * v0 - dl);            // :152
    } else {
        // W^-2 = eta^-2 (2 q q' - J), q = J wbar  
        const double qv = w0 * v0 - dl;
        for (int i = 1 + lane; i < d; i += 32) out[i] = ie2 * (v[i] - 2.0 * wb[i] * qv);
        if (lane == 0) out[0] = ie2 * (2.0 * w0 * qv - v0);
    }
}

// ---------------------------------------------------------------------------
// vprod! / iprod!                      reference src/vectors.jl:58-81, :99-131
// ---------------------------------------------------------------------------
__device__ __forceinline__ void warp_poc_vprod(const double* u, const double* v, double* t, int d, int lane) {
    for (int i = lane; i < d; i += 32) t[i] = u[i] * v[i];
}
__device__ __forceinline__ void warp_soc_vprod(const double* u, const double* v, double* t, int d, int lane) {
    const double u0 = u[0], v0 = v[0];
    double acc = 0.0;
    for (int i = 1 + lane; i < d; i += 32) {
        const double ui = u[i], vi = v[i];
        acc = fma(ui, vi, acc);
        t[i] = u0 * vi + v0 * ui;              // :73-75
    }
    acc = warp_sum(acc) + u0 * v0;             // :66-69
    if (lane == 0) t[0] = acc;
}
__device__ __forceinline__ void warp_poc_iprod(const double* lam, const double* v, double* t, int d, int lane) {
    for (int i = lane; i < d; i += 32) t[i] = v[i] * fast_rcp(lam[i]);   // v/lam  :99-103
}
__device__ __forceinline__ void warp_soc_iprod(const double* lam, const double* v, double* t, int d, int lane) {
    const double l0 = lam[0], v0 = v[0];
    double ll = 0.0, beta = 0.0;
    for (int i = 1 + lane; i < d; i += 32) {
        const double li = lam[i];
        ll = fma(li, li, ll);
        beta = fma(li, v[i], beta);
    }
    ll = warp_sum(ll);
    beta = warp_sum(beta);
    const double a = l0 * l0 - ll;             // :108-111
    const double ia = fast_rcp(a), il0 = fast_rcp(l0);
    // closed form of the double loop :115-124
    for (int i = 1 + lane; i < d; i += 32) {
        const double li = lam[i];
        t[i] = (-v0 * li + (a * v[i] + beta * li) * il0) * ia;
    }
    if (lane == 0) t[0] = (l0 * v0 - beta) * ia;
}

// ---------------------------------------------------------------------------
// max_step / scmax                     reference src/mats.jl:1-28, :42-86
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_poc_max_step(const double* x, int d, int lane) {
    double mn = INFINITY;
    for (int i = lane; i < d; i += 32) mn = fmin(mn, x[i]);
    return -warp_min(mn);
}
__device__ __forceinline__ double warp_soc_max_step(const double* x, int d, int lane) {
    double sq = 0.0;
    for (int i = 1 + lane; i < d; i += 32) { const double xi = x[i]; sq = fma(xi, xi, sq); }
    sq = warp_sum(sq);
    return fast_sqrt(sq) - x[0];
}
__device__ __forceinline__ double warp_poc_scmax(const double* l, const double* x, int d, int lane) {
    double mx = -INFINITY;
    for (int i = lane; i < d; i += 32) mx = fmax(mx, -x[i] * fast_rcp(l[i]));   // -x/l  :53-62
    return warp_max(mx);
}
__device__ __forceinline__ double warp_soc_scmax(const double* l, const double* x, int d, int lane, int* fail) {
    const double l0 = l[0], x0 = x[0];
    double ll = 0.0, lx = 0.0;
    for (int i = 1 + lane; i < d; i += 32) {
        const double li = l[i];
        ll = fma(li, li, ll);
        lx = fma(li, x[i], lx);
    }
    ll = warp_sum(ll);
    lx = warp_sum(lx);
    const double ai = l0 * l0 - ll;                 // :67-70
    *fail |= !(ai >= 0.0);
    const double a = fast_rsqrt(ai);                // 1/sqrt(ai)  :71
    const double r1 = a * l0 * x0 - a * lx;         // :74-77
    const double cst = (r1 + x0) * fast_rcp(a * l0 + 1.0);  // :80
    double r2s = 0.0;
    for (int i = 1 + lane; i < d; i += 32) {
        const double q = a * (x[i] - cst * a * l[i]);   // :83
        r2s = fma(q, q, r2s);
    }
    r2s = warp_sum(r2s);
    return fast_sqrt(r2s) - a * r1;                 // :85
}

// ---------------------------------------------------------------------------
// Dispatch helpers over a ConeLayout: warp w of the CTA owns cones w, w+nw, ...
// ---------------------------------------------------------------------------
#define SOCP_FOR_EACH_CONE(L, c, kind_, offs_, dim_)                                     \
    for (int c = (threadIdx.x >> 5); c < (L).ncones; c += (blockDim.x >> 5))            \
        if (int kind_ = (L).kind[c], offs_ = (L).offs[c], dim_ = (L).dim[c]; true)

template <int MODE>
__device__ __forceinline__ void cta_apply(const ConeLayout& L, const double* wb, const double* iwb, const double* eta,
                                          const double* v, double* out) {
    const int lane = threadIdx.x & 31;
    SOCP_FOR_EACH_CONE(L, c, kind, offs, dim) {
        if (kind == KIND_POC) warp_poc_apply<MODE>(wb + offs, iwb + offs, v + offs, out + offs, dim, lane);
        else warp_soc_apply<MODE>(wb + offs, eta + c, L.ncones, v + offs, out + offs, dim, lane);
    }
}
__device__ __forceinline__ void cta_vprod(const ConeLayout& L, const double* u, const double* v, double* t) {
    const int lane = threadIdx.x & 31;
    SOCP_FOR_EACH_CONE(L, c, kind, offs, dim) {
        if (kind == KIND_POC) warp_poc_vprod(u + offs, v + offs, t + offs, dim, lane);
        else warp_soc_vprod(u + offs, v + offs, t + offs, dim, lane);
    }
}
__device__ __forceinline__ void cta_iprod(const ConeLayout& L, const double* lam, const double* v, double* t) {
    const int lane = threadIdx.x & 31;
    SOCP_FOR_EACH_CONE(L, c, kind, offs, dim) {
        if (kind == KIND_POC) warp_poc_iprod(lam + offs, v + offs, t + offs, dim, lane);
        else warp_soc_iprod(lam + offs, v + offs, t + offs, dim, lane);
    }
}
// returns per-thread partial (max over this warp's cones); combine with block_max
__device__ __forceinline__ double cta_max_step_partial(const ConeLayout& L, const double* x) {
    const int lane = threadIdx.x & 31;
    double mx = -INFINITY;
    SOCP_FOR_EACH_CONE(L, c, kind, offs, dim) {
        const double v = (kind == KIND_POC) ? warp_poc_max_step(x + offs, dim, lane)
                                            : warp_soc_max_step(x + offs, dim, lane);
        mx = fmax(mx, v);
    }
    return mx;
}
__device__ __forceinline__ double cta_scmax_partial(const ConeLayout& L, const double* l, const double* x, int* fail) {
    const int lane = threadIdx.x & 31;
    double mx = -INFINITY;
    SOCP_FOR_EACH_CONE(L, c, kind, offs, dim) {
        const double v = (kind == KIND_POC) ? warp_poc_scmax(l + offs, x + offs, dim, lane)
                                            : warp_soc_scmax(l + offs, x + offs, dim, lane, fail);
        mx = fmax(mx, v);
    }
    return mx;
}
// compute_step(cones, l, ds, dz), reference src/mats.jl:30-40
// step_from_t lives in common.cuh
// make_e!, reference src/vectors.jl:7-24: value of e at index i of a block
__device__ __forceinline__ double e_value(int kind, int i) { return (kind == KIND_POC || i == 0) ? 1.0 : 0.0; }

}  // namespace socp
