// cone_batch.cuh -- batch-wide cone kernels of the step-level API (compute_scaling, scale!, iscale!, W^-2 apply,
// vprod!, iprod!, max_step, compute_step) for [batch][k] vectors in global memory.
//
// These are HBM-bound (24k-40k bytes per problem, a handful of flops per byte), so the
// mapping is chosen for coalescing and bytes in flight rather than for arithmetic:
//   * a group of LPC lanes owns one second-order cone, element e of lane g is index g + e*LPC (4 elements per lane:
//     LPC = 1 for SOC(4) -- one thread per cone, 32 contiguous bytes per thread --, 16 for SOC(50), 32 for SOC(128));
//     consecutive groups take consecutive cones of consecutive problems, so a warp always reads one contiguous run
//   * reductions are xor shuffles inside the group (none for LPC = 1)
//   * positive-orthant rows are elementwise over a flat grid-stride loop
//   * per-problem reductions (max_step, compute_step): one warp per problem when the cones fit in a warp
// Formulas: the closed O(d) forms of SURVEY.md appendix A, as in cone_ops.cuh (reference src/scalings.jl,
// src/vectors.jl, src/mats.jl -- cited per function there).
#pragma once
#include "cone_ops.cuh"

namespace socp {

struct BLayout {
    int k, kpoc, nsoc, lpc;
    int nwork;                 // work cones of the handle (stride of the per-cone scalar block `eta`)
    const int* soc_offs;       // [nsoc] offset of each second-order cone in a k-vector
    const int* soc_dim;        // [nsoc]
    const int* soc_work;       // [nsoc] index of the cone among the work cones (eta slot)
};

struct BLane {
    int b, slot, offs, g, lpc;
    unsigned tm;               // bit e: element g + e*lpc exists and belongs to the tail
    bool valid;
    __device__ __forceinline__ int at(int e) const { return offs + g + e * lpc; }
    __device__ __forceinline__ bool tail(int e) const { return (tm >> e) & 1u; }
    __device__ __forceinline__ bool head() const { return valid && g == 0; }
};
__device__ __forceinline__ BLane b_lane(const BLayout& L, long long gid, long long total, int lane) {
    BLane l;
    l.valid = gid < total;
    // 32-bit division: batch * nsoc < 2^32 (checked by the host when the layout is built)
    const unsigned q = l.valid ? (unsigned)gid : 0u;
    const unsigned ns = (unsigned)max(L.nsoc, 1);  // layouts without second-order cones never have a valid lane
    l.b = (int)(q / ns);
    l.slot = (int)(q - (unsigned)l.b * ns);
    l.offs = L.soc_offs[l.slot];
    const int dim = l.valid ? L.soc_dim[l.slot] : 0;
    l.g = lane & (L.lpc - 1);
    l.lpc = L.lpc;
    l.tm = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int i = l.g + e * L.lpc;
        if (i > 0 && i < dim) l.tm |= 1u << e;
    }
    return l;
}
__device__ __forceinline__ double b_gsum(double v, int lpc) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        if (o < lpc) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
#define B_FOR_E _Pragma("unroll") for (int e = 0; e < 4; ++e)
__device__ __forceinline__ void b_load(const BLane& l, const double* __restrict__ v, double (&r)[4]) {
    B_FOR_E r[e] = l.tail(e) ? v[l.at(e)] : 0.0;
}
__device__ __forceinline__ double b_dot(const double (&a)[4], const double (&b)[4], int lpc) {
    double d = 0.0;
    B_FOR_E d = fma(a[e], b[e], d);
    return b_gsum(d, lpc);
}
// iterate over the second-order-cone slots of the whole batch: f(BLane) with all 32 lanes taking part
template <class F>
__device__ __forceinline__ void b_for_each_group(const BLayout& L, int batch, F f) {
    const int lane = threadIdx.x & 31;
    const int spw = 32 / L.lpc;
    const long long total = (long long)batch * L.nsoc;
    const long long wglobal = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long base = wglobal * spw; base < total; base += wstride * spw) f(b_lane(L, base + lane / L.lpc, total, lane));
}
template <class F>
__device__ __forceinline__ void b_for_each_poc(const BLayout& L, int batch, F f) {
    const unsigned total = (unsigned)batch * (unsigned)L.kpoc;       // < 2^32 (host-checked)
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const unsigned b = q / (unsigned)L.kpoc;
        f((int)b, (int)(q - b * (unsigned)L.kpoc));
    }
}

// ---------------------------------------------------------------------------------------------- compute_scaling
// reference src/scalings.jl:22-30 (POC), :32-99 (SOC).  eta: [batch][4][nwork] = eta, 1/eta, 1/eta^2, 1/(1+wbar0).
__global__ void __launch_bounds__(256, 4)
bk_scaling(BLayout L, int batch, const double* __restrict__ s, const double* __restrict__ z, double* __restrict__ lam,
           double* __restrict__ wb, double* __restrict__ iwb, double* __restrict__ eta, int* __restrict__ fail,
           const int* __restrict__ active) {
    b_for_each_poc(L, batch, [&](int b, int i) {
        if (active && !active[b]) return;
        const size_t o = (size_t)b * L.k + i;
        const double si = s[o], zi = z[o];
        const double q = si * fast_rcp(zi), qi = zi * fast_rcp(si), pz = si * zi;
        if (!(q >= 0.0) | !(pz >= 0.0)) atomicOr(fail + b, 1);
        wb[o] = fast_sqrt(q);
        iwb[o] = fast_sqrt(qi);
        lam[o] = fast_sqrt(pz);
    });
    b_for_each_group(L, batch, [&](const BLane& l) {
        const bool on = l.valid && !(active && !active[l.b]);
        const size_t o = (size_t)l.b * L.k;
        const double* sb_ = s + o;
        const double* zb_ = z + o;
        double sv[4], zv[4];
        B_FOR_E { const bool t = on && l.tail(e); sv[e] = t ? sb_[l.at(e)] : 0.0; zv[e] = t ? zb_[l.at(e)] : 0.0; }
        const double s0 = on ? sb_[l.offs] : 1.0, z0 = on ? zb_[l.offs] : 1.0;
        double ss = 0.0, zz = 0.0, sz = 0.0;
        B_FOR_E { ss = fma(sv[e], sv[e], ss); zz = fma(zv[e], zv[e], zz); sz = fma(sv[e], zv[e], sz); }
        ss = b_gsum(ss, L.lpc); zz = b_gsum(zz, L.lpc); sz = b_gsum(sz, L.lpc);
        const double onrms = s0 * s0 - ss, onrmz = z0 * z0 - zz;          // :39-45
        int f = !(onrms >= 0.0) | !(onrmz >= 0.0);
        const double is = fast_rsqrt(onrms), iz = fast_rsqrt(onrmz);      // :46-49
        const double nrms = onrms * is, nrmz = onrmz * iz;
        const double sb0 = s0 * is, zb0 = z0 * iz;
        const double ns = sz * (is * iz) + zb0 * sb0;                     // :53-56
        const double g2 = (1.0 + ns) / 2.0;
        f |= !(g2 >= 0.0);
        const double rg = fast_rsqrt(g2);
        const double gamma = g2 * rg, ig = 0.5 * rg;                      // :57, :64
        const double et = fast_sqrt(nrms * iz);                           // :68
        const double tmv1 = fast_sqrt(nrms * nrmz);                       // :91
        const double mult = tmv1 * fast_rcp(zb0 + sb0 + 2.0 * gamma);     // :92
        const double csf = gamma + zb0, czf = gamma + sb0;                // :93-94
        if (!on) return;
        double* lb = lam + o;
        double* wbb = wb + o;
        B_FOR_E if (l.tail(e)) {
            const double sb = sv[e] * is, zb = zv[e] * iz;
            wbb[l.at(e)] = (sb - zb) * ig;                                // :62,:64
            lb[l.at(e)] = (sb * csf + zb * czf) * mult;                   // :95-97
        }
        if (l.g == 0) {
            const double w0 = (sb0 + zb0) * ig, ie = fast_rcp(et);
            wbb[l.offs] = w0;                                             // :60
            lb[l.offs] = gamma * tmv1;                                    // :98
            double* es = eta + (size_t)l.b * 4 * L.nwork + L.soc_work[l.slot];
            es[0] = et; es[L.nwork] = ie; es[2 * L.nwork] = ie * ie; es[3 * L.nwork] = fast_rcp(1.0 + w0);
            if (f) atomicOr(fail + l.b, 1);
        }
    });
}

// ---------------------------------------------------------------------------------------------- Gt = W^-1 G
// The scaled copy of G whose Gram matrix is the reduced KKT matrix (setup_iter, reference src/densesolver.jl:41-43):
// iscale! (src/scalings.jl:119-124, :142-156) applied to every column of G.  One lane group per (problem, column,
// cone); positive-orthant rows elementwise.  Gt: [batch][n][ldgt], rows >= k stay zero.
__global__ void __launch_bounds__(256, 4)
bk_build_gt(BLayout L, int batch, int n, const double* __restrict__ G, int64_t sG, const double* __restrict__ wb,
            const double* __restrict__ iwb, const double* __restrict__ eta, double* __restrict__ Gt, int ldgt,
            const int* __restrict__ active) {
    {   // positive-orthant rows
        const unsigned per = (unsigned)n * (unsigned)L.kpoc, total = (unsigned)batch * per;     // < 2^32 (host-checked)
        for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
            const int b = (int)(q / per);
            const int rem = (int)(q - (unsigned)b * per);
            const int col = rem / L.kpoc, r = rem - col * L.kpoc;
            if (active && !active[b]) continue;
            Gt[((size_t)b * n + col) * ldgt + r] = iwb[(size_t)b * L.k + r] * G[(int64_t)b * sG + (int64_t)col * L.k + r];
        }
    }
    // second-order cones: a lane group owns (problem, cone, chunk of GT_CH columns); the cone's wbar and scalars are
    // loaded once per chunk and the columns go two at a time, so that two columns' loads and two shuffle chains are
    // in flight per group (four at a time needs 115 registers: no more bytes in flight per SM at the lower occupancy) (the kernel is bound by bytes in flight, not by arithmetic)
    constexpr int GT_CH = 8, GT_NC = 2;       // columns per chunk, columns in flight
    const int lane = threadIdx.x & 31;
    const int spw = 32 / L.lpc;
    const unsigned nch = (unsigned)(n + GT_CH - 1) / GT_CH;
    const long long total = (long long)batch * L.nsoc * nch;
    const long long wglobal = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long base = wglobal * spw; base < total; base += wstride * spw) {
        const long long gid = base + lane / L.lpc;
        const bool valid = gid < total;
        const unsigned q = valid ? (unsigned)gid : 0u;                                           // < 2^32 (host-checked)
        const unsigned cone = q / nch;                                                           // b * nsoc + slot
        const int c0 = (int)(q - cone * nch) * GT_CH;
        const BLane l = b_lane(L, cone, valid ? (long long)batch * L.nsoc : 0, lane);
        const int bb = l.b;
        const bool on = l.valid && !(active && !active[bb]);
        const double* wbb = wb + (size_t)bb * L.k;
        const double* es = eta + (size_t)bb * 4 * L.nwork + L.soc_work[l.slot];
        double wv[4];
        B_FOR_E wv[e] = (on && l.tail(e)) ? wbb[l.at(e)] : 0.0;
        const double ie = on ? es[L.nwork] : 0.0, e3 = on ? es[3 * L.nwork] : 0.0, w0 = on ? wbb[l.offs] : 0.0;
#pragma unroll 2
        for (int cc = 0; cc < GT_CH; cc += GT_NC) {
            bool onc[GT_NC];
            double gv[GT_NC][4], g0[GT_NC], dl[GT_NC];
            const double* gbase = G + (int64_t)bb * sG + (int64_t)(c0 + cc) * L.k + l.offs;
#pragma unroll
            for (int j = 0; j < GT_NC; ++j) {
                onc[j] = on && c0 + cc + j < n;
                const double* gc = gbase + j * L.k;
                B_FOR_E gv[j][e] = (onc[j] && l.tail(e)) ? gc[l.g + e * l.lpc] : 0.0;
                g0[j] = onc[j] ? gc[0] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < GT_NC; ++j) dl[j] = b_dot(wv, gv[j], L.lpc);                      // src/scalings.jl:145-148
#pragma unroll
            for (int j = 0; j < GT_NC; ++j) {
                if (!onc[j]) continue;
                const double cst = -g0[j] + dl[j] * e3;                                           // :151
                double* oc = Gt + ((size_t)bb * n + (c0 + cc + j)) * ldgt + l.offs;
                B_FOR_E if (l.tail(e)) oc[l.g + e * l.lpc] = ie * (gv[j][e] + cst * wv[e]);       // :153-155
                if (l.g == 0) oc[0] = ie * (w0 * g0[j] - dl[j]);                                  // :152
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- scale! / iscale! / W^-2
// reference src/scalings.jl:112-156, src/densesolver.jl:86.  MODE as ApplyMode of cone_ops.cuh.
template <int MODE>
__global__ void __launch_bounds__(256, 6)
bk_apply(BLayout L, int batch, const double* __restrict__ wb, const double* __restrict__ iwb,
         const double* __restrict__ eta, const double* __restrict__ v, double* __restrict__ out) {
    b_for_each_poc(L, batch, [&](int b, int i) {
        const size_t o = (size_t)b * L.k + i;
        const double vi = v[o];
        double r;
        if (MODE == APPLY_W) r = wb[o] * vi;
        else if (MODE == APPLY_WINV) r = iwb[o] * vi;
        else { const double iw = iwb[o]; r = iw * iw * vi; }
        out[o] = r;
    });
    b_for_each_group(L, batch, [&](const BLane& l) {
        const size_t o = (size_t)l.b * L.k;
        double wv[4], vv[4];
        b_load(l, wb + o, wv);
        b_load(l, v + o, vv);
        const double dl = b_dot(wv, vv, L.lpc);                          // :129-132 / :145-148
        if (!l.valid) return;
        const double* es = eta + (size_t)l.b * 4 * L.nwork + L.soc_work[l.slot];
        const double v0 = v[o + l.offs], w0 = wb[o + l.offs];
        double* ob = out + o;
        if (MODE == APPLY_W) {
            const double et = es[0], cst = v0 + dl * es[3 * L.nwork];    // :135
            B_FOR_E if (l.tail(e)) ob[l.at(e)] = et * (vv[e] + cst * wv[e]);   // :137-139
            if (l.g == 0) ob[l.offs] = et * (w0 * v0 + dl);              // :136
        } else if (MODE == APPLY_WINV) {
            const double ie = es[L.nwork], cst = -v0 + dl * es[3 * L.nwork];   // :151
            B_FOR_E if (l.tail(e)) ob[l.at(e)] = ie * (vv[e] + cst * wv[e]);   // :153-155
            if (l.g == 0) ob[l.offs] = ie * (w0 * v0 - dl);              // :152
        } else {
            const double ie2 = es[2 * L.nwork], qv = w0 * v0 - dl;       // W^-2 = eta^-2 (2 q q' - J)
            B_FOR_E if (l.tail(e)) ob[l.at(e)] = ie2 * (vv[e] - 2.0 * wv[e] * qv);
            if (l.g == 0) ob[l.offs] = ie2 * (2.0 * w0 * qv - v0);
        }
    });
}

// ---------------------------------------------------------------------------------------------- vprod! / iprod!
// reference src/vectors.jl:58-81, :99-131
__global__ void __launch_bounds__(256, 6)
bk_vprod(BLayout L, int batch, const double* __restrict__ u, const double* __restrict__ v, double* __restrict__ t) {
    b_for_each_poc(L, batch, [&](int b, int i) {
        const size_t o = (size_t)b * L.k + i;
        t[o] = u[o] * v[o];
    });
    b_for_each_group(L, batch, [&](const BLane& l) {
        const size_t o = (size_t)l.b * L.k;
        double uv[4], vv[4];
        b_load(l, u + o, uv);
        b_load(l, v + o, vv);
        const double acc = b_dot(uv, vv, L.lpc);
        if (!l.valid) return;
        const double u0 = u[o + l.offs], v0 = v[o + l.offs];
        B_FOR_E if (l.tail(e)) t[o + l.at(e)] = u0 * vv[e] + v0 * uv[e];   // :73-75
        if (l.g == 0) t[o + l.offs] = acc + u0 * v0;                       // :66-69
    });
}
__global__ void __launch_bounds__(256, 6)
bk_iprod(BLayout L, int batch, const double* __restrict__ lam, const double* __restrict__ v, double* __restrict__ t) {
    b_for_each_poc(L, batch, [&](int b, int i) {
        const size_t o = (size_t)b * L.k + i;
        t[o] = v[o] * fast_rcp(lam[o]);                                    // :99-103
    });
    b_for_each_group(L, batch, [&](const BLane& l) {
        const size_t o = (size_t)l.b * L.k;
        double lv[4], vv[4];
        b_load(l, lam + o, lv);
        b_load(l, v + o, vv);
        const double ll = b_dot(lv, lv, L.lpc), beta = b_dot(lv, vv, L.lpc);
        if (!l.valid) return;
        const double l0 = lam[o + l.offs], v0 = v[o + l.offs];
        const double a = l0 * l0 - ll;                                     // :108-111
        const double ia = fast_rcp(a), il0 = fast_rcp(l0);
        B_FOR_E if (l.tail(e)) t[o + l.at(e)] = (-v0 * lv[e] + (a * vv[e] + beta * lv[e]) * il0) * ia;   // :115-124, O(d) form
        if (l.g == 0) t[o + l.offs] = (l0 * v0 - beta) * ia;
    });
}

// ---------------------------------------------------------------------------------------------- max_step / compute_step
// A segment of SW lanes (power of two >= nsoc * lpc, <= 32) per problem, 32 / SW problems per warp.
// reference src/mats.jl:1-28, :30-86.
__device__ __forceinline__ double b_segmax(double v, int sw) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        if (o < sw) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ BLane b_lane_seg(const BLayout& L, int b, bool bvalid, int sl) {
    const int grp = sl / L.lpc;
    return b_lane(L, (long long)b * L.nsoc + grp, bvalid && grp < L.nsoc ? (long long)(b + 1) * L.nsoc : 0, sl);
}
__global__ void __launch_bounds__(256, 6)
bk_max_step(BLayout L, int batch, int sw, const double* __restrict__ x, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, sl = lane & (sw - 1);
    const int b = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / sw) + lane / sw;
    const bool bv = b < batch;
    const size_t o = (size_t)(bv ? b : 0) * L.k;
    double mx = -INFINITY;
    if (bv) for (int i = sl; i < L.kpoc; i += sw) mx = fmax(mx, -x[o + i]);
    const BLane l = b_lane_seg(L, bv ? b : 0, bv, sl);
    double xv[4];
    b_load(l, x + o, xv);
    const double sq = b_dot(xv, xv, L.lpc);
    if (l.valid) mx = fmax(mx, fast_sqrt(sq) - x[o + l.offs]);
    mx = b_segmax(mx, sw);
    if (sl == 0 && bv) out[b] = mx;
}
__device__ __forceinline__ double b_scmax(const BLayout& L, const BLane& l, const double* __restrict__ lam,
                                          const double* __restrict__ x, const double (&lv)[4], double ll, int* fail) {
    double xv[4];
    b_load(l, x, xv);
    const double lx = b_dot(lv, xv, L.lpc);
    const double l0 = l.valid ? lam[l.offs] : 1.0, x0 = l.valid ? x[l.offs] : 0.0;
    const double ai = l0 * l0 - ll;                         // :67-70
    *fail |= l.valid && !(ai >= 0.0);
    const double a = fast_rsqrt(ai);                        // :71
    const double r1 = a * l0 * x0 - a * lx;                 // :74-77
    const double cst = (r1 + x0) * fast_rcp(a * l0 + 1.0);  // :80
    double q = 0.0;
    B_FOR_E if (l.tail(e)) { const double w = a * (xv[e] - cst * a * lv[e]); q = fma(w, w, q); }   // :83
    q = b_gsum(q, L.lpc);
    return l.valid ? fast_sqrt(q) - a * r1 : -INFINITY;     // :85
}
__global__ void __launch_bounds__(256, 6)
bk_compute_step(BLayout L, int batch, int sw, const double* __restrict__ lam, const double* __restrict__ ds,
                const double* __restrict__ dz, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, sl = lane & (sw - 1);
    const int b = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / sw) + lane / sw;
    const bool bv = b < batch;
    const size_t o = (size_t)(bv ? b : 0) * L.k;
    double mx = -INFINITY;
    if (bv)
        for (int i = sl; i < L.kpoc; i += sw) {             // :53-62
            const double il = fast_rcp(lam[o + i]);
            mx = fmax(mx, fmax(-ds[o + i] * il, -dz[o + i] * il));
        }
    const BLane l = b_lane_seg(L, bv ? b : 0, bv, sl);
    double lv[4];
    b_load(l, lam + o, lv);
    const double ll = b_dot(lv, lv, L.lpc);
    int f = 0;
    mx = fmax(mx, b_scmax(L, l, lam + o, ds + o, lv, ll, &f));
    mx = fmax(mx, b_scmax(L, l, lam + o, dz + o, lv, ll, &f));
    mx = b_segmax(mx, sw);
    if (sl == 0 && bv) out[b] = step_from_t(mx);            // :30-40
}

}  // namespace socp
