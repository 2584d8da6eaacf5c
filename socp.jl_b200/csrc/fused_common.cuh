// fused_common.cuh -- pieces shared by the whole-solve kernels (fused_v2.cuh: one-warp teams; fused_v3.cuh: 4- and 8-warp
// teams on the compressed layout) and by the tiled path's panel kernels: shared-memory layout helpers, team barriers
// and reductions, the shared-memory gemvs, the register-resident factor of one 8x8 diagonal tile, and the lane view
// of a second-order cone.  Compiles for the host against tests/simt_emu/simt_emu.h (test infrastructure) as well.
#pragma once
#include "common.cuh"

namespace socp {

constexpr int F2_MAX_SOC = 64;
constexpr int F2_CS = 8;          // scalars kept per second-order cone
enum { CS_ETA = 0, CS_IE, CS_IE2, CS_R1W, CS_W0, CS_LAM0, CS_A, CS_LLT };

__host__ __device__ constexpr int f2_ld(int rows) {      // smallest ld >= rows with ld == 4 (mod 8): DMMA fragment loads
    int ld = rows;
    while (ld % 8 != 4) ++ld;
    return ld;
}
__host__ __device__ constexpr int f2_ldv(int rows) {     // smallest ld >= rows with ld == 8 (mod 16): 128-bit column loads
    int ld = rows;
    while (ld % 16 != 8) ++ld;
    return ld;
}
__host__ __device__ constexpr int f2_split(int units, int nw) {   // lanes per output unit so that one team pass covers them
    int split = 32;
    while (split > 1 && (units + (32 / split) - 1) / (32 / split) > nw) split >>= 1;
    return split;
}
// The inverse factor X = L^-1 is lower triangular: it is stored as a trapezoid, column block cb (8 columns) holding
// only the rows from 8*cb on, with its own leading dimension (== 4 mod 8).  For nbl block columns:
__host__ __device__ constexpr int f2_xld(int nbl, int cb) { return (nbl - cb) * 8 + 4; }
__host__ __device__ constexpr int f2_xbase(int nbl, int cb) { return 8 * cb * (nbl * 8 + 4) - 32 * cb * (cb - 1); }
__host__ __device__ constexpr int f2_xsize(int nbl) { return f2_xbase(nbl, nbl); }
// number of trailing-update tiles of block column b, and where its descriptors start (see f2_trail_desc)
__host__ __device__ constexpr int f2_ntrail(int nbl, int b) { return (nbl * (nbl + 1) - (b + 1) * (b + 2)) / 2; }
__host__ __device__ constexpr int f2_trail_base(int nbl, int b) { return b * (nbl * (nbl + 1) / 2) - b * (b + 1) * (b + 2) / 6; }
__host__ __device__ constexpr int f2_lpc(int maxdim) {   // lanes per second-order cone: 4 elements per lane
    int l = 1;
    while (l * 4 < maxdim) l <<= 1;
    return l;
}

// ------------------------------------------------------------------------------------------------ team helpers
template <int NW>
__device__ __forceinline__ void tsync() {
    if (NW == 1) __syncwarp();
    else __syncthreads();
}
// barrier of one team of NW warps inside a CTA that holds several teams: hardware barrier `bar` (1..15) is the team's
template <int NW>
__device__ __forceinline__ void tsync_t(int bar) {
    if (NW == 1) __syncwarp();
    else bar_sync_id<NW * 32>(bar);
}
__device__ __forceinline__ double grp_sum(double v, int lpc) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        if (o < lpc) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
// Team-wide reduction of NS sums (v[0..NS)), NM maxima (v[NS..NS+NM)) and one flag through a scratch slot
// (8 doubles per warp).  One team barrier (none for a one-warp team); the caller alternates `scr` between two slots.
template <int NW, int NS, int NM>
__device__ __forceinline__ void team_reduce(double (&v)[NS + NM], int& flag, double* scr, int lane, int warp, int bar = -1) {
#pragma unroll
    for (int i = 0; i < NS; ++i) v[i] = warp_sum(v[i]);
#pragma unroll
    for (int i = NS; i < NS + NM; ++i) v[i] = warp_max(v[i]);
    flag = __any_sync(FULL_MASK, flag) ? 1 : 0;
    if (NW == 1) return;
    if (lane == 0) {
        double* q = scr + warp * 8;
#pragma unroll
        for (int i = 0; i < NS + NM; ++i) q[i] = v[i];
        q[7] = (double)flag;
    }
    if (bar < 0) __syncthreads();
    else bar_sync_id<NW * 32>(bar);
#pragma unroll
    for (int i = 0; i < NS; ++i) v[i] = 0.0;
#pragma unroll
    for (int i = NS; i < NS + NM; ++i) v[i] = -INFINITY;
    double f = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const double* q = scr + w * 8;
#pragma unroll
        for (int i = 0; i < NS; ++i) v[i] += q[i];
#pragma unroll
        for (int i = NS; i < NS + NM; ++i) v[i] = fmax(v[i], q[i]);
        f += q[7];
    }
    flag = f != 0.0;
}

// out(c) = sum_r M[c*ld + r] * x[r] over rows [0, rows): eight lanes per column, lane rl takes the row pairs
// {2rl, 2rl+1} + 16m with 128-bit loads (a quarter warp reads 128 contiguous bytes: conflict free for any even ld).
// M rows and x must be zero / finite up to the next even row.  epi(c, acc) runs on one lane per column.
template <int NW, class Epi>
__device__ __forceinline__ void gemv_cols_v(const double* __restrict__ M, int ld, int rows, int cols,
                                            const double* __restrict__ x, int lane, int warp, Epi epi) {
    const int cq = lane >> 3, rl = lane & 7;
    const double2* xv = reinterpret_cast<const double2*>(x);
    const int npair = (rows + 1) >> 1;
    for (int c0 = warp * 4; c0 < cols; c0 += NW * 4) {
        const int c = c0 + cq;
        const bool ok = c < cols;
        const double2* col = reinterpret_cast<const double2*>(M + (ok ? c : 0) * ld);
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
        for (int q = rl; q < npair; q += 8) {
            const double2 g = col[q], w = xv[q];
            a0 = fma(g.x, w.x, a0);
            a1 = fma(g.y, w.y, a1);
        }
        double acc = a0 + a1;
        acc += __shfl_xor_sync(FULL_MASK, acc, 1);
        acc += __shfl_xor_sync(FULL_MASK, acc, 2);
        acc += __shfl_xor_sync(FULL_MASK, acc, 4);
        if (rl == 0 && ok) epi(c, acc);
    }
}
// The same product with the passes of a warp unrolled (at most MAXPASS passes: cols <= 4 NW MAXPASS): the loads, FMA
// chains and shuffle rounds of the passes interleave instead of running one pass after the other -- one pass is a
// dependent chain of ~250 cycles (LDS -> 4 DFMA -> 3 shuffle rounds -> epilogue) that keeps a warp scheduler idle.
// The epilogues run after all the sums, so epi may overwrite x.
template <int NW, int MAXPASS, class Epi>
__device__ __forceinline__ void gemv_cols_vu(const double* __restrict__ M, int ld, int rows, int cols,
                                             const double* __restrict__ x, int lane, int warp, Epi epi) {
    const int cq = lane >> 3, rl = lane & 7;
    const double2* xv = reinterpret_cast<const double2*>(x) + rl;
    const int npair = (rows + 1) >> 1;
    const int nq = (npair + 7) >> 3;
    double acc[MAXPASS];
#pragma unroll
    for (int i = 0; i < MAXPASS; ++i) {
        const int c = warp * 4 + i * NW * 4 + cq;
        const double2* col = reinterpret_cast<const double2*>(M + (c < cols ? c : 0) * ld) + rl;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
        for (int m = 0; m < nq; ++m) {
            if (rl + 8 * m < npair) {
                const double2 g = col[8 * m], w = xv[8 * m];
                a0 = fma(g.x, w.x, a0);
                a1 = fma(g.y, w.y, a1);
            }
        }
        acc[i] = a0 + a1;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
        for (int i = 0; i < MAXPASS; ++i) acc[i] += __shfl_xor_sync(FULL_MASK, acc[i], o);
    }
#pragma unroll
    for (int i = 0; i < MAXPASS; ++i) {
        const int c = warp * 4 + i * NW * 4 + cq;
        if (rl == 0 && c < cols) epi(c, acc[i]);
    }
}
// out(c) = sum_{r >= r_lo(c)} M[c*ld + r] * x[r], four lanes per column, scalar loads (triangular / small operands).
template <int NW, bool TRI, class Epi>
__device__ __forceinline__ void gemv_cols(const double* __restrict__ M, int ld, int rows, int cols,
                                          const double* __restrict__ x, int lane, int warp, Epi epi) {
    const int cq = lane >> 2, rl = lane & 3;
    for (int c0 = warp * 8; c0 < cols; c0 += NW * 8) {
        const int c = c0 + cq;
        const bool ok = c < cols;
        const double* col = M + (ok ? c : 0) * ld;
        double a0 = 0.0, a1 = 0.0;
        int r = rl + (TRI ? (c & ~3) : 0);
        if (TRI && r < c) r += 4;
        if (ok) {
#pragma unroll 2
            for (; r + 4 < rows; r += 8) {
                a0 = fma(col[r], x[r], a0);
                a1 = fma(col[r + 4], x[r + 4], a1);
            }
            if (r < rows) a0 = fma(col[r], x[r], a0);
        }
        double acc = a0 + a1;
        acc += __shfl_xor_sync(FULL_MASK, acc, 1);
        acc += __shfl_xor_sync(FULL_MASK, acc, 2);
        if (rl == 0 && ok) epi(c, acc);
    }
}
// out(r) = sum_c M[c*ld + r] * x[c]: every lane owns the row PAIR (2q, 2q+1) (128-bit loads down a column), `split`
// lanes share a pair and stride the columns.  epi(r, acc) is called for both rows of the pair (r < rows).
template <int NW, class Epi>
__device__ __forceinline__ void gemv_rows_v(const double* __restrict__ M, int ld, int rows, int cols,
                                            const double* __restrict__ x, int split, int lane, int warp, Epi epi) {
    const int ppw = 32 / split;                       // row pairs per warp pass
    const int pr = lane & (ppw - 1), part = lane / ppw;
    const int npair = (rows + 1) >> 1;
    for (int q0 = warp * ppw; q0 < npair; q0 += NW * ppw) {
        const int q = q0 + pr;
        const bool ok = q < npair;
        const double2* row = reinterpret_cast<const double2*>(M) + (ok ? q : 0);
        const int ldv = ld >> 1;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
        for (int c = part; c < cols; c += split) {
            const double2 g = row[c * ldv];
            const double w = x[c];
            a0 = fma(g.x, w, a0);
            a1 = fma(g.y, w, a1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            if (o >= ppw) {
                a0 += __shfl_xor_sync(FULL_MASK, a0, o);
                a1 += __shfl_xor_sync(FULL_MASK, a1, o);
            }
        if (part == 0 && ok) {
            epi(2 * q, a0);
            if (2 * q + 1 < rows) epi(2 * q + 1, a1);
        }
    }
}
// out(r) = sum_{c in [0, c_hi(r))} M[c*ld + r] * x[c*xs]; `split` lanes per row (power of two); scalar loads.
template <int NW, bool TRI, class Epi>
__device__ __forceinline__ void gemv_rows(const double* __restrict__ M, int ld, int rows, int cols,
                                          const double* __restrict__ x, int xs, int split, int lane, int warp, Epi epi) {
    const int rpw = 32 / split;
    const int rr = lane & (rpw - 1), part = lane / rpw;
    for (int r0 = warp * rpw; r0 < rows; r0 += NW * rpw) {
        const int r = r0 + rr;
        const bool ok = r < rows;
        const int chi = TRI ? min(cols, r + 1) : cols;
        double a0 = 0.0, a1 = 0.0;
        if (ok) {
            const double* row = M + r;
            int c = part;
#pragma unroll 2
            for (; c + split < chi; c += 2 * split) {
                a0 = fma(row[c * ld], x[c * xs], a0);
                a1 = fma(row[(c + split) * ld], x[(c + split) * xs], a1);
            }
            if (c < chi) a0 = fma(row[c * ld], x[c * xs], a0);
        }
        double acc = a0 + a1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            if (o >= rpw) acc += __shfl_xor_sync(FULL_MASK, acc, o);
        if (part == 0 && ok) epi(r, acc);
    }
}

// Factor of one 8x8 diagonal tile by one warp.  The Cholesky factor is computed by all lanes redundantly in
// registers (no shuffles); lane c (mod 8) then computes column c of its inverse by forward substitution, so that
// on exit xc[i] = (L^-1)[i][c], c = lane & 7 (zero above the diagonal).  Returns 0 on a pivot that is not > 0.
template <bool WRITE_L>
__device__ __forceinline__ int f2_diag_factor(double* T, int ld, int lane, double (&xc)[8]) {
    double a[36];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) a[i * (i + 1) / 2 + j] = T[j * ld + i];
    double r[8];
    int ok = 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double d = a[j * (j + 1) / 2 + j];
        ok &= (d > 0.0);
        const double rj = fast_rsqrt(d);
        r[j] = rj;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) a[i * (i + 1) / 2 + j] *= rj;
#pragma unroll
        for (int i = j + 1; i < 8; ++i)
#pragma unroll
            for (int c = j + 1; c <= i; ++c)
                a[i * (i + 1) / 2 + c] = fma(-a[i * (i + 1) / 2 + j], a[c * (c + 1) / 2 + j], a[i * (i + 1) / 2 + c]);
    }
    const int c = lane & 7;
    if (WRITE_L) {      // the tiled path needs the factor itself: lane c < 8 writes column c of L (rows i >= c)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double v = a[i * (i + 1) / 2];
#pragma unroll
            for (int cc = 1; cc <= i; ++cc) v = (c == cc) ? a[i * (i + 1) / 2 + cc] : v;
            if (lane < 8 && i >= c) T[c * ld + i] = (i == c) ? v * r[i] : v;      // l_ii = d_i / sqrt(d_i)
        }
    }
    // column c of the inverse: x_i = 0 (i < c), r_c (i == c), -r_i sum_{m<i} l_im x_m (i > c)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double sacc = 0.0;
#pragma unroll
        for (int m = 0; m < i; ++m) sacc = fma(a[i * (i + 1) / 2 + m], xc[m], sacc);
        xc[i] = (i == c) ? r[i] : -r[i] * sacc;      // i < c: sacc == 0
    }
    return ok;
}

// per-thread view of one second-order cone slot: element e of this lane is index g + e*lpc of the cone; bit e of
// `tm` says that this element exists and belongs to the tail (index > 0)
struct SocLane {
    int offs, g, lpc;
    unsigned tm;
    bool valid;
    __device__ __forceinline__ int at(int e) const { return offs + g + e * lpc; }
    __device__ __forceinline__ bool tail(int e) const { return (tm >> e) & 1u; }
    __device__ __forceinline__ bool head() const { return valid && g == 0; }
};

#define F2_FOR_E _Pragma("unroll") for (int e = 0; e < 4; ++e)

}  // namespace socp
