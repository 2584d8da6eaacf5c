// emu_fused3.cpp -- host build of the whole-solve kernel k_fused3 (socp.jl_b200/csrc/fused_v3.cuh) against the SIMT
// emulator.  TEST INFRASTRUCTURE ONLY: built into tests/_build/libsocp_emu.so by tests/simt_emu/build_emu.py and
// loaded by tests/test_emu_fused3.py; nothing under socp.jl_b200/ links it.
#define SOCP_SIMT_EMU 1
#include "fused_v3.cuh"

#include <cstdio>
#include <vector>

using namespace socp;

// row pattern of a batch of G's (column-major k x n each): -1 empty, j single column, -2 dense.  Mirrors the device
// kernel k_row_pattern of solver.cu.
static std::vector<int> row_pattern(const double* G, int64_t sG, int batch, int n, int k) {
    std::vector<int> rc(k, -1);
    const int nb = sG == 0 ? 1 : batch;
    for (int b = 0; b < nb; ++b)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < k; ++i)
                if (G[(int64_t)b * sG + (int64_t)j * k + i] != 0.0) {
                    if (rc[i] == -1) rc[i] = j;
                    else if (rc[i] != j) rc[i] = -2;
                }
    return rc;
}

extern "C" int emu_fused3_plan(int n, int p, int k, int ncones, const int* kind, const int* offs, const int* dim,
                               const int* rowcol, int* out /* fits, smem, ctas_per_sm, d0, kd, nsing, ident, nb */) {
    F3Plan P;
    std::vector<int> tables;
    f3_plan(P, n, p, k, std::vector<int>(kind, kind + ncones), std::vector<int>(offs, offs + ncones),
            std::vector<int>(dim, dim + ncones), std::vector<int>(rowcol, rowcol + k), 227 * 1024, 148, tables);
    out[0] = P.fits; out[1] = (int)P.smem; out[2] = P.ctas_per_sm; out[3] = P.d0; out[4] = P.kd; out[5] = P.nsing;
    out[6] = P.ident; out[7] = P.nb;
    return 0;
}

// flags: bit 0 = generic kernel only, bit 1 = sing_detect, bit 2 = verify, bit 3 = four teams per CTA, bit 4 = re-align
// the teams every iteration (the experiment switch of solve_fused3); rowcol_in may be
// null (detected from G).
// order: fibre schedule of the emulator (0 forward, 1 reverse, 2 random).  Returns 0, or -1 when the plan does not fit.
extern "C" int emu_fused3_solve(int n, int p, int k, int ncones, const int* kind, const int* offs, const int* dim,
                                int batch, const double* c, const double* A, int64_t sA, const double* b,
                                const double* G, int64_t sG, const double* h, const unsigned char* sing,
                                const int* rowcol_in, int flags, int order, int max_iter, double tol, double damp,
                                double init_eps, double* x, double* y, double* z, double* s, int* status, int* iters,
                                double* pobj, double* dobj, unsigned char* sing_out, int* npattern, double* dbg,
                                int dbg_prob, int dbg_iter, int dbg_phase, int grid_cap) {
    std::vector<int> rc = rowcol_in ? std::vector<int>(rowcol_in, rowcol_in + k) : row_pattern(G, sG, batch, n, k);
    F3Plan P;
    std::vector<int> tables;
    f3_plan(P, n, p, k, std::vector<int>(kind, kind + ncones), std::vector<int>(offs, offs + ncones),
            std::vector<int>(dim, dim + ncones), rc, 227 * 1024, 148, tables);
    if (!P.fits) return -1;
    P.d_tables = tables.data();
    int counter = 0;
    std::vector<int> active(batch, 1), fail(batch, 0);
    F3Args a;
    a.g.c = c; a.g.A = A; a.g.b = b; a.g.G = G; a.g.h = h;
    a.g.sA = sA; a.g.sG = sG;
    a.g.sing = sing; a.g.sing_out = sing_out;
    a.g.x = x; a.g.y = y; a.g.z = z; a.g.s = s; a.g.pobj = pobj; a.g.dobj = dobj;
    a.g.status = status; a.g.iters = iters; a.g.active = active.data(); a.g.fail = fail.data();
    int deg = 0;
    for (int i = 0; i < ncones; ++i) deg += (kind[i] == KIND_POC) ? dim[i] : 1;
    a.g.deg = deg;
    a.g.npattern = npattern;
    a.g.dbg = dbg; a.g.dbg_prob = dbg_prob; a.g.dbg_iter = dbg_iter; a.g.dbg_phase = dbg_phase;
    a.P = P;
    a.prm = LoopParams{max_iter, tol, damp, init_eps};
    a.first = 0; a.batch = batch;
    a.counter = &counter;
    a.sing_detect = (flags >> 1) & 1;
    a.verify = (flags >> 2) & 1;
    a.align = (flags >> 4) & 1;
    simt_emu::LaunchCfg cfg;
    const bool teams4 = (flags >> 3) & 1;
    if (teams4 && !P.teams4) return -1;
    cfg.grid = (unsigned)std::max(1, std::min(teams4 ? (batch + 3) / 4 : batch, grid_cap > 0 ? grid_cap : 4));
    cfg.block = teams4 ? 512 : 128;
    cfg.smem = P.smem * (teams4 ? 4 : 1);
    cfg.order = order;
    const bool generic = flags & 1;
    if (teams4) {
        if (!generic && Dims3C2::matches(P)) simt_emu::launch(cfg, [&]() { k_fused3<4, 4, 7, 1, Dims3C2>(a); });
        else if (P.nb <= 4) simt_emu::launch(cfg, [&]() { k_fused3<4, 4, 3, 1, Dims3Dyn>(a); });
        else if (P.nb <= 7) simt_emu::launch(cfg, [&]() { k_fused3<4, 4, 7, 1, Dims3Dyn>(a); });
        else simt_emu::launch(cfg, [&]() { k_fused3<4, 4, 9, 1, Dims3Dyn>(a); });
        return 0;
    }
    if (!generic && Dims3C2::matches(P)) simt_emu::launch(cfg, [&]() { k_fused3<4, 1, 7, 4, Dims3C2>(a); });
    else if (P.nb <= 4) simt_emu::launch(cfg, [&]() { k_fused3<4, 1, 3, 4, Dims3Dyn>(a); });
    else if (P.nb <= 7) simt_emu::launch(cfg, [&]() { k_fused3<4, 1, 7, 4, Dims3Dyn>(a); });
    else simt_emu::launch(cfg, [&]() { k_fused3<4, 1, 9, 3, Dims3Dyn>(a); });
    return 0;
}

// ---- unit harness of the packed-tile kernels: H (n x n, column-major, SPD) -> stage 1: X = L^-1 (lower), stage 2:
// H^-1 = X'X, both returned dense (n x n column-major); y = H^-1 v by the packed symmetric gemv.  Returns the factor's ok flag.
template <int NW, int MAXT>
static void k_tiles_test(int n, const double* Hin, double* Xout, double* Hinv, const double* v, double* y, int* okout) {
    double* sm = reinterpret_cast<double*>(emu_dyn_smem());
    __shared__ int s_fail[2];
    const int tid = (int)(unsigned)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npad = (n + 7) / 8 * 8, nb = npad / 8, ntl = nb * (nb + 1) / 2;
    double* Tt = sm;
    double* Lp = sm + ntl * 64;
    double* xv = Lp + nb * 64;
    unsigned short* tij = reinterpret_cast<unsigned short*>(xv + npad);
    const T3Lane TL = t3_lane(lane);
    for (int t = tid; t < ntl; t += NW * 32) {
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= t) ++i;
        tij[t] = (unsigned short)(i | ((t - i * (i + 1) / 2) << 8));
    }
    for (int e = tid; e < npad * npad; e += NW * 32) {
        const int i = e % npad, j = e / npad;
        if ((i >> 3) < (j >> 3)) continue;
        const double val = (i < n && j < n) ? Hin[j * n + i] : (i == j ? 1.0 : 0.0);
        Tt[t3_idx(i >> 3, j >> 3) * 64 + t3_off(i & 7, j & 7)] = val;
    }
    for (int i = tid; i < npad; i += NW * 32) xv[i] = i < n ? v[i] : 0.0;
    if (tid == 0) s_fail[0] = s_fail[1] = 0;
    tsync<NW>();
    int tl[MAXT];
    f3_my_tiles<NW, MAXT>(nb, warp, tl);
    const int ok = f3_chol_inv<NW>(Tt, Lp, tij, nb, s_fail, lane, warp, TL);
    if (tid == 0) *okout = ok;
    if (!ok) return;
    for (int e = tid; e < n * n; e += NW * 32) {
        const int i = e % n, j = e / n;
        Xout[e] = i >= j ? Tt[t3_idx(i >> 3, j >> 3) * 64 + t3_off(i & 7, j & 7)] : 0.0;
    }
    tsync<NW>();
    f3_xtx<NW, MAXT>(Tt, nb, tl, lane, warp, TL);
    tsync<NW>();
    for (int e = tid; e < n * n; e += NW * 32) Hinv[e] = t3_sym(Tt, e % n, e / n);
    f3_symv<NW>(Tt, nb, xv, lane, warp, TL, [&](int r, double acc) { if (r < n) y[r] = acc; });
}

extern "C" int emu_tiles_test(int n, const double* Hin, double* Xout, double* Hinv, const double* v, double* y, int order) {
    const int npad = (n + 7) / 8 * 8, nb = npad / 8, ntl = nb * (nb + 1) / 2;
    simt_emu::LaunchCfg cfg;
    cfg.grid = 1; cfg.block = 128;
    cfg.smem = (size_t)(ntl * 64 + nb * 64 + npad + ntl + 16) * sizeof(double);
    cfg.order = order;
    int ok = 0;
    if (nb <= 4) simt_emu::launch(cfg, [&]() { k_tiles_test<4, 3>(n, Hin, Xout, Hinv, v, y, &ok); });
    else if (nb <= 7) simt_emu::launch(cfg, [&]() { k_tiles_test<4, 7>(n, Hin, Xout, Hinv, v, y, &ok); });
    else simt_emu::launch(cfg, [&]() { k_tiles_test<4, 9>(n, Hin, Xout, Hinv, v, y, &ok); });
    return ok;
}
