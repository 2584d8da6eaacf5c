/*
 * socp_oracle.c -- CPU oracle (plain C) for the dense KKT hot path of
 * BenChung/Socp.jl.  TEST INFRASTRUCTURE ONLY: nothing under socp.jl_b200/ may
 * link or call this; it is the checker for tests/ and the timed CPU baseline
 * ("port") of bench.py.
 *
 * Same algorithm, step for step, as oracle/socp_oracle.py (which is pinned
 * against the golden vectors of /root/reference/test/runtests.jl); this file is
 * cross-checked against that module in tests/test_oracle_c.py.  It follows the
 * reference's DENSE formulation: dense per-cone blocks of W, W^-1 and
 * iWiW = iW*iW' (src/scalings.jl:70-88,108 -- the k x k matrices are block
 * diagonal, only the blocks are stored), GWiWi = G'*iWiW, H = GWiWi*G, explicit
 * inverse Li = H^-1, AtLi = A*Li, AtLiA = AtLi*A' (src/densesolver.jl:41-52 with
 * the four broken identifiers repaired), solve_kkt as src/densesolver.jl:54-90,
 * driver as src/solver.jl:86-152.  The initial point (src/solver.jl:68-84) uses
 * the block-eliminated form.  "parity unpinned" for
 * dense-solver-specific rounding: the reference never executes its DenseSolver.
 *
 * Julia is not installed, so there is no oracle/_ref build of the reference.
 * Matrices are column-major (Julia layout).  Build: oracle/build_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define POC 0
#define SOC 1
#define ST_CONVERGED 0
#define ST_MAXITER 1
#define ST_NUMERICAL 2

typedef struct {
    int n, p, k, ncones;
    const int *kind, *offs, *dim;
} oc_layout;

/* ------------------------------------------------------------------ dense helpers */
/* C(m x n) = A(m x kk) * B(kk x n) */
static void gemm_nn(int m, int n, int kk, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
    for (int j = 0; j < n; ++j) {
        double* cj = C + (size_t)j * ldc;
        for (int i = 0; i < m; ++i) cj[i] = 0.0;
        for (int l = 0; l < kk; ++l) {
            const double blj = B[(size_t)j * ldb + l];
            const double* al = A + (size_t)l * lda;
            for (int i = 0; i < m; ++i) cj[i] += al[i] * blj;
        }
    }
}
/* C(m x n) = A'(m x kk; A is kk x m) * B(kk x n) */
static void gemm_tn(int m, int n, int kk, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
            const double* ai = A + (size_t)i * lda;
            const double* bj = B + (size_t)j * ldb;
            double acc = 0.0;
            for (int l = 0; l < kk; ++l) acc += ai[l] * bj[l];
            C[(size_t)j * ldc + i] = acc;
        }
}
/* C(m x n) = A(m x kk) * B'(kk x n; B is n x kk) */
static void gemm_nt(int m, int n, int kk, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
    for (int j = 0; j < n; ++j) {
        double* cj = C + (size_t)j * ldc;
        for (int i = 0; i < m; ++i) cj[i] = 0.0;
        for (int l = 0; l < kk; ++l) {
            const double bjl = B[(size_t)l * ldb + j];
            const double* al = A + (size_t)l * lda;
            for (int i = 0; i < m; ++i) cj[i] += al[i] * bjl;
        }
    }
}
static void gemv_n(int m, int n, const double* A, int lda, const double* x, double* y) { /* y = A x */
    for (int i = 0; i < m; ++i) y[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        const double xj = x[j];
        const double* aj = A + (size_t)j * lda;
        for (int i = 0; i < m; ++i) y[i] += aj[i] * xj;
    }
}
static void gemv_t(int m, int n, const double* A, int lda, const double* x, double* y) { /* y = A' x */
    for (int j = 0; j < n; ++j) {
        const double* aj = A + (size_t)j * lda;
        double acc = 0.0;
        for (int i = 0; i < m; ++i) acc += aj[i] * x[i];
        y[j] = acc;
    }
}
/* cholesky!(Hermitian(A)) lower, in place; returns 0, or 1 where LAPACK dpotrf would
 * report info > 0 (Julia: PosDefException).  Left-looking by columns. */
static int chol_lower(int n, double* A, int lda) {
    for (int j = 0; j < n; ++j) {
        double* aj = A + (size_t)j * lda;
        for (int l = 0; l < j; ++l) {
            const double* al = A + (size_t)l * lda;
            const double ajl = al[j];
            for (int i = j; i < n; ++i) aj[i] -= al[i] * ajl;
        }
        const double piv = aj[j];
        if (!(piv > 0.0) || !isfinite(piv)) return 1;
        const double r = sqrt(piv);
        aj[j] = r;
        for (int i = j + 1; i < n; ++i) aj[i] /= r;
    }
    return 0;
}
/* B <- (L L')^-1 B, B is n x nrhs */
static void chol_solve(int n, const double* L, int ldl, double* B, int ldb, int nrhs) {
    for (int c = 0; c < nrhs; ++c) {
        double* x = B + (size_t)c * ldb;
        for (int j = 0; j < n; ++j) {
            const double* lj = L + (size_t)j * ldl;
            const double xj = x[j] / lj[j];
            x[j] = xj;
            for (int i = j + 1; i < n; ++i) x[i] -= lj[i] * xj;
        }
        for (int j = n - 1; j >= 0; --j) {
            const double* lj = L + (size_t)j * ldl;
            double acc = x[j];
            for (int i = j + 1; i < n; ++i) acc -= lj[i] * x[i];
            x[j] = acc / lj[j];
        }
    }
}

/* ------------------------------------------------------------------ src/vectors.jl */
static void make_e(const oc_layout* L, double* r) { /* :7-24 */
    for (int c = 0; c < L->ncones; ++c)
        for (int i = 0; i < L->dim[c]; ++i) r[L->offs[c] + i] = (L->kind[c] == POC || i == 0) ? 1.0 : 0.0;
}
static void vprod(const oc_layout* L, double* t, const double* u, const double* v) { /* :58-81 */
    for (int c = 0; c < L->ncones; ++c) {
        const int o = L->offs[c], d = L->dim[c];
        if (L->kind[c] == POC) {
            for (int i = o; i < o + d; ++i) t[i] = u[i] * v[i];
        } else {
            double acc = 0.0;
            for (int i = o; i < o + d; ++i) acc += u[i] * v[i];
            const double iu = u[o], iv = v[o];
            for (int i = o + 1; i < o + d; ++i) t[i] = iu * v[i] + iv * u[i];
            t[o] = acc;
        }
    }
}
static void iprod(const oc_layout* L, double* t, const double* lam, const double* v) { /* :99-131 */
    for (int c = 0; c < L->ncones; ++c) {
        const int o = L->offs[c], d = L->dim[c];
        if (L->kind[c] == POC) {
            for (int i = o; i < o + d; ++i) t[i] = v[i] / lam[i];
        } else {
            const double l1 = lam[o];
            double a = l1 * l1;
            for (int i = o + 1; i < o + d; ++i) a -= lam[i] * lam[i];
            for (int i = o; i < o + d; ++i) t[i] = 0.0;
            t[o] += v[o] * l1 / a;
            for (int j = o + 1; j < o + d; ++j) t[o] -= v[j] * lam[j] / a;
            for (int i = o + 1; i < o + d; ++i) {
                t[i] -= v[o] * lam[i] / a;
                for (int j = o + 1; j < o + d; ++j)
                    t[i] += v[j] * ((i == j ? a : 0.0) + lam[i] * lam[j]) / (l1 * a);
            }
        }
    }
}
static int deg(const oc_layout* L) { /* :165-179 */
    int dg = 0;
    for (int c = 0; c < L->ncones; ++c) dg += (L->kind[c] == POC) ? L->dim[c] : 1;
    return dg;
}

/* ------------------------------------------------------------------ src/mats.jl */
static double max_step(const oc_layout* L, const double* x) { /* :1-28 */
    double maxim = -INFINITY;
    for (int c = 0; c < L->ncones; ++c) {
        const int o = L->offs[c], d = L->dim[c];
        double val;
        if (L->kind[c] == POC) {
            double minim = INFINITY;
            for (int i = o; i < o + d; ++i)
                if (x[i] < minim) minim = x[i];
            val = -minim;
        } else {
            double sq = 0.0;
            for (int i = o + 1; i < o + d; ++i) sq += x[i] * x[i];
            val = sqrt(sq) - x[o];
        }
        if (val > maxim) maxim = val;
    }
    return maxim;
}
static double scmax(const oc_layout* L, const double* l, const double* x, int* fail) { /* :42-86 */
    double mxv = -INFINITY;
    for (int c = 0; c < L->ncones; ++c) {
        const int o = L->offs[c], d = L->dim[c];
        double val;
        if (L->kind[c] == POC) {
            val = -INFINITY;
            for (int i = o; i < o + d; ++i) {
                const double q = -x[i] / l[i];
                if (q > val) val = q;
            }
        } else {
            double ai = l[o] * l[o];
            for (int i = o + 1; i < o + d; ++i) ai -= l[i] * l[i];
            if (ai < 0.0) *fail = 1;
            const double a = 1.0 / sqrt(ai);
            double r1 = a * l[o] * x[o];
            for (int i = o + 1; i < o + d; ++i) r1 -= a * l[i] * x[i];
            const double cst = (r1 + x[o]) / (a * l[o] + 1.0);
            double r2s = 0.0;
            for (int i = o + 1; i < o + d; ++i) {
                const double q = a * (x[i] - cst * a * l[i]);
                r2s += q * q;
            }
            val = sqrt(r2s) - a * r1;
        }
        if (val > mxv) mxv = val;
    }
    return mxv;
}
static double compute_step(const oc_layout* L, const double* l, const double* ds, const double* dz, int* fail) { /* :30-40 */
    const double mxs = scmax(L, l, ds, fail), mxz = scmax(L, l, dz, fail);
    double t = mxs > mxz ? mxs : mxz;
    if (!(t > 0.0)) t = 0.0;
    if (t == 0.0) return 1.0;
    return 1.0 / t < 1.0 ? 1.0 / t : 1.0;
}

/* ------------------------------------------------------------------ src/scalings.jl */
typedef struct {
    double *l, *wbs, *mu;      /* lambda[k], wbs[k], mu[ncones]          */
    double **Wb, **iWb, **iWiWb; /* dense d x d blocks per cone (POC: diagonal stored dense d x d would be
                                    wasteful -> POC blocks store only the diagonal, length d) */
    double *sik, *zik;
} oc_scaling;

static oc_scaling* scaling_new(const oc_layout* L) {
    oc_scaling* s = (oc_scaling*)calloc(1, sizeof *s);
    int md = 1;
    s->l = (double*)calloc(L->k, sizeof(double));
    s->wbs = (double*)calloc(L->k, sizeof(double));
    s->mu = (double*)calloc(L->ncones, sizeof(double));
    s->Wb = (double**)calloc(L->ncones, sizeof(double*));
    s->iWb = (double**)calloc(L->ncones, sizeof(double*));
    s->iWiWb = (double**)calloc(L->ncones, sizeof(double*));
    for (int c = 0; c < L->ncones; ++c) {
        const size_t d = L->dim[c];
        const size_t sz = (L->kind[c] == POC) ? d : d * d;
        s->Wb[c] = (double*)calloc(sz, sizeof(double));
        s->iWb[c] = (double*)calloc(sz, sizeof(double));
        s->iWiWb[c] = (double*)calloc(sz, sizeof(double));
        if ((int)d > md) md = (int)d;
    }
    s->sik = (double*)calloc(md, sizeof(double));
    s->zik = (double*)calloc(md, sizeof(double));
    return s;
}
static void scaling_free(const oc_layout* L, oc_scaling* s) {
    for (int c = 0; c < L->ncones; ++c) { free(s->Wb[c]); free(s->iWb[c]); free(s->iWiWb[c]); }
    free(s->Wb); free(s->iWb); free(s->iWiWb);
    free(s->l); free(s->wbs); free(s->mu); free(s->sik); free(s->zik);
    free(s);
}
/* returns 1 where Julia's sqrt would throw a DomainError */
static int compute_scaling(const oc_layout* L, oc_scaling* sc, const double* s, const double* z) { /* :101-110 */
    for (int c = 0; c < L->ncones; ++c) {
        const int o = L->offs[c], d = L->dim[c];
        if (L->kind[c] == POC) { /* :22-30 */
            for (int i = 0; i < d; ++i) {
                const int ii = o + i;
                const double q = s[ii] / z[ii], qi = z[ii] / s[ii], pz = s[ii] * z[ii];
                if (q < 0.0 || qi < 0.0 || pz < 0.0) return 1;
                sc->Wb[c][i] = sqrt(q);
                sc->iWb[c][i] = sqrt(qi);
                sc->l[ii] = sqrt(pz);
                sc->wbs[ii] = sqrt(q);
                sc->iWiWb[c][i] = sc->iWb[c][i] * sc->iWb[c][i];   /* (iW*iW')[i,i] */
            }
            sc->mu[c] = 0.0;
            continue;
        }
        /* :32-99 */
        double* sik = sc->sik;
        double* zik = sc->zik;
        for (int i = 0; i < d; ++i) { sik[i] = s[o + i]; zik[i] = z[o + i]; }
        double onrmz = zik[0] * zik[0], onrms = sik[0] * sik[0];
        for (int i = 1; i < d; ++i) { onrmz -= zik[i] * zik[i]; onrms -= sik[i] * sik[i]; }
        if (onrmz < 0.0 || onrms < 0.0) return 1;
        const double nrmz = sqrt(onrmz), nrms = sqrt(onrms);
        const double iz = 1.0 / nrmz, is = 1.0 / nrms;
        for (int i = 0; i < d; ++i) { zik[i] *= iz; sik[i] *= is; }
        double nsum = 0.0;
        for (int i = 0; i < d; ++i) nsum += zik[i] * sik[i];
        if ((1.0 + nsum) / 2.0 < 0.0) return 1;
        const double gamma = sqrt((1.0 + nsum) / 2.0);
        double* wb = sc->wbs + o;
        wb[0] = sik[0] + zik[0];
        for (int i = 1; i < d; ++i) wb[i] = sik[i] - zik[i];
        const double ig = 1.0 / (2.0 * gamma);
        for (int i = 0; i < d; ++i) wb[i] *= ig;
        const double denom = wb[0] + 1.0;
        if (nrms / nrmz < 0.0) return 1;
        const double mu = sqrt(nrms / nrmz);
        sc->mu[c] = mu;
        double* W = sc->Wb[c];
        double* iW = sc->iWb[c];
        for (int j = 1; j < d; ++j)
            for (int i = 1; i < d; ++i) {
                const double cellv = ((i == j) ? 1.0 : 0.0) + wb[i] * wb[j] / denom;
                W[(size_t)j * d + i] = cellv * mu;
                iW[(size_t)j * d + i] = cellv / mu;
            }
        for (int i = 0; i < d; ++i) W[(size_t)i * d + 0] = wb[i] * mu;   /* first row */
        iW[0] = wb[0] / mu;
        for (int i = 1; i < d; ++i) {
            W[i] = wb[i] * mu;                                          /* first column */
            iW[(size_t)i * d + 0] = -wb[i] / mu;
            iW[i] = -wb[i] / mu;
        }
        const double ziv = zik[0], siv = sik[0];
        if (nrms * nrmz < 0.0) return 1;
        const double tmv1 = sqrt(nrms * nrmz);
        const double mult = tmv1 / (ziv + siv + 2.0 * gamma);
        for (int i = 0; i < d; ++i) { sik[i] *= (gamma + ziv); zik[i] *= (gamma + siv); }
        for (int i = 1; i < d; ++i) sc->l[o + i] = (sik[i] + zik[i]) * mult;
        sc->l[o] = gamma * tmv1;
        /* mul!(iWiW, iW, iW')  :108, block by block */
        gemm_nt(d, d, d, iW, d, iW, d, sc->iWiWb[c], d);
    }
    return 0;
}
static void scale(const oc_layout* L, const oc_scaling* sc, const double* s, double* op) { /* :112-117,:126-140 */
    const double* wb = sc->wbs;
    for (int c = 0; c < L->ncones; ++c) {
        const int o = L->offs[c], d = L->dim[c];
        if (L->kind[c] == POC) {
            for (int i = o; i < o + d; ++i) op[i] = wb[i] * s[i];
        } else {
            const double mu = sc->mu[c];
            double del = 0.0;
            for (int i = o + 1; i < o + d; ++i) del += wb[i] * s[i];
            const double cst = s[o] + del / (1.0 + wb[o]);
            const double s0 = s[o];
            for (int i = o + 1; i < o + d; ++i) op[i] = mu * (s[i] + cst * wb[i]);
            op[o] = mu * (wb[o] * s0 + del);
        }
    }
}
static void iscale(const oc_layout* L, const oc_scaling* sc, const double* s, double* op) { /* :119-124,:142-156 */
    const double* wb = sc->wbs;
    for (int c = 0; c < L->ncones; ++c) {
        const int o = L->offs[c], d = L->dim[c];
        if (L->kind[c] == POC) {
            for (int i = o; i < o + d; ++i) op[i] = 1.0 / wb[i] * s[i];
        } else {
            const double mu = sc->mu[c];
            double del = 0.0;
            for (int i = o + 1; i < o + d; ++i) del += wb[i] * s[i];
            const double cst = -s[o] + del / (1.0 + wb[o]);
            const double s0 = s[o];
            for (int i = o + 1; i < o + d; ++i) op[i] = 1.0 / mu * (s[i] + cst * wb[i]);
            op[o] = 1.0 / mu * (wb[o] * s0 - del);
        }
    }
}
/* out = iWiW * v (block diagonal) */
static void iwiw_mul(const oc_layout* L, const oc_scaling* sc, const double* v, double* out) {
    for (int c = 0; c < L->ncones; ++c) {
        const int o = L->offs[c], d = L->dim[c];
        if (L->kind[c] == POC) {
            for (int i = 0; i < d; ++i) out[o + i] = sc->iWiWb[c][i] * v[o + i];
        } else {
            gemv_n(d, d, sc->iWiWb[c], d, v + o, out + o);
        }
    }
}

/* ------------------------------------------------------------------ src/densesolver.jl */
typedef struct {
    double *k0, *k1, *k2, *m0, *n0, *n1;
    double *GWiWi, *H, *Li, *AtLi, *AtLiA, *AA, *eye;
} oc_dense;

static oc_dense* dense_new(const oc_layout* L, const double* A) { /* :21-39 */
    const size_t n = L->n, p = L->p, k = L->k;
    oc_dense* d = (oc_dense*)calloc(1, sizeof *d);
    d->k0 = (double*)calloc(k, 8); d->k1 = (double*)calloc(k, 8); d->k2 = (double*)calloc(k, 8);
    d->m0 = (double*)calloc(p + 1, 8); d->n0 = (double*)calloc(n, 8); d->n1 = (double*)calloc(n, 8);
    d->GWiWi = (double*)calloc(n * k, 8);
    d->H = (double*)calloc(n * n, 8);
    d->Li = (double*)calloc(n * n, 8);
    d->AtLi = (double*)calloc(p * n + 1, 8);
    d->AtLiA = (double*)calloc(p * p + 1, 8);
    d->AA = (double*)calloc(n * n, 8);
    d->eye = NULL;
    if (p > 0) gemm_tn((int)n, (int)n, (int)p, A, (int)p, A, (int)p, d->AA, (int)n);   /* :32 */
    return d;
}
static void dense_free(oc_dense* d) {
    free(d->k0); free(d->k1); free(d->k2); free(d->m0); free(d->n0); free(d->n1);
    free(d->GWiWi); free(d->H); free(d->Li); free(d->AtLi); free(d->AtLiA); free(d->AA);
    free(d);
}
/* identity != 0: W = I (initial point).  returns 1 on PosDefException */
static int setup_iter(const oc_layout* L, oc_dense* ds, const oc_scaling* sc, const double* A, const double* G,
                      int sing, int identity) { /* :41-52 */
    const int n = L->n, p = L->p, k = L->k;
    /* GWiWi (n x k) = G' * iWiW   :42 */
    for (int c = 0; c < L->ncones; ++c) {
        const int o = L->offs[c], d = L->dim[c];
        if (identity) {
            for (int j = 0; j < d; ++j)
                for (int i = 0; i < n; ++i) ds->GWiWi[(size_t)(o + j) * n + i] = G[(size_t)i * k + o + j];
        } else if (L->kind[c] == POC) {
            for (int j = 0; j < d; ++j) {
                const double w = sc->iWiWb[c][j];
                for (int i = 0; i < n; ++i) ds->GWiWi[(size_t)(o + j) * n + i] = G[(size_t)i * k + o + j] * w;
            }
        } else {
            /* block: (G_c)' (d x n)' * iWiW_c (d x d) -> n x d */
            gemm_tn(n, d, d, G + o, k, sc->iWiWb[c], d, ds->GWiWi + (size_t)o * n, n);
        }
    }
    gemm_nn(n, n, k, ds->GWiWi, n, G, k, ds->H, n);                    /* :43 */
    if (sing)
        for (size_t i = 0; i < (size_t)n * n; ++i) ds->H[i] += ds->AA[i];   /* :44-46 */
    for (int j = 0; j < n; ++j)       /* Hermitian(): use one triangle consistently (lower) */
        for (int i = 0; i < j; ++i) ds->H[(size_t)j * n + i] = ds->H[(size_t)i * n + j];
    if (chol_lower(n, ds->H, n)) return 1;                              /* :47 */
    memset(ds->Li, 0, sizeof(double) * (size_t)n * n);
    for (int i = 0; i < n; ++i) ds->Li[(size_t)i * n + i] = 1.0;
    chol_solve(n, ds->H, n, ds->Li, n, n);                              /* :48 */
    if (p > 0) {
        gemm_nn(p, n, n, A, p, ds->Li, n, ds->AtLi, p);                 /* :49 */
        gemm_nt(p, p, n, ds->AtLi, p, A, p, ds->AtLiA, p);              /* :50 */
        if (chol_lower(p, ds->AtLiA, p)) return 1;                      /* :51 */
    }
    return 0;
}
static void solve_kkt(const oc_layout* L, oc_dense* ds, const oc_scaling* sc, const double* A, const double* G,
                      int sing, const double* dx, const double* dy, const double* dz, const double* dsv,
                      double* cx, double* cy, double* cz, double* cs) { /* :54-90 */
    const int n = L->n, p = L->p, k = L->k;
    iprod(L, ds->k0, sc->l, dsv);                                       /* :61 */
    scale(L, sc, ds->k0, ds->k1);                                       /* :62 */
    for (int i = 0; i < k; ++i) ds->k2[i] = dz[i] - ds->k1[i];          /* :65 */
    gemv_n(n, k, ds->GWiWi, n, ds->k2, ds->n0);                         /* :66 */
    for (int i = 0; i < n; ++i) ds->n0[i] += dx[i];                     /* :67 */
    if (sing && p > 0) {                                                /* :69-71 */
        gemv_t(p, n, A, p, dy, ds->n1);
        for (int i = 0; i < n; ++i) ds->n0[i] += ds->n1[i];
    }
    if (p > 0) {
        gemv_n(p, n, ds->AtLi, p, ds->n0, ds->m0);                      /* :73 */
        for (int i = 0; i < p; ++i) ds->m0[i] -= dy[i];                 /* :74 */
        for (int i = 0; i < p; ++i) cy[i] = ds->m0[i];
        chol_solve(p, ds->AtLiA, p, cy, p, 1);                          /* :75 */
        for (int i = 0; i < p; ++i) ds->m0[i] = sing ? dy[i] - cy[i] : -cy[i];   /* :76-80 */
        gemv_t(p, n, A, p, ds->m0, ds->n1);                             /* :81 */
        for (int i = 0; i < n; ++i) ds->n0[i] += ds->n1[i];             /* :82 */
    }
    gemv_n(n, n, ds->Li, n, ds->n0, cx);                                /* :83 */
    gemv_n(k, n, G, k, cx, ds->k1);                                     /* :84 */
    for (int i = 0; i < k; ++i) ds->k1[i] -= ds->k2[i];                 /* :85 */
    iwiw_mul(L, sc, ds->k1, cz);                                        /* :86 */
    scale(L, sc, cz, ds->k1);                                           /* :87 */
    for (int i = 0; i < k; ++i) ds->k0[i] -= ds->k1[i];                 /* :88 */
    scale(L, sc, ds->k0, cs);                                           /* :89 */
}

/* sing of Problem{...,sing}: cholesky(G'G) throws.  src/Socp.jl:49-56 */
int oc_detect_sing(const oc_layout* L, const double* G) {
    const int n = L->n, k = L->k;
    double* H = (double*)malloc(sizeof(double) * (size_t)n * n);
    gemm_tn(n, n, k, G, k, G, k, H, n);
    const int f = chol_lower(n, H, n);
    free(H);
    return f;
}

/* ------------------------------------------------------------------ src/solver.jl */
static int all_finite(const double* v, int n) {
    for (int i = 0; i < n; ++i)
        if (!isfinite(v[i])) return 0;
    return 1;
}

/* solve_socp for one problem.  sing < 0: detect.  Returns status. */
int oc_solve(const oc_layout* L, const double* c, const double* A, const double* b, const double* G,
             const double* h, int sing, int max_iter, double tol, double step_damp, double init_eps,
             double* x, double* y, double* z, double* s, int* iters_out, double* pobj, double* dobj) {
    const int n = L->n, p = L->p, k = L->k;
    if (sing < 0) sing = oc_detect_sing(L, G);
    oc_scaling* sc = scaling_new(L);
    oc_dense* ds = dense_new(L, A);
    double* buf = (double*)calloc((size_t)(4 * n + 4 * p + 12 * k + 16), sizeof(double));
    double *dx = buf, *rx = dx + n, *nt = rx + n, *t0 = nt + n;
    double *dy = t0 + n, *ry = dy + p + 1, *mt = ry + p + 1;
    double *dz = mt + 2 * p + 2, *dsv = dz + k, *rz = dsv + k, *rs = rz + k, *kt1 = rs + k, *kt2 = kt1 + k, *kt3 = kt2 + k,
           *idel = kt3 + k, *iz = idel + k, *tmpk = iz + k;
    int status = ST_MAXITER, iters = 0;
    /* ---- initial point (src/solver.jl:68-84 by block elimination, W = I) */
    memset(x, 0, sizeof(double) * n);
    if (p) memset(y, 0, sizeof(double) * p);
    memset(z, 0, sizeof(double) * k);
    memset(s, 0, sizeof(double) * k);
    if (setup_iter(L, ds, sc, A, G, sing, 1)) { status = ST_NUMERICAL; goto done; }
    gemv_t(k, n, G, k, h, nt);                                  /* G'h */
    for (int i = 0; i < n; ++i) nt[i] -= c[i];                  /* n0 = -c + G'h */
    if (sing && p > 0) {
        gemv_t(p, n, A, p, b, t0);
        for (int i = 0; i < n; ++i) nt[i] += t0[i];
    }
    gemv_n(n, n, ds->Li, n, nt, t0);                            /* t = H^-1 n0 */
    if (p > 0) {
        gemv_n(p, n, A, p, t0, y);
        for (int i = 0; i < p; ++i) y[i] -= b[i];
        chol_solve(p, ds->AtLiA, p, y, p, 1);                   /* y = M^-1 (A t - b) */
        gemv_t(p, n, A, p, y, rx);
        for (int i = 0; i < n; ++i) nt[i] -= rx[i];
        gemv_n(n, n, ds->Li, n, nt, x);                         /* x = H^-1 (n0 - A'y) */
    } else {
        for (int i = 0; i < n; ++i) x[i] = t0[i];
    }
    gemv_n(k, n, G, k, x, iz);
    for (int i = 0; i < k; ++i) iz[i] -= h[i];                  /* z0 = G x - h */
    make_e(L, idel);                                            /* :86 */
    for (int i = 0; i < k; ++i) tmpk[i] = -iz[i];
    {
        const double alphp = max_step(L, tmpk);                 /* :88 */
        const double alphd = max_step(L, iz);                   /* :89 */
        for (int i = 0; i < k; ++i) {
            s[i] = (fabs(alphp) < init_eps) ? -iz[i] : -iz[i] + (1.0 + alphp) * idel[i];   /* :91-95 */
            z[i] = (fabs(alphd) < init_eps) ? iz[i] : iz[i] + (1.0 + alphd) * idel[i];     /* :97-101 */
        }
    }
    {
        const int dg = deg(L);
        for (int it = 0; it < max_iter; ++it) {                 /* :105 */
            int fail = compute_scaling(L, sc, s, z);            /* :106 */
            if (fail) { status = ST_NUMERICAL; break; }
            const double* l = sc->l;
            gemv_t(k, n, G, k, z, dx);                          /* :110-112 */
            if (p > 0) { gemv_t(p, n, A, p, y, nt); for (int i = 0; i < n; ++i) dx[i] += nt[i]; }
            for (int i = 0; i < n; ++i) dx[i] += c[i];
            if (p > 0) { gemv_n(p, n, A, p, x, dy); for (int i = 0; i < p; ++i) dy[i] -= b[i]; }   /* :114-115 */
            gemv_n(k, n, G, k, x, dz);                          /* :117-118 */
            for (int i = 0; i < k; ++i) dz[i] += s[i] - h[i];
            vprod(L, dsv, l, l);                                /* :120 */
            double nx = 0.0, ny = 0.0, gap = 0.0;
            for (int i = 0; i < n; ++i) nx += dx[i] * dx[i];
            for (int i = 0; i < p; ++i) ny += dy[i] * dy[i];
            for (int i = 0; i < k; ++i) gap += z[i] * s[i];
            if (sqrt(nx) + sqrt(ny) + gap < tol) { status = ST_CONVERGED; break; }   /* :122-124 */
            for (int i = 0; i < n; ++i) dx[i] = -dx[i];         /* :125 */
            for (int i = 0; i < p; ++i) dy[i] = -dy[i];
            for (int i = 0; i < k; ++i) { dz[i] = -dz[i]; dsv[i] = -dsv[i]; }
            if (setup_iter(L, ds, sc, A, G, sing, 0)) { status = ST_NUMERICAL; break; }   /* :126 */
            solve_kkt(L, ds, sc, A, G, sing, dx, dy, dz, dsv, rx, ry, rz, rs);            /* :127 */
            scale(L, sc, rz, kt3);                              /* :128 */
            iscale(L, sc, rs, kt2);                             /* :129 */
            const double t = compute_step(L, l, kt3, kt2, &fail);   /* :130 */
            double ll = 0.0, d23 = 0.0;
            for (int i = 0; i < k; ++i) { ll += l[i] * l[i]; d23 += kt2[i] * kt3[i]; }
            const double rho = 1.0 - t - t * t * d23 / ll;      /* :132 */
            double cl = rho < 1.0 ? rho : 1.0;
            if (!(cl > 0.0)) cl = 0.0;
            const double sig = cl * cl * cl;                    /* :133 */
            const double mu = ll / dg;                          /* :134 */
            const double scf = 1.0 - sig;                       /* :136 */
            vprod(L, kt1, kt2, kt3);                            /* :137 */
            for (int i = 0; i < k; ++i) dsv[i] += sig * mu * idel[i] - kt1[i];   /* :138-139 */
            for (int i = 0; i < n; ++i) dx[i] *= scf;           /* :140 */
            for (int i = 0; i < p; ++i) dy[i] *= scf;
            for (int i = 0; i < k; ++i) dz[i] *= scf;
            solve_kkt(L, ds, sc, A, G, sing, dx, dy, dz, dsv, rx, ry, rz, rs);   /* :141 */
            scale(L, sc, rz, kt3);                              /* :143 */
            iscale(L, sc, rs, kt2);                             /* :144 */
            double step = compute_step(L, l, kt3, kt2, &fail);  /* :145 */
            step *= step_damp;                                  /* :146 */
            if (fail || !all_finite(rx, n) || !all_finite(ry, p) || !all_finite(rz, k) || !all_finite(rs, k) ||
                !isfinite(step)) {
                status = ST_NUMERICAL;
                break;
            }
            for (int i = 0; i < n; ++i) x[i] += rx[i] * step;   /* :147-150 */
            for (int i = 0; i < p; ++i) y[i] += ry[i] * step;
            for (int i = 0; i < k; ++i) { z[i] += rz[i] * step; s[i] += rs[i] * step; }
            ++iters;
        }
    }
done:
    if (iters_out) *iters_out = iters;
    if (pobj) { double a = 0.0; for (int i = 0; i < n; ++i) a += c[i] * x[i]; *pobj = a; }
    if (dobj) {
        double a = 0.0;
        for (int i = 0; i < p; ++i) a -= b[i] * y[i];
        for (int i = 0; i < k; ++i) a -= h[i] * z[i];
        *dobj = a;
    }
    free(buf);
    dense_free(ds);
    scaling_free(L, sc);
    return status;
}

/* Batch: one problem per OpenMP thread (the timed CPU baseline).  sing may be NULL (detect).
 * strideA / strideG = 0 for a matrix shared by the batch. */
int oc_solve_batch(const oc_layout* L, int64_t batch, const double* c, const double* A, int64_t strideA,
                   const double* b, const double* G, int64_t strideG, const double* h, const uint8_t* sing,
                   int max_iter, double tol, double step_damp, double init_eps, int nthreads,
                   double* x, double* y, double* z, double* s, int32_t* status, int32_t* iters,
                   double* pobj, double* dobj) {
    const int n = L->n, p = L->p, k = L->k;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t q = 0; q < batch; ++q) {
        int it = 0;
        double po = 0.0, dobjv = 0.0;
        const int st = oc_solve(L, c + q * n, A + q * strideA, b + q * p, G + q * strideG, h + q * k,
                                sing ? (int)sing[q] : -1, max_iter, tol, step_damp, init_eps,
                                x + q * n, y + q * p, z + q * k, s + q * k, &it, &po, &dobjv);
        status[q] = st;
        iters[q] = it;
        if (pobj) pobj[q] = po;
        if (dobj) dobj[q] = dobjv;
    }
    return 0;
}

/* Step-level entry for tests: compute_scaling + setup_iter + solve_kkt from given (s, z) and rhs. */
int oc_kkt_step(const oc_layout* L, const double* A, const double* G, int sing, const double* s, const double* z,
                const double* dx, const double* dy, const double* dz, const double* dsv,
                double* cx, double* cy, double* cz, double* cs, double* lambda, double* wbs, double* mu) {
    oc_scaling* sc = scaling_new(L);
    oc_dense* ds = dense_new(L, A);
    int rc = compute_scaling(L, sc, s, z);
    if (!rc) rc = 2 * setup_iter(L, ds, sc, A, G, sing, 0);
    if (!rc) solve_kkt(L, ds, sc, A, G, sing, dx, dy, dz, dsv, cx, cy, cz, cs);
    if (lambda) memcpy(lambda, sc->l, sizeof(double) * L->k);
    if (wbs) memcpy(wbs, sc->wbs, sizeof(double) * L->k);
    if (mu) memcpy(mu, sc->mu, sizeof(double) * L->ncones);
    dense_free(ds);
    scaling_free(L, sc);
    return rc;
}

int oc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
