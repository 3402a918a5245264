/*
 * acopf_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 * See acopf_oracle.h for the pinning status. Compile with
 *   gcc -O2 -ffp-contract=off -fPIC -shared [-fopenmp] acopf_oracle.c -lm
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the reference repository root, src/models/acopf/ unless stated otherwise).
 * The TRON routines follow SURVEY.md Appendix B (public TRON 1.2 algorithm;
 * ExaTron.jl itself is not vendored in the reference).
 *
 * State is held in the REFERENCE layout (one vector of nvar doubles,
 * generators first, 8 entries per line) with 0-based indices.
 */
#include "acopf_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_internal.h"

/* ------------------------------------------------------------------------ */
/* small helpers                                                            */
/* ------------------------------------------------------------------------ */
static double *dup_d(const double *src, int64_t n) {
    double *p = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (p && src && n > 0) memcpy(p, src, sizeof(double) * (size_t)n);
    return p;
}
static int64_t *dup_i_0based(const int64_t *src, int64_t n) {
    int64_t *p = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    if (p && src) for (int64_t i = 0; i < n; ++i) p[i] = src[i] - 1;
    return p;
}
static double wall(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static double nrm2(int n, const double *x) {
    double s = 0.0; for (int i = 0; i < n; ++i) s += x[i] * x[i]; return sqrt(s);
}
static double dot(int n, const double *x, const double *y) {
    double s = 0.0; for (int i = 0; i < n; ++i) s += x[i] * y[i]; return s;
}
static double vnorm(const double *x, int64_t n) {
    double s = 0.0; for (int64_t i = 0; i < n; ++i) s += x[i] * x[i]; return sqrt(s);
}
static double dmin(double a, double b) { return a < b ? a : b; }
static double dmax(double a, double b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------ */
/* branch objective: f, grad, Hessian                                       */
/* (acopf_eval_linelimit_kernel_cpu.jl:1-605, restated in the compact form  */
/*  of SURVEY.md Appendix A.4: every flow is a*vi^2 + b*vj^2 + vi*vj*P(t))  */
/* ------------------------------------------------------------------------ */
typedef struct { double a, b, P, Q, F; } flow_t;

#include "portable_sincos.h"
static int g_portable_sincos = 0;
void orc_set_portable_sincos(int on) { g_portable_sincos = on; }
static void orc_sincos(double t, double *s, double *c) {
    if (g_portable_sincos) psincos(t, s, c);
    else { *s = sin(t); *c = cos(t); }
}

static void flow_terms(const double x[6], const double Y[8], flow_t fl[4]) {
    const double vi = x[0], vj = x[1], t = x[2] - x[3];
    double c, s;
    orc_sincos(t, &s, &c);
    /* rows: pij, qij, pji, qji (eval_cpu.jl:13-16) */
    const double a[4] = { Y[0], -Y[1], 0.0, 0.0 };
    const double b[4] = { 0.0, 0.0, Y[4], -Y[5] };
    const double ga[4] = { Y[2], -Y[3], Y[6], -Y[7] };
    const double de[4] = { Y[3], Y[2], -Y[7], -Y[6] };
    for (int k = 0; k < 4; ++k) {
        fl[k].a = a[k]; fl[k].b = b[k];
        fl[k].P = ga[k] * c + de[k] * s;
        fl[k].Q = -ga[k] * s + de[k] * c;
        fl[k].F = a[k] * (vi * vi) + b[k] * (vj * vj) + (vi * vj) * fl[k].P;
    }
}

double orc_eval_f(const double x[6], const double *p, const double Y[8], double scale) {
    flow_t fl[4];
    flow_terms(x, Y, fl);
    const double h[8] = { fl[0].F, fl[1].F, fl[2].F, fl[3].F, x[0] * x[0], x[1] * x[1], x[2], x[3] };
    double f = 0.0;
    for (int k = 0; k < 8; ++k) {
        const double d = h[k] - p[16 + k];
        f += p[k] * h[k] + 0.5 * (p[8 + k] * (d * d));
    }
    const double c1 = fl[0].F * fl[0].F + fl[1].F * fl[1].F + x[4];
    const double c2 = fl[2].F * fl[2].F + fl[3].F * fl[3].F + x[5];
    f += p[24] * c1 + p[25] * c2 + 0.5 * (p[26] * (c1 * c1)) + 0.5 * (p[26] * (c2 * c2));
    return scale * f;
}

/* Same function, accumulated term by term in the reference's order
 * (acopf_eval_linelimit_kernel_cpu.jl:9-43). Used to cross-check orc_eval_f. */
double orc_eval_f_reforder(const double x[6], const double *p, const double Y[8], double scale) {
    const double cc = x[0] * x[1] * cos(x[2] - x[3]);
    const double ss = x[0] * x[1] * sin(x[2] - x[3]);
    const double pij = Y[0] * (x[0] * x[0]) + Y[2] * cc + Y[3] * ss;
    const double qij = -Y[1] * (x[0] * x[0]) - Y[3] * cc + Y[2] * ss;
    const double pji = Y[4] * (x[1] * x[1]) + Y[6] * cc - Y[7] * ss;
    const double qji = -Y[5] * (x[1] * x[1]) - Y[7] * cc - Y[6] * ss;
    const double h[8] = { pij, qij, pji, qji, x[0] * x[0], x[1] * x[1], x[2], x[3] };
    double f = 0.0;
    for (int k = 0; k < 8; ++k) f += p[k] * h[k];
    for (int k = 0; k < 8; ++k) f += 0.5 * (p[8 + k] * ((h[k] - p[16 + k]) * (h[k] - p[16 + k])));
    const double c1 = pij * pij + qij * qij + x[4];
    const double c2 = pji * pji + qji * qji + x[5];
    f += p[24] * c1;
    f += p[25] * c2;
    f += 0.5 * (p[26] * (c1 * c1));
    f += 0.5 * (p[26] * (c2 * c2));
    return f * scale;
}

void orc_eval_gh(const double x[6], const double *p, const double Y[8], double scale,
                 double g[6], double H[36]) {
    flow_t fl[4];
    flow_terms(x, Y, fl);
    const double vi = x[0], vj = x[1];
    const double mu = p[26];
    const double c1 = fl[0].F * fl[0].F + fl[1].F * fl[1].F + x[4];
    const double c2 = fl[2].F * fl[2].F + fl[3].F * fl[3].F + x[5];
    const double m[2] = { p[24] + mu * c1, p[25] + mu * c2 };   /* multiplier estimates */

    double gy[3] = { 0, 0, 0 };        /* gradient in y = (vi, vj, t = ti - tj) */
    double Hy[3][3] = { { 0 } };
    double d[2][3] = { { 0 } };        /* d_j = grad_y (p^2 + q^2) of side j */
    for (int k = 0; k < 4; ++k) {
        const int j = k >> 1;
        const double G[3] = { 2.0 * fl[k].a * vi + vj * fl[k].P,
                              2.0 * fl[k].b * vj + vi * fl[k].P,
                              (vi * vj) * fl[k].Q };
        const double r = p[k] + p[8 + k] * (fl[k].F - p[16 + k]);
        const double w = r + 2.0 * m[j] * fl[k].F;
        const double kap = p[8 + k] + 2.0 * m[j];
        for (int i = 0; i < 3; ++i) {
            gy[i] += w * G[i];
            d[j][i] += 2.0 * fl[k].F * G[i];
            for (int q = 0; q < 3; ++q) Hy[i][q] += kap * G[i] * G[q];
        }
        Hy[0][0] += w * (2.0 * fl[k].a);
        Hy[1][1] += w * (2.0 * fl[k].b);
        Hy[0][1] += w * fl[k].P;          Hy[1][0] += w * fl[k].P;
        Hy[0][2] += w * (vj * fl[k].Q);   Hy[2][0] += w * (vj * fl[k].Q);
        Hy[1][2] += w * (vi * fl[k].Q);   Hy[2][1] += w * (vi * fl[k].Q);
        Hy[2][2] += w * (-(vi * vj) * fl[k].P);
    }
    for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 3; ++i)
            for (int q = 0; q < 3; ++q) Hy[i][q] += mu * d[j][i] * d[j][q];
    /* w_i = vi^2, w_j = vj^2 consensus terms (eval_cpu.jl:236,246) */
    const double ri = p[4] + p[12] * (vi * vi - p[20]);
    const double rj = p[5] + p[13] * (vj * vj - p[21]);
    gy[0] += 2.0 * vi * ri;
    gy[1] += 2.0 * vj * rj;
    Hy[0][0] += 2.0 * ri + 4.0 * p[12] * (vi * vi);
    Hy[1][1] += 2.0 * rj + 4.0 * p[13] * (vj * vj);

    g[0] = scale * gy[0];
    g[1] = scale * gy[1];
    g[2] = scale * (gy[2] + p[6] + p[14] * (x[2] - p[22]));
    g[3] = scale * (-gy[2] + p[7] + p[15] * (x[3] - p[23]));
    g[4] = scale * m[0];
    g[5] = scale * m[1];

    double A[6][6];
    memset(A, 0, sizeof(A));
    A[0][0] = Hy[0][0]; A[0][1] = Hy[0][1]; A[0][2] = Hy[0][2]; A[0][3] = -Hy[0][2];
    A[1][1] = Hy[1][1]; A[1][2] = Hy[1][2]; A[1][3] = -Hy[1][2];
    A[2][2] = Hy[2][2] + p[14]; A[2][3] = -Hy[2][2];
    A[3][3] = Hy[2][2] + p[15];
    A[0][4] = mu * d[0][0]; A[0][5] = mu * d[1][0];
    A[1][4] = mu * d[0][1]; A[1][5] = mu * d[1][1];
    A[2][4] = mu * d[0][2]; A[2][5] = mu * d[1][2];
    A[3][4] = -mu * d[0][2]; A[3][5] = -mu * d[1][2];
    A[4][4] = mu; A[5][5] = mu;          /* eval_cpu.jl:593-600; d2/dsij dsji = 0 */
    for (int i = 0; i < 6; ++i)
        for (int q = i; q < 6; ++q) {
            H[i * 6 + q] = scale * A[i][q];
            H[q * 6 + i] = scale * A[i][q];
        }
}

/* ------------------------------------------------------------------------ */
/* TRON (dense n x n, n <= NV)                          */
/* ------------------------------------------------------------------------ */

static void symv(int n, const double *A, int lda, const double *x, double *y) {
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += A[i * lda + j] * x[j];
        y[i] = s;
    }
}
static void dmid(int n, double *x, const double *xl, const double *xu) {
    for (int i = 0; i < n; ++i) x[i] = dmax(xl[i], dmin(x[i], xu[i]));
}
/* B.1 dgpstep */
static void dgpstep(int n, const double *x, const double *xl, const double *xu,
                    double alpha, const double *w, double *s) {
    for (int i = 0; i < n; ++i) {
        if (x[i] + alpha * w[i] < xl[i]) s[i] = xl[i] - x[i];
        else if (x[i] + alpha * w[i] > xu[i]) s[i] = xu[i] - x[i];
        else s[i] = alpha * w[i];
    }
}
/* B.1 dbreakpt */
static void dbreakpt(int n, const double *x, const double *xl, const double *xu, const double *w,
                     int *nbrpt, double *brptmin, double *brptmax) {
    *nbrpt = 0; *brptmin = 0.0; *brptmax = 0.0;
    for (int i = 0; i < n; ++i) {
        double brpt;
        if (x[i] < xu[i] && w[i] > 0.0) brpt = (xu[i] - x[i]) / w[i];
        else if (x[i] > xl[i] && w[i] < 0.0) brpt = (xl[i] - x[i]) / w[i];
        else continue;
        (*nbrpt)++;
        if (*nbrpt == 1) { *brptmin = brpt; *brptmax = brpt; }
        else { *brptmin = dmin(*brptmin, brpt); *brptmax = dmax(*brptmax, brpt); }
    }
    if (*nbrpt == 0) { *brptmin = 0.0; *brptmax = 0.0; }
}
/* B.1 dgpnorm */
static double dgpnorm(int n, const double *x, const double *xl, const double *xu, const double *g) {
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) {
        if (xl[i] == xu[i]) continue;
        if (x[i] == xl[i]) nrm = dmax(nrm, fabs(dmin(g[i], 0.0)));
        else if (x[i] == xu[i]) nrm = dmax(nrm, fabs(dmax(g[i], 0.0)));
        else nrm = dmax(nrm, fabs(g[i]));
    }
    return nrm;
}
/* B.1 dtrqsol */
static double dtrqsol(int n, const double *x, const double *p, double delta) {
    const double ptx = dot(n, p, x), ptp = dot(n, p, p), xtx = dot(n, x, x);
    const double dsq = delta * delta;
    const double rad = sqrt(dmax(ptx * ptx + ptp * (dsq - xtx), 0.0));
    if (ptx > 0.0) return (dsq - xtx) / (ptx + rad);
    if (rad > 0.0) return (rad - ptx) / ptp;
    return 0.0;
}
/* B.2 dcauchy */
static double dcauchy(int n, const double *x, const double *xl, const double *xu, const double *A,
                      const double *g, double delta, double alpha, double *s) {
    const double mu0 = 0.01, interpf = 0.1, extrapf = 10.0;
    double wa[NV], q, gts, brptmin, brptmax;
    int nbrpt, interp;
    for (int i = 0; i < n; ++i) wa[i] = -g[i];
    dbreakpt(n, x, xl, xu, wa, &nbrpt, &brptmin, &brptmax);
    dgpstep(n, x, xl, xu, -alpha, g, s);
    if (nrm2(n, s) > delta) interp = 1;
    else {
        symv(n, A, NV, s, wa);
        gts = dot(n, g, s);
        q = 0.5 * dot(n, s, wa) + gts;
        interp = (q >= mu0 * gts);
    }
    if (interp) {
        int search = 1;
        while (search) {
            alpha = interpf * alpha;
            dgpstep(n, x, xl, xu, -alpha, g, s);
            if (nrm2(n, s) <= delta) {
                symv(n, A, NV, s, wa);
                gts = dot(n, g, s);
                q = 0.5 * dot(n, s, wa) + gts;
                search = (q > mu0 * gts);
            }
        }
    } else {
        int search = 1;
        double alphas = alpha;
        while (search && alpha <= brptmax) {
            alpha = extrapf * alpha;
            dgpstep(n, x, xl, xu, -alpha, g, s);
            if (nrm2(n, s) <= delta) {
                symv(n, A, NV, s, wa);
                gts = dot(n, g, s);
                q = 0.5 * dot(n, s, wa) + gts;
                if (q < mu0 * gts) { search = 1; alphas = alpha; }
            } else search = 0;
        }
        alpha = alphas;
        dgpstep(n, x, xl, xu, -alpha, g, s);
    }
    return alpha;
}
/* forward (L r = b) and backward (L' r = b) solves, L lower triangular n x n */
static void lsolve(int n, const double *L, double *r) {
    for (int i = 0; i < n; ++i) {
        double s = r[i];
        for (int k = 0; k < i; ++k) s -= L[i * NV + k] * r[k];
        r[i] = s / L[i * NV + i];
    }
}
static void ltsolve(int n, const double *L, double *r) {
    for (int i = n - 1; i >= 0; --i) {
        double s = r[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * NV + i] * r[k];
        r[i] = s / L[i * NV + i];
    }
}
/* B.5 dicfs, dense: scaled Cholesky with diagonal shift */
static int dense_chol(int n, double *L) {
    for (int j = 0; j < n; ++j) {
        double djj = L[j * NV + j];
        for (int k = 0; k < j; ++k) djj -= L[j * NV + k] * L[j * NV + k];
        if (!(djj > 0.0)) return -1;
        djj = sqrt(djj);
        L[j * NV + j] = djj;
        for (int i = j + 1; i < n; ++i) {
            double s = L[i * NV + j];
            for (int k = 0; k < j; ++k) s -= L[i * NV + k] * L[j * NV + k];
            L[i * NV + j] = s / djj;
        }
    }
    return 0;
}
static void dicfs(int n, const double *B, double *L, tron_stats_t *st) {
    const double alpham = 1e-3;
    const int nbmax = 3;
    const double nbfactor = 512.0;
    double wa1[NV], wa2[NV];
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += B[k * NV + i] * B[k * NV + i];
        wa1[i] = sqrt(s);
    }
    for (int i = 0; i < n; ++i)
        wa2[i] = (B[i * NV + i] > 0.0) ? 1.0 / sqrt(B[i * NV + i]) : 1.0 / sqrt(wa1[i]);
    double alphas = alpham, alpha = 0.0;
    for (int i = 0; i < n; ++i) {
        if (B[i * NV + i] == 0.0) alpha = alphas;
        else alpha = dmax(alpha, -B[i * NV + i] * (wa2[i] * wa2[i]));
    }
    if (alpha > 0.0) alpha = dmax(alpha, alphas);
    int nb = 1, shifted = 0;
    for (;;) {
        for (int i = 0; i < n; ++i)
            for (int j = 0; j <= i; ++j) L[i * NV + j] = B[i * NV + j] * wa2[i] * wa2[j];
        for (int j = 0; j < n; ++j) L[j * NV + j] += alpha;
        if (alpha > 0.0) shifted = 1;
        if (dense_chol(n, L) == 0) {
            if (alpha == alphas && nb < nbmax) {
                alphas = alphas / nbfactor; alpha = alphas; nb++;
            } else {
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j <= i; ++j) L[i * NV + j] /= wa2[i];
                break;
            }
        } else alpha = dmax(2.0 * alpha, alphas);
    }
    if (shifted) st->shifts++;
}
/* B.4 dtrpcg */
static void dtrpcg(int n, const double *A, const double *g, double delta, const double *L,
                   double tol, double stol, int itermax, double *w, int *iters, int *info) {
    double t[NV], r[NV], p[NV], q[NV], z[NV];
    for (int i = 0; i < n; ++i) { w[i] = 0.0; t[i] = -g[i]; r[i] = t[i]; }
    lsolve(n, L, r);
    for (int i = 0; i < n; ++i) p[i] = r[i];
    double rho = dot(n, r, r);
    if (sqrt(rho) == 0.0) { *iters = 0; *info = 1; return; }
    for (int it = 1; it <= itermax; ++it) {
        for (int i = 0; i < n; ++i) z[i] = p[i];
        ltsolve(n, L, z);
        symv(n, A, NV, z, q);
        for (int i = 0; i < n; ++i) z[i] = q[i];
        lsolve(n, L, q);
        const double ptq = dot(n, p, q);
        const double alpha = (ptq > 0.0) ? rho / ptq : 0.0;
        const double sigma = dtrqsol(n, w, p, delta);
        if (ptq <= 0.0 || alpha >= sigma) {
            for (int i = 0; i < n; ++i) w[i] += sigma * p[i];
            *iters = it; *info = (ptq <= 0.0) ? 3 : 4;
            return;
        }
        for (int i = 0; i < n; ++i) { w[i] += alpha * p[i]; r[i] -= alpha * q[i]; t[i] -= alpha * z[i]; }
        const double rtr = dot(n, r, r);
        if (nrm2(n, t) <= tol) { *iters = it; *info = 1; return; }
        if (sqrt(rtr) <= stol) { *iters = it; *info = 2; return; }
        const double beta = rtr / rho;
        for (int i = 0; i < n; ++i) p[i] = r[i] + beta * p[i];
        rho = rtr;
    }
    *iters = itermax; *info = 5;
}
/* B.6 dprsrch: on exit x is the new point and w the step actually taken */
static void dprsrch(int n, double *x, const double *xl, const double *xu, const double *A,
                    const double *g, double *w) {
    const double mu0 = 0.01, interpf = 0.5;
    double wa1[NV], wa2[NV], brptmin, brptmax, alpha = 1.0;
    int nbrpt, search = 1;
    dbreakpt(n, x, xl, xu, w, &nbrpt, &brptmin, &brptmax);
    while (search && alpha > brptmin) {
        dgpstep(n, x, xl, xu, alpha, w, wa1);
        symv(n, A, NV, wa1, wa2);
        const double gts = dot(n, g, wa1);
        const double q = 0.5 * dot(n, wa1, wa2) + gts;
        if (q <= mu0 * gts) search = 0;
        else alpha = interpf * alpha;
    }
    if (alpha < 1.0 && alpha < brptmin) alpha = brptmin;
    dgpstep(n, x, xl, xu, alpha, w, wa1);
    for (int i = 0; i < n; ++i) x[i] += alpha * w[i];
    dmid(n, x, xl, xu);
    for (int i = 0; i < n; ++i) w[i] = wa1[i];
}
/* B.3 dspcg */
static void dspcg(int n, double *x, const double *xl, const double *xu, const double *A,
                  const double *g, double delta, double rtol, double *s, int itermax,
                  tron_stats_t *st) {
    double w[NV], B[NV * NV], L[NV * NV], gfree[NV], wa[NV], xf[NV], xlf[NV], xuf[NV], wf[NV];
    int indfree[NV];
    symv(n, A, NV, s, w);
    for (int i = 0; i < n; ++i) x[i] += s[i];
    dmid(n, x, xl, xu);
    int iters = 0;
    for (int nfaces = 1; nfaces <= n; ++nfaces) {
        int nfree = 0;
        for (int j = 0; j < n; ++j)
            if (xl[j] < x[j] && x[j] < xu[j]) indfree[nfree++] = j;
        if (nfree == 0) return;
        for (int a = 0; a < nfree; ++a)
            for (int b = 0; b < nfree; ++b) B[a * NV + b] = A[indfree[a] * NV + indfree[b]];
        dicfs(nfree, B, L, st);
        for (int j = 0; j < nfree; ++j) { gfree[j] = w[indfree[j]] + g[indfree[j]]; wa[j] = g[indfree[j]]; }
        const double gfnorm = nrm2(nfree, wa);
        int itertr, infotr;
        dtrpcg(nfree, B, gfree, delta, L, rtol * gfnorm, 0.0, itermax, wf, &itertr, &infotr);
        iters += itertr;
        st->cg += itertr;
        ltsolve(nfree, L, wf);
        for (int j = 0; j < nfree; ++j) { xf[j] = x[indfree[j]]; xlf[j] = xl[indfree[j]]; xuf[j] = xu[indfree[j]]; }
        dprsrch(nfree, xf, xlf, xuf, B, gfree, wf);
        for (int j = 0; j < nfree; ++j) { x[indfree[j]] = xf[j]; s[indfree[j]] += wf[j]; }
        symv(n, A, NV, s, w);
        for (int j = 0; j < nfree; ++j) gfree[j] = w[indfree[j]] + g[indfree[j]];
        const double gfnormf = nrm2(nfree, gfree);
        if (gfnormf <= rtol * gfnorm) return;
        if (infotr == 3 || infotr == 4) return;
        if (iters > itermax) return;
    }
}

/* B.0 driver: acopf_tron_linelimit_kernel.jl:44-148 around ExaTron.dtron,
 * == ExaTron.solveProblem on the CPU path (acopf_auglag_linelimit_kernel_cpu.jl:104-117). */
int orc__tron_cb_g(int n, double *x, const double *xl, const double *xu, orc__f_fn evalf, orc__gh_fn evalgh,
                   const void *ctx, int max_feval, int max_minor, double gtol, int *minor_out, tron_stats_t *st,
                   double *g_out) {
    const double eta0 = 1e-4, eta1 = 0.25, eta2 = 0.75, sigma1 = 0.25, sigma2 = 0.5, sigma3 = 4.0;
    const double frtol = 1e-12, fatol = 0.0, fmin = -1e32, cgtol = 0.1;
    const int cg_itermax = n;
    double g[NV], A[NV * NV], s[NV], xc[NV], wa[NV];
    double f = evalf(x, ctx);
    evalgh(x, ctx, g, A);
    int nfev = 1, minor = 1, iter = 1, status = 0;
    st->nfev++; st->ngev++;
    double delta = nrm2(n, g), alphac = 1.0;
    for (;;) {
        int task; /* 1 F, 2 GH, 4 CONV, 10 WARN */
        do {
            const double fc = f;
            memcpy(xc, x, sizeof(double) * (size_t)n);
            alphac = dcauchy(n, x, xl, xu, A, g, delta, alphac, s);
            dspcg(n, x, xl, xu, A, g, delta, cgtol, s, cg_itermax, st);
            symv(n, A, NV, s, wa);
            const double prered = -(dot(n, s, g) + 0.5 * dot(n, s, wa));
            f = evalf(x, ctx);
            nfev++; st->nfev++;
            if (nfev >= max_feval) { *minor_out = minor; if (g_out) memcpy(g_out, g, sizeof(double) * (size_t)n); return status; }
            const double actred = fc - f;
            const double snorm = nrm2(n, s);
            if (iter == 1) delta = dmin(delta, snorm);
            const double g0 = dot(n, g, s);
            double alpha;
            if (f - fc - g0 <= 0.0) alpha = sigma3;
            else alpha = dmax(sigma1, -0.5 * (g0 / (f - fc - g0)));
            if (actred < eta0 * prered) delta = dmin(dmax(alpha, sigma1) * snorm, sigma2 * delta);
            else if (actred < eta1 * prered) delta = dmax(sigma1 * delta, dmin(alpha * snorm, sigma2 * delta));
            else if (actred < eta2 * prered) delta = dmax(sigma1 * delta, dmin(alpha * snorm, sigma3 * delta));
            else delta = dmax(delta, dmin(alpha * snorm, sigma3 * delta));
            if (actred > eta0 * prered) { task = 2; iter++; }
            else { task = 1; memcpy(x, xc, sizeof(double) * (size_t)n); f = fc; st->rejected++; }
            if (f < fmin) task = 10;
            if (fabs(actred) <= fatol && prered <= fatol) task = 4;
            if (fabs(actred) <= frtol * fabs(f) && prered <= frtol * fabs(f)) task = 4;
        } while (task == 1);
        if (task == 4 || task == 10) break;
        evalgh(x, ctx, g, A);
        minor++; st->ngev++;
        if (dgpnorm(n, x, xl, xu, g) <= gtol) break;
        if (minor >= max_minor) { status = 1; break; }
    }
    *minor_out = minor;
    if (g_out) memcpy(g_out, g, sizeof(double) * (size_t)n);   /* `tron.g`: the last gradient evaluated */
    return status;
}
int orc__tron_cb(int n, double *x, const double *xl, const double *xu, orc__f_fn evalf, orc__gh_fn evalgh,
                 const void *ctx, int max_feval, int max_minor, double gtol, int *minor_out, tron_stats_t *st) {
    return orc__tron_cb_g(n, x, xl, xu, evalf, evalgh, ctx, max_feval, max_minor, gtol, minor_out, st, NULL);
}

/* the branch instance of the driver: objective of acopf_eval_linelimit_kernel_cpu.jl */
typedef struct { const double *param, *Y; double scale; } branch_ctx_t;
static double branch_f(const double *x, const void *c) {
    const branch_ctx_t *b = (const branch_ctx_t *)c;
    return orc_eval_f(x, b->param, b->Y, b->scale);
}
static void branch_gh(const double *x, const void *c, double *g, double *A) {
    const branch_ctx_t *b = (const branch_ctx_t *)c;
    orc_eval_gh(x, b->param, b->Y, b->scale, g, A);
}
static int tron_solve(double x[NV], const double xl[NV], const double xu[NV], const double *param,
                      const double Y[8], double scale, int max_feval, int max_minor, double gtol,
                      int *minor_out, tron_stats_t *st) {
    const branch_ctx_t ctx = { param, Y, scale };
    return orc__tron_cb(NV, x, xl, xu, branch_f, branch_gh, &ctx, max_feval, max_minor, gtol, minor_out, st);
}

int orc_tron_solve(double x[6], const double xl[6], const double xu[6], const double *param,
                   const double Y[8], double scale, int max_feval, int max_minor, double gtol,
                   int *minor_iter, int *nfev) {
    tron_stats_t st; memset(&st, 0, sizeof(st));
    int minor = 0;
    int status = tron_solve(x, xl, xu, param, Y, scale, max_feval, max_minor, gtol, &minor, &st);
    if (minor_iter) *minor_iter = minor;
    if (nfev) *nfev = (int)st.nfev;
    return status;
}

/* ------------------------------------------------------------------------ */
/* model                                                                    */
/* ------------------------------------------------------------------------ */
int orc_create(const ea_grid_t *G, orc_model_t **out) {
    if (!G || !out) return EA_ERR_ARG;
    orc_model_t *m = (orc_model_t *)calloc(1, sizeof(*m));
    if (!m) return EA_ERR_ALLOC;
    m->ngen = G->ngen; m->nline = G->nline; m->nbus = G->nbus; m->baseMVA = G->baseMVA;
    m->nvar = 2 * G->ngen + 8 * G->nline;                      /* acopf_model.jl:55 */
    m->pgmin = dup_d(G->pgmin, G->ngen); m->pgmax = dup_d(G->pgmax, G->ngen);
    m->qgmin = dup_d(G->qgmin, G->ngen); m->qgmax = dup_d(G->qgmax, G->ngen);
    m->c2 = dup_d(G->c2, G->ngen); m->c1 = dup_d(G->c1, G->ngen); m->c0 = dup_d(G->c0, G->ngen);
    m->pgmin_curr = dup_d(G->pgmin, G->ngen);                  /* acopf_model.jl:61-64 */
    m->pgmax_curr = dup_d(G->pgmax, G->ngen);
    m->YshR = dup_d(G->YshR, G->nbus); m->YshI = dup_d(G->YshI, G->nbus);
    const double *ys[8] = { G->YffR, G->YffI, G->YftR, G->YftI, G->YttR, G->YttI, G->YtfR, G->YtfI };
    for (int k = 0; k < 8; ++k) m->Y[k] = dup_d(ys[k], G->nline);
    m->FrVmBound = dup_d(G->FrVmBound, 2 * G->nline); m->ToVmBound = dup_d(G->ToVmBound, 2 * G->nline);
    m->FrVaBound = dup_d(G->FrVaBound, 2 * G->nline); m->ToVaBound = dup_d(G->ToVaBound, 2 * G->nline);
    m->rateA = dup_d(G->rateA, G->nline);
    m->FrStart = dup_i_0based(G->FrStart, G->nbus + 1); m->FrIdx = dup_i_0based(G->FrIdx, G->nline);
    m->ToStart = dup_i_0based(G->ToStart, G->nbus + 1); m->ToIdx = dup_i_0based(G->ToIdx, G->nline);
    m->GenStart = dup_i_0based(G->GenStart, G->nbus + 1); m->GenIdx = dup_i_0based(G->GenIdx, G->ngen);
    m->brBusIdx = dup_i_0based(G->brBusIdx, 2 * G->nline);
    m->Pd = dup_d(G->Pd, G->nbus); m->Qd = dup_d(G->Qd, G->nbus);
    m->Vmin = dup_d(G->Vmin, G->nbus); m->Vmax = dup_d(G->Vmax, G->nbus);
    for (int f = 0; f < EA_NUM_FIELDS; ++f) m->vec[f] = (double *)calloc((size_t)(m->nvar > 0 ? m->nvar : 1), sizeof(double));
    m->membuf = (double *)calloc((size_t)(MEMROWS * (m->nline > 0 ? m->nline : 1)), sizeof(double));
    for (int64_t l = 0; l < m->nline; ++l) m->membuf[l * MEMROWS + 28] = m->rateA[l];  /* acopf_model.jl:89 */
    m->nthreads = 1;
    *out = m;
    return EA_OK;
}

void orc_destroy(orc_model_t *m) {
    if (!m) return;
    free(m->pgmin); free(m->pgmax); free(m->qgmin); free(m->qgmax); free(m->c2); free(m->c1); free(m->c0);
    free(m->pgmin_curr); free(m->pgmax_curr); free(m->YshR); free(m->YshI);
    for (int k = 0; k < 8; ++k) free(m->Y[k]);
    free(m->FrVmBound); free(m->ToVmBound); free(m->FrVaBound); free(m->ToVaBound); free(m->rateA);
    free(m->FrStart); free(m->FrIdx); free(m->ToStart); free(m->ToIdx); free(m->GenStart); free(m->GenIdx);
    free(m->brBusIdx); free(m->Pd); free(m->Qd); free(m->Vmin); free(m->Vmax);
    for (int f = 0; f < EA_NUM_FIELDS; ++f) free(m->vec[f]);
    free(m->membuf);
    free(m);
}

void orc_set_threads(orc_model_t *m, int n) { m->nthreads = n > 0 ? n : 1; }
int  orc_get_threads(const orc_model_t *m) {
#ifdef _OPENMP
    return m->nthreads;
#else
    (void)m; return 1;
#endif
}
int64_t orc_nvar(const orc_model_t *m) { return m->nvar; }
double *orc_vector(orc_model_t *m, int field) { return (field >= 0 && field < EA_NUM_FIELDS) ? m->vec[field] : NULL; }
double *orc_membuf(orc_model_t *m) { return m->membuf; }
void orc_set_load(orc_model_t *m, const double *Pd, const double *Qd) {
    memcpy(m->Pd, Pd, sizeof(double) * (size_t)m->nbus); memcpy(m->Qd, Qd, sizeof(double) * (size_t)m->nbus);
}
void orc_set_pg_bounds(orc_model_t *m, const double *lo, const double *hi) {
    memcpy(m->pgmin_curr, lo, sizeof(double) * (size_t)m->ngen); memcpy(m->pgmax_curr, hi, sizeof(double) * (size_t)m->ngen);
}
void orc_get_counters(const orc_model_t *m, ea_counters_t *out) { *out = m->cnt; }
void orc_reset_counters(orc_model_t *m) { memset(&m->cnt, 0, sizeof(m->cnt)); }
void orc_set_eval_trace(orc_model_t *m, int32_t *buf) { m->eval_trace = buf; }

/* acopf_init_solution_cpu.jl:1-45 */
void orc_init_solution(orc_model_t *m, double rho_pq, double rho_va) {
    for (int f = 0; f < EA_NUM_FIELDS; ++f) memset(m->vec[f], 0, sizeof(double) * (size_t)m->nvar);
    double *v = m->vec[EA_V_CURR], *rho = m->vec[EA_RHO];
    for (int64_t i = 0; i < m->nvar; ++i) rho[i] = rho_pq;
    for (int64_t g = 0; g < m->ngen; ++g) {
        v[2 * g] = 0.5 * (m->pgmin[g] + m->pgmax[g]);
        v[2 * g + 1] = 0.5 * (m->qgmin[g] + m->qgmax[g]);
    }
    const int64_t ls = 2 * m->ngen;
    for (int64_t l = 0; l < m->nline; ++l) {
        const int64_t fb = m->brBusIdx[2 * l], tb = m->brBusIdx[2 * l + 1];
        const double wij0 = (m->Vmax[fb] * m->Vmax[fb] + m->Vmin[fb] * m->Vmin[fb]) / 2;
        const double wji0 = (m->Vmax[tb] * m->Vmax[tb] + m->Vmin[tb] * m->Vmin[tb]) / 2;
        const double wR0 = sqrt(wij0 * wji0);
        double *p = v + ls + 8 * l;
        p[0] = m->Y[0][l] * wij0 + m->Y[2][l] * wR0;
        p[1] = -m->Y[1][l] * wij0 - m->Y[3][l] * wR0;
        p[2] = m->Y[4][l] * wji0 + m->Y[6][l] * wR0;
        p[3] = -m->Y[5][l] * wji0 - m->Y[7][l] * wR0;
        p[4] = wij0; p[5] = wji0; p[6] = 0.0; p[7] = 0.0;
        for (int k = 4; k < 8; ++k) rho[ls + 8 * l + k] = rho_va;
    }
    /* membuf is NOT touched here: it is zeroed once by the model constructor
     * (acopf_model.jl:87-89 -> orc_create) and persists across init_solution! calls. */
}

/* acopf_admm_prepoststep_cpu.jl:1-10 */
double orc_outer_prestep(orc_model_t *m) { return vnorm(m->vec[EA_Z_CURR], m->nvar); }
/* acopf_admm_prepoststep_cpu.jl:15-23 */
void orc_inner_prestep(orc_model_t *m) {
    memcpy(m->vec[EA_Z_PREV], m->vec[EA_Z_CURR], sizeof(double) * (size_t)m->nvar);
}

/* acopf_generator_kernel_cpu.jl:1-20 */
void orc_update_x_gen(orc_model_t *m) {
    double *u = m->vec[EA_U_CURR];
    const double *x = m->vec[EA_V_CURR], *z = m->vec[EA_Z_CURR], *l = m->vec[EA_L_CURR], *rho = m->vec[EA_RHO];
    const double B = m->baseMVA;
    for (int64_t I = 0; I < m->ngen; ++I) {
        const int64_t pg = 2 * I, qg = 2 * I + 1;
        u[pg] = dmax(m->pgmin_curr[I], dmin(m->pgmax_curr[I],
                 (-(m->c1[I] * B + l[pg] + rho[pg] * (-x[pg] + z[pg]))) / (2 * m->c2[I] * (B * B) + rho[pg])));
        u[qg] = dmax(m->qgmin[I], dmin(m->qgmax[I], (-(l[qg] + rho[qg] * (-x[qg] + z[qg]))) / rho[qg]));
    }
}

/* acopf_auglag_linelimit_kernel_cpu.jl:1-172 — one branch */
static void solve_branch(orc_model_t *m, int64_t I, int64_t major_iter, int32_t max_auglag,
                         double mu_max, double scale, ea_counters_t *cnt) {
    const int64_t pij = 2 * m->ngen + 8 * I;
    double *u = m->vec[EA_U_CURR];
    const double *xbar = m->vec[EA_V_CURR], *z = m->vec[EA_Z_CURR], *l = m->vec[EA_L_CURR], *rho = m->vec[EA_RHO];
    double *param = m->membuf + I * MEMROWS;
    double Y[8], x[NV], xl[NV], xu[NV];
    for (int k = 0; k < 8; ++k) Y[k] = m->Y[k][I];
    xl[0] = m->FrVmBound[2 * I]; xu[0] = m->FrVmBound[2 * I + 1];
    xl[1] = m->ToVmBound[2 * I]; xu[1] = m->ToVmBound[2 * I + 1];
    xl[2] = m->FrVaBound[2 * I]; xu[2] = m->FrVaBound[2 * I + 1];
    xl[3] = m->ToVaBound[2 * I]; xu[3] = m->ToVaBound[2 * I + 1];
    xl[4] = -param[28]; xu[4] = 0.0;
    xl[5] = -param[28]; xu[5] = 0.0;
    x[0] = dmin(xu[0], dmax(xl[0], sqrt(u[pij + 4])));
    x[1] = dmin(xu[1], dmax(xl[1], sqrt(u[pij + 5])));
    x[2] = dmin(xu[2], dmax(xl[2], u[pij + 6]));
    x[3] = dmin(xu[3], dmax(xl[3], u[pij + 7]));
    x[4] = dmin(xu[4], dmax(xl[4], -(u[pij] * u[pij] + u[pij + 1] * u[pij + 1])));
    x[5] = dmin(xu[5], dmax(xl[5], -(u[pij + 2] * u[pij + 2] + u[pij + 3] * u[pij + 3])));
    for (int k = 0; k < 8; ++k) {
        param[k] = l[pij + k];
        param[8 + k] = rho[pij + k];
        param[16 + k] = xbar[pij + k] - z[pij + k];
    }
    double mu;
    if (major_iter == 1) { param[26] = 10.0; mu = 10.0; } else mu = param[26];
    double eta = 1.0 / pow(mu, 0.1);
    int it = 0, terminate = 0;
    tron_stats_t st; memset(&st, 0, sizeof(st));
    while (!terminate) {
        it++;
        int minor;
        tron_solve(x, xl, xu, param, Y, scale, 500, 200, 1e-6, &minor, &st);
        double sn, cs;
        orc_sincos(x[2] - x[3], &sn, &cs);
        const double cc = x[0] * x[1] * cs;
        const double ss = x[0] * x[1] * sn;
        const double fpij = Y[0] * (x[0] * x[0]) + Y[2] * cc + Y[3] * ss;
        const double fqij = -Y[1] * (x[0] * x[0]) - Y[3] * cc + Y[2] * ss;
        const double fpji = Y[4] * (x[1] * x[1]) + Y[6] * cc - Y[7] * ss;
        const double fqji = -Y[5] * (x[1] * x[1]) - Y[7] * cc - Y[6] * ss;
        const double cviol1 = fpij * fpij + fqij * fqij + x[4];
        const double cviol2 = fpji * fpji + fqji * fqji + x[5];
        const double cnorm = dmax(fabs(cviol1), fabs(cviol2));
        if (cnorm <= eta) {
            if (cnorm <= 1e-6) terminate = 1;
            else {
                param[24] += mu * cviol1;
                param[25] += mu * cviol2;
                eta = eta / pow(mu, 0.9);
            }
        } else {
            mu = dmin(mu_max, mu * 10);
            eta = 1.0 / pow(mu, 0.1);
            param[26] = mu;
        }
        if (it >= max_auglag) { if (!terminate) cnt->max_auglag_hits++; terminate = 1; }
    }
    double sn, cs;
    orc_sincos(x[2] - x[3], &sn, &cs);
    const double cc = x[0] * x[1] * cs;
    const double ss = x[0] * x[1] * sn;
    u[pij] = Y[0] * (x[0] * x[0]) + Y[2] * cc + Y[3] * ss;
    u[pij + 1] = -Y[1] * (x[0] * x[0]) - Y[3] * cc + Y[2] * ss;
    u[pij + 2] = Y[4] * (x[1] * x[1]) + Y[6] * cc - Y[7] * ss;
    u[pij + 3] = -Y[5] * (x[1] * x[1]) - Y[7] * cc - Y[6] * ss;
    u[pij + 4] = x[0] * x[0];
    u[pij + 5] = x[1] * x[1];
    u[pij + 6] = x[2];
    u[pij + 7] = x[3];
    param[26] = mu;
    cnt->line_calls++;
    cnt->auglag_iters += it;
    cnt->tron_evals += st.nfev;   /* 1 at the start point + 1 per trial point */
    cnt->cg_iters += st.cg;
    cnt->chol_shifts += st.shifts;
    cnt->rejected_steps += st.rejected;
    if (st.nfev > cnt->max_evals_lane) cnt->max_evals_lane = st.nfev;
    if (m->eval_trace) m->eval_trace[I] = (int32_t)st.nfev;
}

static void merge_counters(ea_counters_t *dst, const ea_counters_t *src) {
    dst->line_calls += src->line_calls; dst->auglag_iters += src->auglag_iters;
    dst->tron_evals += src->tron_evals; dst->cg_iters += src->cg_iters;
    dst->chol_shifts += src->chol_shifts; dst->rejected_steps += src->rejected_steps;
    dst->max_auglag_hits += src->max_auglag_hits;
    if (src->max_evals_lane > dst->max_evals_lane) dst->max_evals_lane = src->max_evals_lane;
}

void orc_update_x_line(orc_model_t *m, int64_t inner, int32_t max_auglag, double mu_max, double scale) {
#ifdef _OPENMP
    if (m->nthreads > 1) {
        #pragma omp parallel num_threads(m->nthreads)
        {
            ea_counters_t local; memset(&local, 0, sizeof(local));
            #pragma omp for schedule(dynamic, 64)
            for (int64_t I = 0; I < m->nline; ++I) solve_branch(m, I, inner, max_auglag, mu_max, scale, &local);
            #pragma omp critical
            merge_counters(&m->cnt, &local);
        }
        return;
    }
#endif
    ea_counters_t local; memset(&local, 0, sizeof(local));
    for (int64_t I = 0; I < m->nline; ++I) solve_branch(m, I, inner, max_auglag, mu_max, scale, &local);
    merge_counters(&m->cnt, &local);
}

/* acopf_bus_kernel_cpu.jl:1-116 — one bus */
static void solve_bus(orc_model_t *m, int64_t I) {
    const double *u = m->vec[EA_U_CURR], *z = m->vec[EA_Z_CURR], *l = m->vec[EA_L_CURR], *rho = m->vec[EA_RHO];
    double *v = m->vec[EA_V_CURR];
    const int64_t ls = 2 * m->ngen;
    double common_wi = 0, common_ti = 0, inv_p = 0, inv_q = 0, rs_w = 0, rs_t = 0;
    for (int64_t k = m->FrStart[I]; k < m->FrStart[I + 1]; ++k) {
        const int64_t p = ls + 8 * m->FrIdx[k];
        common_wi += l[p + 4] + rho[p + 4] * (u[p + 4] + z[p + 4]);
        common_ti += l[p + 6] + rho[p + 6] * (u[p + 6] + z[p + 6]);
        inv_p += 1.0 / rho[p]; inv_q += 1.0 / rho[p + 1];
        rs_w += rho[p + 4]; rs_t += rho[p + 6];
    }
    for (int64_t k = m->ToStart[I]; k < m->ToStart[I + 1]; ++k) {
        const int64_t p = ls + 8 * m->ToIdx[k];
        common_wi += l[p + 5] + rho[p + 5] * (u[p + 5] + z[p + 5]);
        common_ti += l[p + 7] + rho[p + 7] * (u[p + 7] + z[p + 7]);
        inv_p += 1.0 / rho[p + 2]; inv_q += 1.0 / rho[p + 3];
        rs_w += rho[p + 5]; rs_t += rho[p + 7];
    }
    common_wi /= rs_w;
    double rhs1 = 0, rhs2 = 0, inv_pg = 0, inv_qg = 0;
    for (int64_t k = m->GenStart[I]; k < m->GenStart[I + 1]; ++k) {
        const int64_t p = 2 * m->GenIdx[k];
        rhs1 += (u[p] + z[p]) + (l[p] / rho[p]);
        rhs2 += (u[p + 1] + z[p + 1]) + (l[p + 1] / rho[p + 1]);
        inv_pg += 1.0 / rho[p]; inv_qg += 1.0 / rho[p + 1];
    }
    rhs1 -= (m->Pd[I] / m->baseMVA);
    rhs2 -= (m->Qd[I] / m->baseMVA);
    for (int64_t k = m->FrStart[I]; k < m->FrStart[I + 1]; ++k) {
        const int64_t p = ls + 8 * m->FrIdx[k];
        rhs1 -= (u[p] + z[p]) + (l[p] / rho[p]);
        rhs2 -= (u[p + 1] + z[p + 1]) + (l[p + 1] / rho[p + 1]);
    }
    for (int64_t k = m->ToStart[I]; k < m->ToStart[I + 1]; ++k) {
        const int64_t p = ls + 8 * m->ToIdx[k];
        rhs1 -= (u[p + 2] + z[p + 2]) + (l[p + 2] / rho[p + 2]);
        rhs2 -= (u[p + 3] + z[p + 3]) + (l[p + 3] / rho[p + 3]);
    }
    const double gr = m->YshR[I], gi = m->YshI[I];
    rhs1 -= gr * common_wi;
    rhs2 += gi * common_wi;
    const double A11 = (inv_pg + inv_p) + (gr * gr / rs_w);
    const double A12 = -gr * (gi / rs_w);
    const double A21 = A12;
    const double A22 = (inv_qg + inv_q) + (gi * gi / rs_w);
    const double mu2 = (rhs2 - (A21 / A11) * rhs1) / (A22 - (A21 / A11) * A12);
    const double mu1 = (rhs1 - A12 * mu2) / A11;
    const double wi = common_wi + ((gr * mu1 - gi * mu2) / rs_w);
    const double ti = common_ti / rs_t;
    for (int64_t k = m->GenStart[I]; k < m->GenStart[I + 1]; ++k) {
        const int64_t p = 2 * m->GenIdx[k];
        v[p] = (u[p] + z[p]) + (l[p] - mu1) / rho[p];
        v[p + 1] = (u[p + 1] + z[p + 1]) + (l[p + 1] - mu2) / rho[p + 1];
    }
    for (int64_t k = m->FrStart[I]; k < m->FrStart[I + 1]; ++k) {
        const int64_t p = ls + 8 * m->FrIdx[k];
        v[p] = (u[p] + z[p]) + (l[p] + mu1) / rho[p];
        v[p + 1] = (u[p + 1] + z[p + 1]) + (l[p + 1] + mu2) / rho[p + 1];
        v[p + 4] = wi; v[p + 6] = ti;
    }
    for (int64_t k = m->ToStart[I]; k < m->ToStart[I + 1]; ++k) {
        const int64_t p = ls + 8 * m->ToIdx[k];
        v[p + 2] = (u[p + 2] + z[p + 2]) + (l[p + 2] + mu1) / rho[p + 2];
        v[p + 3] = (u[p + 3] + z[p + 3]) + (l[p + 3] + mu2) / rho[p + 3];
        v[p + 5] = wi; v[p + 7] = ti;
    }
}

void orc_update_xbar(orc_model_t *m) {
#ifdef _OPENMP
    if (m->nthreads > 1) {
        #pragma omp parallel for num_threads(m->nthreads) schedule(static)
        for (int64_t I = 0; I < m->nbus; ++I) solve_bus(m, I);
        return;
    }
#endif
    for (int64_t I = 0; I < m->nbus; ++I) solve_bus(m, I);
}

/* acopf_admm_update_z_cpu.jl:10 */
void orc_update_z(orc_model_t *m, double beta) {
    double *z = m->vec[EA_Z_CURR];
    const double *lz = m->vec[EA_LZ], *l = m->vec[EA_L_CURR], *rho = m->vec[EA_RHO], *u = m->vec[EA_U_CURR], *v = m->vec[EA_V_CURR];
    for (int64_t i = 0; i < m->nvar; ++i) z[i] = (-(lz[i] + l[i] + rho[i] * (u[i] - v[i]))) / (beta + rho[i]);
}
/* acopf_admm_update_l_cpu.jl:10 */
void orc_update_l(orc_model_t *m, double beta) {
    double *l = m->vec[EA_L_CURR];
    const double *lz = m->vec[EA_LZ], *z = m->vec[EA_Z_CURR];
    for (int64_t i = 0; i < m->nvar; ++i) l[i] = -(lz[i] + beta * z[i]);
}
/* acopf_admm_update_lz_cpu.jl:10 */
void orc_update_lz(orc_model_t *m, double beta, double M) {
    double *lz = m->vec[EA_LZ];
    const double *z = m->vec[EA_Z_CURR];
    for (int64_t i = 0; i < m->nvar; ++i) lz[i] = dmax(-M, dmin(M, lz[i] + (beta * z[i])));
}
/* acopf_admm_update_residual_cpu.jl:20-27 (norms; the in-loop objval/auglag of
 * :28-35 is not part of the GPU path and is skipped) */
void orc_update_residual(orc_model_t *m, double out[4]) {
    double *rp = m->vec[EA_RP], *rd = m->vec[EA_RD], *ab = m->vec[EA_AX_PLUS_BY];
    const double *u = m->vec[EA_U_CURR], *v = m->vec[EA_V_CURR], *z = m->vec[EA_Z_CURR], *zp = m->vec[EA_Z_PREV];
    for (int64_t i = 0; i < m->nvar; ++i) {
        rp[i] = u[i] - v[i] + z[i];
        rd[i] = z[i] - zp[i];
        ab[i] = rp[i] - z[i];
    }
    out[0] = vnorm(rp, m->nvar); out[1] = vnorm(rd, m->nvar);
    out[2] = vnorm(z, m->nvar);  out[3] = vnorm(ab, m->nvar);
}
/* acopf_admm_prepoststep_cpu.jl:42-45 */
double orc_poststep(orc_model_t *m) {
    const double *u = m->vec[EA_U_CURR];
    double obj = 0.0;
    for (int64_t g = 0; g < m->ngen; ++g) {
        const double pg = m->baseMVA * u[2 * g];
        obj += m->c2[g] * (pg * pg) + m->c1[g] * pg + m->c0[g];
    }
    return obj;
}

/* src/algorithms/admm_two_level.jl:1-88 */
int orc_admm_two_level(orc_model_t *m, const ea_params_t *par, ea_info_t *info) {
    const double sqrt_d = sqrt((double)m->nvar);
    const double OUTER_TOL = sqrt_d * par->outer_eps;
    memset(info, 0, sizeof(*info));
    info->mismatch = INFINITY; info->norm_z_prev = INFINITY; info->norm_z_curr = INFINITY;
    double beta = par->initial_beta;
    double res[4];
    if (par->verbose > 0) {
        orc_update_residual(m, res);
        info->primres = res[0]; info->dualres = res[1]; info->norm_z_curr = res[2]; info->mismatch = res[3];
    }
    info->status = EA_STATUS_ITERATION_LIMIT;
    const double t0 = wall();
    while (info->outer < par->outer_iterlim) {
        info->outer++;
        info->norm_z_prev = orc_outer_prestep(m);
        info->inner = 0;
        while (info->inner < par->inner_iterlim) {
            info->inner++; info->cumul++;
            orc_inner_prestep(m);
            double t = wall();
            orc_update_x_gen(m);
            double t1 = wall(); info->time_generators += t1 - t;
            orc_update_x_line(m, info->inner, par->max_auglag, par->mu_max, par->scale);
            double t2 = wall(); info->time_branches += t2 - t1; info->time_x_update += t2 - t;
            orc_update_xbar(m);
            double t3 = wall(); info->time_buses += t3 - t2; info->time_xbar_update += t3 - t2;
            orc_update_z(m, beta);
            double t4 = wall(); info->time_z_update += t4 - t3;
            orc_update_l(m, beta);
            double t5 = wall(); info->time_l_update += t5 - t4;
            orc_update_residual(m, res);
            info->primres = res[0]; info->dualres = res[1]; info->norm_z_curr = res[2]; info->mismatch = res[3];
            info->eps_pri = sqrt_d / (2500.0 * (double)info->outer);
            if (par->verbose > 1)
                printf("%8ld  %8ld  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e\n",
                       (long)info->outer, (long)info->inner, info->primres, info->eps_pri, info->dualres,
                       info->norm_z_curr, info->mismatch, OUTER_TOL, beta);
            if (info->primres <= info->eps_pri) break;
        }
        if (info->mismatch <= OUTER_TOL) { info->status = EA_STATUS_SOLVED; break; }
        double t = wall();
        orc_update_lz(m, beta, par->MAX_MULTIPLIER);
        info->time_lz_update += wall() - t;
        if (info->norm_z_curr > par->theta * info->norm_z_prev) beta = dmin(par->inc_c * beta, 1e24);
    }
    info->time_overall = wall() - t0;
    info->beta = beta;
    info->objval = orc_poststep(m);
    return EA_OK;
}
