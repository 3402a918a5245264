/*
 * qpsub_oracle.c — CPU ORACLE (test infrastructure, NOT product code) for `ModelQpsub` + admm_one_level.
 * See qpsub_oracle.h for the pinning status.
 *
 * Restates the reference's `use_gpu=false` path of src/models/qpsub/: generators and buses reuse the ACOPF
 * kernels of acopf_oracle.c on shifted bounds / costs / loads; a branch solves, by an augmented-Lagrangian loop
 * around TRON, the box-constrained QP in x = (t_ij, t_ji, w_i, w_j, theta_i, theta_j) obtained from the SQP
 * Hessian block after eliminating (w_ijR, w_ijI) through the linearised 1h / 1i. Indices are 0-based.
 */
#include "qpsub_oracle.h"
#include "oracle_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

struct orc_qp {
    orc_model_t *m;           /* Solution vectors, grid, generator / bus kernels (with the qpsub_* data installed) */
    int64_t nline, ngen, nbus, nvar;
    double *Hs;               /* nline x 36 row-major */
    double *LH_1h, *RH_1h, *LH_1i, *RH_1i, *LH_1j, *RH_1j, *LH_1k, *RH_1k;
    double *ls, *us, *line_res;
    double *c1, *c2;          /* qpsub_c1, qpsub_c2 */
    double *sqp_line;         /* 6 x nline column-major */
    double *membuf;           /* 5 x nline column-major (qpsub_membuf) */
    double *lambda;           /* 4 x nline column-major */
    int nthreads;
    int64_t cnt[4];
};

static double qp_wall(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static double *qdup(const double *src, size_t n) {
    double *p = (double *)calloc(n ? n : 1, sizeof(double));
    if (p && src) memcpy(p, src, n * sizeof(double));
    return p;
}
static double qnorm(const double *x, int64_t n) {
    double s = 0.0; for (int64_t i = 0; i < n; ++i) s += x[i] * x[i]; return sqrt(s);
}

/* qpsub_model.jl:63-140 + the field copies of solve_qpsub.jl:83-104 */
int orc_qp_create(const ea_grid_t *G, const ea_qpsub_data_t *D, orc_qp_t **out) {
    if (!G || !D || !out || !D->Hs || !D->LH_1h || !D->RH_1h || !D->LH_1i || !D->RH_1i || !D->LH_1j || !D->RH_1j ||
        !D->LH_1k || !D->RH_1k || !D->ls || !D->us || !D->pgmax || !D->pgmin || !D->qgmax || !D->qgmin || !D->c1 ||
        !D->c2 || !D->Pd || !D->Qd) return EA_ERR_ARG;
    orc_qp_t *q = (orc_qp_t *)calloc(1, sizeof(*q));
    if (!q) return EA_ERR_ALLOC;
    const int rc = orc_create(G, &q->m);
    if (rc != EA_OK) { free(q); return rc; }
    const size_t nl = (size_t)G->nline, ng = (size_t)G->ngen;
    q->nline = G->nline; q->ngen = G->ngen; q->nbus = G->nbus; q->nvar = orc_nvar(q->m); q->nthreads = 1;
    q->Hs = qdup(D->Hs, 36 * nl);
    q->LH_1h = qdup(D->LH_1h, 4 * nl); q->RH_1h = qdup(D->RH_1h, nl);
    q->LH_1i = qdup(D->LH_1i, 4 * nl); q->RH_1i = qdup(D->RH_1i, nl);
    q->LH_1j = qdup(D->LH_1j, 2 * nl); q->RH_1j = qdup(D->RH_1j, nl);
    q->LH_1k = qdup(D->LH_1k, 2 * nl); q->RH_1k = qdup(D->RH_1k, nl);
    q->ls = qdup(D->ls, 6 * nl); q->us = qdup(D->us, 6 * nl);
    q->line_res = qdup(D->line_res, 4 * nl);
    q->c1 = qdup(D->c1, ng); q->c2 = qdup(D->c2, ng);
    q->sqp_line = qdup(NULL, 6 * nl); q->membuf = qdup(NULL, 5 * nl); q->lambda = qdup(NULL, 4 * nl);
    /* the generator kernel runs on qpsub_pg/qg bounds and qpsub_c1/c2 (qpsub_generator_kernel_cpu.jl:7-15), the bus
     * kernel on qpsub_Pd / qpsub_Qd (qpsub_admm_update_xbar_cpu.jl:18-21) */
    orc_model_t *m = q->m;
    memcpy(m->pgmin_curr, D->pgmin, ng * sizeof(double)); memcpy(m->pgmax_curr, D->pgmax, ng * sizeof(double));
    memcpy(m->pgmin, D->pgmin, ng * sizeof(double));      memcpy(m->pgmax, D->pgmax, ng * sizeof(double));
    memcpy(m->qgmin, D->qgmin, ng * sizeof(double));      memcpy(m->qgmax, D->qgmax, ng * sizeof(double));
    memcpy(m->c1, D->c1, ng * sizeof(double));            memcpy(m->c2, D->c2, ng * sizeof(double));
    orc_set_load(m, D->Pd, D->Qd);
    *out = q;
    return EA_OK;
}

void orc_qp_destroy(orc_qp_t *q) {
    if (!q) return;
    orc_destroy(q->m);
    free(q->Hs); free(q->LH_1h); free(q->RH_1h); free(q->LH_1i); free(q->RH_1i); free(q->LH_1j); free(q->RH_1j);
    free(q->LH_1k); free(q->RH_1k); free(q->ls); free(q->us); free(q->line_res); free(q->c1); free(q->c2);
    free(q->sqp_line); free(q->membuf); free(q->lambda);
    free(q);
}
void orc_qp_set_threads(orc_qp_t *q, int n) { q->nthreads = n > 0 ? n : 1; orc_set_threads(q->m, n); }
int64_t orc_qp_nvar(const orc_qp_t *q) { return q->nvar; }
double *orc_qp_vector(orc_qp_t *q, int field) { return orc_vector(q->m, field); }
double *orc_qp_line_array(orc_qp_t *q, int which) {
    return which == EA_QP_SQP_LINE ? q->sqp_line : which == EA_QP_MEMBUF ? q->membuf : which == EA_QP_LAMBDA ? q->lambda : NULL;
}
void orc_qp_counters(const orc_qp_t *q, int64_t out[4]) { memcpy(out, q->cnt, sizeof(q->cnt)); }

/* rows of supY in the Hessian's variable order (w_ijR, w_ijI, w_i, w_j, theta_i, theta_j):
 * p_ij, q_ij, p_ji, q_ji as linear functions (qpsub_eval_Ab_linelimit_kernel_cpu.jl:37-40);
 * Y = {YffR,YffI,YftR,YftI,YttR,YttI,YtfR,YtfI} */
static void sup_rows(const double Y[8], double S[4][6]) {
    const double r[4][6] = { {  Y[2],  Y[3],  Y[0], 0, 0, 0 },
                             { -Y[3],  Y[2], -Y[1], 0, 0, 0 },
                             {  Y[6], -Y[7], 0,  Y[4], 0, 0 },
                             { -Y[7], -Y[6], 0, -Y[5], 0, 0 } };
    memcpy(S, r, sizeof(r));
}

/* qpsub_init_solution_cpu.jl:9-66 */
void orc_qp_init_solution(orc_qp_t *q, double rho_pq, double rho_va) {
    orc_model_t *m = q->m;
    for (int f = 0; f < EA_NUM_FIELDS; ++f) memset(m->vec[f], 0, sizeof(double) * (size_t)m->nvar);
    memset(q->lambda, 0, sizeof(double) * 4 * (size_t)q->nline);
    double *v = m->vec[EA_V_CURR], *rho = m->vec[EA_RHO];
    for (int64_t i = 0; i < m->nvar; ++i) rho[i] = rho_pq;
    for (int64_t g = 0; g < q->ngen; ++g) {
        v[2 * g] = 0.5 * (m->pgmin[g] + m->pgmax[g]);
        v[2 * g + 1] = 0.5 * (m->qgmin[g] + m->qgmax[g]);
    }
    for (int64_t l = 0; l < q->nline; ++l) {
        double Y[8], S[4][6], *sq = q->sqp_line + 6 * l, *p = v + 2 * q->ngen + 8 * l;
        for (int k = 0; k < 8; ++k) Y[k] = m->Y[k][l];
        sup_rows(Y, S);
        for (int k = 0; k < 6; ++k) sq[k] = (q->ls[6 * l + k] + q->us[6 * l + k]) / 2;
        for (int r = 0; r < 4; ++r) { double s = 0.0; for (int k = 0; k < 6; ++k) s += S[r][k] * sq[k]; p[r] = s; }
        p[4] = sq[2]; p[5] = sq[3]; p[6] = sq[4]; p[7] = sq[5];
        for (int k = 0; k < 8; ++k) rho[2 * q->ngen + 8 * l + k] = rho_va;
    }
}

/* eval_A_branch_kernel_cpu_qpsub / eval_b_branch_kernel_cpu_qpsub (qpsub_eval_Ab_linelimit_kernel_cpu.jl:24-57,
 * 138-165): the ADMM-augmented branch QP in the Hessian's variables, unscaled */
static void branch_Ab(const double H[36], const double l[8], const double rho[8], const double v[8], const double z[8],
                      const double Y[8], const double res[4], double Hbr[36], double bbr[6]) {
    double S[4][6];
    sup_rows(Y, S);
    memcpy(Hbr, H, 36 * sizeof(double));
    for (int r = 0; r < 4; ++r)
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Hbr[6 * i + j] += rho[r] * S[r][i] * S[r][j];
    Hbr[6 * 2 + 2] += rho[4]; Hbr[6 * 3 + 3] += rho[5]; Hbr[6 * 4 + 4] += rho[6]; Hbr[6 * 5 + 5] += rho[7];
    for (int i = 0; i < 6; ++i) bbr[i] = 0.0;
    for (int r = 0; r < 4; ++r) {
        const double c = l[r] - rho[r] * (v[r] - z[r] - res[r]);
        for (int i = 0; i < 6; ++i) bbr[i] += c * S[r][i];
    }
    bbr[2] += l[4] - rho[4] * (v[4] - z[4]);
    bbr[3] += l[5] - rho[5] * (v[5] - z[5]);
    bbr[4] += l[6] - rho[6] * (v[6] - z[6]);
    bbr[5] += l[7] - rho[7] * (v[7] - z[7]);
}

/* the line-limit rows 1j / 1k in TRON's 8 variables (t_ij, t_ji, w_ijR, w_ijI, w_i, w_j, theta_i, theta_j)
 * (qpsub_auglag_Ab_linelimit_kernel_red_cpu.jl:73-80) */
static void limit_rows(const double Y[8], const double LH_1j[2], const double LH_1k[2], double v1j[8], double v1k[8],
                       double S8[4][8]) {
    double S[4][6];
    sup_rows(Y, S);
    for (int r = 0; r < 4; ++r) { S8[r][0] = S8[r][1] = 0.0; for (int k = 0; k < 6; ++k) S8[r][2 + k] = S[r][k]; }
    for (int k = 0; k < 8; ++k) {
        v1j[k] = (k == 0 ? 1.0 : 0.0) + LH_1j[0] * S8[0][k] + LH_1j[1] * S8[1][k];
        v1k[k] = (k == 1 ? 1.0 : 0.0) + LH_1k[0] * S8[2][k] + LH_1k[1] * S8[3][k];
    }
}

/* eval_A_auglag_branch_kernel_cpu_qpsub_red + eval_b_auglag_branch_kernel_cpu_qpsub_red
 * (qpsub_eval_Ab_linelimit_kernel_cpu.jl:97-136, 203-253): x8 = C x6 + d eliminates (w_ijR, w_ijI) */
static void reduced_qp(const double Hbr[36], const double bbr[6], const double v1j[8], const double v1k[8],
                       const double LH_1h[4], double RH_1h, const double LH_1i[4], double RH_1i,
                       double RH_1j, double RH_1k, const double membuf5[5], double scale,
                       double A[36], double b[6], double C[48], double d[8]) {
    const double mu = membuf5[4];
    double A8[64], b8[8];
    memset(A8, 0, sizeof(A8)); memset(b8, 0, sizeof(b8));
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A8[8 * (i + 2) + (j + 2)] = Hbr[6 * i + j];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) A8[8 * i + j] += mu * v1j[i] * v1j[j];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) A8[8 * i + j] += mu * v1k[i] * v1k[j];
    for (int i = 0; i < 6; ++i) b8[i + 2] = bbr[i];
    for (int i = 0; i < 8; ++i) b8[i] += (membuf5[2] - mu * RH_1j) * v1j[i];
    for (int i = 0; i < 8; ++i) b8[i] += (membuf5[3] - mu * RH_1k) * v1k[i];
    /* inv([LH_1h[1] LH_1h[2]; LH_1i[1] LH_1i[2]]) in closed form (as qpsub_auglag_Ab_linelimit_kernel_red_gpu.jl:96-100) */
    const double prod = LH_1h[0] * LH_1i[1] - LH_1h[1] * LH_1i[0];
    const double i11 = LH_1i[1] / prod, i12 = -LH_1h[1] / prod, i21 = -LH_1i[0] / prod, i22 = LH_1h[0] / prod;
    memset(C, 0, 48 * sizeof(double)); memset(d, 0, 8 * sizeof(double));
    C[6 * 0 + 0] = 1.0; C[6 * 1 + 1] = 1.0;
    C[6 * 2 + 2] = -i11 * LH_1h[2]; C[6 * 2 + 3] = -i11 * LH_1h[3]; C[6 * 2 + 4] = -i12 * LH_1i[2]; C[6 * 2 + 5] = -i12 * LH_1i[3];
    C[6 * 3 + 2] = -i21 * LH_1h[2]; C[6 * 3 + 3] = -i21 * LH_1h[3]; C[6 * 3 + 4] = -i22 * LH_1i[2]; C[6 * 3 + 5] = -i22 * LH_1i[3];
    C[6 * 4 + 2] = 1.0; C[6 * 5 + 3] = 1.0; C[6 * 6 + 4] = 1.0; C[6 * 7 + 5] = 1.0;
    d[2] = i11 * RH_1h + i12 * RH_1i;
    d[3] = i21 * RH_1h + i22 * RH_1i;
    double AC[48], t[8];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 6; ++j) {
        double s = 0.0; for (int k = 0; k < 8; ++k) s += A8[8 * i + k] * C[6 * k + j]; AC[6 * i + j] = s;
    }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
        double s = 0.0; for (int k = 0; k < 8; ++k) s += C[6 * k + i] * AC[6 * k + j]; A[6 * i + j] = s * scale;
    }
    for (int i = 0; i < 8; ++i) { double s = 0.0; for (int k = 0; k < 8; ++k) s += A8[8 * i + k] * d[k]; t[i] = s + b8[i]; }
    for (int i = 0; i < 6; ++i) { double s = 0.0; for (int k = 0; k < 8; ++k) s += C[6 * k + i] * t[k]; b[i] = s * scale; }
}

void orc_qp_branch_qp(const double H[36], const double l[8], const double rho[8], const double v[8],
                      const double z[8], const double Y[8], const double res[4], const double LH_1h[4], double RH_1h,
                      const double LH_1i[4], double RH_1i, const double LH_1j[2], double RH_1j,
                      const double LH_1k[2], double RH_1k, const double membuf5[5], double scale,
                      double A[36], double b[6], double C[48], double d[8]) {
    double Hbr[36], bbr[6], v1j[8], v1k[8], S8[4][8];
    branch_Ab(H, l, rho, v, z, Y, res, Hbr, bbr);
    limit_rows(Y, LH_1j, LH_1k, v1j, v1k, S8);
    reduced_qp(Hbr, bbr, v1j, v1k, LH_1h, RH_1h, LH_1i, RH_1i, RH_1j, RH_1k, membuf5, scale, A, b, C, d);
}

/* build_QP_DS (qpsub_auglag_tron_linelimit_kernel_cpu.jl:41-71): f and g use the full matrix, the Hessian handed to
 * TRON is its lower triangle */
typedef struct { const double *A, *b; } qp_ctx_t;
static double qp_f(const double *x, const void *c) {
    const qp_ctx_t *q = (const qp_ctx_t *)c;
    double xAx = 0.0, bx = 0.0;
    for (int j = 0; j < 6; ++j) { double s = 0.0; for (int i = 0; i < 6; ++i) s += x[i] * q->A[6 * i + j]; xAx += s * x[j]; }
    for (int i = 0; i < 6; ++i) bx += q->b[i] * x[i];
    return 0.5 * xAx + bx;
}
static void qp_gh(const double *x, const void *c, double *g, double *Aout) {
    const qp_ctx_t *q = (const qp_ctx_t *)c;
    for (int i = 0; i < 6; ++i) { double s = 0.0; for (int j = 0; j < 6; ++j) s += q->A[6 * i + j] * x[j]; g[i] = s + q->b[i]; }
    for (int j = 0; j < 6; ++j) for (int i = j; i < 6; ++i) Aout[NV * i + j] = Aout[NV * j + i] = q->A[6 * i + j];
}

/* auglag_Ab_linelimit_two_level_alternative_qpsub_ij_red (qpsub_auglag_Ab_linelimit_kernel_red_cpu.jl:29-157), one
 * branch on explicit inputs. sq = sqp_line column (in: start point, out: solution), mb = qpsub_membuf column (in / out),
 * u (8) and lam (4) out, work = { AL iterations, objective evaluations }. */
void orc_qp_branch_solve(const double H[36], const double l[8], const double rho[8], const double v[8], const double z[8],
                         const double Y[8], const double res[4], const double LH_1h[4], double RH_1h,
                         const double LH_1i[4], double RH_1i, const double LH_1j[2], double RH_1j,
                         const double LH_1k[2], double RH_1k, const double ls[6], const double us[6], double sq[6],
                         double mb[5], int64_t major_iter, int32_t max_auglag, double mu_max, double scale,
                         double u[8], double lam[4], int32_t work[2]) {
    double Hbr[36], bbr[6], v1j[8], v1k[8], S8[4][8];
    branch_Ab(H, l, rho, v, z, Y, res, Hbr, bbr);
    limit_rows(Y, LH_1j, LH_1k, v1j, v1k, S8);

    double x[6] = { 0.0, 0.0, sq[2], sq[3], sq[4], sq[5] };
    const double xl[6] = { 0.0, 0.0, ls[2], ls[3], ls[4], ls[5] };
    const double xu[6] = { 200000.0, 200000.0, us[2], us[3], us[4], us[5] };
    double trg[6] = { 0, 0, 0, 0, 0, 0 }, x8[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    double mu;
    if (major_iter == 1) { mb[4] = 10.0; mu = 10.0; } else mu = mb[4];
    double eta = 1 / pow(mu, 0.1);
    int it = 0, terminate = 0;
    work[1] = 0;
    while (!terminate) {
        it++;
        double A[36], b[6], C[48], d[8];
        reduced_qp(Hbr, bbr, v1j, v1k, LH_1h, RH_1h, LH_1i, RH_1i, RH_1j, RH_1k, mb, scale, A, b, C, d);
        const qp_ctx_t ctx = { A, b };
        tron_stats_t st; memset(&st, 0, sizeof(st));
        int minor = 0;
        orc__tron_cb_g(6, x, xl, xu, qp_f, qp_gh, &ctx, 500, 200, 1e-6, &minor, &st, trg);
        work[1] += (int32_t)st.nfev;
        for (int i = 0; i < 8; ++i) { double s = 0.0; for (int k = 0; k < 6; ++k) s += C[6 * i + k] * x[k]; x8[i] = s + d[i]; }
        for (int k = 0; k < 6; ++k) sq[k] = x8[2 + k];
        double c3 = 0.0, c4 = 0.0;
        for (int k = 0; k < 8; ++k) { c3 += v1j[k] * x8[k]; c4 += v1k[k] * x8[k]; }
        c3 -= RH_1j; c4 -= RH_1k;
        const double cnorm = fmax(fabs(c3), fabs(c4));
        if (cnorm <= eta) {
            if (cnorm <= 1e-6) terminate = 1;
            else { mb[2] += mu * c3; mb[3] += mu * c4; eta = eta / pow(mu, 0.9); }
        } else {
            mu = fmin(mu_max, mu * 10);
            eta = 1 / pow(mu, 0.1);
            mb[4] = mu;
        }
        if (it >= max_auglag && cnorm > 1e-6) terminate = 1;
    }
    work[0] = it;
    for (int r = 0; r < 4; ++r) { double s = 0.0; for (int k = 0; k < 8; ++k) s += S8[r][k] * x8[k]; u[r] = s + res[r]; }
    u[4] = x[2]; u[5] = x[3]; u[6] = x[4]; u[7] = x[5];
    /* multipliers of 14h / 14i (:138-149): tmpH = inv([LH_1h[1] LH_1i[1]; LH_1h[2] LH_1i[2]]) */
    const double prod = LH_1h[0] * LH_1i[1] - LH_1i[0] * LH_1h[1];
    const double t11 = LH_1i[1] / prod, t12 = -LH_1i[0] / prod, t21 = -LH_1h[1] / prod, t22 = LH_1h[0] / prod;
    const double ti[2] = { 2 * u[0] * Y[2] + 2 * u[1] * (-Y[3]), 2 * u[0] * Y[3] + 2 * u[1] * Y[2] };
    const double th[2] = { 2 * u[2] * Y[6] + 2 * u[3] * (-Y[7]), 2 * u[2] * (-Y[7]) + 2 * u[3] * (-Y[6]) };
    double w[2];
    for (int r = 0; r < 2; ++r) {
        double s = trg[0] * ti[r] + trg[1] * th[r];
        for (int k = 0; k < 6; ++k) s += Hbr[6 * r + k] * sq[k];
        w[r] = s + bbr[r];
    }
    lam[0] = -(t11 * w[0] + t12 * w[1]);
    lam[1] = -(t21 * w[0] + t22 * w[1]);
    lam[2] = -fabs(trg[0]);
    lam[3] = -fabs(trg[1]);
}

static void solve_branch_qp(orc_qp_t *q, int64_t I, int64_t major_iter, int32_t max_auglag, double mu_max, double scale,
                            int64_t cnt[4]) {
    orc_model_t *m = q->m;
    const int64_t p = 2 * q->ngen + 8 * I;
    double Y[8];
    int32_t work[2];
    for (int k = 0; k < 8; ++k) Y[k] = m->Y[k][I];
    orc_qp_branch_solve(q->Hs + 36 * I, m->vec[EA_L_CURR] + p, m->vec[EA_RHO] + p, m->vec[EA_V_CURR] + p,
                        m->vec[EA_Z_CURR] + p, Y, q->line_res + 4 * I, q->LH_1h + 4 * I, q->RH_1h[I], q->LH_1i + 4 * I,
                        q->RH_1i[I], q->LH_1j + 2 * I, q->RH_1j[I], q->LH_1k + 2 * I, q->RH_1k[I], q->ls + 6 * I,
                        q->us + 6 * I, q->sqp_line + 6 * I, q->membuf + 5 * I, major_iter, max_auglag, mu_max, scale,
                        m->vec[EA_U_CURR] + p, q->lambda + 4 * I, work);
    cnt[0]++; cnt[1] += work[0]; cnt[2] += work[1];
    if (work[0] > cnt[3]) cnt[3] = work[0];
}

/* admm_update_x: acopf_admm_update_x_gen (qpsub bounds / costs) + acopf_admm_update_x_line (qpsub_admm_update_x_cpu.jl) */
void orc_qp_update_x(orc_qp_t *q, int64_t inner, int32_t max_auglag, double mu_max, double scale) {
    orc_update_x_gen(q->m);
#ifdef _OPENMP
    if (q->nthreads > 1) {
        #pragma omp parallel num_threads(q->nthreads)
        {
            int64_t c[4] = { 0, 0, 0, 0 };
            #pragma omp for schedule(dynamic, 16)
            for (int64_t I = 0; I < q->nline; ++I) solve_branch_qp(q, I, inner, max_auglag, mu_max, scale, c);
            #pragma omp critical
            { q->cnt[0] += c[0]; q->cnt[1] += c[1]; q->cnt[2] += c[2]; if (c[3] > q->cnt[3]) q->cnt[3] = c[3]; }
        }
        return;
    }
#endif
    for (int64_t I = 0; I < q->nline; ++I) solve_branch_qp(q, I, inner, max_auglag, mu_max, scale, q->cnt);
}

/* qpsub_admm_update_xbar_cpu.jl:11-24 */
void orc_qp_update_xbar(orc_qp_t *q) {
    memcpy(q->m->vec[EA_V_PREV], q->m->vec[EA_V_CURR], sizeof(double) * (size_t)q->nvar);
    orc_update_xbar(q->m);
}

/* qpsub_admm_update_l_single_cpu.jl:9-19 */
void orc_qp_update_l_single(orc_qp_t *q) {
    double *l = q->m->vec[EA_L_CURR];
    const double *rho = q->m->vec[EA_RHO], *u = q->m->vec[EA_U_CURR], *v = q->m->vec[EA_V_CURR];
    for (int64_t i = 0; i < q->nvar; ++i) l[i] = l[i] + rho[i] * (u[i] - v[i]);
}

static double qp_objective(const orc_qp_t *q) {
    const double *u = q->m->vec[EA_U_CURR];
    double obj = 0.0, quad = 0.0;
    for (int64_t g = 0; g < q->ngen; ++g) {
        const double pg = q->m->baseMVA * u[2 * g];
        obj += q->c2[g] * (pg * pg) + q->c1[g] * pg;
    }
    for (int64_t l = 0; l < q->nline; ++l) {
        const double *x = q->sqp_line + 6 * l, *H = q->Hs + 36 * l;
        double xHx = 0.0;
        for (int j = 0; j < 6; ++j) { double s = 0.0; for (int i = 0; i < 6; ++i) s += x[i] * H[6 * i + j]; xHx += s * x[j]; }
        quad += 0.5 * xHx;
    }
    return obj + quad;
}
static double qp_auglag(const orc_qp_t *q, double objval, double beta) {
    const orc_model_t *m = q->m;
    const double *lz = m->vec[EA_LZ], *z = m->vec[EA_Z_CURR], *l = m->vec[EA_L_CURR], *rp = m->vec[EA_RP], *rho = m->vec[EA_RHO];
    double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    for (int64_t i = 0; i < q->nvar; ++i) { s1 += lz[i] * z[i]; s2 += z[i] * z[i]; s3 += l[i] * rp[i]; s4 += rho[i] * (rp[i] * rp[i]); }
    return objval + s1 + 0.5 * beta * s2 + s3 + 0.5 * s4;
}

/* qpsub_admm_update_residual_cpu.jl:11-40 (par.beta = 0 on this path, admm_one_level.jl:17-18) */
void orc_qp_update_residual(orc_qp_t *q, double out[5]) {
    orc_model_t *m = q->m;
    double *rp = m->vec[EA_RP], *rd = m->vec[EA_RD], *ab = m->vec[EA_AX_PLUS_BY];
    const double *u = m->vec[EA_U_CURR], *v = m->vec[EA_V_CURR], *vp = m->vec[EA_V_PREV], *rho = m->vec[EA_RHO];
    for (int64_t i = 0; i < q->nvar; ++i) {
        rp[i] = u[i] - v[i];
        rd[i] = rho[i] * (v[i] - vp[i]);
        ab[i] = rp[i];
    }
    out[0] = qnorm(rp, q->nvar); out[1] = qnorm(rd, q->nvar); out[2] = qnorm(ab, q->nvar);
    out[3] = qp_objective(q);
    out[4] = qp_auglag(q, out[3], 0.0);
}

/* qpsub_admm_prepoststep_cpu.jl:9-84 */
void orc_qp_poststep(orc_qp_t *q, double *objval, double *auglag, double *dw_sol, double *dtheta_sol, double *dual_infeas) {
    const orc_model_t *m = q->m;
    const double obj = qp_objective(q);
    if (objval) *objval = obj;
    if (auglag) *auglag = qp_auglag(q, obj, 0.0);
    if (dual_infeas) {
        const double *u = m->vec[EA_U_CURR];
        for (int64_t g = 0; g < q->ngen; ++g) dual_infeas[g] = 2 * q->c2[g] * (m->baseMVA * m->baseMVA) * u[2 * g];
        for (int64_t l = 0; l < q->nline; ++l)
            for (int i = 0; i < 6; ++i) {
                double s = 0.0;
                for (int j = 0; j < 6; ++j) s += q->Hs[36 * l + 6 * i + j] * q->sqp_line[6 * l + j];
                dual_infeas[q->ngen + 6 * l + i] = s;
            }
    }
    if (dw_sol || dtheta_sol)
        for (int64_t b = 0; b < q->nbus; ++b) {
            double ws = 0, ts = 0; int ct = 0;
            for (int64_t k = m->FrStart[b]; k < m->FrStart[b + 1]; ++k) {
                ws += q->sqp_line[6 * m->FrIdx[k] + 2]; ts += q->sqp_line[6 * m->FrIdx[k] + 4]; ct++;
            }
            for (int64_t k = m->ToStart[b]; k < m->ToStart[b + 1]; ++k) {
                ws += q->sqp_line[6 * m->ToIdx[k] + 3]; ts += q->sqp_line[6 * m->ToIdx[k] + 5]; ct++;
            }
            if (dw_sol) dw_sol[b] = ws / ct;
            if (dtheta_sol) dtheta_sol[b] = ts / ct;
        }
}

/* src/algorithms/admm_one_level.jl:1-81 */
int orc_qp_admm_one_level(orc_qp_t *q, const ea_params_t *par, ea_info_t *info) {
    orc_model_t *m = q->m;
    const double sqrt_d = sqrt((double)q->nvar);
    const double OUTER_TOL = sqrt_d * par->outer_eps;
    memset(info, 0, sizeof(*info));
    info->mismatch = INFINITY;
    memset(m->vec[EA_LZ], 0, sizeof(double) * (size_t)q->nvar);
    memset(m->vec[EA_Z_CURR], 0, sizeof(double) * (size_t)q->nvar);
    memset(m->vec[EA_Z_PREV], 0, sizeof(double) * (size_t)q->nvar);
    double res[5];
    if (par->verbose > 0) {
        orc_qp_update_residual(q, res);
        info->primres = res[0]; info->dualres = res[1]; info->mismatch = res[2]; info->objval = res[3]; info->auglag = res[4];
    }
    info->status = EA_STATUS_ITERATION_LIMIT;
    const double norm_rho = qnorm(m->vec[EA_RHO], q->nvar);
    const double t0 = qp_wall();
    while (info->outer < par->outer_iterlim) {
        info->outer++;
        info->inner = 0;
        while (info->inner < 1) {                  /* par.inner_iterlim = 1 (admm_one_level.jl:22) */
            info->inner++; info->cumul++;
            double t = qp_wall();
            orc_qp_update_x(q, info->inner, par->max_auglag, par->mu_max, par->scale);
            double t1 = qp_wall(); info->time_x_update += t1 - t;
            orc_qp_update_xbar(q);
            double t2 = qp_wall(); info->time_xbar_update += t2 - t1;
            orc_qp_update_l_single(q);
            info->time_l_update += qp_wall() - t2;
            orc_qp_update_residual(q, res);
            info->primres = res[0]; info->dualres = res[1]; info->mismatch = res[2]; info->objval = res[3]; info->auglag = res[4];
            if (par->verbose > 1)
                printf("%8ld  %10.3e  %10.3e  %10.3e  %10.3e %10.3e  %10.3e\n", (long)info->outer, info->objval,
                       info->auglag, info->mismatch, OUTER_TOL, info->dualres, OUTER_TOL * norm_rho / sqrt_d);
        }
        if (info->mismatch <= OUTER_TOL && info->dualres <= OUTER_TOL * norm_rho / sqrt_d) {
            info->status = EA_STATUS_SOLVED;
            break;
        }
    }
    info->time_overall = qp_wall() - t0;
    orc_qp_poststep(q, &info->objval, &info->auglag, NULL, NULL, NULL);
    return EA_OK;
}
