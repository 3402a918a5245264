/*
 * mpacopf_oracle.c — CPU ORACLE (test infrastructure, NOT product code) for the
 * multi-period ACOPF model. See mpacopf_oracle.h for the pinning status.
 *
 * Restates the reference's `use_gpu=false` path of src/models/mpacopf/ (`ModelMpacopf`,
 * the "tight" coupling; `ModelMpacopfLoose` is not reachable from solve_mpacopf and
 * its constructor calls an undefined `Model{...}`): per-period ACOPF models from
 * acopf_oracle.c plus the ramp-coupling vectors and the n = 3 generator sub-problem
 * (p_t, phat_{t-1}, s_t) solved by the same AL + TRON scheme as a branch.
 * Indices are 0-based; period t here is the reference's i = t + 1.
 */
#include "mpacopf_oracle.h"
#include "oracle_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define GEN_MEMROWS 8   /* rows used of gen_membuf (mpacopf_model.jl:73-79) */

struct orc_mp {
    int32_t T;
    int64_t ngen, nvar;
    orc_model_t **m;
    double *ramp_rate;                              /* ngen */
    double **r;                                     /* T x EA_RAMP_NUM_FIELDS vectors of ngen */
    double **gen_membuf;                            /* T x (8 x ngen) */
    double (*res)[4];                               /* per-period residual norms (models[i].info) */
    double *norm_z_prev;                            /* per period */
    int64_t gen_cnt[5];                             /* generator sub-problems: calls, AL iterations, f-evaluations, max AL its, max evals of one call */
};

static double mp_dmin(double a, double b) { return a < b ? a : b; }
static double mp_dmax(double a, double b) { return a > b ? a : b; }
static double mp_wall(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static double mp_sumsq(const double *x, int64_t n) {
    double s = 0.0; for (int64_t i = 0; i < n; ++i) s += x[i] * x[i]; return s;
}
static double *R(orc_mp_t *mp, int t, int f) { return mp->r[t * EA_RAMP_NUM_FIELDS + f]; }

/* mpacopf_model.jl:56-109 */
int orc_mp_create(const ea_grid_t *G, int32_t T, const double *Pd, const double *Qd, double ramp_ratio, orc_mp_t **out) {
    if (!G || !out || T < 1 || !Pd || !Qd) return EA_ERR_ARG;
    orc_mp_t *mp = (orc_mp_t *)calloc(1, sizeof(*mp));
    if (!mp) return EA_ERR_ALLOC;
    mp->T = T; mp->ngen = G->ngen;
    mp->m = (orc_model_t **)calloc((size_t)T, sizeof(*mp->m));
    mp->r = (double **)calloc((size_t)T * EA_RAMP_NUM_FIELDS, sizeof(double *));
    mp->gen_membuf = (double **)calloc((size_t)T, sizeof(double *));
    mp->res = calloc((size_t)T, sizeof(*mp->res));
    mp->norm_z_prev = (double *)calloc((size_t)T, sizeof(double));
    mp->ramp_rate = (double *)calloc((size_t)(G->ngen > 0 ? G->ngen : 1), sizeof(double));
    for (int64_t g = 0; g < G->ngen; ++g) mp->ramp_rate[g] = ramp_ratio * G->pgmax[g];
    for (int t = 0; t < T; ++t) {
        const int rc = orc_create(G, &mp->m[t]);
        if (rc != EA_OK) { orc_mp_destroy(mp); return rc; }
        orc_set_load(mp->m[t], Pd + (size_t)t * (size_t)G->nbus, Qd + (size_t)t * (size_t)G->nbus);
        for (int f = 0; f < EA_RAMP_NUM_FIELDS; ++f)
            mp->r[t * EA_RAMP_NUM_FIELDS + f] = (double *)calloc((size_t)(G->ngen > 0 ? G->ngen : 1), sizeof(double));
        mp->gen_membuf[t] = (double *)calloc((size_t)GEN_MEMROWS * (size_t)(G->ngen > 0 ? G->ngen : 1), sizeof(double));
    }
    mp->nvar = (T == 1) ? orc_nvar(mp->m[0]) : orc_nvar(mp->m[0]) + G->ngen;      /* mpacopf_model.jl:97-102 */
    *out = mp;
    return EA_OK;
}

void orc_mp_destroy(orc_mp_t *mp) {
    if (!mp) return;
    for (int t = 0; t < mp->T; ++t) {
        if (mp->m && mp->m[t]) orc_destroy(mp->m[t]);
        if (mp->r) for (int f = 0; f < EA_RAMP_NUM_FIELDS; ++f) free(mp->r[t * EA_RAMP_NUM_FIELDS + f]);
        if (mp->gen_membuf) free(mp->gen_membuf[t]);
    }
    free(mp->m); free(mp->r); free(mp->gen_membuf); free(mp->res); free(mp->norm_z_prev); free(mp->ramp_rate);
    free(mp);
}

void orc_mp_set_threads(orc_mp_t *mp, int n) { for (int t = 0; t < mp->T; ++t) orc_set_threads(mp->m[t], n); }
int32_t orc_mp_len_horizon(const orc_mp_t *mp) { return mp->T; }
orc_model_t *orc_mp_period(orc_mp_t *mp, int32_t t) { return (t >= 0 && t < mp->T) ? mp->m[t] : NULL; }
double *orc_mp_ramp_vector(orc_mp_t *mp, int32_t t, int f) {
    return (t >= 0 && t < mp->T && f >= 0 && f < EA_RAMP_NUM_FIELDS) ? R(mp, t, f) : NULL;
}
double *orc_mp_gen_membuf(orc_mp_t *mp, int32_t t) { return (t >= 0 && t < mp->T) ? mp->gen_membuf[t] : NULL; }
int64_t orc_mp_nvar(const orc_mp_t *mp) { return mp->nvar; }
void orc_mp_gen_counters(const orc_mp_t *mp, int64_t out[5]) { memcpy(out, mp->gen_cnt, sizeof(mp->gen_cnt)); }

/* mpacopf_init_solution_cpu.jl:1-21 */
void orc_mp_init_solution(orc_mp_t *mp, double rho_pq, double rho_va) {
    for (int t = 0; t < mp->T; ++t) {
        orc_init_solution(mp->m[t], rho_pq, rho_va);
        for (int f = 0; f < EA_RAMP_NUM_FIELDS; ++f) memset(R(mp, t, f), 0, sizeof(double) * (size_t)mp->ngen);
        for (int64_t g = 0; g < mp->ngen; ++g) R(mp, t, EA_RAMP_RHO)[g] = rho_pq;
        if (t > 0) {
            const double *vprev = orc_vector(mp->m[t - 1], EA_V_CURR), *u = orc_vector(mp->m[t], EA_U_CURR);
            for (int64_t g = 0; g < mp->ngen; ++g) {
                R(mp, t, EA_RAMP_U_CURR)[g] = vprev[2 * g];
                R(mp, t, EA_RAMP_S_CURR)[g] = u[2 * g] - R(mp, t, EA_RAMP_U_CURR)[g];
            }
        }
    }
}

/* mpacopf_admm_prepoststep_cpu.jl:1-19 */
double orc_mp_outer_prestep(orc_mp_t *mp) {
    double top = 0.0;
    for (int t = 0; t < mp->T; ++t) {
        double nz = orc_outer_prestep(mp->m[t]);
        if (t > 0) nz = sqrt(nz * nz + mp_sumsq(R(mp, t, EA_RAMP_Z_CURR), mp->ngen));
        mp->norm_z_prev[t] = nz;
        top = mp_dmax(top, nz);
    }
    return top;
}

/* mpacopf_admm_prepoststep_cpu.jl:21-36 */
void orc_mp_inner_prestep(orc_mp_t *mp) {
    for (int t = 0; t < mp->T; ++t) {
        orc_inner_prestep(mp->m[t]);
        if (t > 0) memcpy(R(mp, t, EA_RAMP_Z_PREV), R(mp, t, EA_RAMP_Z_CURR), sizeof(double) * (size_t)mp->ngen);
    }
}

/* ---- generator sub-problem with ramping: objective (mpacopf_eval_generator_kernel_cpu.jl:1-100) ---- */
typedef struct { const double *p; double scale, c2, c1, c0, baseMVA; } gen_ctx_t;

static double gen_f(const double *x, const void *c) {
    const gen_ctx_t *k = (const gen_ctx_t *)c;
    const double *p = k->p;
    double f = 0.0;
    f += k->c2 * ((x[0] * k->baseMVA) * (x[0] * k->baseMVA)) + k->c1 * (x[0] * k->baseMVA) + k->c0;
    f += p[0] * (x[0] - p[4]) + (0.5 * p[2]) * ((x[0] - p[4]) * (x[0] - p[4]));
    f += p[1] * (x[1] - p[5]) + (0.5 * p[3]) * ((x[1] - p[5]) * (x[1] - p[5]));
    f += p[6] * (x[0] - x[1] - x[2]) + (0.5 * p[7]) * ((x[0] - x[1] - x[2]) * (x[0] - x[1] - x[2]));
    return f * k->scale;
}
static void gen_gh(const double *x, const void *c, double *g, double *A) {
    const gen_ctx_t *k = (const gen_ctx_t *)c;
    const double *p = k->p;
    const double B = k->baseMVA, s = k->scale;
    g[0] = 2 * k->c2 * (B * B) * x[0] + k->c1 * B;
    g[0] += p[0] + p[2] * (x[0] - p[4]);
    g[0] += p[6] + p[7] * (x[0] - x[1] - x[2]);
    g[0] *= s;
    g[1] = p[1] + p[3] * (x[1] - p[5]);
    g[1] += -(p[6] + p[7] * (x[0] - x[1] - x[2]));
    g[1] *= s;
    g[2] = -(p[6] + p[7] * (x[0] - x[1] - x[2]));
    g[2] *= s;
    double H[3][3];
    H[0][0] = s * (2 * k->c2 * (B * B) + p[2] + p[7]);
    H[1][0] = s * (-p[7]);
    H[2][0] = s * (-p[7]);
    H[1][1] = s * (p[3] + p[7]);
    H[2][1] = s * (p[7]);
    H[2][2] = s * (p[7]);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j <= i; ++j) { A[i * NV + j] = H[i][j]; A[j * NV + i] = H[i][j]; }
}

/* The AL loop around TRON for ONE generator (mpacopf_auglag_generator_kernel_cpu.jl:66-127).
 * param: gen_membuf column (8 doubles, rows 1-6 set by the caller, [6] = multiplier, [7] = xi; both
 * updated in place). work (optional): {auglag iterations, f-evaluations, cg iterations}. */
void orc_gen_ramp_solve(double x[3], const double xl[3], const double xu[3], double *param, double c2, double c1,
                        double c0, double baseMVA, double scale, int32_t max_auglag, double xi_max, int32_t *work) {
    const gen_ctx_t ctx = { param, scale, c2, c1, c0, baseMVA };
    double xi = param[7];
    double eta = 1 / pow(xi, 0.1);
    int it = 0, terminate = 0;
    tron_stats_t st; memset(&st, 0, sizeof(st));
    while (!terminate) {
        it++;
        int minor;
        orc__tron_cb(3, x, xl, xu, gen_f, gen_gh, &ctx, 500, 200, 1e-6, &minor, &st);
        const double cviol = x[0] - x[1] - x[2];
        const double cnorm = fabs(cviol);
        if (cnorm <= eta) {
            if (cnorm <= 1e-6) terminate = 1;
            else {
                param[6] += xi * cviol;
                eta = eta / pow(xi, 0.9);
            }
        } else {
            xi = mp_dmin(xi_max, xi * 10);
            eta = 1 / pow(xi, 0.1);
            param[7] = xi;
        }
        if (it >= max_auglag) terminate = 1;
    }
    if (work) { work[0] = it; work[1] = (int32_t)st.nfev; work[2] = (int32_t)st.cg; }
}

/* mpacopf_auglag_generator_kernel_cpu.jl:1-129 — generators of period t >= 1 */
static void gen_ramp_update(orc_mp_t *mp, int t, int64_t major_iter, int32_t max_auglag, double xi_max, double scale) {
    orc_model_t *m = mp->m[t];
    double *u = orc_vector(m, EA_U_CURR);
    const double *v = orc_vector(m, EA_V_CURR), *z = orc_vector(m, EA_Z_CURR), *l = orc_vector(m, EA_L_CURR),
                 *rho = orc_vector(m, EA_RHO);
    double *r_u = R(mp, t, EA_RAMP_U_CURR), *r_s = R(mp, t, EA_RAMP_S_CURR);
    const double *r_v = orc_vector(mp->m[t - 1], EA_V_CURR);    /* the previous period's whole v_curr */
    const double *r_z = R(mp, t, EA_RAMP_Z_CURR), *r_l = R(mp, t, EA_RAMP_L_CURR), *r_rho = R(mp, t, EA_RAMP_RHO);
    for (int64_t I = 0; I < mp->ngen; ++I) {
        double *param = mp->gen_membuf[t] + GEN_MEMROWS * I;
        const int64_t pg = 2 * I, qg = 2 * I + 1;
        u[qg] = mp_dmax(m->qgmin[I], mp_dmin(m->qgmax[I], (-(l[qg] + rho[qg] * (-v[qg] + z[qg]))) / rho[qg]));
        double x[3], xl[3], xu[3];
        xl[0] = xl[1] = m->pgmin[I];
        xu[0] = xu[1] = m->pgmax[I];
        xl[2] = -mp->ramp_rate[I];
        xu[2] = mp->ramp_rate[I];
        x[0] = mp_dmin(xu[0], mp_dmax(xl[0], u[pg]));
        x[1] = mp_dmin(xu[1], mp_dmax(xl[1], r_u[I]));
        x[2] = mp_dmin(xu[2], mp_dmax(xl[2], r_s[I]));
        param[0] = l[pg];
        param[1] = r_l[I];
        param[2] = rho[pg];
        param[3] = r_rho[I];
        param[4] = v[pg] - z[pg];
        param[5] = r_v[pg] - r_z[I];
        if (major_iter <= 1) param[7] = 10.0;
        int32_t work[3];
        orc_gen_ramp_solve(x, xl, xu, param, m->c2[I], m->c1[I], m->c0[I], m->baseMVA, scale, max_auglag, xi_max, work);
        mp->gen_cnt[0]++; mp->gen_cnt[1] += work[0]; mp->gen_cnt[2] += work[1];
        if (work[0] > mp->gen_cnt[3]) mp->gen_cnt[3] = work[0];
        if (work[1] > mp->gen_cnt[4]) mp->gen_cnt[4] = work[1];
        u[pg] = x[0];
        r_u[I] = x[1];
        r_s[I] = x[2];
    }
}

/* mpacopf_admm_update_x_cpu.jl:5-45: generators (t = 0 plain, t >= 1 with ramping; the generator
 * kernel gets scale = 1.0, not par.scale), then every period's branches */
void orc_mp_update_x(orc_mp_t *mp, int64_t inner, int32_t max_auglag, double mu_max, double scale) {
    orc_update_x_gen(mp->m[0]);
    for (int t = 1; t < mp->T; ++t) gen_ramp_update(mp, t, inner, max_auglag, mu_max, 1.0);
    for (int t = 0; t < mp->T; ++t) orc_update_x_line(mp->m[t], inner, max_auglag, mu_max, scale);
}

/* mpacopf_bus_kernel_cpu.jl:1-121 — one bus of period t < T-1: the consensus value of pg also
 * serves period t+1's copy phat_t */
static void solve_bus_ramp(orc_model_t *m, int64_t I, const double *r_u, const double *r_z, const double *r_l,
                           const double *r_rho) {
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
        const int64_t gid = m->GenIdx[k], p = 2 * gid;
        rhs1 += ((l[p] + rho[p] * (u[p] + z[p])) + (r_l[gid] + r_rho[gid] * (r_u[gid] + r_z[gid]))) / (rho[p] + r_rho[gid]);
        rhs2 += (u[p + 1] + z[p + 1]) + (l[p + 1] / rho[p + 1]);
        inv_pg += 1.0 / (rho[p] + r_rho[gid]);
        inv_qg += 1.0 / rho[p + 1];
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
        const int64_t gid = m->GenIdx[k], p = 2 * gid;
        v[p] = ((l[p] + rho[p] * (u[p] + z[p])) + (r_l[gid] + r_rho[gid] * (r_u[gid] + r_z[gid])) - mu1) / (rho[p] + r_rho[gid]);
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

/* mpacopf_admm_update_xbar_cpu.jl:1-20 */
void orc_mp_update_xbar(orc_mp_t *mp) {
    for (int t = 0; t + 1 < mp->T; ++t) {
        orc_model_t *m = mp->m[t];
        const double *r_u = R(mp, t + 1, EA_RAMP_U_CURR), *r_z = R(mp, t + 1, EA_RAMP_Z_CURR),
                     *r_l = R(mp, t + 1, EA_RAMP_L_CURR), *r_rho = R(mp, t + 1, EA_RAMP_RHO);
        for (int64_t I = 0; I < m->nbus; ++I) solve_bus_ramp(m, I, r_u, r_z, r_l, r_rho);
    }
    orc_update_xbar(mp->m[mp->T - 1]);
}

/* mpacopf_admm_update_z_cpu.jl:1-17 */
void orc_mp_update_z(orc_mp_t *mp, double beta) {
    for (int t = 0; t < mp->T; ++t) orc_update_z(mp->m[t], beta);
    for (int t = 1; t < mp->T; ++t) {
        const double *vp = orc_vector(mp->m[t - 1], EA_V_CURR);
        double *z = R(mp, t, EA_RAMP_Z_CURR);
        const double *lz = R(mp, t, EA_RAMP_LZ), *l = R(mp, t, EA_RAMP_L_CURR), *rho = R(mp, t, EA_RAMP_RHO),
                     *u = R(mp, t, EA_RAMP_U_CURR);
        for (int64_t g = 0; g < mp->ngen; ++g) z[g] = (-(lz[g] + l[g] + rho[g] * (u[g] - vp[2 * g]))) / (beta + rho[g]);
    }
}

/* mpacopf_admm_update_l_cpu.jl:1-16 */
void orc_mp_update_l(orc_mp_t *mp, double beta) {
    for (int t = 0; t < mp->T; ++t) orc_update_l(mp->m[t], beta);
    for (int t = 1; t < mp->T; ++t) {
        double *l = R(mp, t, EA_RAMP_L_CURR);
        const double *lz = R(mp, t, EA_RAMP_LZ), *z = R(mp, t, EA_RAMP_Z_CURR);
        for (int64_t g = 0; g < mp->ngen; ++g) l[g] = -(lz[g] + beta * z[g]);
    }
}

/* mpacopf_admm_update_lz_cpu.jl:1-16 */
void orc_mp_update_lz(orc_mp_t *mp, double beta, double M) {
    for (int t = 0; t < mp->T; ++t) orc_update_lz(mp->m[t], beta, M);
    for (int t = 1; t < mp->T; ++t) {
        double *lz = R(mp, t, EA_RAMP_LZ);
        const double *z = R(mp, t, EA_RAMP_Z_CURR);
        for (int64_t g = 0; g < mp->ngen; ++g) lz[g] = mp_dmax(-M, mp_dmin(M, lz[g] + (beta * z[g])));
    }
}

/* mpacopf_admm_update_residual_cpu.jl:1-55: per-period norms, the ramp vectors of period t are
 * added to period t's norms, the model-level value is the MAX over periods */
void orc_mp_update_residual(orc_mp_t *mp, double out[4]) {
    for (int t = 0; t < mp->T; ++t) orc_update_residual(mp->m[t], mp->res[t]);
    for (int t = 1; t < mp->T; ++t) {
        const double *vp = orc_vector(mp->m[t - 1], EA_V_CURR);
        double *rp = R(mp, t, EA_RAMP_RP), *rd = R(mp, t, EA_RAMP_RD), *ab = R(mp, t, EA_RAMP_AX_PLUS_BY);
        const double *u = R(mp, t, EA_RAMP_U_CURR), *z = R(mp, t, EA_RAMP_Z_CURR), *zp = R(mp, t, EA_RAMP_Z_PREV);
        for (int64_t g = 0; g < mp->ngen; ++g) {
            rp[g] = u[g] - vp[2 * g] + z[g];
            rd[g] = z[g] - zp[g];
            ab[g] = rp[g] - z[g];
        }
        const double n0 = sqrt(mp_sumsq(rp, mp->ngen)), n1 = sqrt(mp_sumsq(rd, mp->ngen)),
                     n2 = sqrt(mp_sumsq(z, mp->ngen)), n3 = sqrt(mp_sumsq(ab, mp->ngen));
        mp->res[t][0] = sqrt(mp->res[t][0] * mp->res[t][0] + n0 * n0);
        mp->res[t][1] = sqrt(mp->res[t][1] * mp->res[t][1] + n1 * n1);
        mp->res[t][2] = sqrt(mp->res[t][2] * mp->res[t][2] + n2 * n2);
        mp->res[t][3] = sqrt(mp->res[t][3] * mp->res[t][3] + n3 * n3);
    }
    for (int k = 0; k < 4; ++k) out[k] = 0.0;
    for (int t = 0; t < mp->T; ++t)
        for (int k = 0; k < 4; ++k) out[k] = mp_dmax(out[k], mp->res[t][k]);
}

/* mpacopf_admm_prepoststep_cpu.jl:38-74 */
double orc_mp_poststep(orc_mp_t *mp, double *err_ramp) {
    double obj = 0.0, err = 0.0;
    for (int t = 0; t < mp->T; ++t) obj += orc_poststep(mp->m[t]);
    for (int t = 1; t < mp->T; ++t) {
        const double *uc = orc_vector(mp->m[t], EA_U_CURR), *up = orc_vector(mp->m[t - 1], EA_U_CURR);
        for (int64_t g = 0; g < mp->ngen; ++g)
            err = mp_dmax(err, mp_dmax(0.0, -(mp->ramp_rate[g] - fabs(uc[2 * g] - up[2 * g]))));
    }
    if (err_ramp) *err_ramp = err;
    return obj;
}

/* src/algorithms/admm_two_level.jl:1-88 with mod::ModelMpacopf */
int orc_mp_admm_two_level(orc_mp_t *mp, const ea_params_t *par, ea_info_t *info, double *err_ramp) {
    const double sqrt_d = sqrt((double)mp->nvar);
    const double OUTER_TOL = sqrt_d * par->outer_eps;
    memset(info, 0, sizeof(*info));
    info->mismatch = INFINITY; info->norm_z_prev = INFINITY; info->norm_z_curr = INFINITY;
    double beta = par->initial_beta;
    double res[4];
    if (par->verbose > 0) {
        orc_mp_update_residual(mp, res);
        info->primres = res[0]; info->dualres = res[1]; info->norm_z_curr = res[2]; info->mismatch = res[3];
    }
    info->status = EA_STATUS_ITERATION_LIMIT;
    const double t0 = mp_wall();
    while (info->outer < par->outer_iterlim) {
        info->outer++;
        info->norm_z_prev = orc_mp_outer_prestep(mp);
        info->inner = 0;
        while (info->inner < par->inner_iterlim) {
            info->inner++; info->cumul++;
            orc_mp_inner_prestep(mp);
            double t = mp_wall();
            orc_mp_update_x(mp, info->inner, par->max_auglag, par->mu_max, par->scale);
            double t1 = mp_wall(); info->time_x_update += t1 - t;
            orc_mp_update_xbar(mp);
            double t2 = mp_wall(); info->time_xbar_update += t2 - t1; info->time_buses += t2 - t1;
            orc_mp_update_z(mp, beta);
            double t3 = mp_wall(); info->time_z_update += t3 - t2;
            orc_mp_update_l(mp, beta);
            info->time_l_update += mp_wall() - t3;
            orc_mp_update_residual(mp, res);
            info->primres = res[0]; info->dualres = res[1]; info->norm_z_curr = res[2]; info->mismatch = res[3];
            info->eps_pri = sqrt_d / (2500.0 * (double)info->outer);
            if (par->verbose > 1)
                printf("%8ld  %8ld  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e\n",
                       (long)info->outer, (long)info->inner, info->primres, info->eps_pri, info->dualres,
                       info->norm_z_curr, info->mismatch, OUTER_TOL, beta);
            if (info->primres <= info->eps_pri) break;
        }
        if (info->mismatch <= OUTER_TOL) { info->status = EA_STATUS_SOLVED; break; }
        orc_mp_update_lz(mp, beta, par->MAX_MULTIPLIER);
        if (info->norm_z_curr > par->theta * info->norm_z_prev) beta = mp_dmin(par->inc_c * beta, 1e24);
    }
    info->time_overall = mp_wall() - t0;
    info->beta = beta;
    info->objval = orc_mp_poststep(mp, err_ramp);
    return EA_OK;
}
