// host_harness.cu — TEST TOOL. Compiles the product's device routines
// (exaadmm.jl_b200/csrc/tron.cuh, branch.cuh) for the HOST so that the CPU test
// suite can drive exactly the code the GPU runs (flattened AL/TRON loop, masked
// free-set) against the oracle without a GPU. Not part of the product library.
#include "../../exaadmm.jl_b200/csrc/branch.cuh"
#include "../../exaadmm.jl_b200/csrc/genramp.cuh"
#include "../../exaadmm.jl_b200/csrc/qpsub.cuh"
#include <cmath>
#include <cstring>

static void make_pow_table(branch::PowTable &T, double mu_max) {
    T.n = 0;
    double m = 10.0;
    for (int k = 0; k < 24; ++k) {
        T.mu[k] = m; T.inv_p01[k] = 1.0 / std::pow(m, 0.1); T.p09[k] = std::pow(m, 0.9); T.n = k + 1;
        double nx = std::fmin(mu_max, m * 10); if (nx == m) break; m = nx;
    }
}

// The ORACLE's f / grad / Hessian plugged into the product's state machine
// (function pointers into oracle/_build/libacopf_oracle.so). Built with -DEA_NO_FMA the
// arithmetic of the TRON routines is then bit-identical to the oracle's restatement, so
// iterates and evaluation counts must match EXACTLY - any difference is a logic error.
typedef double (*orc_f_t)(const double *, const double *, const double *, double);
typedef void (*orc_gh_t)(const double *, const double *, const double *, double, double *, double *);

struct OracleEval {
    orc_f_t fn; orc_gh_t ghn; double *param; const double *Y; double scale;
    void operator()(const double (&x)[6], const double (&ls)[2], double mu, double &f, double (&g)[6],
                    branch::Sym6 &A, double (&F)[4]) const {
        param[24] = ls[0]; param[25] = ls[1]; param[26] = mu;
        double H[36];
        f = fn(x, param, Y, scale);
        ghn(x, param, Y, scale, g, H);
        A.from_dense(H);
        const double cc = x[0] * x[1] * cos(x[2] - x[3]), ss = x[0] * x[1] * sin(x[2] - x[3]);
        F[0] = Y[0] * (x[0] * x[0]) + Y[2] * cc + Y[3] * ss;        // as the oracle's AL loop computes them
        F[1] = -Y[1] * (x[0] * x[0]) - Y[3] * cc + Y[2] * ss;
        F[2] = Y[4] * (x[1] * x[1]) + Y[6] * cc - Y[7] * ss;
        F[3] = -Y[5] * (x[1] * x[1]) - Y[7] * cc - Y[6] * ss;
    }
};

template <class Eval>
static void run(const Eval &eval, double *x, const double *xl, const double *xu, double *param, long long major_iter,
                int max_auglag, double mu_max, double *F, int *work) {
    branch::PowTable T;
    make_pow_table(T, mu_max);
    branch::Lane L;
    double cold[branch::COLD_ROWS];
    L.cold = cold; L.cs = 1;
    double l[6], u[6];
    for (int k = 0; k < 6; ++k) { L.x[k] = x[k]; l[k] = xl[k]; u[k] = xu[k]; }
    L.ls[0] = param[24]; L.ls[1] = param[25];
    L.mu = (major_iter == 1) ? 10.0 : param[26];
    branch::solve(L, eval, l, u, max_auglag, mu_max, T);
    for (int k = 0; k < 6; ++k) x[k] = L.x[k];
    if (F) for (int k = 0; k < 4; ++k) F[k] = L.Fc(k);
    param[24] = L.ls[0]; param[25] = L.ls[1]; param[26] = L.mu;
    work[0] = L.it_al; work[1] = L.evals; work[2] = L.cg; work[3] = L.shifts; work[4] = L.rejected; work[5] = L.hit_max;
}

extern "C" {

#ifdef EA_STATS
// analysis build (tools/step_stats.py): outcome counters of tron::newton_step, [later steps | first step] x reason
void hh_stats(long long *out, int reset) {
    for (int f = 0; f < 2; ++f) for (int k = 0; k < 16; ++k) { out[16 * f + k] = tron::g_stat[f][k]; if (reset) tron::g_stat[f][k] = 0; }
}
#endif

// param: 31 doubles (membuf column, 0-based rows); rows 24-26 (lambda_s, mu) updated in place.
// x: 6 doubles in/out. Y: 8. xl/xu: 6. work[6]: auglag, evals, cg, shifts, rejected, hit_max.
void hh_solve_branch(double *x, const double *xl, const double *xu, double *param, const double *Y,
                     long long major_iter, int max_auglag, double mu_max, double scale, double *F, int *work) {
    branch::Data D;
    for (int k = 0; k < 8; ++k) { D.lam[k] = param[k]; D.rho[k] = param[8 + k]; D.xt[k] = param[16 + k]; D.Y[k] = Y[k]; }
    const branch::Objective<branch::StructView> eval{ { &D }, scale };
    run(eval, x, xl, xu, param, major_iter, max_auglag, mu_max, F, work);
}

void hh_solve_branch_oracle_eval(void *fn, void *ghn, double *x, const double *xl, const double *xu, double *param,
                                 const double *Y, long long major_iter, int max_auglag, double mu_max, double scale,
                                 int *work) {
    const OracleEval eval{ (orc_f_t)fn, (orc_gh_t)ghn, param, Y, scale };
    run(eval, x, xl, xu, param, major_iter, max_auglag, mu_max, nullptr, work);
}

void hh_eval(const double *x, const double *param, const double *Y, double scale, double *f, double *g, double *H) {
    branch::Data D;
    for (int k = 0; k < 8; ++k) { D.lam[k] = param[k]; D.rho[k] = param[8 + k]; D.xt[k] = param[16 + k]; D.Y[k] = Y[k]; }
    double ls[2] = { param[24], param[25] }, xx[6], gg[6], F[4];
    for (int k = 0; k < 6; ++k) xx[k] = x[k];
    branch::Sym6 A;
    branch::eval_fgh(branch::StructView{ &D }, ls, param[26], scale, xx, *f, gg, A, F);
    for (int a = 0; a < 6; ++a) { g[a] = gg[a]; for (int b = 0; b < 6; ++b) H[6 * a + b] = A.at(a, b); }
}

// One generator of the multi-period model (genramp.cuh). param: gen_membuf column (8 doubles; [6] multiplier and
// [7] xi updated in place). work[3]: auglag iterations, f-evaluations, cg iterations.
void hh_solve_gen(double *x, const double *xl, const double *xu, double *param, double c2, double c1, double c0,
                  double baseMVA, double scale, int max_auglag, double xi_max, int *work) {
    branch::PowTable T;
    make_pow_table(T, xi_max);
    const genramp::Problem P = { param[0], param[1], param[2], param[3], param[4], param[5], c2, c1, c0, baseMVA, scale };
    double xx[3] = { x[0], x[1], x[2] }, l[3] = { xl[0], xl[1], xl[2] }, u[3] = { xu[0], xu[1], xu[2] };
    double mu = param[6], xi = param[7];
    int evals = 0, cg = 0, it = 0;
    genramp::solve(P, mu, xi, xx, l, u, max_auglag, xi_max, T, evals, cg, it);
    for (int k = 0; k < 3; ++k) x[k] = xx[k];
    param[6] = mu; param[7] = xi;
    work[0] = it; work[1] = evals; work[2] = cg;
}


// One branch of the QP sub-problem model (qpsub.cuh). H: 6 x 6 row-major; l, rho, v, z: 8; Y: 8; res: 4;
// lin = LH_1h[4], RH_1h, LH_1i[4], RH_1i, LH_1j[2], RH_1j, LH_1k[2], RH_1k; ls / us: 6; sq: sqp_line column (in / out);
// mb: qpsub_membuf column (in / out); u: 8 out; lam: 4 out; work[3]: AL iterations, evaluations, cg iterations.
void hh_solve_qp_branch(const double *H, const double *l, const double *rho, const double *v, const double *z,
                        const double *Y, const double *res, const double *lin, const double *ls, const double *us,
                        double *sq, double *mb, long long major_iter, int max_auglag, double mu_max, double scale,
                        double *u, double *lam, int *work) {
    branch::PowTable T;
    make_pow_table(T, mu_max);
    qpsub::Inputs in;
    for (int i = 0; i < 6; ++i) for (int j = 0; j <= i; ++j) in.H[tron::tri(i, j)] = H[6 * i + j];
    for (int k = 0; k < 8; ++k) { in.lam[k] = l[k]; in.rho[k] = rho[k]; in.xt[k] = v[k] - z[k]; in.Y[k] = Y[k]; }
    for (int k = 0; k < 4; ++k) { in.res[k] = res[k]; in.LH_1h[k] = lin[k]; in.LH_1i[k] = lin[5 + k]; }
    in.RH_1h = lin[4]; in.RH_1i = lin[9];
    in.LH_1j[0] = lin[10]; in.LH_1j[1] = lin[11]; in.RH_1j = lin[12];
    in.LH_1k[0] = lin[13]; in.LH_1k[1] = lin[14]; in.RH_1k = lin[15];
    const double x0[4] = { sq[2], sq[3], sq[4], sq[5] }, xl[4] = { ls[2], ls[3], ls[4], ls[5] }, xu[4] = { us[2], us[3], us[4], us[5] };
    double lam_j = mb[2], lam_k = mb[3], mu = (major_iter == 1) ? 10.0 : mb[4];
    qpsub::ArrStore st;
    qpsub::Result R;
    qpsub::solve(in, st, x0, xl, xu, lam_j, lam_k, mu, max_auglag, mu_max, scale, T, R);
    for (int k = 0; k < 8; ++k) u[k] = R.u[k];
    for (int k = 0; k < 6; ++k) sq[k] = R.sqp[k];
    for (int k = 0; k < 4; ++k) lam[k] = R.lambda[k];
    mb[2] = lam_j; mb[3] = lam_k; mb[4] = mu;
    work[0] = R.it; work[1] = R.evals; work[2] = R.cg;
}

}
