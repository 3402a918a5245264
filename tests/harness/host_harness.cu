// host_harness.cu — TEST TOOL. Compiles the product's device routines
// (exaadmm.jl_b200/csrc/tron.cuh, branch.cuh) for the HOST so that the CPU test
// suite can drive exactly the code the GPU runs (flattened AL/TRON loop, masked
// free-set) against the oracle without a GPU. Not part of the product library.
#include "../../exaadmm.jl_b200/csrc/branch.cuh"
#include <cmath>
#include <cstring>

extern "C" {

// param: 31 doubles (membuf column, 0-based rows); rows 24-26 (lambda_s, mu) updated in place.
// x: 6 doubles in/out. Y: 8. xl/xu: 6. work[6]: auglag, evals, cg, shifts, rejected, hit_max.
void hh_solve_branch(double *x, const double *xl, const double *xu, double *param, const double *Y,
                     long long major_iter, int max_auglag, double mu_max, double scale, double *F, int *work) {
    branch::Data D;
    for (int k = 0; k < 8; ++k) { D.lam[k] = param[k]; D.rho[k] = param[8 + k]; D.xt[k] = param[16 + k]; D.Y[k] = Y[k]; }
    double ls[2] = { param[24], param[25] };
    double mu = (major_iter == 1) ? 10.0 : param[26];
    branch::PowTable T; T.n = 0;
    double m = 10.0;
    for (int k = 0; k < 24; ++k) {
        T.mu[k] = m; T.inv_p01[k] = 1.0 / std::pow(m, 0.1); T.p09[k] = std::pow(m, 0.9); T.n = k + 1;
        double nx = std::fmin(mu_max, m * 10); if (nx == m) break; m = nx;
    }
    double xx[6], l[6], u[6], FF[4];
    for (int k = 0; k < 6; ++k) { xx[k] = x[k]; l[k] = xl[k]; u[k] = xu[k]; }
    branch::Work wk;
    branch::Objective obj{ D, { ls[0], ls[1] }, mu, scale };
    branch::solve(obj, l, u, xx, max_auglag, mu_max, T, FF, wk);
    for (int k = 0; k < 6; ++k) x[k] = xx[k];
    for (int k = 0; k < 4; ++k) F[k] = FF[k];
    param[24] = obj.ls[0]; param[25] = obj.ls[1]; param[26] = obj.mu;
    work[0] = wk.auglag; work[1] = wk.evals; work[2] = wk.cg; work[3] = wk.shifts; work[4] = wk.rejected; work[5] = wk.hit_max;
}

// The same flattened AL/TRON loop, but with the ORACLE's f / grad / Hessian plugged in
// (function pointers into oracle/_build/libacopf_oracle.so). Built with -DEA_NO_FMA the
// arithmetic of the TRON routines is then bit-identical to the oracle's restatement, so
// iterates and evaluation counts must match EXACTLY - any difference is a logic error.
typedef double (*orc_f_t)(const double *, const double *, const double *, double);
typedef void (*orc_gh_t)(const double *, const double *, const double *, double, double *, double *);

struct OracleObjective {
    orc_f_t fn; orc_gh_t ghn; double *param; const double *Y; double scale;
    double ls[2]; double mu;
    void eval(const double (&x)[6], double &f, double (&g)[6], branch::Sym6 &A, double (&F)[4]) {
        param[24] = ls[0]; param[25] = ls[1]; param[26] = mu;
        double H[36];
        f = fn(x, param, Y, scale);
        ghn(x, param, Y, scale, g, H);
        for (int a = 0; a < 6; ++a) for (int b = 0; b <= a; ++b) A.a[tron::tri(a, b)] = H[6 * a + b];
        const double cc = x[0] * x[1] * cos(x[2] - x[3]), ss = x[0] * x[1] * sin(x[2] - x[3]);
        F[0] = Y[0] * (x[0] * x[0]) + Y[2] * cc + Y[3] * ss;        // as the oracle's AL loop computes them
        F[1] = -Y[1] * (x[0] * x[0]) - Y[3] * cc + Y[2] * ss;
        F[2] = Y[4] * (x[1] * x[1]) + Y[6] * cc - Y[7] * ss;
        F[3] = -Y[5] * (x[1] * x[1]) - Y[7] * cc - Y[6] * ss;
    }
};

void hh_solve_branch_oracle_eval(void *fn, void *ghn, double *x, const double *xl, const double *xu, double *param,
                                 const double *Y, long long major_iter, int max_auglag, double mu_max, double scale,
                                 int *work) {
    branch::PowTable T; T.n = 0;
    double m = 10.0;
    for (int k = 0; k < 24; ++k) {
        T.mu[k] = m; T.inv_p01[k] = 1.0 / std::pow(m, 0.1); T.p09[k] = std::pow(m, 0.9); T.n = k + 1;
        double nx = std::fmin(mu_max, m * 10); if (nx == m) break; m = nx;
    }
    OracleObjective obj{ (orc_f_t)fn, (orc_gh_t)ghn, param, Y, scale, { param[24], param[25] },
                         (major_iter == 1) ? 10.0 : param[26] };
    double xx[6], l[6], u[6], FF[4];
    for (int k = 0; k < 6; ++k) { xx[k] = x[k]; l[k] = xl[k]; u[k] = xu[k]; }
    branch::Work wk;
    branch::solve(obj, l, u, xx, max_auglag, mu_max, T, FF, wk);
    for (int k = 0; k < 6; ++k) x[k] = xx[k];
    param[24] = obj.ls[0]; param[25] = obj.ls[1]; param[26] = obj.mu;
    work[0] = wk.auglag; work[1] = wk.evals; work[2] = wk.cg; work[3] = wk.shifts; work[4] = wk.rejected; work[5] = wk.hit_max;
}

void hh_eval(const double *x, const double *param, const double *Y, double scale, double *f, double *g, double *H) {
    branch::Data D;
    for (int k = 0; k < 8; ++k) { D.lam[k] = param[k]; D.rho[k] = param[8 + k]; D.xt[k] = param[16 + k]; D.Y[k] = Y[k]; }
    double ls[2] = { param[24], param[25] }, xx[6], gg[6], F[4];
    for (int k = 0; k < 6; ++k) xx[k] = x[k];
    branch::Sym6 A;
    branch::eval_fgh(D, ls, param[26], scale, xx, *f, gg, A, F);
    for (int a = 0; a < 6; ++a) { g[a] = gg[a]; for (int b = 0; b < 6; ++b) H[6 * a + b] = A.a[tron::tri(a, b)]; }
}

}
