// genramp.cuh — the generator sub-problem of the multi-period model: for a generator in
// period t >= 2 the x-update is a bound-constrained problem in
//     x = (p_t, phat_{t-1}, s_t),   pgmin <= p_t, phat_{t-1} <= pgmax,  |s_t| <= ramp_rate,
// with the ramp equality p_t - phat_{t-1} - s_t = 0 handled by an augmented Lagrangian
// around TRON, exactly like the line limits of a branch.
//
// Replaces auglag_generator_kernel + tron_generator_kernel + eval_*_generator_kernel
// (src/models/mpacopf/mpacopf_auglag_generator_kernel_gpu.jl:1-116,
//  mpacopf_tron_generator_kernel.jl:1-122, mpacopf_eval_generator_kernel_gpu.jl:1-73).
// One generator per LANE (the reference uses a 32-thread block per generator); the TRON
// routines are the N-templated ones of tron.cuh. The objective is a convex quadratic, so a
// solve takes one or two Newton steps; no flattened state machine is needed here.
#pragma once
#include "branch.cuh"
#include "tron.cuh"

namespace genramp {

constexpr int N = 3;
using Sym3 = tron::Sym<N>;

struct Problem {             // gen_membuf rows 1-6 (auglag_generator_kernel_gpu.jl:23-33) + cost data
    double lam_p, lam_ph;    // lambda of p_t and of phat_{t-1}
    double rho_p, rho_ph;
    double xt_p, xt_ph;      // pbar_t - z_p,  pbar_{t-1} - z_phat
    double c2, c1, c0, baseMVA, scale;
};

// eval_f_generator_kernel (mpacopf_eval_generator_kernel_gpu.jl:1-24); mu = multiplier, xi = penalty
EA_DEV double eval_f(const Problem &P, double mu, double xi, const double (&x)[N]) {
    const double pb = x[0] * P.baseMVA;
    const double dp = x[0] - P.xt_p, dh = x[1] - P.xt_ph, c = x[0] - x[1] - x[2];
    double f = 0.0;
    f += P.c2 * (pb * pb) + P.c1 * pb + P.c0;
    f += P.lam_p * dp + (0.5 * P.rho_p) * (dp * dp);
    f += P.lam_ph * dh + (0.5 * P.rho_ph) * (dh * dh);
    f += mu * c + (0.5 * xi) * (c * c);
    return f * P.scale;
}

// gradient (:26-43) and the constant Hessian (:45-73), packed lower triangle
EA_DEV void eval_gh(const Problem &P, double mu, double xi, const double (&x)[N], double (&g)[N], Sym3 &A) {
    const double B = P.baseMVA, s = P.scale;
    const double c = x[0] - x[1] - x[2];
    const double al = mu + xi * c;
    g[0] = (2 * P.c2 * (B * B) * x[0] + P.c1 * B + (P.lam_p + P.rho_p * (x[0] - P.xt_p)) + al) * s;
    g[1] = ((P.lam_ph + P.rho_ph * (x[1] - P.xt_ph)) + -al) * s;
    g[2] = -al * s;
    A.a[tron::tri(0, 0)] = s * (2 * P.c2 * (B * B) + P.rho_p + xi);
    A.a[tron::tri(1, 0)] = s * (-xi);
    A.a[tron::tri(2, 0)] = s * (-xi);
    A.a[tron::tri(1, 1)] = s * (P.rho_ph + xi);
    A.a[tron::tri(2, 1)] = s * (xi);
    A.a[tron::tri(2, 2)] = s * (xi);
}

// The TRON driver (mpacopf_tron_generator_kernel.jl:44-121 around ExaTron.dtron), one lane.
EA_DEV void tron_solve(const Problem &P, double mu, double xi, double (&x)[N], const double (&xl)[N],
                       const double (&xu)[N], int &evals, int &cg) {
    const int max_feval = 500, max_minor = 200;      // call site auglag_generator_kernel_gpu.jl:77
    const double gtol = 1e-6;
    double g[N];
    Sym3 A;
    double f = eval_f(P, mu, xi, x);
    eval_gh(P, mu, xi, x, g, A);
    int nfev = 1, minor = 1, iter = 1;
    evals++;
    double delta = tron::nrm2<N>(g), alphac = 1.0;
#pragma unroll 1
    for (;;) {
        int task;
#pragma unroll 1
        do {
            const double fc = f;
            double xc[N];
#pragma unroll
            for (int i = 0; i < N; ++i) xc[i] = x[i];
            double prered, g0, snorm;
            tron::Stats st;
            tron::compute_step_auto<N>(x, xl, xu, A, g, delta, alphac, prered, g0, snorm, st, evals);
            cg += st.cg;
            const double fn = eval_f(P, mu, xi, x);
            nfev++; evals++;
            if (nfev >= max_feval) return;
            bool accepted;
            task = tron::judge_step(fn, fc, g0, snorm, prered, iter == 1, delta, accepted);
            if (accepted) { iter++; f = fn; }
            else {
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = xc[i];
                f = fc;
            }
        } while (task == 0);
        if (task == 2) return;
        eval_gh(P, mu, xi, x, g, A);
        minor++;
        if (tron::gpnorm<N>(x, xl, xu, g) <= gtol) return;
        if (minor >= max_minor) return;
    }
}

// The augmented-Lagrangian loop (auglag_generator_kernel_gpu.jl:66-109). mu (multiplier) and xi
// (penalty) are the persistent gen_membuf rows 7-8; xi restarts at 10 on the first inner
// iteration of an outer iteration (the caller passes it in).
EA_DEV void solve(const Problem &P, double &mu, double &xi, double (&x)[N], const double (&xl)[N],
                  const double (&xu)[N], int max_auglag, double xi_max, const branch::PowTable &T,
                  int &evals, int &cg, int &it_out) {
    double inv_p01, p09;
    branch::mu_powers(T, xi, inv_p01, p09);
    double eta = inv_p01;
    int it = 0;
    bool terminate = false;
#pragma unroll 1
    while (!terminate) {
        it++;
        tron_solve(P, mu, xi, x, xl, xu, evals, cg);
        const double cviol = x[0] - x[1] - x[2];
        const double cnorm = fabs(cviol);
        if (cnorm <= eta) {
            if (cnorm <= 1e-6) terminate = true;
            else {
                mu += xi * cviol;
                eta = eta / p09;
            }
        } else {
            xi = tron::dmin(xi_max, xi * 10.0);
            branch::mu_powers(T, xi, inv_p01, p09);
            eta = inv_p01;
        }
        if (it >= max_auglag) terminate = true;
    }
    it_out = it;
}

}  // namespace genramp
