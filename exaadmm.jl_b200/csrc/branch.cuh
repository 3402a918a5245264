// branch.cuh — the per-branch sub-problem of the ADMM x-update:
//   min over x = (vi, vj, ti, tj, s_ij, s_ji) in a box of the scaled augmented
//   Lagrangian of the branch flow consensus terms + line-limit constraints,
//   wrapped in the outer augmented-Lagrangian loop on the two line limits.
//
// Replaces /root/reference/src/models/acopf/acopf_auglag_linelimit_kernel_gpu.jl:1-151
// (AL loop), acopf_tron_linelimit_kernel.jl:4-149 (TRON driver) and
// acopf_eval_linelimit_kernel_gpu.jl:1-594 (f, grad, Hessian).
//
// One thread owns one branch. The objective uses the structure every flow
// shares, F = a vi^2 + b vj^2 + vi vj P(t), t = ti - tj:
// gradient and Hessian are assembled in the 3 variables (vi, vj, t) from
// aggregated flow weights and then expanded to the 6x6 packed matrix, one
// sincos per evaluation, f/grad/Hessian fused. The AL loop and the TRON
// reverse-communication loop are flattened into ONE loop whose body is
//   [compute trial step]  ->  [evaluate f,g,H]  ->  [judge / converge / AL update]
// so lanes of a warp that are in different TRON iterations or different AL
// iterations still run the same instructions.
#pragma once
#include "tron.cuh"

namespace branch {

constexpr int N = 6;
using Sym6 = tron::Sym<N>;

struct Data {          // per-branch inputs, reference rows in comments (membuf rows 1-24, acopf_auglag..gpu.jl:50-73)
    double lam[8];     // rows 1-8   lambda   (pij,qij,pji,qji,wi,wj,ti,tj)
    double rho[8];     // rows 9-16  rho
    double xt[8];      // rows 17-24 xbar - z
    double Y[8];       // YffR,YffI,YftR,YftI,YttR,YttI,YtfR,YtfI
};

struct PowTable {      // host-computed (glibc) 1/mu^0.1 and mu^0.9 for the mu sequence 10, 100, ... <= mu_max
    int n;
    double mu[24], inv_p01[24], p09[24];
};

struct Work { int auglag = 0, evals = 0, cg = 0, shifts = 0, rejected = 0, hit_max = 0; };

EA_DEV void mu_powers(const PowTable &T, double mu, double &inv_p01, double &p09) {
#pragma unroll 1
    for (int k = 0; k < T.n; ++k)
        if (T.mu[k] == mu) { inv_p01 = T.inv_p01[k]; p09 = T.p09[k]; return; }
    inv_p01 = 1.0 / pow(mu, 0.1);
    p09 = pow(mu, 0.9);
}

// flows at x: F[0..3] = pij, qij, pji, qji (acopf_eval_linelimit_kernel_gpu.jl:17-22)
EA_DEV void flows(const double (&x)[N], const double (&Y)[8], double (&F)[4]) {
    double s, c;
    sincos(x[2] - x[3], &s, &c);
    const double vv = x[0] * x[1], vi2 = x[0] * x[0], vj2 = x[1] * x[1];
    F[0] = Y[0] * vi2 + vv * (Y[2] * c + Y[3] * s);
    F[1] = -Y[1] * vi2 + vv * (-Y[3] * c + Y[2] * s);
    F[2] = Y[4] * vj2 + vv * (Y[6] * c - Y[7] * s);
    F[3] = -Y[5] * vj2 + vv * (-Y[7] * c - Y[6] * s);
}

// Fused f, grad f, Hessian (packed lower) of the scaled branch AL objective, and the four flows.
EA_DEV void eval_fgh(const Data &D, const double (&ls)[2], double mu, double scale,
                                         const double (&x)[N], double &f, double (&g)[N], Sym6 &A, double (&F)[4]) {
    const double vi = x[0], vj = x[1];
    double s, c;
    sincos(x[2] - x[3], &s, &c);
    const double vv = vi * vj, vi2 = vi * vi, vj2 = vj * vj;
    // per-flow a, b, gamma, delta
    const double a[4] = { D.Y[0], -D.Y[1], 0.0, 0.0 };
    const double b[4] = { 0.0, 0.0, D.Y[4], -D.Y[5] };
    const double ga[4] = { D.Y[2], -D.Y[3], D.Y[6], -D.Y[7] };
    const double de[4] = { D.Y[3], D.Y[2], -D.Y[7], -D.Y[6] };
    double P[4], Q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        P[k] = ga[k] * c + de[k] * s;
        Q[k] = de[k] * c - ga[k] * s;
        F[k] = a[k] * vi2 + b[k] * vj2 + vv * P[k];
    }
    const double c1 = F[0] * F[0] + F[1] * F[1] + x[4];
    const double c2 = F[2] * F[2] + F[3] * F[3] + x[5];
    const double m[2] = { ls[0] + mu * c1, ls[1] + mu * c2 };

    // objective
    double fv = 0.0;
    {
        const double h[8] = { F[0], F[1], F[2], F[3], vi2, vj2, x[2], x[3] };
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double d = h[k] - D.xt[k];
            fv += D.lam[k] * h[k] + 0.5 * (D.rho[k] * (d * d));
        }
        fv += ls[0] * c1 + ls[1] * c2 + 0.5 * (mu * (c1 * c1)) + 0.5 * (mu * (c2 * c2));
    }
    f = scale * fv;

    // reduced (vi, vj, t) gradient / Hessian
    double As = 0.0, Bs = 0.0, Ps = 0.0, Qs = 0.0;        // sum_k w_k * (a,b,P,Q)_k
    double H00 = 0.0, H01 = 0.0, H02 = 0.0, H11 = 0.0, H12 = 0.0, H22 = 0.0;
    double d[2][3] = { { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 } };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = k >> 1;
        const double G0 = 2.0 * a[k] * vi + vj * P[k];
        const double G1 = 2.0 * b[k] * vj + vi * P[k];
        const double G2 = vv * Q[k];
        const double r = D.lam[k] + D.rho[k] * (F[k] - D.xt[k]);
        const double w = r + 2.0 * m[j] * F[k];
        const double kap = D.rho[k] + 2.0 * m[j];
        As += w * a[k]; Bs += w * b[k]; Ps += w * P[k]; Qs += w * Q[k];
        const double tF = 2.0 * F[k];
        d[j][0] += tF * G0; d[j][1] += tF * G1; d[j][2] += tF * G2;
        const double k0 = kap * G0, k1 = kap * G1, k2 = kap * G2;
        H00 += k0 * G0; H01 += k0 * G1; H02 += k0 * G2;
        H11 += k1 * G1; H12 += k1 * G2; H22 += k2 * G2;
    }
    H00 += 2.0 * As; H11 += 2.0 * Bs; H01 += Ps;
    H02 += vj * Qs;  H12 += vi * Qs;  H22 -= vv * Ps;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const double m0 = mu * d[j][0], m1 = mu * d[j][1], m2 = mu * d[j][2];
        H00 += m0 * d[j][0]; H01 += m0 * d[j][1]; H02 += m0 * d[j][2];
        H11 += m1 * d[j][1]; H12 += m1 * d[j][2]; H22 += m2 * d[j][2];
    }
    const double ri = D.lam[4] + D.rho[4] * (vi2 - D.xt[4]);
    const double rj = D.lam[5] + D.rho[5] * (vj2 - D.xt[5]);
    const double gy0 = 2.0 * As * vi + vj * Ps + 2.0 * vi * ri;
    const double gy1 = 2.0 * Bs * vj + vi * Ps + 2.0 * vj * rj;
    const double gy2 = vv * Qs;
    H00 += 2.0 * ri + 4.0 * D.rho[4] * vi2;
    H11 += 2.0 * rj + 4.0 * D.rho[5] * vj2;

    g[0] = scale * gy0;
    g[1] = scale * gy1;
    g[2] = scale * (gy2 + D.lam[6] + D.rho[6] * (x[2] - D.xt[6]));
    g[3] = scale * (-gy2 + D.lam[7] + D.rho[7] * (x[3] - D.xt[7]));
    g[4] = scale * m[0];
    g[5] = scale * m[1];

    using tron::tri;
    const double smu = scale * mu;
    A.a[tri(0, 0)] = scale * H00;
    A.a[tri(1, 0)] = scale * H01;
    A.a[tri(1, 1)] = scale * H11;
    A.a[tri(2, 0)] = scale * H02;
    A.a[tri(2, 1)] = scale * H12;
    A.a[tri(2, 2)] = scale * (H22 + D.rho[6]);
    A.a[tri(3, 0)] = -(scale * H02);
    A.a[tri(3, 1)] = -(scale * H12);
    A.a[tri(3, 2)] = -(scale * H22);
    A.a[tri(3, 3)] = scale * (H22 + D.rho[7]);
    A.a[tri(4, 0)] = smu * d[0][0];
    A.a[tri(4, 1)] = smu * d[0][1];
    A.a[tri(4, 2)] = smu * d[0][2];
    A.a[tri(4, 3)] = -(smu * d[0][2]);
    A.a[tri(4, 4)] = smu;
    A.a[tri(5, 0)] = smu * d[1][0];
    A.a[tri(5, 1)] = smu * d[1][1];
    A.a[tri(5, 2)] = smu * d[1][2];
    A.a[tri(5, 3)] = -(smu * d[1][2]);
    A.a[tri(5, 4)] = 0.0;
    A.a[tri(5, 5)] = smu;
}

// The branch objective with its augmented-Lagrangian state (membuf rows 25-27).
struct Objective {
    const Data &D;
    double ls[2];      // line-limit multipliers
    double mu;         // AL penalty
    double scale;
    EA_DEV void eval(const double (&x)[N], double &f, double (&g)[N], Sym6 &A, double (&F)[4]) const {
        eval_fgh(D, ls, mu, scale, x, f, g, A, F);
    }
};

// Solve one branch sub-problem: AL loop on the two line limits around TRON, flattened
// into one loop. x: start point in, solution out. obj.ls / obj.mu are updated in place.
// Returns the flows at the solution in Fout. `Obj` only needs eval(), ls[2] and mu, so
// the test harness can run this same loop on another evaluator.
template <class Obj>
EA_DEV void solve(Obj &obj, const double (&xl)[N], const double (&xu)[N], double (&x)[N],
                  int max_auglag, double mu_max, const PowTable &T, double (&Fout)[4], Work &wk) {
    double (&ls)[2] = obj.ls;
    double &mu = obj.mu;
    const int max_feval = 500, max_minor = 200;     // call site acopf_auglag_linelimit_kernel_gpu.jl:94
    const double gtol = 1e-6;
    double inv_p01, p09;
    mu_powers(T, mu, inv_p01, p09);
    double eta = inv_p01;                            // eta = 1/mu^0.1 (:84)

    // TRON state. g and A always belong to the last evaluated point; a rejected
    // step (rare) re-evaluates them at x_c (phase RESTORE) instead of keeping a copy.
    double f = 0.0, fc = 0.0, delta = 0.0, alphac = 1.0, prered = 0.0, g0 = 0.0, snorm = 0.0;
    double g[N], xc[N];
    Sym6 A;
    double Fc[4] = { 0.0, 0.0, 0.0, 0.0 };           // flows at the current accepted point
    int nfev = 0, minor = 0, iter = 1, it_al = 0;
    enum { START = 0, TRIAL = 1, RESTORE = 2 };
    int phase = START;
    bool step_pending = false;
    tron::Stats st;
#pragma unroll
    for (int i = 0; i < N; ++i) { g[i] = 0.0; xc[i] = x[i]; }

#pragma unroll 1
    for (;;) {
        if (step_pending) {
            // dtron COMPUTE: Cauchy point + projected CG -> trial point in x
            fc = f;
#pragma unroll
            for (int i = 0; i < N; ++i) xc[i] = x[i];
            tron::compute_step<N>(x, xl, xu, A, g, delta, alphac, prered, g0, snorm, st);
            phase = TRIAL;
            step_pending = false;
        }

        double fn, Fn[4];
        obj.eval(x, fn, g, A, Fn);
        if (phase != RESTORE) wk.evals++;              // counted like the reference's f-evaluations
        bool tron_done = false;
        if (phase == TRIAL) {
            nfev++;
            if (nfev >= max_feval) {
                tron_done = true;                    // driver stops, trial point kept (tron_kernel.jl:72-75)
#pragma unroll
                for (int k = 0; k < 4; ++k) Fc[k] = Fn[k];
            } else {
                bool accepted;
                const int task = tron::judge_step(fn, fc, g0, snorm, prered, iter == 1, delta, accepted);
                if (accepted) {
                    iter++;
                    f = fn;
#pragma unroll
                    for (int k = 0; k < 4; ++k) Fc[k] = Fn[k];
                    if (task == 2) tron_done = true;
                    else {
                        minor++;                      // the reference evaluates g,H here (task GH)
                        if (tron::gpnorm<N>(x, xl, xu, g) <= gtol) tron_done = true;          // NEWX test (:121-130)
                        else if (minor >= max_minor) tron_done = true;
                        else step_pending = true;
                    }
                } else {
                    wk.rejected++;
#pragma unroll
                    for (int i = 0; i < N; ++i) x[i] = xc[i];
                    f = fc;
                    if (task == 2) tron_done = true;   // Fc still holds the flows at xc
                    else phase = RESTORE;              // g, A must be re-evaluated at xc before the next step
                }
            }
        } else {
            f = fn;
#pragma unroll
            for (int k = 0; k < 4; ++k) Fc[k] = Fn[k];
            if (phase == START) {                      // task 0: a fresh TRON solve
                nfev = 1; minor = 1; iter = 1; alphac = 1.0;
                delta = tron::nrm2<N>(g);              // tron_kernel.jl:102-105
            }
            step_pending = true;
        }

        if (tron_done) {
            // augmented-Lagrangian update on the line limits (auglag_gpu.jl:96-131)
            it_al++;
            const double cviol1 = Fc[0] * Fc[0] + Fc[1] * Fc[1] + x[4];
            const double cviol2 = Fc[2] * Fc[2] + Fc[3] * Fc[3] + x[5];
            const double cnorm = fmax(fabs(cviol1), fabs(cviol2));
            bool terminate = false;
            if (cnorm <= eta) {
                if (cnorm <= 1e-6) terminate = true;
                else {
                    ls[0] += mu * cviol1;
                    ls[1] += mu * cviol2;
                    eta = eta / p09;
                }
            } else {
                mu = fmin(mu_max, mu * 10.0);
                mu_powers(T, mu, inv_p01, p09);
                eta = inv_p01;
            }
            if (it_al >= max_auglag) { if (!terminate) wk.hit_max = 1; terminate = true; }
            if (terminate) break;
            phase = START;
            step_pending = false;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) Fout[k] = Fc[k];
    wk.auglag = it_al;
    wk.cg = st.cg;
    wk.shifts = st.shifts;
}

}  // namespace branch
