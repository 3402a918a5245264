// qpsub.cuh — the branch sub-problem of the one-level ADMM on the SQP sub-problem (`ModelQpsub`).
//
// Replaces auglag_linelimit_qpsub + eval_A_b_branch_kernel_gpu_qpsub + eval_A_b_auglag_branch_kernel_gpu_qpsub_red +
// tron_gpu_test (src/models/qpsub/qpsub_auglag_Ab_linelimit_kernel_red_gpu.jl:1-281,
// qpsub_eval_Ab_linelimit_kernel_gpu.jl:1-180, qpsub_tron_linelimit_kernel.jl:1-170).
//
// Per branch the reference minimises, by an augmented-Lagrangian loop around TRON, a box-constrained QP in
//     x = (t_ij, t_ji, w_i, w_j, theta_i, theta_j)
// obtained from the 8-variable problem in (t_ij, t_ji, w_ijR, w_ijI, w_i, w_j, theta_i, theta_j) by eliminating
// (w_ijR, w_ijI) through the two linearised equalities 1h / 1i:  x8 = C x + d.  It forms the 8 x 8 matrix, C and the
// products C' A8 C, C'(A8 d + b8) again in every AL iteration. Here the algebra is done once per branch, using what the
// matrices look like: with y = (w_ijR, w_ijI, w_i, w_j, theta_i, theta_j) = E x' + e, x' = x[2..5],
//     A(mu)  = scale * ( [0 0; 0 M] + mu (c_j c_j' + c_k c_k') ),   M = E' Hbr E,  c_j = (1, 0, g_j),  c_k = (0, 1, g_k)
//     b(mu)  = scale * ( (0, 0, q) + (lambda_j + mu d_j) c_j + (lambda_k + mu d_k) c_k ),   q = E'(Hbr e + bbr)
// and the violation of the line limits is c_j.x + d_j, c_k.x + d_k. Only mu, lambda_j, lambda_k change in the AL loop.
// One branch per LANE (the reference uses a 32-thread block per branch); the TRON routines are the N-templated ones
// of tron.cuh at N = 6. The per-branch constants (Static, 47 doubles) live in a shared-memory column per lane.
#pragma once
#include "branch.cuh"
#include "tron.cuh"

namespace qpsub {

constexpr int N = 6;
using Sym6 = tron::Sym<N>;

struct Inputs {                 // one branch, reference order
    double H[21];               // Hs block of the branch, packed lower triangle (tron::tri)
    double lam[8], rho[8], xt[8];   // lambda, rho, xbar - z   (pij,qij,pji,qji,wi,wj,ti,tj)
    double Y[8];                // YffR,YffI,YftR,YftI,YttR,YttI,YtfR,YtfI
    double res[4];              // line_res
    double LH_1h[4], RH_1h, LH_1i[4], RH_1i, LH_1j[2], RH_1j, LH_1k[2], RH_1k;
};

// layout of the per-branch constants
enum { S_M = 0, S_Q = 10, S_GJ = 14, S_GK = 18, S_DJ = 22, S_DK = 23, S_CR = 24, S_CI = 28, S_DR = 32, S_DI = 33,
       S_H0 = 34 /* Hbr[0][0..5] */, S_H1 = 40 /* Hbr[1][1..5] */, S_B0 = 45, S_B1 = 46, S_ROWS = 47 };

struct ArrStore {               // host harness / tests
    double a[S_ROWS];
    EA_DEV double &operator[](int k) { return a[k]; }
    EA_DEV double operator[](int k) const { return a[k]; }
};
template <int STRIDE> struct TileStore {      // kernel: element k of this lane at base[k * STRIDE]
    double *base;
    EA_DEV double &operator[](int k) const { return base[k * STRIDE]; }
};

// rows of supY over (w_ijR, w_ijI, w_i, w_j): p_ij, q_ij, p_ji, q_ji (qpsub_eval_Ab_linelimit_kernel_gpu.jl:33-64)
EA_DEV void sup_rows(const double (&Y)[8], double (&S)[4][4]) {
    S[0][0] = Y[2];  S[0][1] = Y[3];  S[0][2] = Y[0];  S[0][3] = 0.0;
    S[1][0] = -Y[3]; S[1][1] = Y[2];  S[1][2] = -Y[1]; S[1][3] = 0.0;
    S[2][0] = Y[6];  S[2][1] = -Y[7]; S[2][2] = 0.0;   S[2][3] = Y[4];
    S[3][0] = -Y[7]; S[3][1] = -Y[6]; S[3][2] = 0.0;   S[3][3] = -Y[5];
}

// Everything that stays fixed during the AL loop of one branch.
template <class Store> EA_DEV void setup(const Inputs &in, Store &st) {
    double S[4][4];
    sup_rows(in.Y, S);
    // Hbr = Hs + sum_r rho_r S_r S_r' + diag(0, 0, rho_wi, rho_wj, rho_ti, rho_tj); bbr likewise (eval_A_b_branch_kernel)
    double Hbr[21], bbr[6];
#pragma unroll
    for (int k = 0; k < 21; ++k) Hbr[k] = in.H[k];
#pragma unroll
    for (int i = 0; i < 6; ++i) bbr[i] = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const double c = in.lam[r] - in.rho[r] * (in.xt[r] - in.res[r]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bbr[i] = EA_FMA(c, S[r][i], bbr[i]);
#pragma unroll
            for (int j = 0; j <= i; ++j) Hbr[tron::tri(i, j)] = EA_FMA(in.rho[r] * S[r][i], S[r][j], Hbr[tron::tri(i, j)]);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        Hbr[tron::tri(2 + k, 2 + k)] += in.rho[4 + k];
        bbr[2 + k] += in.lam[4 + k] - in.rho[4 + k] * in.xt[4 + k];
    }
    // elimination of (w_ijR, w_ijI): inverse of [LH_1h[0] LH_1h[1]; LH_1i[0] LH_1i[1]] (..red_gpu.jl:96-100)
    const double prod = in.LH_1h[0] * in.LH_1i[1] - in.LH_1h[1] * in.LH_1i[0];
    const double i11 = in.LH_1i[1] / prod, i12 = -in.LH_1h[1] / prod, i21 = -in.LH_1i[0] / prod, i22 = in.LH_1h[0] / prod;
    const double cR[4] = { -i11 * in.LH_1h[2], -i11 * in.LH_1h[3], -i12 * in.LH_1i[2], -i12 * in.LH_1i[3] };
    const double cI[4] = { -i21 * in.LH_1h[2], -i21 * in.LH_1h[3], -i22 * in.LH_1i[2], -i22 * in.LH_1i[3] };
    const double dR = i11 * in.RH_1h + i12 * in.RH_1i, dI = i21 * in.RH_1h + i22 * in.RH_1i;
    // t = Hbr e + bbr, e = (dR, dI, 0, 0, 0, 0)
    double t[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) t[a] = EA_FMA(Hbr[tron::tri(a, 0)], dR, EA_FMA(Hbr[tron::tri(a, 1)], dI, bbr[a]));
    const double h00 = Hbr[tron::tri(0, 0)], h01 = Hbr[tron::tri(1, 0)], h11 = Hbr[tron::tri(1, 1)];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        st[S_Q + i] = t[2 + i] + cR[i] * t[0] + cI[i] * t[1];
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double m = Hbr[tron::tri(2 + i, 2 + j)];
            m += cR[i] * Hbr[tron::tri(2 + j, 0)] + cI[i] * Hbr[tron::tri(2 + j, 1)];
            m += Hbr[tron::tri(2 + i, 0)] * cR[j] + Hbr[tron::tri(2 + i, 1)] * cI[j];
            m += cR[i] * cR[j] * h00 + (cR[i] * cI[j] + cI[i] * cR[j]) * h01 + cI[i] * cI[j] * h11;
            st[S_M + tron::tri(i, j)] = m;
        }
    }
    // line-limit rows: v_1j = e_tij + LH_1j[0] S_pij + LH_1j[1] S_qij over (w_ijR, w_ijI, w_i, w_j) (..red_gpu.jl:76-93)
    double aj[4], ak[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        aj[k] = in.LH_1j[0] * S[0][k] + in.LH_1j[1] * S[1][k];
        ak[k] = in.LH_1k[0] * S[2][k] + in.LH_1k[1] * S[3][k];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        st[S_GJ + i] = aj[0] * cR[i] + aj[1] * cI[i] + (i < 2 ? aj[2 + i] : 0.0);
        st[S_GK + i] = ak[0] * cR[i] + ak[1] * cI[i] + (i < 2 ? ak[2 + i] : 0.0);
        st[S_CR + i] = cR[i];
        st[S_CI + i] = cI[i];
    }
    st[S_DJ] = aj[0] * dR + aj[1] * dI - in.RH_1j;
    st[S_DK] = ak[0] * dR + ak[1] * dI - in.RH_1k;
    st[S_DR] = dR; st[S_DI] = dI;
#pragma unroll
    for (int k = 0; k < 6; ++k) st[S_H0 + k] = Hbr[tron::tri(k, 0)];
#pragma unroll
    for (int k = 1; k < 6; ++k) st[S_H1 + k - 1] = Hbr[tron::tri(k, 1)];
    st[S_B0] = bbr[0]; st[S_B1] = bbr[1];
}

struct Qp { Sym6 A; double b[N]; };

// the reduced, scaled QP of one AL iteration (eval_A_b_auglag_branch_kernel_gpu_qpsub_red)
template <class Store> EA_DEV void build_qp(const Store &st, double lam_j, double lam_k, double mu, double scale, Qp &P) {
    const double wj = lam_j + mu * st[S_DJ], wk = lam_k + mu * st[S_DK];
    P.A.a[tron::tri(0, 0)] = scale * mu;
    P.A.a[tron::tri(1, 0)] = 0.0;
    P.A.a[tron::tri(1, 1)] = scale * mu;
    P.b[0] = scale * wj;
    P.b[1] = scale * wk;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double gji = st[S_GJ + i], gki = st[S_GK + i];
        P.A.a[tron::tri(2 + i, 0)] = scale * (mu * gji);
        P.A.a[tron::tri(2 + i, 1)] = scale * (mu * gki);
        P.b[2 + i] = scale * (st[S_Q + i] + wj * gji + wk * gki);
#pragma unroll
        for (int j = 0; j <= i; ++j)
            P.A.a[tron::tri(2 + i, 2 + j)] = scale * (st[S_M + tron::tri(i, j)] + mu * (gji * st[S_GJ + j] + gki * st[S_GK + j]));
    }
}

// eval_f_kernel / eval_g_kernel (qpsub_tron_linelimit_kernel.jl:121-150)
EA_DEV void eval_fg(const Qp &P, const double (&x)[N], double &f, double (&g)[N]) {
    double w[N];
    tron::symv<N>(P.A, x, w);
    f = EA_FMA(0.5, tron::dot<N>(x, w), tron::dot<N>(P.b, x));
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = w[i] + P.b[i];
}

struct Result {
    double u[8];            // pij, qij, pji, qji, wi, wj, ti, tj
    double sqp[6];          // w_ijR, w_ijI, w_i, w_j, theta_i, theta_j
    double lambda[4];       // multipliers of 14h, 14i, 14j, 14k
    int it, evals, cg;
};

// The AL loop (auglag_linelimit_qpsub, ..red_gpu.jl:150-240) around tron_gpu_test (qpsub_tron_linelimit_kernel.jl:8-117,
// ExaTron.dtron), flattened like the ACOPF branch solver (branch.cuh) into a state machine that the lanes of a warp, each
// at a different stage of a different branch, advance in lock step:
//     start_pass   lanes at the start of an AL iteration: the QP of this (lambda_j, lambda_k, mu), f and g at x
//     step_pass    every live lane: one TRON step (dtron COMPUTE), f and g at the trial point, its judgement; when the
//                  TRON solve ends, the AL update - the branch is finished or goes back to start_pass
// The Hessian is constant within an AL iteration, so a rejected step needs no re-evaluation: the next round retries
// with the smaller trust region. The kernel refills finished lanes from a work queue between the rounds; the host
// harness drives the same functions for one branch (solve).
enum Phase : int { NEED = 0, START = 1, RUN = 2, DONE = 3 };

struct Lane {
    double x[N], g[N];          // g: the last gradient the driver accepted (`trg`: read back for the multipliers)
    Qp P;
    double f, delta, alphac;
    double lam_j, lam_k, mu;    // qpsub_membuf rows 3-5
    double eta, inv_p01, p09;
    int nfev, minor, iter, it, evals, cg;
    int phase;
};

// x0 = sqp_line[2..5] of the previous call; mu restarts at 10 when info.inner == 1 (the caller passes it in).
template <class Store>
EA_DEV void begin(Lane &L, const Inputs &in, Store &st, const double (&x0)[4], double lam_j, double lam_k, double mu,
                  const branch::PowTable &T) {
    setup(in, st);
    L.x[0] = 0.0; L.x[1] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) L.x[2 + i] = x0[i];
#pragma unroll
    for (int i = 0; i < N; ++i) L.g[i] = 0.0;
    L.lam_j = lam_j; L.lam_k = lam_k; L.mu = mu;
    branch::mu_powers(T, mu, L.inv_p01, L.p09);
    L.eta = L.inv_p01;
    L.it = 0; L.evals = 0; L.cg = 0;
    L.f = 0.0; L.delta = 0.0; L.alphac = 1.0; L.nfev = 0; L.minor = 0; L.iter = 1;
    L.phase = START;
}

template <class Store>
EA_DEV void start_pass(Lane &L, const Store &st, double scale) {
    if (L.phase != START) return;
    L.it++;
    build_qp(st, L.lam_j, L.lam_k, L.mu, scale, L.P);
    eval_fg(L.P, L.x, L.f, L.g);
    L.nfev = 1; L.minor = 1; L.iter = 1;
    L.evals++;
    L.delta = tron::nrm2<N>(L.g);
    L.alphac = 1.0;
    L.phase = RUN;
}

// Returns true when the branch is finished (x, g, lam_j, lam_k, mu hold the result: finish()).
template <class Store>
EA_DEV bool step_pass(Lane &L, const Store &st, const double (&xl)[N], const double (&xu)[N], int max_auglag,
                      double mu_max, const branch::PowTable &T) {
    const int max_feval = 500, max_minor = 200;
    const double gtol = 1e-6;
    if (L.phase != RUN) return false;
    const double fc = L.f;
    double xc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) xc[i] = L.x[i];
    double prered, g0, snorm;
    tron::Stats s;
    tron::compute_step_auto<N>(L.x, xl, xu, L.P.A, L.g, L.delta, L.alphac, prered, g0, snorm, s, L.evals);
    L.cg += s.cg;
    double fn, gn[N];
    eval_fg(L.P, L.x, fn, gn);
    L.nfev++; L.evals++;
    bool tron_done = false;
    if (L.nfev >= max_feval) tron_done = true;       // the driver stops with the trial point and the old gradient
    else {
        bool accepted;
        const int task = tron::judge_step(fn, fc, g0, snorm, prered, L.iter == 1, L.delta, accepted);
        if (accepted) { L.iter++; L.f = fn; }
        else {
#pragma unroll
            for (int i = 0; i < N; ++i) L.x[i] = xc[i];
            L.f = fc;
        }
        if (task == 0) return false;                  // rejected: the next round retries with the smaller trust region
        if (task == 2) tron_done = true;              // converged on the function-value test: g stays the old gradient
        else {
#pragma unroll
            for (int i = 0; i < N; ++i) L.g[i] = gn[i];
            L.minor++;
            if (tron::gpnorm<N>(L.x, xl, xu, L.g) <= gtol) tron_done = true;
            else if (L.minor >= max_minor) tron_done = true;
        }
    }
    if (!tron_done) return false;
    // AL update on the two linearised line limits
    double c3 = L.x[0] + st[S_DJ], c4 = L.x[1] + st[S_DK];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c3 = EA_FMA(st[S_GJ + i], L.x[2 + i], c3); c4 = EA_FMA(st[S_GK + i], L.x[2 + i], c4); }
    const double cnorm = tron::dmax(fabs(c3), fabs(c4));
    bool terminate = false;
    if (cnorm <= L.eta) {
        if (cnorm <= 1e-6) terminate = true;
        else { L.lam_j += L.mu * c3; L.lam_k += L.mu * c4; L.eta = L.eta / L.p09; }
    } else {
        L.mu = tron::dmin(mu_max, L.mu * 10.0);
        branch::mu_powers(T, L.mu, L.inv_p01, L.p09);
        L.eta = L.inv_p01;
    }
    if (L.it >= max_auglag && cnorm > 1e-6) terminate = true;
    if (terminate) { L.phase = DONE; return true; }
    L.phase = START;
    return false;
}

// What the branch hands back: u = supY (C x + d) + res, sqp_line, and the multipliers of 14h-14k (..red_gpu.jl:256-277).
// Y, res and the leading 2 x 2 blocks of 1h / 1i are re-read by the caller (they are not kept during the solve).
template <class Store>
EA_DEV void finish(const Lane &L, const Store &st, const double (&Y)[8], const double (&res)[4], const double (&LH_1h)[2],
                   const double (&LH_1i)[2], Result &R) {
    const double (&x)[N] = L.x;
    R.it = L.it; R.evals = L.evals; R.cg = L.cg;
    // x8 = C x + d
    double wR = st[S_DR], wI = st[S_DI];
#pragma unroll
    for (int i = 0; i < 4; ++i) { wR = EA_FMA(st[S_CR + i], x[2 + i], wR); wI = EA_FMA(st[S_CI + i], x[2 + i], wI); }
    R.sqp[0] = wR; R.sqp[1] = wI;
#pragma unroll
    for (int i = 0; i < 4; ++i) R.sqp[2 + i] = x[2 + i];
    double S[4][4];
    sup_rows(Y, S);
#pragma unroll
    for (int r = 0; r < 4; ++r)
        R.u[r] = (S[r][0] * wR + S[r][1] * wI + S[r][2] * x[2] + S[r][3] * x[3]) + res[r];
#pragma unroll
    for (int i = 0; i < 4; ++i) R.u[4 + i] = x[2 + i];
    // multipliers handed back to the SQP: tmpH = inv([LH_1h[0] LH_1i[0]; LH_1h[1] LH_1i[1]])
    const double prod = LH_1h[0] * LH_1i[1] - LH_1i[0] * LH_1h[1];
    const double t11 = LH_1i[1] / prod, t12 = -LH_1i[0] / prod, t21 = -LH_1h[1] / prod, t22 = LH_1h[0] / prod;
    const double ti0 = 2 * R.u[0] * Y[2] + 2 * R.u[1] * (-Y[3]), ti1 = 2 * R.u[0] * Y[3] + 2 * R.u[1] * Y[2];
    const double th0 = 2 * R.u[2] * Y[6] + 2 * R.u[3] * (-Y[7]), th1 = 2 * R.u[2] * (-Y[7]) + 2 * R.u[3] * (-Y[6]);
    const double (&trg)[N] = L.g;
    double w0 = trg[0] * ti0 + trg[1] * th0, w1 = trg[0] * ti1 + trg[1] * th1;
#pragma unroll
    for (int k = 0; k < 6; ++k) w0 = EA_FMA(st[S_H0 + k], R.sqp[k], w0);
    w1 = EA_FMA(st[S_H0 + 1], R.sqp[0], w1);
#pragma unroll
    for (int k = 1; k < 6; ++k) w1 = EA_FMA(st[S_H1 + k - 1], R.sqp[k], w1);
    w0 += st[S_B0]; w1 += st[S_B1];
    R.lambda[0] = -(t11 * w0 + t12 * w1);
    R.lambda[1] = -(t21 * w0 + t22 * w1);
    R.lambda[2] = -fabs(trg[0]);
    R.lambda[3] = -fabs(trg[1]);
}

// One branch to completion on one lane (host harness).
template <class Store>
EA_DEV void solve(const Inputs &in, Store &st, const double (&x0)[4], const double (&xl4)[4], const double (&xu4)[4],
                  double &lam_j, double &lam_k, double &mu, int max_auglag, double mu_max, double scale,
                  const branch::PowTable &T, Result &R) {
    const double xl[N] = { 0.0, 0.0, xl4[0], xl4[1], xl4[2], xl4[3] };
    const double xu[N] = { 200000.0, 200000.0, xu4[0], xu4[1], xu4[2], xu4[3] };
    Lane L;
    begin(L, in, st, x0, lam_j, lam_k, mu, T);
#pragma unroll 1
    for (;;) {
        start_pass(L, st, scale);
        if (step_pass(L, st, xl, xu, max_auglag, mu_max, T)) break;
    }
    const double h2[2] = { in.LH_1h[0], in.LH_1h[1] }, i2[2] = { in.LH_1i[0], in.LH_1i[1] };
    finish(L, st, in.Y, in.res, h2, i2, R);
    lam_j = L.lam_j; lam_k = L.lam_k; mu = L.mu;
}

}  // namespace qpsub
