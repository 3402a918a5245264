// chain.cuh — the branch sub-problem as straight nested loops, one augmented-Lagrangian iteration per call.
//
// Why a second driver next to the state machine of branch.cuh: the x-update lasts as long as its slowest branch. 96 %
// of the branches need ONE augmented-Lagrangian iteration (2-3 objective evaluations); a branch whose line limit has
// just become active walks the penalty ladder of acopf_auglag_linelimit_kernel_gpu.jl:84-126 - 20 to 30 TRON solves
// in a row, each a START evaluation, one Newton step and a trial evaluation - and those few chains used to be three
// quarters of the kernel. The state machine is built for throughput (every lane of a warp in a different phase of a
// different branch, three predicated passes per round); a chain wants latency. Here one AL iteration is one call:
//   * the START evaluation after a multiplier / penalty update re-uses the x-only part of the last evaluation
//     (branch::eval_base) and only re-runs branch::combine - same bits as a fresh evaluation, a third of the work;
//   * the Cauchy search runs in closed form on the straight part of the projected path (tron::cauchy_straight);
//   * no phase bookkeeping: TRON's reverse-communication loop is a plain loop.
// The arithmetic is the state machine's (same functions, explicit FMAs), so a branch gives the same bits whichever
// driver solves it; tests/test_device_code_on_host.py checks that on the host.
//
// The kernel (kernels.cuh) hands a branch over to the chain workers when its first AL iteration does not terminate.
#pragma once
#include "branch.cuh"

namespace chain {

using branch::N;
using branch::Sym6;
using branch::Base;
using branch::PowTable;

struct State {            // what travels from the bulk lane to the chain worker (one queue entry)
    double x[N];
    double ls[2], mu, eta;
    int it_al;
    int evals, cg, shifts, rejected, hit_max;
};

// Evaluation at S.x for the start of an AL iteration (fresh: new branch, hand-over, or after a rejected last step).
template <class View>
EA_DEV void start_eval(const View &D, double scale, const State &S, double &f, double (&g)[N], Sym6 &A, double (&Fc)[4]) {
    Base B;
    branch::eval_base(D, S.x, B);
    branch::combine(D, B, S.ls, S.mu, scale, f, g, A);
#pragma unroll
    for (int k = 0; k < 4; ++k) Fc[k] = B.F[k];
}

// One AL iteration: a TRON solve from the START state (f, g, A evaluated at S.x with S.ls, S.mu; Fc = flows at S.x)
// followed by the multiplier / penalty update (acopf_auglag_linelimit_kernel_gpu.jl:90-133,
// acopf_tron_linelimit_kernel.jl:61-144). Returns true when the branch is finished (S.x, Fc, S.ls, S.mu hold the
// result); otherwise f, g, A, Fc are the START state of the next iteration.
template <class View>
EA_DEV bool al_iteration(const View &D, double scale, const double (&xl)[N], const double (&xu)[N], State &S,
                         double &f, double (&g)[N], Sym6 &A, double (&Fc)[4], int max_auglag, double mu_max,
                         const PowTable &T) {
    const int max_feval = 500, max_minor = 200;     // call site acopf_auglag_linelimit_kernel_gpu.jl:94
    const double gtol = 1e-6;
    int nfev = 1, minor = 1, iter = 1;
    double alphac = 1.0;
    double delta = tron::nrm2<N>(g);                // tron_kernel.jl:102-105
    S.evals++;                                      // the START evaluation (counted like the reference's f-evaluations)
    Base B;
    bool base_at_x = false;                         // B belongs to S.x
#pragma unroll 1
    for (;;) {
        const double fc = f;
        double xc[N];
#pragma unroll
        for (int i = 0; i < N; ++i) xc[i] = S.x[i];
        double prered, g0, snorm;
        tron::Stats st;
#ifdef EA_CAUCHY_LOOP
        tron::compute_step<N>(S.x, xl, xu, A, g, delta, alphac, prered, g0, snorm, st);
#else
        tron::compute_step_fast<N>(S.x, xl, xu, A, g, delta, alphac, prered, g0, snorm, st);
#endif
        S.cg += st.cg; S.shifts += st.shifts;
        double fn;
        branch::eval_base(D, S.x, B);
        branch::combine(D, B, S.ls, S.mu, scale, fn, g, A);     // g, A always belong to the last evaluated point
        S.evals++;
        nfev++;
        if (nfev >= max_feval) {                    // driver stops, trial point kept (tron_kernel.jl:72-75)
#pragma unroll
            for (int k = 0; k < 4; ++k) Fc[k] = B.F[k];
            base_at_x = true;
            break;
        }
        bool accepted;
        const int task = tron::judge_step(fn, fc, g0, snorm, prered, iter == 1, delta, accepted);
        if (accepted) {
            iter++;
            f = fn;
#pragma unroll
            for (int k = 0; k < 4; ++k) Fc[k] = B.F[k];
            base_at_x = true;
            if (task == 2) break;
            minor++;                                 // the reference evaluates g, H here (task GH)
            if (tron::gpnorm<N>(S.x, xl, xu, g) <= gtol) break;            // NEWX test (:121-130)
            if (minor >= max_minor) break;
        } else {
            S.rejected++;
#pragma unroll
            for (int i = 0; i < N; ++i) S.x[i] = xc[i];
            f = fc;
            base_at_x = false;
            if (task == 2) break;                    // Fc still holds the flows at xc
            branch::eval_base(D, S.x, B);            // g, A back at xc before the next step (not an f-evaluation)
            branch::combine(D, B, S.ls, S.mu, scale, f, g, A);
#pragma unroll
            for (int k = 0; k < 4; ++k) Fc[k] = B.F[k];
            base_at_x = true;
        }
    }

    // augmented-Lagrangian update on the line limits (auglag_gpu.jl:96-131)
    S.it_al++;
    const double cviol1 = Fc[0] * Fc[0] + Fc[1] * Fc[1] + S.x[4];
    const double cviol2 = Fc[2] * Fc[2] + Fc[3] * Fc[3] + S.x[5];
    const double cnorm = tron::dmax(fabs(cviol1), fabs(cviol2));
    bool terminate = false;
    if (cnorm <= S.eta) {
        if (cnorm <= 1e-6) terminate = true;
        else {
            double inv_p01, p09;
            branch::mu_powers(T, S.mu, inv_p01, p09);
            S.ls[0] += S.mu * cviol1;
            S.ls[1] += S.mu * cviol2;
            S.eta = S.eta / p09;
        }
    } else {
        S.mu = tron::dmin(mu_max, S.mu * 10.0);
        double inv_p01, p09;
        branch::mu_powers(T, S.mu, inv_p01, p09);
        S.eta = inv_p01;
    }
    if (S.it_al >= max_auglag) { if (!terminate) S.hit_max = 1; terminate = true; }
    if (terminate) return true;
    // START evaluation of the next AL iteration: same point, new (ls, mu)
    if (!base_at_x) {
        branch::eval_base(D, S.x, B);
#pragma unroll
        for (int k = 0; k < 4; ++k) Fc[k] = B.F[k];
    }
    branch::combine(D, B, S.ls, S.mu, scale, f, g, A);
    return false;
}

// A whole branch on one lane (host harness, probes).
template <class View>
EA_DEV void solve(const View &D, double scale, const double (&xl)[N], const double (&xu)[N], State &S, double (&Fc)[4],
                  int max_auglag, double mu_max, const PowTable &T) {
    double f, g[N];
    Sym6 A;
    start_eval(D, scale, S, f, g, A, Fc);
#pragma unroll 1
    while (!al_iteration(D, scale, xl, xu, S, f, g, A, Fc, max_auglag, mu_max, T)) {}
}

EA_DEV void init_state(State &S, const PowTable &T) {   // x, ls, mu set by the caller (branch::begin's counterpart)
    double inv_p01, p09;
    branch::mu_powers(T, S.mu, inv_p01, p09);
    S.eta = inv_p01;                                     // eta = 1/mu^0.1 (auglag_gpu.jl:84)
    S.it_al = 0;
    S.evals = S.cg = S.shifts = S.rejected = S.hit_max = 0;
}

}  // namespace chain
