// branch.cuh — the per-branch sub-problem of the ADMM x-update:
//   min over x = (vi, vj, ti, tj, s_ij, s_ji) in a box of the scaled augmented
//   Lagrangian of the branch flow consensus terms + line-limit constraints,
//   wrapped in the outer augmented-Lagrangian loop on the two line limits.
//
// Replaces /root/reference/src/models/acopf/acopf_auglag_linelimit_kernel_gpu.jl:1-151
// (AL loop), acopf_tron_linelimit_kernel.jl:4-149 (TRON driver) and
// acopf_eval_linelimit_kernel_gpu.jl:1-594 (f, grad, Hessian).
//
// One lane owns one branch at a time. The objective uses the structure every flow
// shares, F = a vi^2 + b vj^2 + vi vj P(t), t = ti - tj:
// gradient and Hessian are assembled in the 3 variables (vi, vj, t) from
// aggregated flow weights and then expanded to the 6x6 matrix (kept in its 15
// independent entries, Hess), one sincos per evaluation. The evaluation comes in
// three parts (flows / f and gradient / Hessian) so that a trial point costs only
// what its judgement needs (eval_pass).
//
// The AL loop and TRON's reverse-communication loop are flattened into a state
// machine (`Lane`) advanced by three uniform phases per round,
//     eval_pass(0)  lanes holding a trial point: evaluate, judge, converge, AL update
//     eval_pass(1)  lanes starting a TRON solve (new branch, next AL iteration, or a
//                   rejected step): evaluate at the current point
//     compute()     every live lane: dtron's COMPUTE (Cauchy point + projected CG), taken directly where it runs its
//                   common course (tron::newton_step) -> next trial point
// so that the lanes of a warp, each at a different stage of a different branch,
// execute the same instructions. The kernel (kernels.cuh) refills finished lanes
// with new branches between the phases; the host test harness drives the very same
// three functions for one branch.
#pragma once
#include "tron.cuh"

#ifdef EA_CYC
__host__ __device__ __forceinline__ long long ea_clock() {
#ifdef __CUDA_ARCH__
    return clock64();
#else
    return 0;
#endif
}
#define EA_CYC_T(t) const long long t = ea_clock()
#define EA_CYC_ADD(L, k, a, b) (L).cyc[k] += (b) - (a)
#else
#define EA_CYC_T(t)
#define EA_CYC_ADD(L, k, a, b)
#endif

namespace branch {

constexpr int N = 6;
// The Hessian of the branch objective in its 15 independent entries. Of the 21 entries of the packed lower triangle
// six follow from the others - the objective depends on the angles through t_i - t_j only, and each slack enters one
// constraint linearly: a30 = -a20, a31 = -a21, a43 = -a42, a53 = -a52, a54 = 0, a55 = a44 - and the TRON routines read a
// matrix through at(i, j) with compile-time indices, so the negations become operand modifiers and the zero drops out:
// twelve registers less in the lane state (the branch kernel sits at the 255-register cap) and a few multiply-adds less
// per matrix-vector product, with bit-identical results.
struct Hess {
    double a00, a10, a11, a20, a21, a22, a32, a33, a40, a41, a42, a44, a50, a51, a52;
    EA_DEV double at(int i, int j) const {
        const int k = tron::tri(i, j);
        return k == 0 ? a00 : k == 1 ? a10 : k == 2 ? a11 : k == 3 ? a20 : k == 4 ? a21 : k == 5 ? a22
             : k == 6 ? -a20 : k == 7 ? -a21 : k == 8 ? a32 : k == 9 ? a33
             : k == 10 ? a40 : k == 11 ? a41 : k == 12 ? a42 : k == 13 ? -a42 : k == 14 ? a44
             : k == 15 ? a50 : k == 16 ? a51 : k == 17 ? a52 : k == 18 ? -a52 : k == 19 ? 0.0 : a44;
    }
    // from a full 6 x 6 matrix with that structure (row-major; the oracle's Hessian in the host harness)
    EA_DEV void from_dense(const double *H) {
        a00 = H[0]; a10 = H[6]; a11 = H[7]; a20 = H[12]; a21 = H[13]; a22 = H[14]; a32 = H[20]; a33 = H[21];
        a40 = H[24]; a41 = H[25]; a42 = H[26]; a44 = H[28]; a50 = H[30]; a51 = H[31]; a52 = H[32];
    }
};
using Sym6 = Hess;

struct Data {          // per-branch inputs, reference rows in comments (membuf rows 1-24, acopf_auglag..gpu.jl:50-73)
    double lam[8];     // rows 1-8   lambda   (pij,qij,pji,qji,wi,wj,ti,tj)
    double rho[8];     // rows 9-16  rho
    double xt[8];      // rows 17-24 xbar - z
    double Y[8];       // YffR,YffI,YftR,YftI,YttR,YttI,YtfR,YtfI
};

// Views of the per-branch data: a plain struct (host harness) or one column of a
// shared-memory tile (kernel; element k of lane t lives at base[k * STRIDE], which is
// bank-conflict free for a warp).
struct StructView {
    const Data *d;
    EA_DEV double lam(int k) const { return d->lam[k]; }
    EA_DEV double rho(int k) const { return d->rho[k]; }
    EA_DEV double xt(int k) const { return d->xt[k]; }
    EA_DEV double Y(int k) const { return d->Y[k]; }
};
template <int STRIDE> struct TileView {
    const double *base;                                  // &tile[0][lane]
    EA_DEV double lam(int k) const { return base[k * STRIDE]; }
    EA_DEV double rho(int k) const { return base[(8 + k) * STRIDE]; }
    EA_DEV double xt(int k) const { return base[(16 + k) * STRIDE]; }
    EA_DEV double Y(int k) const { return base[(24 + k) * STRIDE]; }
};
constexpr int COLD_ROWS = 18;                            // Lane::xc (6), Fc (4), eta, inv_p01, p09, f, fc, prered, g0, snorm
constexpr int TILE_ROWS = 32 + 9 + COLD_ROWS;            // lam, rho, xt, Y + xl0..3, xu0..3, rateA + the lane's cold state
constexpr int TILE_COLD = 32 + 9;                        // first cold row

struct PowTable {      // host-computed (glibc) 1/mu^0.1 and mu^0.9 for the mu sequence 10, 100, ... <= mu_max
    int n;
    double mu[24], inv_p01[24], p09[24];
};

__host__ __device__ __noinline__ inline void mu_powers_general(double mu, double *inv_p01, double *p09) {
    *inv_p01 = 1.0 / pow(mu, 0.1);          // only if mu was set from outside to something off the 10^k ladder
    *p09 = pow(mu, 0.9);
}
// returns the index of mu in the table, -1 if it is not there
EA_DEV int mu_powers(const PowTable &T, double mu, double &inv_p01, double &p09) {
#pragma unroll 1
    for (int k = 0; k < T.n; ++k)
        if (T.mu[k] == mu) { inv_p01 = T.inv_p01[k]; p09 = T.p09[k]; return k; }
    double a, b;
    mu_powers_general(mu, &a, &b);
    inv_p01 = a; p09 = b;
    return -1;
}

#ifdef EA_PARITY
// sin / cos for the parity build: Cody-Waite reduction by pi/2 in three parts, then the fdlibm polynomial kernels in
// Horner form, every operation an individually rounded multiply or add. The oracle can be switched to the same
// formulas (oracle/portable_sincos.h), which removes the last difference between its arithmetic and this build's:
// libm's and libdevice's sin / cos disagree in the last bit for a few percent of the arguments.
EA_DEV void psincos(double t, double *sn, double *cs) {
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00, pio2_2 = 6.07710050630396597660e-11, pio2_3 = 2.02226624871116645580e-21;
    const double pio2_3t = 8.47842766036889956997e-32;
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double fn = rint(t * invpio2);
    double r = t - fn * pio2_1;
    r = r - fn * pio2_2;
    r = r - fn * pio2_3;
    r = r - fn * pio2_3t;
    const double z = r * r;
    const double ps = S1 + z * (S2 + z * (S3 + z * (S4 + z * (S5 + z * S6))));
    const double pc = C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6))));
    const double s = r + r * (z * ps);
    const double c = (1.0 - 0.5 * z) + z * (z * pc);
    const long long q = (long long)fn;
    const int k = (int)(q & 3);
    *sn = (k == 0) ? s : ((k == 1) ? c : ((k == 2) ? -s : -c));
    *cs = (k == 0) ? c : ((k == 1) ? -s : ((k == 2) ? -c : s));
}

// The oracle's formulas in the oracle's operation order, every operation rounded separately (oracle/acopf_oracle.c:
// orc_eval_f, orc_eval_gh - themselves restatements of acopf_eval_linelimit_kernel_cpu.jl in compact form).
template <class View>
EA_DEV void eval_fgh_ref(const View &D, const double (&ls)[2], double mu, double scale,
                         const double (&x)[N], double &f, double (&g)[N], Sym6 &A) {
    const double vi = x[0], vj = x[1], t = x[2] - x[3];
    double c, s;
    psincos(t, &s, &c);
    const double Y0 = D.Y(0), Y1 = D.Y(1), Y2 = D.Y(2), Y3 = D.Y(3), Y4 = D.Y(4), Y5 = D.Y(5), Y6 = D.Y(6), Y7 = D.Y(7);
    const double a[4] = { Y0, -Y1, 0.0, 0.0 };
    const double b[4] = { 0.0, 0.0, Y4, -Y5 };
    const double ga[4] = { Y2, -Y3, Y6, -Y7 };
    const double de[4] = { Y3, Y2, -Y7, -Y6 };
    double P[4], Q[4], F[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        P[k] = ga[k] * c + de[k] * s;
        Q[k] = -ga[k] * s + de[k] * c;
        F[k] = a[k] * (vi * vi) + b[k] * (vj * vj) + (vi * vj) * P[k];
    }
    {   // orc_eval_f
        const double h[8] = { F[0], F[1], F[2], F[3], x[0] * x[0], x[1] * x[1], x[2], x[3] };
        double fv = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double dd = h[k] - D.xt(k);
            fv += D.lam(k) * h[k] + 0.5 * (D.rho(k) * (dd * dd));
        }
        const double c1 = F[0] * F[0] + F[1] * F[1] + x[4];
        const double c2 = F[2] * F[2] + F[3] * F[3] + x[5];
        fv += ls[0] * c1 + ls[1] * c2 + 0.5 * (mu * (c1 * c1)) + 0.5 * (mu * (c2 * c2));
        f = scale * fv;
    }
    // orc_eval_gh
    const double c1 = F[0] * F[0] + F[1] * F[1] + x[4];
    const double c2 = F[2] * F[2] + F[3] * F[3] + x[5];
    const double m[2] = { ls[0] + mu * c1, ls[1] + mu * c2 };
    double gy[3] = { 0, 0, 0 }, Hy[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } }, d[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = k >> 1;
        const double G[3] = { 2.0 * a[k] * vi + vj * P[k], 2.0 * b[k] * vj + vi * P[k], (vi * vj) * Q[k] };
        const double r = D.lam(k) + D.rho(k) * (F[k] - D.xt(k));
        const double w = r + 2.0 * m[j] * F[k];
        const double kap = D.rho(k) + 2.0 * m[j];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            gy[i] += w * G[i];
            d[j][i] += 2.0 * F[k] * G[i];
#pragma unroll
            for (int q = 0; q < 3; ++q) Hy[i][q] += kap * G[i] * G[q];
        }
        Hy[0][0] += w * (2.0 * a[k]);
        Hy[1][1] += w * (2.0 * b[k]);
        Hy[0][1] += w * P[k];
        Hy[0][2] += w * (vj * Q[k]);
        Hy[1][2] += w * (vi * Q[k]);
        Hy[2][2] += w * (-(vi * vj) * P[k]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int q = i; q < 3; ++q) Hy[i][q] += mu * d[j][i] * d[j][q];
    const double ri = D.lam(4) + D.rho(4) * (vi * vi - D.xt(4));
    const double rj = D.lam(5) + D.rho(5) * (vj * vj - D.xt(5));
    gy[0] += 2.0 * vi * ri;
    gy[1] += 2.0 * vj * rj;
    Hy[0][0] += 2.0 * ri + 4.0 * D.rho(4) * (vi * vi);
    Hy[1][1] += 2.0 * rj + 4.0 * D.rho(5) * (vj * vj);
    g[0] = scale * gy[0];
    g[1] = scale * gy[1];
    g[2] = scale * (gy[2] + D.lam(6) + D.rho(6) * (x[2] - D.xt(6)));
    g[3] = scale * (-gy[2] + D.lam(7) + D.rho(7) * (x[3] - D.xt(7)));
    g[4] = scale * m[0];
    g[5] = scale * m[1];
    A.a00 = scale * Hy[0][0];
    A.a10 = scale * Hy[0][1];
    A.a11 = scale * Hy[1][1];
    A.a20 = scale * Hy[0][2];
    A.a21 = scale * Hy[1][2];
    A.a22 = scale * (Hy[2][2] + D.rho(6));
    A.a32 = scale * (-Hy[2][2]);
    A.a33 = scale * (Hy[2][2] + D.rho(7));
    A.a40 = scale * (mu * d[0][0]);
    A.a41 = scale * (mu * d[0][1]);
    A.a42 = scale * (mu * d[0][2]);
    A.a44 = scale * mu;
    A.a50 = scale * (mu * d[1][0]);
    A.a51 = scale * (mu * d[1][1]);
    A.a52 = scale * (mu * d[1][2]);
    // (a30 = scale * (-Hy02), a31 = scale * (-Hy12), a43 = scale * (-mu d02), a53 = scale * (-mu d12), a54 = 0,
    //  a55 = scale * mu in the oracle: exactly the negatives / copies that Hess::at returns)
}
#endif

// f, grad f, Hessian (packed lower) of the scaled branch AL objective, and the four flows F = (pij, qij, pji, qji)
// (acopf_eval_linelimit_kernel_gpu.jl:17-22), in three parts so that the driver can stop after the part it needs:
//   eval_pre   what depends on the point only: sin / cos, the flows F_k = a_k vi^2 + b_k vj^2 + vi vj P_k(t), P_k and
//              Q_k = dP_k/dt, the two line-limit constraint values;
//   eval_fg    f and the gradient for given multipliers / penalty (ls, mu): the aggregated flow weights
//              sum_k w_k (a, b, P, Q)_k, which the Hessian needs again, are handed on in Wsum;
//   eval_hess  the Hessian.
// A trial point whose TRON solve ends there needs f and g only (acceptance and convergence tests); and when the AL
// update that follows keeps the point and changes (ls, mu), the START evaluation of the next solve is eval_fg + eval_hess
// on the SAME Pre - bit for bit what a fresh evaluation at that point gives (same code, same inputs), without the
// sincos and the flows. Flows 0, 1 are the from-side (b_k = 0), 2, 3 the to-side (a_k = 0): the zero terms are left out.
struct Pre { double P[4], Q[4], F[4], c1, c2; };
struct Wsum { double As, Bs, Ps, Qs, ri, rj; };

template <class View>
EA_DEV void eval_pre(const View &D, const double (&x)[N], Pre &p) {
    const double vi = x[0], vj = x[1];
    double s, c;
    sincos(x[2] - x[3], &s, &c);
    const double vv = vi * vj, vi2 = vi * vi, vj2 = vj * vj;
    const double Y2 = D.Y(2), Y3 = D.Y(3), Y6 = D.Y(6), Y7 = D.Y(7);
    const double ga[4] = { Y2, -Y3, Y6, -Y7 };
    const double de[4] = { Y3, Y2, -Y7, -Y6 };
    const double ab[4] = { D.Y(0), -D.Y(1), D.Y(4), -D.Y(5) };     // a_0, a_1, b_2, b_3
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        p.P[k] = ga[k] * c + de[k] * s;
        p.Q[k] = de[k] * c - ga[k] * s;
        p.F[k] = ab[k] * (k < 2 ? vi2 : vj2) + vv * p.P[k];
    }
    p.c1 = p.F[0] * p.F[0] + p.F[1] * p.F[1] + x[4];
    p.c2 = p.F[2] * p.F[2] + p.F[3] * p.F[3] + x[5];
}

template <class View>
EA_DEV void eval_fg(const View &D, const Pre &p, const double (&ls)[2], double mu, double scale,
                    const double (&x)[N], double &f, double (&g)[N], Wsum &W) {
    const double vi = x[0], vj = x[1];
    const double vv = vi * vj, vi2 = vi * vi, vj2 = vj * vj;
    const double ab[4] = { D.Y(0), -D.Y(1), D.Y(4), -D.Y(5) };
    const double m[2] = { ls[0] + mu * p.c1, ls[1] + mu * p.c2 };
    double fv = ls[0] * p.c1 + ls[1] * p.c2 + 0.5 * (mu * (p.c1 * p.c1)) + 0.5 * (mu * (p.c2 * p.c2));
    double As = 0.0, Bs = 0.0, Ps = 0.0, Qs = 0.0;        // sum_k w_k * (a,b,P,Q)_k
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double lam = D.lam(k), rho = D.rho(k);
        const double dev = p.F[k] - D.xt(k);
        fv += lam * p.F[k] + 0.5 * (rho * (dev * dev));
        const double w = (lam + rho * dev) + 2.0 * m[k >> 1] * p.F[k];
        if (k < 2) As += w * ab[k]; else Bs += w * ab[k];
        Ps += w * p.P[k]; Qs += w * p.Q[k];
    }
    // consensus terms on w_i = vi^2, w_j = vj^2, t_i, t_j
    const double rho4 = D.rho(4), rho5 = D.rho(5), rho6 = D.rho(6), rho7 = D.rho(7);
    const double dwi = vi2 - D.xt(4), dwj = vj2 - D.xt(5), dti = x[2] - D.xt(6), dtj = x[3] - D.xt(7);
    const double lam4 = D.lam(4), lam5 = D.lam(5), lam6 = D.lam(6), lam7 = D.lam(7);
    fv += lam4 * vi2 + 0.5 * (rho4 * (dwi * dwi)) + lam5 * vj2 + 0.5 * (rho5 * (dwj * dwj))
        + lam6 * x[2] + 0.5 * (rho6 * (dti * dti)) + lam7 * x[3] + 0.5 * (rho7 * (dtj * dtj));
    f = scale * fv;
    const double ri = lam4 + rho4 * dwi;
    const double rj = lam5 + rho5 * dwj;
    const double gy2 = vv * Qs;
    g[0] = scale * (2.0 * As * vi + vj * Ps + 2.0 * vi * ri);
    g[1] = scale * (2.0 * Bs * vj + vi * Ps + 2.0 * vj * rj);
    g[2] = scale * (gy2 + lam6 + rho6 * dti);
    g[3] = scale * (-gy2 + lam7 + rho7 * dtj);
    g[4] = scale * m[0];
    g[5] = scale * m[1];
    W.As = As; W.Bs = Bs; W.Ps = Ps; W.Qs = Qs; W.ri = ri; W.rj = rj;
}

template <class View>
EA_DEV void eval_hess(const View &D, const Pre &p, const Wsum &W, const double (&ls)[2], double mu, double scale,
                      const double (&x)[N], Sym6 &A) {
    const double vi = x[0], vj = x[1];
    const double vv = vi * vj, vi2 = vi * vi, vj2 = vj * vj;
    const double ab[4] = { D.Y(0), -D.Y(1), D.Y(4), -D.Y(5) };
    const double m[2] = { ls[0] + mu * p.c1, ls[1] + mu * p.c2 };
    // reduced (vi, vj, t) Hessian
    double H00 = 0.0, H01 = 0.0, H02 = 0.0, H11 = 0.0, H12 = 0.0, H22 = 0.0;
    double d[2][3] = { { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 } };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = k >> 1;
        const double G0 = (k < 2) ? 2.0 * ab[k] * vi + vj * p.P[k] : vj * p.P[k];
        const double G1 = (k < 2) ? vi * p.P[k] : 2.0 * ab[k] * vj + vi * p.P[k];
        const double G2 = vv * p.Q[k];
        const double kap = D.rho(k) + 2.0 * m[j];
        const double tF = 2.0 * p.F[k];
        d[j][0] += tF * G0; d[j][1] += tF * G1; d[j][2] += tF * G2;
        const double k0 = kap * G0, k1 = kap * G1, k2 = kap * G2;
        H00 += k0 * G0; H01 += k0 * G1; H02 += k0 * G2;
        H11 += k1 * G1; H12 += k1 * G2; H22 += k2 * G2;
    }
    H00 += 2.0 * W.As; H11 += 2.0 * W.Bs; H01 += W.Ps;
    H02 += vj * W.Qs;  H12 += vi * W.Qs;  H22 -= vv * W.Ps;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const double m0 = mu * d[j][0], m1 = mu * d[j][1], m2 = mu * d[j][2];
        H00 += m0 * d[j][0]; H01 += m0 * d[j][1]; H02 += m0 * d[j][2];
        H11 += m1 * d[j][1]; H12 += m1 * d[j][2]; H22 += m2 * d[j][2];
    }
    const double rho4 = D.rho(4), rho5 = D.rho(5), rho6 = D.rho(6), rho7 = D.rho(7);
    H00 += 2.0 * W.ri + 4.0 * rho4 * vi2;
    H11 += 2.0 * W.rj + 4.0 * rho5 * vj2;

    const double smu = scale * mu;
    A.a00 = scale * H00;
    A.a10 = scale * H01;
    A.a11 = scale * H11;
    A.a20 = scale * H02;
    A.a21 = scale * H12;
    A.a22 = scale * (H22 + rho6);
    A.a32 = -(scale * H22);
    A.a33 = scale * (H22 + rho7);
    A.a40 = smu * d[0][0];
    A.a41 = smu * d[0][1];
    A.a42 = smu * d[0][2];
    A.a44 = smu;
    A.a50 = smu * d[1][0];
    A.a51 = smu * d[1][1];
    A.a52 = smu * d[1][2];
}

// The whole evaluation at once (diagnostics, unit tests; the parity build's only form).
template <class View>
EA_DEV void eval_fgh(const View &D, const double (&ls)[2], double mu, double scale,
                     const double (&x)[N], double &f, double (&g)[N], Sym6 &A, double (&F)[4]) {
#ifdef EA_PARITY
    eval_fgh_ref(D, ls, mu, scale, x, f, g, A);
    {   // flows as the AL loop of the reference computes them (acopf_auglag_linelimit_kernel_gpu.jl:98-104, 135-138)
        double sn, cs;
        psincos(x[2] - x[3], &sn, &cs);
        const double cc = x[0] * x[1] * cs, ss = x[0] * x[1] * sn;
        F[0] = D.Y(0) * (x[0] * x[0]) + D.Y(2) * cc + D.Y(3) * ss;
        F[1] = -D.Y(1) * (x[0] * x[0]) - D.Y(3) * cc + D.Y(2) * ss;
        F[2] = D.Y(4) * (x[1] * x[1]) + D.Y(6) * cc - D.Y(7) * ss;
        F[3] = -D.Y(5) * (x[1] * x[1]) - D.Y(7) * cc - D.Y(6) * ss;
    }
#else
    Pre p;
    Wsum W;
    eval_pre(D, x, p);
    eval_fg(D, p, ls, mu, scale, x, f, g, W);
    eval_hess(D, p, W, ls, mu, scale, x, A);
#pragma unroll
    for (int k = 0; k < 4; ++k) F[k] = p.F[k];
#endif
}

// The branch objective bound to a data view (what the kernel and the harness plug into Lane). An evaluator that
// declares the types Pre / Wsum and the three-part interface is driven part by part (eval_pass); one that only has
// operator() (the parity build, the oracle's functions in the host harness) is called whole.
template <class View> struct Objective {
    View D;
    double scale;
    EA_DEV void operator()(const double (&x)[N], const double (&ls)[2], double mu, double &f, double (&g)[N],
                           Sym6 &A, double (&F)[4]) const {
        eval_fgh(D, ls, mu, scale, x, f, g, A, F);
    }
#ifndef EA_PARITY
    using Pre = branch::Pre;
    using Wsum = branch::Wsum;
    EA_DEV void pre(const double (&x)[N], Pre &p) const { eval_pre(D, x, p); }
    EA_DEV void fg(const Pre &p, const double (&x)[N], const double (&ls)[2], double mu, double &f, double (&g)[N],
                   Wsum &W) const { eval_fg(D, p, ls, mu, scale, x, f, g, W); }
    EA_DEV void hess(const Pre &p, const Wsum &W, const double (&x)[N], const double (&ls)[2], double mu,
                     Sym6 &A) const { eval_hess(D, p, W, ls, mu, scale, x, A); }
#endif
};
template <class E, class = void> struct is_split { static constexpr bool value = false; };
template <class E> struct is_split<E, decltype((void)sizeof(typename E::Pre))> { static constexpr bool value = true; };

enum Phase : int { NEED = 0, START = 1, RESTORE = 2, TRIAL = 3, DONE = 4 };

// State of one branch solve (registers).
struct Lane {
    double x[N], g[N];
    Sym6 A;
    double ls[2], mu;                       // AL state (membuf rows 25-27)
    double delta, alphac;
    // State that is written once per TRON step and read once per step or per solve stays out of the register file - the
    // lone lane's round is bound by it. It lives at cold[k * cs] (a shared-memory column in the kernels, a plain array
    // in the host harness): xc (the point the step started from), Fc (flows at the current accepted point), the AL
    // thresholds eta / 1/mu^0.1 / mu^0.9, and the scalars that travel from the step to its evaluation (f, fc, prered,
    // g's, |s|).
    double *cold;
    int cs;
    EA_DEV double &xc(int i) const { return cold[i * cs]; }
    EA_DEV double &Fc(int k) const { return cold[(N + k) * cs]; }
    EA_DEV double &eta() const { return cold[(N + 4) * cs]; }
    EA_DEV double &inv_p01() const { return cold[(N + 5) * cs]; }
    EA_DEV double &p09() const { return cold[(N + 6) * cs]; }
    EA_DEV double &f() const { return cold[(N + 7) * cs]; }
    EA_DEV double &fc() const { return cold[(N + 8) * cs]; }
    EA_DEV double &prered() const { return cold[(N + 9) * cs]; }
    EA_DEV double &g0() const { return cold[(N + 10) * cs]; }
    EA_DEV double &snorm() const { return cold[(N + 11) * cs]; }
    int nfev, minor, iter, it_al;
    int kmu;                                // index of mu in the PowTable (-1: off the ladder)
    int phase;
    bool step_pending;
    // work of the current branch
    int evals, cg, shifts, rejected, hit_max;
#ifdef EA_CYC
    long long cyc[8];                       // diagnostics build: SM cycles per section (tools/probe_branches.py)
#endif
};

// Start a branch: x, ls, mu must be set by the caller (mu = 10 on the first inner
// iteration of an outer iteration, acopf_auglag_linelimit_kernel_gpu.jl:75-80).
EA_DEV void begin(Lane &L, const PowTable &T) {
    L.kmu = mu_powers(T, L.mu, L.inv_p01(), L.p09());
    L.eta() = L.inv_p01();                       // eta = 1/mu^0.1 (:84)
    L.f() = L.fc() = L.delta = L.prered() = L.g0() = L.snorm() = 0.0;
    L.alphac = 1.0;
    L.nfev = 0; L.minor = 0; L.iter = 1; L.it_al = 0;
    L.phase = START;
    L.step_pending = false;
    L.evals = L.cg = L.shifts = L.rejected = L.hit_max = 0;
#ifdef EA_CYC
    for (int k = 0; k < 8; ++k) L.cyc[k] = 0;
#endif
#pragma unroll
    for (int i = 0; i < N; ++i) { L.g[i] = 0.0; L.xc(i) = L.x[i]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) L.Fc(k) = 0.0;
}

// The evaluation in progress: the three-part form keeps Pre / Wsum between its parts, the whole form has nothing to keep.
template <class Eval, bool SPLIT> struct EvalState;
template <class Eval> struct EvalState<Eval, true> {
    typename Eval::Pre p;
    typename Eval::Wsum W;
    EA_DEV void pre(const Eval &eval, const double (&x)[N]) { eval.pre(x, p); }
    EA_DEV void fg(const Eval &eval, Lane &L, double &fn, double (&Fn)[4]) {
        eval.fg(p, L.x, L.ls, L.mu, fn, L.g, W);
#pragma unroll
        for (int k = 0; k < 4; ++k) Fn[k] = p.F[k];
    }
    EA_DEV void hess(const Eval &eval, Lane &L) { eval.hess(p, W, L.x, L.ls, L.mu, L.A); }
};
template <class Eval> struct EvalState<Eval, false> {
    EA_DEV void pre(const Eval &, const double (&)[N]) {}
#pragma nv_exec_check_disable                        // (the host harness plugs host-only evaluators in)
    EA_DEV void fg(const Eval &eval, Lane &L, double &fn, double (&Fn)[4]) { eval(L.x, L.ls, L.mu, fn, L.g, L.A, Fn); }
    EA_DEV void hess(const Eval &, Lane &) {}
};

// mu moved one rung up the ladder: the powers of the next table entry, if mu is on the ladder (else the search).
EA_DEV void mu_powers_next(const PowTable &T, Lane &L) {
    const int k = L.kmu + 1;
    if (L.kmu >= 0 && k < T.n && T.mu[k] == L.mu) { L.kmu = k; L.inv_p01() = T.inv_p01[k]; L.p09() = T.p09[k]; return; }
    if (L.kmu >= 0 && T.mu[L.kmu] == L.mu) return;                 // mu_max reached: unchanged
    L.kmu = mu_powers(T, L.mu, L.inv_p01(), L.p09());
}

// One evaluation phase. pass 0 serves lanes holding a trial point, pass 1 lanes at
// the start of a TRON solve (or re-evaluating after a rejected step). Returns true
// when the branch is finished (x, Fc, ls, mu hold the result).
//
// With a three-part evaluator (Objective in the fast build) a trial point is evaluated as far as its judgement needs
// - f and g -; the Hessian follows only if the lane goes on from that point: the TRON solve continues, or the solve
// ended, the AL update changed (ls, mu) and the next solve starts where this one stopped. In that last case the START
// evaluation of the next solve is f, g, A for the new (ls, mu) on the flows already computed ("stage 1" below): same
// values as a fresh evaluation, counted like one, without a second pass over the point.
template <class Eval>
EA_DEV bool eval_pass(Lane &L, const Eval &eval, int pass, const double (&xl)[N], const double (&xu)[N],
                      int max_auglag, double mu_max, const PowTable &T) {
    const int max_feval = 500, max_minor = 200;     // call site acopf_auglag_linelimit_kernel_gpu.jl:94
    const double gtol = 1e-6;
    const bool mine = (pass == 0) ? (L.phase == TRIAL && !L.step_pending)
                                  : ((L.phase == START || L.phase == RESTORE) && !L.step_pending);
    if (!mine) return false;
    constexpr bool SPLIT = is_split<Eval>::value;

    double fn, Fn[4];
    EvalState<Eval, SPLIT> es;
    EA_CYC_T(c0);
    es.pre(eval, L.x);
    EA_CYC_T(c1);
    EA_CYC_ADD(L, 0, c0, c1);
    int stage = pass;                                // 0: judge the trial point, 1: start of a TRON solve
#pragma unroll 1
    for (;;) {
        EA_CYC_T(c2);
        es.fg(eval, L, fn, Fn);                      // g (and without the split A) belong to the last evaluated point
        EA_CYC_T(c3);
        EA_CYC_ADD(L, 1, c2, c3);
        if (L.phase != RESTORE) L.evals++;           // counted like the reference's f-evaluations

        if (stage == 1) {
            L.f() = fn;
#pragma unroll
            for (int k = 0; k < 4; ++k) L.Fc(k) = Fn[k];
            if (L.phase == START) {                  // task 0: a fresh TRON solve
                L.nfev = 1; L.minor = 1; L.iter = 1; L.alphac = 1.0;
                L.delta = tron::nrm2<N>(L.g);        // tron_kernel.jl:102-105
            }
            L.step_pending = true;
            break;
        }

        bool tron_done = false, at_trial = true;     // at_trial: x is still the point just evaluated
        L.nfev++;
        if (L.nfev >= max_feval) {
            tron_done = true;                        // driver stops, trial point kept (tron_kernel.jl:72-75)
#pragma unroll
            for (int k = 0; k < 4; ++k) L.Fc(k) = Fn[k];
        } else {
            bool accepted;
            const int task = tron::judge_step(fn, L.fc(), L.g0(), L.snorm(), L.prered(), L.iter == 1, L.delta, accepted);
            if (accepted) {
                L.iter++;
                L.f() = fn;
#pragma unroll
                for (int k = 0; k < 4; ++k) L.Fc(k) = Fn[k];
                if (task == 2) tron_done = true;
                else {
                    L.minor++;                        // the reference evaluates g,H here (task GH)
                    if (tron::gpnorm<N>(L.x, xl, xu, L.g) <= gtol) tron_done = true;          // NEWX test (:121-130)
                    else if (L.minor >= max_minor) tron_done = true;
                    else L.step_pending = true;
                }
            } else {
                L.rejected++;
                at_trial = false;
#pragma unroll
                for (int i = 0; i < N; ++i) L.x[i] = L.xc(i);
                L.f() = L.fc();
                if (task == 2) tron_done = true;      // Fc still holds the flows at xc
                else L.phase = RESTORE;               // g, A must be re-evaluated at xc before the next step
            }
        }
        if (!tron_done) {
            if (L.step_pending) break;                // goes on from the trial point: needs its Hessian
            return false;                             // rejected: pass 1 re-evaluates at xc
        }

        // augmented-Lagrangian update on the line limits (auglag_gpu.jl:96-131)
        L.it_al++;
        const double cviol1 = L.Fc(0) * L.Fc(0) + L.Fc(1) * L.Fc(1) + L.x[4];
        const double cviol2 = L.Fc(2) * L.Fc(2) + L.Fc(3) * L.Fc(3) + L.x[5];
        const double cnorm = tron::dmax(fabs(cviol1), fabs(cviol2));
        bool terminate = false;
        if (cnorm <= L.eta()) {
            if (cnorm <= 1e-6) terminate = true;
            else {
                L.ls[0] += L.mu * cviol1;
                L.ls[1] += L.mu * cviol2;
                L.eta() = L.eta() / L.p09();
            }
        } else {
            L.mu = tron::dmin(mu_max, L.mu * 10.0);
            mu_powers_next(T, L);
            L.eta() = L.inv_p01();
        }
        if (L.it_al >= max_auglag) { if (!terminate) L.hit_max = 1; terminate = true; }
        if (terminate) { L.phase = DONE; return true; }
        L.phase = START;
        if (!SPLIT || !at_trial) return false;        // pass 1 evaluates at x
        stage = 1;                                    // same point, new (ls, mu): f, g again, then the Hessian
    }
    EA_CYC_T(c4);
    es.hess(eval, L);
    EA_CYC_T(c5);
    EA_CYC_ADD(L, 2, c4, c5);
    return false;
}

// dtron COMPUTE for lanes with a pending step: Cauchy point + projected CG -> trial point in x.
EA_DEV void compute(Lane &L, const double (&xl)[N], const double (&xu)[N]) {
    if (!L.step_pending) return;
    L.fc() = L.f();
#pragma unroll
    for (int i = 0; i < N; ++i) L.xc(i) = L.x[i];
    tron::Stats st;
    // tron::compute_step_auto: the step is taken directly where dtron's COMPUTE runs its common course (tron::newton_step:
    // > 99.9 % of the steps of a solve of the BASELINE grids, tools/step_stats.py), by the literal algorithm otherwise.
    // Which of the two runs depends on the branch's own data only, so a branch gives the same bits wherever and whenever
    // it is solved.
#ifdef EA_CYC
    {
        const long long t0 = ea_clock();
        const bool ok = EA_FAST_EVALS > 0 && L.evals >= EA_FAST_EVALS &&
                        tron::newton_step<N>(L.x, xl, xu, L.A, L.g, L.delta, L.alphac, L.prered(), L.g0(), L.snorm(), st);
        const long long t1 = ea_clock();
        if (ok) { L.cyc[4] += t1 - t0; L.cyc[7] += 1; }
        else {
            L.cyc[5] += t1 - t0;
            tron::compute_step<N>(L.x, xl, xu, L.A, L.g, L.delta, L.alphac, L.prered(), L.g0(), L.snorm(), st);
            L.cyc[6] += ea_clock() - t1;
        }
    }
#else
    tron::compute_step_auto<N>(L.x, xl, xu, L.A, L.g, L.delta, L.alphac, L.prered(), L.g0(), L.snorm(), st, L.evals);
#endif
    L.cg += st.cg;
    L.shifts += st.shifts;
    L.phase = TRIAL;
    L.step_pending = false;
}

// Solve one branch to completion on one lane (host harness and diagnostics).
template <class Eval>
EA_DEV void solve(Lane &L, const Eval &eval, const double (&xl)[N], const double (&xu)[N],
                  int max_auglag, double mu_max, const PowTable &T) {
    begin(L, T);
#pragma unroll 1
    for (;;) {
        if (eval_pass(L, eval, 0, xl, xu, max_auglag, mu_max, T)) break;
        eval_pass(L, eval, 1, xl, xu, max_auglag, mu_max, T);
        compute(L, xl, xu);
    }
}

}  // namespace branch
