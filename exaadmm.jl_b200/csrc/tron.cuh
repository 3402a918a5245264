// tron.cuh — bound-constrained trust-region Newton (TRON) for tiny dense
// problems, one problem per thread, everything in registers.
//
// Replaces the ExaTron.jl device routines the reference calls from its branch
// kernel (ExaTron.dtron / dcauchy / dspcg / dtrpcg / dicfs / dprsrch / dgpnorm;
// call sites /root/reference/src/models/acopf/acopf_tron_linelimit_kernel.jl:103-122).
// The algorithm is TRON 1.2 (Lin & More', SIOPT 1999) as summarised in
// SURVEY.md Appendix B. Design differences from the reference's 32-thread-block
// version:
//   * N is a template parameter, every loop over variables is fully unrolled and
//     every vector / the packed symmetric matrix live in registers;
//   * the free-variable sub-problem of dspcg is not compacted (no indfree
//     gather): fixed variables are masked, i.e. the reduced matrix is embedded
//     as diag(1) on the fixed rows. The free block sees exactly the same
//     operations in the same order (adding exact zeros), so iterates agree with
//     the compacted form;
//   * the reverse-communication protocol is flattened into "compute trial step"
//     and "judge trial step" so that a warp's lanes, each at a different stage
//     of its own solve, still execute the same code (branch.cuh);
//   * "compute trial step" has two forms: the literal algorithm (compute_step: dcauchy, dspcg with dicfs / dtrpcg /
//     dprsrch) and newton_step, which takes the step directly where the literal algorithm runs its common course
//     (> 99.9 % of the steps) and declines otherwise (compute_step_auto);
//   * matrices are read through at(i, j), so a caller can hand in a structured type (branch::Hess).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// Device code that can also be compiled for the host: tests/host_harness.cu runs
// these exact routines on the CPU against the oracle (no GPU needed).
#ifndef EA_DEV
#define EA_DEV __host__ __device__ __forceinline__
#endif
// EA_NO_FMA (host test harness only) selects "exact" arithmetic: every fused
// multiply-add becomes the separately rounded a*b + c and the triangular solves /
// Cholesky keep true divisions and sqrt, which makes the arithmetic bit-identical
// to the oracle's (compiled with -ffp-contract=off), so the control logic can be
// compared exactly. The device build uses FMA and reciprocal / rsqrt forms
// (fewer instructions on the critical path, results equal to within rounding).
#ifdef EA_NO_FMA
#define EA_FMA(a, b, c) ((a) * (b) + (c))
#define EA_EXACT 1
#else
#define EA_FMA(a, b, c) fma((a), (b), (c))
#ifdef EA_EXACT_DIV
#define EA_EXACT 1
#else
#define EA_EXACT 0
#endif
#endif

namespace tron {

// EA_STATS (host analysis build only, tools/step_stats.py): why newton_step declines, counted per reason.
#ifdef EA_STATS
static long long g_stat[2][16];
static int g_stat_first = 0;
#endif
#if defined(EA_STATS) && !defined(__CUDA_ARCH__)
#define EA_STAT(k) (tron::g_stat[tron::g_stat_first][k]++)
#else
#define EA_STAT(k) ((void)0)
#endif

// ---- scalar helpers ------------------------------------------------------------------
// min / max without the NaN bookkeeping of fmin / fmax (inputs are finite)
EA_DEV double dmin(double a, double b) { return a < b ? a : b; }
EA_DEV double dmax(double a, double b) { return a > b ? a : b; }

// Division, reciprocal square root and square root. The device build uses the hardware
// approximations refined by Newton steps (full double precision to within an ulp, no
// slow-path branches: operands here are finite, normal and of the right sign by
// construction); EA_EXACT and the host use the correctly rounded operations.
EA_DEV double ddiv(double a, double b) {
#if defined(__CUDA_ARCH__) && !EA_EXACT
    // PRECONDITION: b is finite, normal and 1 / b does not overflow (rcp.approx flushes subnormals to zero; the Newton
    // steps then produce NaN where IEEE division gives a huge value). Every call site divides by a quantity that is
    // bounded away from zero by construction - curvatures p'Ap > 0, rho > 0, beta + rho, squared norms behind an
    // explicit > 0 test, gradient components behind a != 0 test (a SUBNORMAL gradient component would need |g_i| <
    // 2.3e-308 with g scaled by 1e-4 ... 1e-5 of O(1) data) - except the bus update of a bus without branch ends, which
    // is guarded there (kernels.cuh: bus_solve). A range check here costs 14 % of the branch kernel (measured).
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(fma(-b, r, 1.0), r, r);
    r = fma(fma(-b, r, 1.0), r, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
#else
    return a / b;
#endif
}
EA_DEV double drsqrt(double a) {            // a > 0
#if defined(__CUDA_ARCH__) && !EA_EXACT
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    const double h = 0.5 * a;
    r = fma(fma(-h * r, r, 0.5), r, r);
    r = fma(fma(-h * r, r, 0.5), r, r);
    return r;
#else
    return 1.0 / sqrt(a);
#endif
}
EA_DEV double dsqrt(double a) {             // a >= 0
#if defined(__CUDA_ARCH__) && !EA_EXACT
    if (!(a > 0.0)) return 0.0;
    const double r = drsqrt(a);
    const double s = a * r;
    return fma(fma(-s, s, a), 0.5 * r, s);
#else
    return sqrt(a);
#endif
}

__host__ __device__ constexpr int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// Symmetric matrices are read through at(i, j) (compile-time indices in the unrolled loops), so that a problem whose
// Hessian has structure can hand in a type that stores only its independent entries (branch::Hess: 15 of 21).
template <int N> struct Sym {                                 // packed lower triangle
    double a[N * (N + 1) / 2];
    EA_DEV double at(int i, int j) const { return a[tri(i, j)]; }
};
// Cholesky factor: packed lower triangle + reciprocal diagonal (so that the many
// triangular solves multiply instead of divide)
template <int N> struct Chol { double a[N * (N + 1) / 2]; double rd[N]; };

template <int N> EA_DEV double dot(const double (&x)[N], const double (&y)[N]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) s = EA_FMA(x[i], y[i], s);
    return s;
}
template <int N> EA_DEV double nrm2(const double (&x)[N]) { return dsqrt(dot<N>(x, x)); }
// ||x|| <= bound (bound >= 0) without the square root, for the stopping tests of the CG: they sit on the serial chain of
// the step (a Newton-refined sqrt is ~12 dependent operations). EA_EXACT keeps the root so that the host harness stays
// bit-identical to the oracle. (Not for the Cauchy search: at START delta IS ||g||, an exact tie.)
EA_DEV bool norm_le(double sumsq, double bound) {
#if EA_EXACT
    return sqrt(sumsq) <= bound;
#else
    return sumsq <= bound * bound;
#endif
}

// y = A x  (A packed symmetric)
template <int N, class M> EA_DEV void symv(const M &A, const double (&x)[N], double (&y)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) s = EA_FMA(A.at(i, j), x[j], s);
        y[i] = s;
    }
}

template <int N> EA_DEV void mid(double (&x)[N], const double (&xl)[N], const double (&xu)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = dmax(xl[i], dmin(x[i], xu[i]));
}

// s = P[x + alpha*w] - x   (Appendix B.1 dgpstep)
template <int N> EA_DEV void gpstep(const double (&x)[N], const double (&xl)[N], const double (&xu)[N],
                                                        double alpha, const double (&w)[N], double (&s)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double aw = alpha * w[i];
        const double xt = x[i] + aw;
        s[i] = (xt < xl[i]) ? (xl[i] - x[i]) : ((xt > xu[i]) ? (xu[i] - x[i]) : aw);
    }
}

// break points along w (B.1 dbreakpt). sgn = +1 uses w, sgn = -1 uses -w. Branch-free: the N
// quotients are independent chains the scheduler can overlap.
template <int N> EA_DEV void breakpt(const double (&x)[N], const double (&xl)[N], const double (&xu)[N],
                                                         const double (&w)[N], double sgn, double &brptmin, double &brptmax) {
    double b[N];
    bool has[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double wi = sgn * w[i];
        const bool up = (x[i] < xu[i]) && (wi > 0.0);
        const bool dn = (x[i] > xl[i]) && (wi < 0.0);
        has[i] = up || dn;
        b[i] = ddiv((up ? xu[i] : xl[i]) - x[i], has[i] ? wi : 1.0);
    }
    bool any = false;
    brptmin = 0.0; brptmax = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double lo = any ? dmin(brptmin, b[i]) : b[i];
        const double hi = any ? dmax(brptmax, b[i]) : b[i];
        brptmin = has[i] ? lo : brptmin;
        brptmax = has[i] ? hi : brptmax;
        any = any || has[i];
    }
}

// projected gradient sup-norm (B.1 dgpnorm)
template <int N> EA_DEV double gpnorm(const double (&x)[N], const double (&xl)[N], const double (&xu)[N],
                                                          const double (&g)[N]) {
    double nrm = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double at_l = fabs(dmin(g[i], 0.0)), at_u = fabs(dmax(g[i], 0.0)), in = fabs(g[i]);
        const double v = (x[i] == xl[i]) ? at_l : ((x[i] == xu[i]) ? at_u : in);
        nrm = (xl[i] != xu[i]) ? dmax(nrm, v) : nrm;
    }
    return nrm;
}

// B.1 dtrqsol
template <int N> EA_DEV double trqsol(const double (&x)[N], const double (&p)[N], double delta) {
    const double ptx = dot<N>(p, x), ptp = dot<N>(p, p), xtx = dot<N>(x, x);
    const double dsq = delta * delta;
    const double rad = dsqrt(dmax(EA_FMA(ptx, ptx, ptp * (dsq - xtx)), 0.0));
    const bool first = ptx > 0.0, second = rad > 0.0;
    const double num = first ? (dsq - xtx) : (rad - ptx);
    const double den = first ? (ptx + rad) : (second ? ptp : 1.0);
    const double q = ddiv(num, den);
    return (first || second) ? q : 0.0;
}

// q(s) = 0.5 s'As + g's and g's
template <int N, class M> EA_DEV void quad(const M &A, const double (&g)[N], const double (&s)[N],
                                                      double &q, double &gts) {
    double w[N];
    symv<N>(A, s, w);
    gts = dot<N>(g, s);
    q = EA_FMA(0.5, dot<N>(s, w), gts);
}

// Cauchy step (B.2 dcauchy). Returns the new alpha; s is the step. One loop with a
// single projected-step / quadratic-model site (the three phases of the original:
// first trial, interpolation, extrapolation) to keep the instruction footprint small.
template <int N, class M> EA_DEV double cauchy(const double (&x)[N], const double (&xl)[N], const double (&xu)[N],
                                                          const M &A, const double (&g)[N], double delta,
                                                          double alpha, double (&s)[N]) {
    const double mu0 = 0.01, interpf = 0.1, extrapf = 10.0;
    double brptmin, brptmax;
    breakpt<N>(x, xl, xu, g, -1.0, brptmin, brptmax);
    int mode = 0;                 // 0 first trial, 1 interpolating, 2 extrapolating, 3 recover last good, 4 done
    double alphas = alpha;
#pragma unroll 1
    while (mode != 4) {
        gpstep<N>(x, xl, xu, -alpha, g, s);
        if (mode == 3) break;
        const bool within = nrm2<N>(s) <= delta;
        double q, gts;
        quad<N>(A, g, s, q, gts);                              // only used when `within`; cheaper than a branch
        if (mode == 0) {
            const bool interp = !within || (q >= mu0 * gts);
            if (interp) { mode = 1; alpha = interpf * alpha; }
            else {
                alphas = alpha;
                if (alpha <= brptmax) { mode = 2; alpha = extrapf * alpha; }
                else mode = 3;                                   // loop body of the original never runs
            }
        } else if (mode == 1) {
            if (within && !(q > mu0 * gts)) mode = 4;
            else alpha = interpf * alpha;
        } else {                                                 // mode 2
            bool search = true;
            if (within) { if (q < mu0 * gts) alphas = alpha; }
            else search = false;
            if (search && alpha <= brptmax) alpha = extrapf * alpha;
            else { alpha = alphas; mode = 3; }
        }
    }
    return alpha;
}

// Lower-triangular solves with the free-set mask folded into L (fixed rows are e_i).
template <int N> EA_DEV void lsolve(const Chol<N> &L, double (&r)[N]) {      // L r = b
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s = r[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s = EA_FMA(-L.a[tri(i, k)], r[k], s);
#if EA_EXACT
        r[i] = s / L.a[tri(i, i)];
#else
        r[i] = s * L.rd[i];
#endif
    }
}
template <int N> EA_DEV void ltsolve(const Chol<N> &L, double (&r)[N]) {     // L' r = b
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        double s = r[i];
#pragma unroll
        for (int k = i + 1; k < N; ++k) s = EA_FMA(-L.a[tri(k, i)], r[k], s);
#if EA_EXACT
        r[i] = s / L.a[tri(i, i)];
#else
        r[i] = s * L.rd[i];
#endif
    }
}

// In-place dense Cholesky of the packed lower triangle; false if a pivot is <= 0.
template <int N> EA_DEV bool cholesky(Chol<N> &L) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double d = L.a[tri(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) d = EA_FMA(-L.a[tri(j, k)], L.a[tri(j, k)], d);
        ok = ok && (d > 0.0);
#if EA_EXACT
        d = sqrt(d);
        L.a[tri(j, j)] = d;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double s = L.a[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) s = EA_FMA(-L.a[tri(i, k)], L.a[tri(j, k)], s);
            L.a[tri(i, j)] = s / d;
        }
#else
        const double r = drsqrt(d);
        L.a[tri(j, j)] = d * r;
        L.rd[j] = r;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double s = L.a[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) s = EA_FMA(-L.a[tri(i, k)], L.a[tri(j, k)], s);
            L.a[tri(i, j)] = s * r;
        }
#endif
    }
    return ok;
}

// B.5 dicfs for a dense matrix: scaled Cholesky of the masked matrix with the shift
// search; the general algorithm (non-positive diagonal, failed pivots, shifts). Kept
// out of line: it runs for a handful of factorizations per million.
template <int N> __host__ __device__ __noinline__ void icfs_general(const Sym<N> &A, unsigned freemask, Chol<N> &L) {
    const double alpham = 1e-3, nbfactor = 512.0;
    const int nbmax = 3;
    double wa2[N];        // the scaling D = diag(1/sqrt(a_ii))
    double alphas = alpham, alpha = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const bool fr = (freemask >> i) & 1u;
        const double aii = A.at(i, i);
        wa2[i] = 1.0;
        if (fr && aii > 0.0) {
#if EA_EXACT
            wa2[i] = 1.0 / sqrt(aii);
#else
            wa2[i] = drsqrt(aii);
#endif
        } else if (fr) {                       // non-positive diagonal entry: scale by the column norm
            double cs = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double akj = ((freemask >> k) & 1u) ? A.at(k, i) : 0.0;
                cs = EA_FMA(akj, akj, cs);
            }
            wa2[i] = 1.0 / sqrt(sqrt(cs));
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if ((freemask >> i) & 1u) {
            const double aii = A.at(i, i);
            if (aii == 0.0) alpha = alphas;
            else alpha = dmax(alpha, -aii * (wa2[i] * wa2[i]));
        }
    }
    if (alpha > 0.0) alpha = dmax(alpha, alphas);
    int nb = 1;
#pragma unroll 1
    for (;;) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const bool fi = (freemask >> i) & 1u;
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const bool fj = (freemask >> j) & 1u;
                double v;
                if (fi && fj) v = A.at(i, j) * wa2[i] * wa2[j] + ((i == j) ? alpha : 0.0);
                else v = (i == j) ? 1.0 : 0.0;
                L.a[tri(i, j)] = v;
            }
        }
        if (cholesky<N>(L)) {
            if (alpha == alphas && nb < nbmax) { alphas = alphas / nbfactor; alpha = alphas; nb++; }
            else {
                // undo the scaling: L <- D^-1 L
#pragma unroll
                for (int i = 0; i < N; ++i) {
#pragma unroll
                    for (int j = 0; j <= i; ++j) L.a[tri(i, j)] = L.a[tri(i, j)] / wa2[i];
                    L.rd[i] = L.rd[i] * wa2[i];
                }
                break;
            }
        } else alpha = dmax(2.0 * alpha, alphas);
    }
}

// dicfs, common case first: every free diagonal entry is positive and the unshifted
// factorization succeeds (then the general algorithm does exactly this single attempt).
// Returns true if the general path (shift search) had to run.
template <int N, class M> EA_DEV bool icfs(const M &A, unsigned freemask, Chol<N> &L) {
    double wa2[N];
    bool pos = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const bool fr = (freemask >> i) & 1u;
        const double aii = A.at(i, i);
        pos = pos && (!fr || aii > 0.0);
#if EA_EXACT
        wa2[i] = fr ? 1.0 / sqrt(aii) : 1.0;
#else
        wa2[i] = fr ? drsqrt(aii) : 1.0;
#endif
    }
    if (pos) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const bool fi = (freemask >> i) & 1u;
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const bool fj = (freemask >> j) & 1u;
                L.a[tri(i, j)] = (fi && fj) ? A.at(i, j) * wa2[i] * wa2[j] + 0.0 : ((i == j) ? 1.0 : 0.0);
            }
        }
        if (cholesky<N>(L)) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
#if EA_EXACT
#pragma unroll
                for (int j = 0; j <= i; ++j) L.a[tri(i, j)] = L.a[tri(i, j)] / wa2[i];
#else
                const bool fi = (freemask >> i) & 1u;
                const double si = fi ? A.at(i, i) * wa2[i] : 1.0;     // sqrt(a_ii) = a_ii * rsqrt(a_ii)
#pragma unroll
                for (int j = 0; j <= i; ++j) L.a[tri(i, j)] = L.a[tri(i, j)] * si;
                L.rd[i] = L.rd[i] * wa2[i];
#endif
            }
            return false;
        }
    }
    {   // rare: go through addressable copies so that A and L themselves stay in registers
        Sym<N> A2;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) A2.a[tri(i, j)] = A.at(i, j);
        Chol<N> L2;
        icfs_general<N>(A2, freemask, L2);
        L = L2;
    }
    return true;
}

// B.4 dtrpcg on the masked system. w is the solution in the L-transformed space.
template <int N, class M> EA_DEV void trpcg(const M &A, unsigned freemask, const double (&g)[N], double delta,
                                                       const Chol<N> &L, double tol, double stol, int itermax,
                                                       double (&w)[N], int &iters, int &info) {
    double t[N], r[N], p[N], q[N], z[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { w[i] = 0.0; t[i] = -g[i]; r[i] = t[i]; }
    lsolve<N>(L, r);
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = r[i];
    double rho = dot<N>(r, r);
    if (rho == 0.0) { iters = 0; info = 1; return; }
    iters = itermax; info = 5;
#pragma unroll 1
    for (int it = 1; it <= itermax; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) z[i] = p[i];
        ltsolve<N>(L, z);
        symv<N>(A, z, q);
#pragma unroll
        for (int i = 0; i < N; ++i) { q[i] = ((freemask >> i) & 1u) ? q[i] : 0.0; z[i] = q[i]; }
        lsolve<N>(L, q);
        const double ptq = dot<N>(p, q);
        const double alpha = (ptq > 0.0) ? ddiv(rho, ptq) : 0.0;
        const double sigma = trqsol<N>(w, p, delta);
        if (ptq <= 0.0 || alpha >= sigma) {
#pragma unroll
            for (int i = 0; i < N; ++i) w[i] = EA_FMA(sigma, p[i], w[i]);
            iters = it; info = (ptq <= 0.0) ? 3 : 4;
            return;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) { w[i] = EA_FMA(alpha, p[i], w[i]); r[i] = EA_FMA(-alpha, q[i], r[i]); t[i] = EA_FMA(-alpha, z[i], t[i]); }
        const double rtr = dot<N>(r, r);
        if (norm_le(dot<N>(t, t), tol)) { iters = it; info = 1; return; }
        if (norm_le(rtr, stol)) { iters = it; info = 2; return; }
        const double beta = ddiv(rtr, rho);
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = EA_FMA(beta, p[i], r[i]);
        rho = rtr;
    }
}

// B.6 dprsrch on the masked system (w is zero on fixed variables).
template <int N, class M> EA_DEV void prsrch(double (&x)[N], const double (&xl)[N], const double (&xu)[N],
                                                        const M &A, const double (&g)[N], double (&w)[N]) {
    const double mu0 = 0.01, interpf = 0.5;
    double wa1[N], brptmin, brptmax, alpha = 1.0, q, gts;
    bool search = true;
    breakpt<N>(x, xl, xu, w, 1.0, brptmin, brptmax);
#pragma unroll 1
    while (search && alpha > brptmin) {
        gpstep<N>(x, xl, xu, alpha, w, wa1);
        quad<N>(A, g, wa1, q, gts);
        if (q <= mu0 * gts) search = false;
        else alpha = interpf * alpha;
    }
    if (alpha < 1.0 && alpha < brptmin) alpha = brptmin;
    gpstep<N>(x, xl, xu, alpha, w, wa1);
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = EA_FMA(alpha, w[i], x[i]);
    mid<N>(x, xl, xu);
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = wa1[i];
}

struct Stats { int cg = 0; int shifts = 0; };

// B.3 dspcg: from the Cauchy step s, move x to the trial point; s becomes x_trial - x_0.
template <int N, class M> EA_DEV void spcg(double (&x)[N], const double (&xl)[N], const double (&xu)[N],
                                                      const M &A, const double (&g)[N], double delta, double rtol,
                                                      double (&s)[N], int itermax, Stats &st) {
    double w[N];
    symv<N>(A, s, w);
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] += s[i];
    mid<N>(x, xl, xu);
    int iters = 0;
#pragma unroll 1
    for (int nfaces = 1; nfaces <= N; ++nfaces) {
        unsigned freemask = 0;
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (xl[j] < x[j] && x[j] < xu[j]) freemask |= (1u << j);
        if (freemask == 0) return;
        Chol<N> L;
        if (icfs<N>(A, freemask, L)) st.shifts++;
        double gfree[N], wf[N];
        double gsq = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const bool fr = (freemask >> j) & 1u;
            gfree[j] = fr ? (w[j] + g[j]) : 0.0;
            const double gj = fr ? g[j] : 0.0;
            gsq = EA_FMA(gj, gj, gsq);
        }
        const double gfnorm = dsqrt(gsq);
        int itertr, infotr;
        trpcg<N>(A, freemask, gfree, delta, L, rtol * gfnorm, 0.0, itermax, wf, itertr, infotr);
        iters += itertr;
        st.cg += itertr;
        ltsolve<N>(L, wf);
        prsrch<N>(x, xl, xu, A, gfree, wf);
#pragma unroll
        for (int j = 0; j < N; ++j) s[j] += wf[j];
        symv<N>(A, s, w);
        double gf2 = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double v = ((freemask >> j) & 1u) ? (w[j] + g[j]) : 0.0;
            gf2 = EA_FMA(v, v, gf2);
        }
        if (norm_le(gf2, rtol * gfnorm)) return;
        if (infotr == 3 || infotr == 4) return;
        if (iters > itermax) return;
    }
}

// One "COMPUTE" of dtron (B.0): Cauchy point + projected CG. On entry x is the
// current iterate; on exit x is the trial point, prered the predicted reduction,
// gts = g's and snorm = |s| for the step s = trial - x_c.
template <int N, class M> EA_DEV void compute_step(double (&x)[N], const double (&xl)[N], const double (&xu)[N],
                                                              const M &A, const double (&g)[N], double delta,
                                                              double &alphac, double &prered, double &gts, double &snorm,
                                                              Stats &st) {
    const double cgtol = 0.1;
    double s[N];
    alphac = cauchy<N>(x, xl, xu, A, g, delta, alphac, s);
    spcg<N>(x, xl, xu, A, g, delta, cgtol, s, N, st);
    double q;
    quad<N>(A, g, s, q, gts);
    prered = -q;
    snorm = nrm2<N>(s);
}

// Plain Cholesky of the masked matrix (fixed rows / columns replaced by the identity), no scaling, no shift search.
// false if a pivot is not positive.
template <int N, class M> EA_DEV bool chol_masked(const M &A, unsigned freemask, Chol<N> &L) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const bool fi = (freemask >> i) & 1u;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            const bool fj = (freemask >> j) & 1u;
            L.a[tri(i, j)] = (fi && fj) ? A.at(i, j) : ((i == j) ? 1.0 : 0.0);
        }
    }
    return cholesky<N>(L);
}

// One dtron COMPUTE taken directly, where it runs its common course.
//
// On this problem class the Hessian restricted to the free variables is positive definite on all but a handful of steps
// per million, so the unshifted Cholesky factor is an EXACT preconditioner and the conjugate-gradient loop of a face
// ends after ONE step - the Newton correction w = -A_FF^-1 (A s + g)_F if it stays inside the trust region, the same
// direction cut at the boundary otherwise (info = 4). newton_step computes that course without the machinery around it:
//   * the Cauchy search on scalars while the trial stays on the straight part of the projected path (up to the first
//     break point s(alpha) = -alpha gh, so |s| = alpha |gh|, g's = -alpha gh'gh, q = alpha (alpha/2 gh'A gh - gh'gh):
//     one matrix-vector product for the whole search); trials beyond the first break point as dcauchy evaluates them;
//   * per face of dspcg one masked Cholesky factorization without scaling or shift search, two triangular solves, the
//     trust-region cut; the projected search reduced to "the full step fits" where no break point of w lies below 1
//     (then dprsrch does nothing else), dprsrch itself otherwise; dspcg's own exits (face optimal, boundary step, next
//     face after a clipped step, at most three faces).
// Every condition the literal algorithm would have tested on the way is re-checked; if one fails (pivot <= 0, zero
// residual, an unclipped step that leaves a residual - the CG loop would go on -, a fourth face) newton_step returns
// false WITHOUT touching its arguments and the caller runs the literal algorithm (compute_step). It covers > 99.9 % of
// the steps of a solve of the BASELINE grids (tools/step_stats.py), first steps of a branch included. Results agree
// with the literal path to rounding (the same vectors computed with fewer operations), decisions except within rounding
// of a tie (tests/test_device_code_on_host.py::test_direct_step_follows_the_literal_algorithm); EA_EXACT builds never
// take it.
template <int N, class M> EA_DEV bool newton_step(double (&x)[N], const double (&xl)[N], const double (&xu)[N],
                                         const M &A, const double (&g)[N], double delta,
                                         double &alphac, double &prered, double &gts, double &snorm, Stats &st) {
#if EA_EXACT
    return false;
#else
    const double mu0 = 0.01, interpf = 0.1, extrapf = 10.0, cgtol = 0.1;
    // ---- Cauchy search in closed form on the straight part of the projected path: up to the first break point the
    // projected step is s(alpha) = -alpha gh (gh = g on the variables that can move, 0 on those held by a bound), so
    // |s| = alpha |gh|, g's = -alpha gh'gh, q(s) = alpha (alpha/2 gh'A gh - gh'gh): the trials of dcauchy are scalar ----
    double gh[N], bp[N];
    bool has[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double wi = -g[i];
        const bool up = (x[i] < xu[i]) && (wi > 0.0);
        const bool dn = (x[i] > xl[i]) && (wi < 0.0);
        has[i] = up || dn;
        bp[i] = ddiv((up ? xu[i] : xl[i]) - x[i], has[i] ? wi : 1.0);
        gh[i] = has[i] ? g[i] : 0.0;
    }
    bool any = false;
    double brptmin = 0.0, brptmax = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double lo = any ? dmin(brptmin, bp[i]) : bp[i];
        const double hi = any ? dmax(brptmax, bp[i]) : bp[i];
        brptmin = has[i] ? lo : brptmin;
        brptmax = has[i] ? hi : brptmax;
        any = any || has[i];
    }
    if (!any) { EA_STAT(0); return false; }
    double Ag[N];
    symv<N>(A, gh, Ag);
    const double gg = dot<N>(gh, gh), gAg = dot<N>(gh, Ag);
    const double gnorm = dsqrt(gg);
    double alpha = alphac, alphas = alphac;
    int mode = 0;
#pragma unroll 1
    while (mode < 3) {
        bool within;
        double tgts, q;
        if (alpha < brptmin) {                               // on the straight part: scalars
            within = alpha * gnorm <= delta;
            tgts = -(alpha * gg);
            q = alpha * EA_FMA(0.5 * alpha, gAg, -gg);
        } else {                                             // a trial beyond the first break point: as dcauchy does it
            EA_STAT(1);
            double sp[N];
            gpstep<N>(x, xl, xu, -alpha, g, sp);
            within = nrm2<N>(sp) <= delta;
            quad<N>(A, g, sp, q, tgts);
        }
        if (mode == 0) {
            const bool interp = !within || (q >= mu0 * tgts);
            if (interp) { mode = 1; alpha = interpf * alpha; }
            else {
                alphas = alpha;
                if (alpha <= brptmax) { mode = 2; alpha = extrapf * alpha; }
                else mode = 3;
            }
        } else if (mode == 1) {
            if (within && !(q > mu0 * tgts)) mode = 4;
            else alpha = interpf * alpha;
        } else {
            bool search = true;
            if (within) { if (q < mu0 * tgts) alphas = alpha; }
            else search = false;
            if (search && alpha <= brptmax) alpha = extrapf * alpha;
            else { alpha = alphas; mode = 3; }
        }
    }
    // ---- the Cauchy step s, A s and the point after it (dspcg) ----
    double sv[N], As[N], x1[N];
    if (alpha < brptmin) {
#pragma unroll
        for (int i = 0; i < N; ++i) { sv[i] = -alpha * gh[i]; As[i] = -alpha * Ag[i]; }
    } else {                                                 // the Cauchy point lies beyond the first break point
        EA_STAT(2);
        gpstep<N>(x, xl, xu, -alpha, g, sv);
        symv<N>(A, sv, As);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) x1[i] = dmax(xl[i], dmin(x[i] + sv[i], xu[i]));
    // ---- the faces of dspcg, each taken in one go: exact Cholesky factor of A on the free variables -> the conjugate
    // gradient loop is ONE step, the Newton correction w = -A_FF^-1 (A s + g)_F, cut at the trust-region boundary if it
    // leaves it; projected search with its first trial (alpha = 1): the full step, clipped where a variable reaches a
    // bound, if it passes the decrease test. Then dspcg's own tests: face optimal -> done; boundary step -> done; a
    // clipped step leaves a smaller face -> once more. Anything else (pivot <= 0, zero residual, decrease test failed,
    // an unclipped interior step that leaves a residual, more than three faces) is left to the literal algorithm. ----
    const double dsq = delta * delta;
    int ncg = 0;
    bool done = false;
#pragma unroll 1
    for (int face = 0; face < 3; ++face) {
        unsigned freemask = 0;
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (xl[i] < x1[i] && x1[i] < xu[i]) freemask |= (1u << i);
        if (freemask == 0) {
            if (face == 0) { EA_STAT(3); return false; }
            done = true;                                     // dspcg returns with the step it has
            break;
        }
        Chol<N> L;
        if (!chol_masked<N>(A, freemask, L)) { EA_STAT(4); return false; }   // not positive definite: dicfs decides
        double gfree[N], w[N];
        double gsq = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const bool fr = (freemask >> i) & 1u;
            gfree[i] = fr ? (As[i] + g[i]) : 0.0;
            const double gi = fr ? g[i] : 0.0;
            gsq = EA_FMA(gi, gi, gsq);
            w[i] = -gfree[i];
        }
        lsolve<N>(L, w);
        const double pp = dot<N>(w, w);                       // |p|^2, p = L^-1 (-gfree): the first CG direction
        if (!(pp > 0.0)) { EA_STAT(5); return false; }       // zero residual: dtrpcg's own exit
        // dtrpcg, first iteration: with the exact factor p'(L^-1 A L^-T)p = p'p, so alpha = 1 and the step is p unless
        // it leaves the trust region (alpha >= sigma = delta / |p|): then the step is sigma p (info = 4)
        const bool boundary = !(pp < dsq);
        if (boundary) {
            const double sigma = delta * drsqrt(pp);         // dtrqsol with w = 0: delta / |p|
#pragma unroll
            for (int i = 0; i < N; ++i) w[i] = sigma * w[i];
        }
        ltsolve<N>(L, w);
        // dprsrch: no break point of w below 1 -> the full step (what dprsrch does then, without its divisions);
        // otherwise dprsrch itself
        bool inside = true;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const bool fr = (freemask >> i) & 1u;
            const double wi = fr ? w[i] : 0.0;
            const bool ok = (wi > 0.0) ? (xu[i] - x1[i] >= wi) : ((wi < 0.0) ? (xl[i] - x1[i] <= wi) : true);
            inside = inside & ok;
            w[i] = wi;
        }
        if (inside) {
#pragma unroll
            for (int i = 0; i < N; ++i) x1[i] = dmax(xl[i], dmin(x1[i] + w[i], xu[i]));
        } else {
            EA_STAT(6);
            prsrch<N>(x1, xl, xu, A, gfree, w);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) sv[i] += w[i];
        symv<N>(A, sv, As);
        ncg++;
        double gf2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double v = ((freemask >> i) & 1u) ? (As[i] + g[i]) : 0.0;
            gf2 = EA_FMA(v, v, gf2);
        }
        if (gf2 <= (cgtol * cgtol) * gsq) { done = true; break; }     // the face is optimal
        if (boundary) { done = true; break; }                          // dspcg returns after a boundary step
        if (inside) { EA_STAT(7); return false; }                      // the CG loop would go on
    }
    if (!done) { EA_STAT(11); return false; }
    // ---- accept: outputs of compute_step ----
    gts = dot<N>(g, sv);
    prered = -EA_FMA(0.5, dot<N>(sv, As), gts);
    snorm = nrm2<N>(sv);
    alphac = alpha;
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = x1[i];
    st.cg += ncg;
    EA_STAT(ncg > 1 ? 9 : 8);
    return true;
#endif
}

// compute_step, with the direct step tried first once the problem has needed EA_FAST_EVALS evaluations (1: on every step;
// 0: never, the literal algorithm alone - tests). A rule on the problem's own history only.
#ifndef EA_FAST_EVALS
#define EA_FAST_EVALS 1
#endif
template <int N, class M> EA_DEV void compute_step_auto(double (&x)[N], const double (&xl)[N], const double (&xu)[N],
                                               const M &A, const double (&g)[N], double delta,
                                               double &alphac, double &prered, double &gts, double &snorm,
                                               Stats &st, int evals_so_far) {
#if defined(EA_STATS) && !defined(__CUDA_ARCH__)
    g_stat_first = (evals_so_far < 2) ? 1 : 0;        // [1]: the first step of a problem
#endif
    if (EA_FAST_EVALS > 0 && (EA_FAST_EVALS == 1 || evals_so_far >= EA_FAST_EVALS) &&
        newton_step<N>(x, xl, xu, A, g, delta, alphac, prered, gts, snorm, st)) return;
    compute_step<N>(x, xl, xu, A, g, delta, alphac, prered, gts, snorm, st);
}

// The "EVALUATE" part of dtron (B.0): trust-region update and acceptance.
// Returns: 0 = rejected (task F), 1 = accepted (task GH), 2 = converged/warn.
EA_DEV int judge_step(double f_trial, double fc, double g0, double snorm, double prered,
                                          bool first_iter, double &delta, bool &accepted) {
    const double eta0 = 1e-4, eta1 = 0.25, eta2 = 0.75, sigma1 = 0.25, sigma2 = 0.5, sigma3 = 4.0;
    const double frtol = 1e-12, fatol = 0.0, fmin_ = -1e32;
    const double actred = fc - f_trial;
    if (first_iter) delta = dmin(delta, snorm);
    const double den = f_trial - fc - g0;
    const double alpha = (den <= 0.0) ? sigma3 : dmax(sigma1, -0.5 * ddiv(g0, den));
    const double as = alpha * snorm;
    const double d0 = dmin(dmax(alpha, sigma1) * snorm, sigma2 * delta);
    const double d1 = dmax(sigma1 * delta, dmin(as, sigma2 * delta));
    const double d2 = dmax(sigma1 * delta, dmin(as, sigma3 * delta));
    const double d3 = dmax(delta, dmin(as, sigma3 * delta));
    delta = (actred < eta0 * prered) ? d0 : ((actred < eta1 * prered) ? d1 : ((actred < eta2 * prered) ? d2 : d3));
    accepted = actred > eta0 * prered;
    const double f = accepted ? f_trial : fc;
    int task = accepted ? 1 : 0;
    if (f < fmin_) task = 2;
    if (fabs(actred) <= fatol && prered <= fatol) task = 2;
    if (fabs(actred) <= frtol * fabs(f) && prered <= frtol * fabs(f)) task = 2;
    return task;
}

}  // namespace tron
