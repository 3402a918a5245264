/* TEST INFRASTRUCTURE (part of the CPU oracle). An optional replacement of libm's sin / cos in the branch objective:
 * the PARITY build of the CUDA library (exaadmm.jl_b200/csrc/branch.cuh, -DEA_PARITY) evaluates the same formulas in
 * the same order, so that oracle and device then round every operation of a branch solve identically and their
 * iterates can be compared bit for bit (tests/test_gpu_baseline_configs.py). Default off: the oracle uses libm.
 * libm's and libdevice's sin / cos differ in the last bit for a few percent of the arguments, Julia's own (the
 * reference's) from both - enough to flip a trust-region decision that sits within rounding of its threshold. */
#pragma once
#include <math.h>

/* Portable sin / cos of a double: Cody-Waite reduction by pi/2 in three parts, then the polynomial kernels of
 * fdlibm (k_sin.c / k_cos.c coefficients) in Horner form. Every operation is an individually rounded IEEE
 * multiply or add (compile without contraction), so the result is the same bits on any IEEE machine. Valid for
 * |t| < ~1e5 (the reduction is exact enough there); accuracy ~1 ulp. */
static inline void psincos(double t, double *sn, double *cs) {
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
    switch (q & 3) {
    case 0: *sn = s; *cs = c; break;
    case 1: *sn = c; *cs = -s; break;
    case 2: *sn = -s; *cs = -c; break;
    default: *sn = -c; *cs = s; break;
    }
}
