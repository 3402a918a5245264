/*
 * oracle_internal.h — shared internals of the CPU ORACLE (test infrastructure, NOT product code):
 * the model struct and the generic TRON driver, used by acopf_oracle.c and mpacopf_oracle.c.
 */
#ifndef ORACLE_INTERNAL_H
#define ORACLE_INTERNAL_H

#include "acopf_oracle.h"

#define NV 6          /* branch sub-problem size with line limits (acopf_model.jl:46) */
#define MEMROWS 31    /* acopf_model.jl:87 */

struct orc_model {
    int64_t ngen, nline, nbus, nvar;
    double baseMVA;
    double *pgmin, *pgmax, *qgmin, *qgmax, *c2, *c1, *c0, *pgmin_curr, *pgmax_curr;
    double *YshR, *YshI, *Y[8];
    double *FrVmBound, *ToVmBound, *FrVaBound, *ToVaBound, *rateA;
    int64_t *FrStart, *FrIdx, *ToStart, *ToIdx, *GenStart, *GenIdx, *brBusIdx; /* 0-based */
    double *Pd, *Qd, *Vmin, *Vmax;
    double *vec[EA_NUM_FIELDS];
    double *membuf;
    int nthreads;
    ea_counters_t cnt;
    int32_t *eval_trace;   /* optional: per-line evaluation count of the last x-update (diagnostics) */
};

typedef struct { int64_t nfev, ngev, cg, shifts, rejected; } tron_stats_t;

/* objective callbacks of the generic TRON driver: A is n x n with leading dimension NV */
typedef double (*orc__f_fn)(const double *x, const void *ctx);
typedef void (*orc__gh_fn)(const double *x, const void *ctx, double *g, double *A);
int orc__tron_cb(int n, double *x, const double *xl, const double *xu, orc__f_fn evalf, orc__gh_fn evalgh,
                 const void *ctx, int max_feval, int max_minor, double gtol, int *minor_out, tron_stats_t *st);
/* same; g_out (may be NULL) receives the last gradient the driver evaluated (`tron.g` after solveProblem) */
int orc__tron_cb_g(int n, double *x, const double *xl, const double *xu, orc__f_fn evalf, orc__gh_fn evalgh,
                   const void *ctx, int max_feval, int max_minor, double gtol, int *minor_out, tron_stats_t *st,
                   double *g_out);

#endif
