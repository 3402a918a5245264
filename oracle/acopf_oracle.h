/*
 * acopf_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's `use_gpu=false` path for the
 * two-level ADMM ACOPF inner loop. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (exaadmm.jl_b200/) never does.
 *
 * Pinning status: PINNED at path level by the reference's own golden vectors
 * (test/algorithms/acopf_update_cpu.jl:28-151) and its end-to-end known answer
 * (case9: Solved / outer 20 / cumul 705 / objval 5303.435, :168-172), see
 * tests/test_oracle_golden.py. The TRON solver itself lives in the un-vendored
 * dependency ExaTron.jl ("3", Project.toml:20) and is restated from the
 * published TRON 1.2 algorithm (Lin & More', SIOPT 1999): at the ExaTron unit
 * level parity is UNPINNED (no reference test isolates dtron); the Cauchy
 * extrapolation branch and the Cholesky shift loop are not exercised by case9.
 */
#ifndef ACOPF_ORACLE_H
#define ACOPF_ORACLE_H

#include "../include/exaadmm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_model orc_model_t;

int    orc_create(const ea_grid_t *grid, orc_model_t **out);
void   orc_destroy(orc_model_t *m);
void   orc_set_threads(orc_model_t *m, int nthreads);   /* 1 = faithful serial loops */
/* 1: the branch objective uses the portable sin / cos of portable_sincos.h (process-wide), 0: libm (default) */
void   orc_set_portable_sincos(int on);
int    orc_get_threads(const orc_model_t *m);

void   orc_init_solution(orc_model_t *m, double rho_pq, double rho_va);
double orc_outer_prestep(orc_model_t *m);
void   orc_inner_prestep(orc_model_t *m);
void   orc_update_x_gen(orc_model_t *m);
void   orc_update_x_line(orc_model_t *m, int64_t inner, int32_t max_auglag, double mu_max, double scale);
void   orc_update_xbar(orc_model_t *m);
void   orc_update_z(orc_model_t *m, double beta);
void   orc_update_l(orc_model_t *m, double beta);
void   orc_update_lz(orc_model_t *m, double beta, double max_multiplier);
void   orc_update_residual(orc_model_t *m, double out[4]);
double orc_poststep(orc_model_t *m);
int    orc_admm_two_level(orc_model_t *m, const ea_params_t *par, ea_info_t *info);

int64_t orc_nvar(const orc_model_t *m);
double *orc_vector(orc_model_t *m, int field);      /* borrowed pointer, reference layout, nvar  */
double *orc_membuf(orc_model_t *m);                 /* borrowed pointer, 31 x nline column-major */
void    orc_set_load(orc_model_t *m, const double *Pd, const double *Qd);
void    orc_set_pg_bounds(orc_model_t *m, const double *lo, const double *hi);
void    orc_get_counters(const orc_model_t *m, ea_counters_t *out);
void    orc_reset_counters(orc_model_t *m);
/* diagnostics: buf[nline] receives the per-line evaluation count of every x-update (NULL = off) */
void    orc_set_eval_trace(orc_model_t *m, int32_t *buf);

/* Unit-level entry points used by the tests.
 * param = one membuf column (31 doubles, 0-based index k <-> reference row k+1);
 * Y = {YffR,YffI,YftR,YftI,YttR,YttI,YtfR,YtfI}. H is 6x6 row-major (symmetric). */
double orc_eval_f(const double x[6], const double *param, const double Y[8], double scale);
/* the same objective, summed in the reference's own term order
 * (acopf_eval_linelimit_kernel_cpu.jl:1-46) */
double orc_eval_f_reforder(const double x[6], const double *param, const double Y[8], double scale);
void   orc_eval_gh(const double x[6], const double *param, const double Y[8], double scale,
                   double g[6], double H[36]);
/* one TRON solve on the branch objective: returns status (0 ok, 1 max_minor), x updated in place */
int    orc_tron_solve(double x[6], const double xl[6], const double xu[6], const double *param,
                      const double Y[8], double scale, int max_feval, int max_minor, double gtol,
                      int *minor_iter, int *nfev);

#ifdef __cplusplus
}
#endif
#endif
