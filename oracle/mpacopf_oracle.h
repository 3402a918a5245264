/*
 * mpacopf_oracle.h — CPU ORACLE (test infrastructure, NOT product code) for the
 * multi-period ACOPF model (src/models/mpacopf/, `ModelMpacopf`): T single-period
 * models coupled through generator ramp constraints, solved by the same two-level
 * ADMM loop. Only tests/, __graft_entry__.smoke() and bench.py's baselines may load it.
 *
 * Pinning status: PINNED against the reference's golden vectors for one iteration on
 * case9 x 3 periods (test/algorithms/mpacopf_update_cpu.jl:28-395) and its end-to-end
 * known answer (Solved / 20 / 729 / 15901.48, :431-435). The reference's load profile
 * (ExaData artifact, mp_demand/case9_onehour_60) is not in /root/reference; the three
 * period scale factors were recovered from the golden xbar values, see
 * tests/golden/make_mpacopf_golden.py.
 */
#ifndef MPACOPF_ORACLE_H
#define MPACOPF_ORACLE_H

#include "acopf_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_mp orc_mp_t;

/* Pd, Qd: T x nbus (period-major), MW / MVAr as in GridData. ramp_rate = ramp_ratio * pgmax
 * (acopf_model.jl:66-67). */
int    orc_mp_create(const ea_grid_t *grid, int32_t T, const double *Pd, const double *Qd, double ramp_ratio,
                     orc_mp_t **out);
void   orc_mp_destroy(orc_mp_t *mp);
void   orc_mp_set_threads(orc_mp_t *mp, int nthreads);
int32_t orc_mp_len_horizon(const orc_mp_t *mp);
orc_model_t *orc_mp_period(orc_mp_t *mp, int32_t t);            /* 0-based period, borrowed */
double *orc_mp_ramp_vector(orc_mp_t *mp, int32_t t, int field); /* ngen doubles, borrowed    */
double *orc_mp_gen_membuf(orc_mp_t *mp, int32_t t);             /* 8 x ngen column-major     */
int64_t orc_mp_nvar(const orc_mp_t *mp);
/* work of the generator sub-problems so far: calls, AL iterations, f-evaluations, largest AL count and largest
 * f-evaluation count of one call */
void   orc_mp_gen_counters(const orc_mp_t *mp, int64_t out[5]);

void   orc_mp_init_solution(orc_mp_t *mp, double rho_pq, double rho_va);
double orc_mp_outer_prestep(orc_mp_t *mp);
void   orc_mp_inner_prestep(orc_mp_t *mp);
void   orc_mp_update_x(orc_mp_t *mp, int64_t inner, int32_t max_auglag, double mu_max, double scale);
void   orc_mp_update_xbar(orc_mp_t *mp);
void   orc_mp_update_z(orc_mp_t *mp, double beta);
void   orc_mp_update_l(orc_mp_t *mp, double beta);
void   orc_mp_update_lz(orc_mp_t *mp, double beta, double max_multiplier);
void   orc_mp_update_residual(orc_mp_t *mp, double out[4]);
double orc_mp_poststep(orc_mp_t *mp, double *err_ramp);
int    orc_mp_admm_two_level(orc_mp_t *mp, const ea_params_t *par, ea_info_t *info, double *err_ramp);

/* unit level: one generator's AL + TRON solve (see mpacopf_oracle.c) */
void   orc_gen_ramp_solve(double x[3], const double xl[3], const double xu[3], double *param, double c2, double c1,
                          double c0, double baseMVA, double scale, int32_t max_auglag, double xi_max, int32_t *work);

#ifdef __cplusplus
}
#endif
#endif
