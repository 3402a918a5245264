/*
 * qpsub_oracle.h — CPU ORACLE (test infrastructure, NOT product code) for the one-level ADMM on the
 * SQP sub-problem (src/models/qpsub/, `ModelQpsub`; src/algorithms/admm_one_level.jl;
 * src/interface/solve_qpsub.jl). Only tests/, __graft_entry__.smoke() and bench.py's baselines may load it.
 *
 * Pinning status: PINNED against the reference's golden vectors for one ADMM iteration on case9
 * (test/algorithms/qpsub_update_cpu.jl:160-197: u, v, l, rp, rd at 2e-6), its end-to-end known answer
 * (Solved / 5107 / 5107 / objval -21.92744641968529, :224-237) and the step / KKT error / multipliers it
 * hands back to the SQP driver after that solve (test/algorithms/qpsub_update_gpu.jl:228-346), see
 * tests/test_oracle_qpsub.py. The TRON solver is the restatement shared with acopf_oracle.c.
 */
#ifndef QPSUB_ORACLE_H
#define QPSUB_ORACLE_H

#include "acopf_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_qp orc_qp_t;

/* data: see ea_qpsub_data_t (row-major Hs nline x 6 x 6, LH_* nline x k, ls / us nline x 6, line_res nline x 4). */
int    orc_qp_create(const ea_grid_t *grid, const ea_qpsub_data_t *data, orc_qp_t **out);
void   orc_qp_destroy(orc_qp_t *q);
void   orc_qp_set_threads(orc_qp_t *q, int nthreads);
int64_t orc_qp_nvar(const orc_qp_t *q);
double *orc_qp_vector(orc_qp_t *q, int field);          /* enum ea_field; EA_V_PREV = mod.v_prev; borrowed */
double *orc_qp_line_array(orc_qp_t *q, int which);      /* enum ea_qp_array; rows x nline column-major; borrowed */

void   orc_qp_init_solution(orc_qp_t *q, double rho_pq, double rho_va);
void   orc_qp_update_x(orc_qp_t *q, int64_t inner, int32_t max_auglag, double mu_max, double scale);
void   orc_qp_update_xbar(orc_qp_t *q);
void   orc_qp_update_l_single(orc_qp_t *q);
void   orc_qp_update_residual(orc_qp_t *q, double out[5]);   /* primres, dualres, mismatch, objval, auglag */
void   orc_qp_poststep(orc_qp_t *q, double *objval, double *auglag, double *dw_sol, double *dtheta_sol,
                       double *dual_infeas);
int    orc_qp_admm_one_level(orc_qp_t *q, const ea_params_t *par, ea_info_t *info);
/* work so far: branch calls, AL iterations (TRON solves), objective evaluations, largest AL count of one call */
void   orc_qp_counters(const orc_qp_t *q, int64_t out[4]);

/* unit level: the reduced QP of one AL iteration of one branch (qpsub_eval_Ab_linelimit_kernel_cpu.jl):
 * A 6 x 6 row-major and b 6, both scaled; C 8 x 6 row-major, d 8. membuf5 = one column of qpsub_membuf. */
void   orc_qp_branch_qp(const double H[36], const double l[8], const double rho[8], const double v[8],
                        const double z[8], const double Y[8], const double res[4], const double LH_1h[4], double RH_1h,
                        const double LH_1i[4], double RH_1i, const double LH_1j[2], double RH_1j,
                        const double LH_1k[2], double RH_1k, const double membuf5[5], double scale,
                        double A[36], double b[6], double C[48], double d[8]);

/* unit level: the whole AL + TRON solve of one branch on explicit inputs (see qpsub_oracle.c) */
void   orc_qp_branch_solve(const double H[36], const double l[8], const double rho[8], const double v[8],
                           const double z[8], const double Y[8], const double res[4], const double LH_1h[4],
                           double RH_1h, const double LH_1i[4], double RH_1i, const double LH_1j[2], double RH_1j,
                           const double LH_1k[2], double RH_1k, const double ls[6], const double us[6], double sq[6],
                           double mb[5], int64_t major_iter, int32_t max_auglag, double mu_max, double scale,
                           double u[8], double lam[4], int32_t work[2]);

#ifdef __cplusplus
}
#endif
#endif
