/*
 * exaadmm_b200.h — C ABI of the B200-native two-level ADMM ACOPF hot path.
 *
 * This is the drop-in boundary for ExaAdmm.jl's operator API on the
 * `solve_acopf(...; use_gpu=true)` path: every entry point below replaces the
 * body of one Julia method that the reference dispatches on
 * (AdmmEnv{T,TD,TI,TM}, AbstractOPFModel{T,TD,TI,TM}) — the citation next to
 * each declaration is the reference method it stands in for (paths relative to
 * the reference repository root). A Julia maintainer binds them with `ccall`
 * (see INTEGRATION.md and exaadmm.jl_b200/julia/); in this repository they are
 * exercised through Python ctypes (exaadmm.jl_b200/capi.py).
 *
 * Conventions
 *   - plain C types only; all floating point is IEEE binary64;
 *   - integer index arrays handed in by the caller are int64 and 1-BASED,
 *     exactly as the Julia side holds them (src/utils/opfdata.jl:613-618);
 *   - host pointers are borrowed for the duration of the call only; the
 *     library owns all device memory, streams, events and graphs;
 *   - vectors cross the boundary in the reference layout
 *       [ (pg,qg) x ngen | (pij,qij,pji,qji,wi,wj,ti,tj) x nline ]
 *     (docs/src/dev.md:157-162) regardless of the layout used in HBM;
 *   - every function returns 0 on success and a negative EA_ERR_* code on
 *     failure; ea_last_error() gives the message. Nothing throws across the ABI;
 *   - a handle is single-writer: one host thread at a time. Distinct handles
 *     are independent;
 *   - there is NO CPU fallback: without a CUDA device ea_create() fails with
 *     EA_ERR_CUDA.
 */
#ifndef EXAADMM_B200_H
#define EXAADMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EA_ABI_VERSION 1

/* ---- error codes ------------------------------------------------------- */
#define EA_OK            0
#define EA_ERR_ARG      (-1)   /* bad argument (NULL, size mismatch, bad field id) */
#define EA_ERR_CUDA     (-2)   /* CUDA runtime error (incl. "no device")           */
#define EA_ERR_NCCL     (-3)   /* NCCL error / NCCL not loadable                    */
#define EA_ERR_STATE    (-4)   /* call not valid in the current state              */
#define EA_ERR_ALLOC    (-5)

/* ---- status codes (mod.info.status, src/utils/environment.jl:324)  ------ */
#define EA_STATUS_NOT_SPECIFIED   0
#define EA_STATUS_ITERATION_LIMIT 1   /* :IterationLimit, admm_two_level.jl:26 */
#define EA_STATUS_SOLVED          2   /* :Solved,         admm_two_level.jl:67 */

/* ---- Solution fields (src/utils/environment.jl:177-226) ----------------- */
enum ea_field {
    EA_U_CURR = 0, EA_V_CURR = 1, EA_L_CURR = 2, EA_RHO = 3,
    EA_Z_CURR = 4, EA_Z_PREV = 5, EA_LZ = 6,
    EA_RP = 7, EA_RD = 8, EA_AX_PLUS_BY = 9,
    /* allocated-but-unused in the reference's two-level path; kept so that
       every field of `Solution` can be read back (always zero): */
    EA_U_PREV = 10, EA_V_PREV = 11, EA_L_PREV = 12, EA_RP_PREV = 13, EA_Z_OUTER = 14,
    EA_NUM_FIELDS = 15
};

/* Multi-period model: the ramp-coupling vectors of period t >= 2, one entry per generator
 * (SolutionRamping, src/models/mpacopf/mpacopf_model.jl:1-16). */
enum ea_ramp_field {
    EA_RAMP_U_CURR = 0,   /* phat_{t-1}: period t's copy of the previous period's pg */
    EA_RAMP_V_CURR = 1,   /* allocated by the reference, never used (the consensus value is v_curr[pg] of period t-1) */
    EA_RAMP_L_CURR = 2, EA_RAMP_RHO = 3, EA_RAMP_RD = 4, EA_RAMP_RP = 5, EA_RAMP_Z_OUTER = 6,
    EA_RAMP_Z_CURR = 7, EA_RAMP_Z_PREV = 8, EA_RAMP_LZ = 9, EA_RAMP_AX_PLUS_BY = 10,
    EA_RAMP_S_CURR = 11,  /* s_t = pg_t - phat_{t-1}, |s_t| <= ramp_rate */
    EA_RAMP_NUM_FIELDS = 12
};

/* GridData as consumed by the hot path (src/utils/grid_data.jl:61-83,
 * src/utils/opfdata.jl:542-831). Lengths in comments. */
typedef struct ea_grid {
    int64_t ngen, nline, nbus;
    double  baseMVA;
    const double *pgmin, *pgmax, *qgmin, *qgmax;      /* ngen, p.u.                       */
    const double *c2, *c1, *c0;                       /* ngen, UNSCALED       */
    const double *YshR, *YshI;                        /* nbus                             */
    const double *YffR, *YffI, *YftR, *YftI;          /* nline                            */
    const double *YttR, *YttI, *YtfR, *YtfI;          /* nline                            */
    const double *FrVmBound, *ToVmBound;              /* 2*nline, interleaved (min,max)   */
    const double *FrVaBound, *ToVaBound;              /* 2*nline                          */
    const double *rateA;                              /* nline, squared p.u. rating       */
    const int64_t *FrStart, *ToStart, *GenStart;      /* nbus+1, 1-based CSR pointers     */
    const int64_t *FrIdx, *ToIdx;                     /* nline, 1-based line ids          */
    const int64_t *GenIdx;                            /* ngen,  1-based generator ids     */
    const double *Pd, *Qd;                            /* nbus, MW / MVAr (not p.u.)       */
    const double *Vmin, *Vmax;                        /* nbus                             */
    const int64_t *brBusIdx;                          /* 2*nline, 1-based (from,to) bus   */
} ea_grid_t;

/* The subset of `Parameters` (src/utils/environment.jl:6-76) the hot path reads. */
typedef struct ea_params {
    double  mu_max;          /* 1e8   */
    int32_t max_auglag;      /* 50    */
    int32_t verbose;         /* 0: silent; >0: per-iteration table like admm_two_level.jl:47-57 */
    double  initial_beta;    /* 1e3   */
    double  inc_c;           /* 6.0   */
    double  theta;           /* 0.8   */
    double  outer_eps;       /* 2e-4  */
    double  MAX_MULTIPLIER;  /* 1e12  */
    double  scale;           /* 1e-4  */
    double  obj_scale;       /* accepted and stored; inert on this path */
    int64_t outer_iterlim;   /* 20    */
    int64_t inner_iterlim;   /* 1000  */
} ea_params_t;

/* `IterationInformation` + `ComponentInformation` (src/utils/environment.jl:277-358). */
typedef struct ea_info {
    int32_t status;
    int32_t _pad;
    int64_t inner, outer, cumul;
    double  objval, primres, dualres, mismatch, auglag, eps_pri;
    double  norm_z_curr, norm_z_prev;
    double  beta;                                   /* par.beta after the run */
    double  time_x_update, time_xbar_update, time_z_update, time_l_update,
            time_lz_update, time_projection, time_overall;
    double  time_generators, time_branches, time_buses;
} ea_info_t;

/* Work counters of the branch kernel, accumulated since the last reset. */
typedef struct ea_counters {
    int64_t line_calls;      /* branch sub-problems solved                           */
    int64_t auglag_iters;    /* augmented-Lagrangian iterations (TRON solves)        */
    int64_t tron_evals;      /* (f, grad, Hessian) evaluations                       */
    int64_t cg_iters;        /* projected-CG iterations                              */
    int64_t chol_shifts;     /* Cholesky factorizations that needed a diagonal shift */
    int64_t rejected_steps;  /* trust-region steps rejected                          */
    int64_t max_auglag_hits; /* lines that stopped on max_auglag                     */
    int64_t max_evals_lane;  /* largest per-line evaluation count seen in one call   */
} ea_counters_t;

typedef struct ea_handle ea_handle_t;

int  ea_abi_version(void);
/* Message of the last error on this handle (h may be NULL: last ea_create error). */
const char *ea_last_error(const ea_handle_t *h);
/* Number of CUDA devices visible (0 without a GPU; never fails). */
int  ea_device_count(void);

/* ModelAcopf constructor, device part: GridData H2D copies, Solution and membuf
 * allocation (src/models/acopf/acopf_model.jl:41-94). Does NOT call
 * init_solution (call ea_init_solution). `device` = CUDA ordinal (gpu_no). */
int  ea_create(const ea_grid_t *grid, int device, ea_handle_t **out);
void ea_destroy(ea_handle_t *h);

/* init_solution! (src/models/acopf/acopf_init_solution_gpu.jl:49-67). As in the
 * reference, membuf (rows 25-27: line-limit multipliers and mu) is zeroed once by
 * the constructor (ea_create, acopf_model.jl:87-89) and is NOT reset here. */
int  ea_init_solution(ea_handle_t *h, double rho_pq, double rho_va);

/* admm_outer_prestep: norm_z_prev = ||z_curr|| (acopf_admm_prepoststep_gpu.jl:1-9). */
int  ea_outer_prestep(ea_handle_t *h, double *norm_z_prev);
/* admm_inner_prestep: z_prev <- z_curr (acopf_admm_prepoststep_gpu.jl:11-21). */
int  ea_inner_prestep(ea_handle_t *h);

/* acopf_admm_update_x_gen -> generator_kernel_two_level
 * (acopf_admm_update_x_gpu.jl:1-12, acopf_generator_kernel_gpu.jl:1-34). */
int  ea_update_x_gen(ea_handle_t *h);
/* acopf_admm_update_x_line -> auglag_linelimit_two_level_alternative + TRON
 * (acopf_admm_update_x_gpu.jl:14-46, acopf_auglag_linelimit_kernel_gpu.jl:1-151,
 * acopf_tron_linelimit_kernel.jl:4-149). `inner` is info.inner (mu resets to 10
 * when it equals 1). */
int  ea_update_x_line(ea_handle_t *h, int64_t inner, int32_t max_auglag,
                      double mu_max, double scale);
/* admm_update_x = gen then line (acopf_admm_update_x_gpu.jl:48-56). */
int  ea_update_x(ea_handle_t *h, int64_t inner, int32_t max_auglag,
                 double mu_max, double scale);
/* admm_update_xbar -> bus_kernel_two_level_alternative
 * (acopf_admm_update_xbar_gpu.jl:1-13, acopf_bus_kernel_gpu.jl:1-119). */
int  ea_update_xbar(ea_handle_t *h);
/* admm_update_z (acopf_admm_update_z_gpu.jl:1-24). */
int  ea_update_z(ea_handle_t *h, double beta);
/* admm_update_l (acopf_admm_update_l_gpu.jl:1-26). */
int  ea_update_l(ea_handle_t *h, double beta);
/* admm_update_residual (acopf_admm_update_residual_gpu.jl:11-29):
 * out = { primres, dualres, norm_z_curr, mismatch }. */
int  ea_update_residual(ea_handle_t *h, double out[4]);
/* admm_update_lz (acopf_admm_update_lz_gpu.jl:1-20). */
int  ea_update_lz(ea_handle_t *h, double beta, double max_multiplier);
/* admm_poststep: objval from the unscaled cost (acopf_admm_prepoststep_gpu.jl:23-42). */
int  ea_poststep(ea_handle_t *h, double *objval);

/* One whole inner iteration (inner_prestep, update_x, update_xbar, update_z,
 * update_l, update_residual of admm_two_level.jl:36-42) in two fused launches.
 * Produces the same iterate as the step-wise calls. */
int  ea_inner_iteration(ea_handle_t *h, int64_t inner, double beta,
                        int32_t max_auglag, double mu_max, double scale,
                        double out[4]);

/* The whole `while inner < inner_iterlim` loop of one outer iteration
 * (admm_two_level.jl:33-63) with the primres <= eps_pri test evaluated on the
 * device; the host polls once per `chunk` iterations (chunk <= 0: default).
 * `outer` is info.outer (eps_pri = sqrt(nvar)/(2500*outer)). On return
 * *inner_done is info.inner and out = last {primres,dualres,norm_z_curr,mismatch}. */
int  ea_run_inner(ea_handle_t *h, int64_t outer, double beta,
                  int64_t inner_iterlim, int32_t max_auglag, double mu_max,
                  double scale, int32_t chunk, int64_t *inner_done, double out[4]);

/* Same loop, resumable: counts inner iterations from `inner_start` (so that the mu reset
 * keyed on info.inner == 1 fires only at the true start) up to `inner_limit`. */
int  ea_run_inner_from(ea_handle_t *h, int64_t outer, double beta, int64_t inner_start,
                       int64_t inner_limit, int32_t max_auglag, double mu_max, double scale,
                       int32_t chunk, int64_t *inner_done, double out[4]);

/* admm_two_level (src/algorithms/admm_two_level.jl:1-88) end to end, including
 * admm_poststep. Starts from the current Solution (warm start is implicit, as
 * in the reference). */
int  ea_admm_two_level(ea_handle_t *h, const ea_params_t *par, ea_info_t *info);

/* copyto!(host, mod.solution.<field>) / copyto!(mod.solution.<field>, host);
 * n must equal nvar = 2*ngen + 8*nline. */
int64_t ea_nvar(const ea_handle_t *h);
int  ea_get_vector(ea_handle_t *h, int field, double *host, int64_t n);
int  ea_set_vector(ea_handle_t *h, int field, const double *host, int64_t n);
/* mod.membuf[row, :] with the reference's 1-based row numbers (acopf_model.jl:87-89):
 * rows 25,26 = line-limit multipliers, 27 = mu, 29 = rateA are live; rows 1-24 are
 * re-derived from the Solution on read (they are staging in the reference); set is
 * accepted for rows 25, 26, 27 and 29. n must equal nline. */
int  ea_get_membuf(ea_handle_t *h, int row, double *host, int64_t n);
int  ea_set_membuf(ea_handle_t *h, int row, const double *host, int64_t n);

/* Rolling horizon / scenarios: replace bus loads (acopf_admm_rolling_gpu.jl:42-43)
 * and the ramp-limited generator bounds (acopf_admm_rolling_gpu.jl:1-14). */
int  ea_set_load(ea_handle_t *h, const double *Pd, const double *Qd, int64_t nbus);
int  ea_set_pg_bounds(ea_handle_t *h, const double *pgmin_curr,
                      const double *pgmax_curr, int64_t ngen);

/* ---- bus-partitioned multi-GPU mode (one process per GPU; new design, the reference is
 * single-GPU: SURVEY.md F6 / section 8e) -----------------------------------------------------
 * The handle is created on the RANK-LOCAL grid (owned buses first, then the ghost buses at
 * the far end of cut branches; exaadmm.jl_b200/partition.py builds it). ea_set_partition
 * declares which branch ends are sent / received; indices here are 0-BASED local ids,
 * end = 0 (from) or 1 (to). Every rank's send list has at most max_send entries; a ghost end
 * takes its xbar from position ghost_src_pos of rank ghost_src_rank's list. Afterwards the
 * fused entry points (ea_inner_iteration, ea_run_inner*, ea_admm_two_level) run the
 * partitioned iteration: x-update (cut branches redundantly on both ranks), bus update of the
 * owned buses, ONE ncclAllGather per inner iteration carrying the xbar halves of the cut
 * branch ends and each rank's 4 residual partial sums, then the ghost z / lambda update and
 * the termination test (identical on every rank). */
int  ea_set_partition(ea_handle_t *h, int32_t rank, int32_t nranks, int64_t n_owned_bus,
                      int64_t n_send, const int64_t *send_line, const int64_t *send_end,
                      int64_t n_ghost, const int64_t *ghost_line, const int64_t *ghost_end,
                      const int64_t *ghost_src_rank, const int64_t *ghost_src_pos, int64_t max_send,
                      int64_t nvar_global /* 2*ngen + 8*nline of the WHOLE case: eps_pri and OUTER_TOL scale with its sqrt */);
/* NCCL is dlopen'ed (nccl_lib = path, or NULL: $EXAADMM_NCCL_LIB, then libnccl.so.2). Rank 0
 * makes the id, the caller broadcasts the 128 bytes (e.g. torch.distributed), every rank
 * calls ea_comm_init. */
int  ea_nccl_unique_id(const char *nccl_lib, char out[128]);
int  ea_comm_init(ea_handle_t *h, const char *nccl_lib, const char id[128]);
/* Peer-memory exchange (one node, NVLink/NVSwitch): every rank exports its exchange buffer as a CUDA IPC handle
 * (64 bytes), the caller all-gathers the handles (rank order) and every rank imports them. From then on the
 * fused loop does NOT call NCCL: the last block of the bus kernel stores this rank's segment into every peer's
 * buffer and raises a flag there; k_finish waits on its own flags (double-buffered by iteration parity). */
int  ea_peer_export(ea_handle_t *h, char out[64]);
int  ea_peer_import(ea_handle_t *h, const char *handles);
/* Loopback form of one partitioned iteration for single-GPU tests (the caller carries the
 * message through the host): begin = x-update + bus kernel; get_message = this rank's
 * segment (stride = 4 + 4*max_send doubles); put_gathered = all nranks segments;
 * end = ghost update + norms + termination bookkeeping, out = the 4 norms. */
int  ea_part_begin(ea_handle_t *h, int64_t inner, double beta, int32_t max_auglag, double mu_max, double scale);
int  ea_part_get_message(ea_handle_t *h, double *host, int64_t n);
int  ea_part_put_gathered(ea_handle_t *h, const double *host, int64_t n);
int  ea_part_end(ea_handle_t *h, double out[4]);

/* ---- multi-period ACOPF: `ModelMpacopf` (src/models/mpacopf/) --------
 * T copies of the single-period model on one device, period t = 0..T-1 with its own loads, coupled
 * through the generator ramp constraints |pg_t - pg_{t-1}| <= ramp_rate. The periods are ordinary
 * handles (ea_mp_period; borrowed, do not destroy) so every single-period accessor works on them
 * (`mod.models[i].solution`, membuf, ...); the ramp vectors (`mod.solution[i]`, SolutionRamping)
 * are read with ea_mp_get_ramp_vector. One host thread per model, as in the reference. */
typedef struct ea_mp_handle ea_mp_handle_t;
const char *ea_mp_last_error(const ea_mp_handle_t *h);   /* h may be NULL: last ea_mp_create error */
/* ModelMpacopf constructor (mpacopf_model.jl:56-109) + init_solution!. Pd, Qd: T x nbus (period-major),
 * MW / MVAr; ramp_rate = ramp_ratio * pgmax (acopf_model.jl:66-67). */
int  ea_mp_create(const ea_grid_t *grid, int device, int32_t T, const double *Pd, const double *Qd,
                  double ramp_ratio, double rho_pq, double rho_va, ea_mp_handle_t **out);
void ea_mp_destroy(ea_mp_handle_t *h);
int32_t ea_mp_len_horizon(const ea_mp_handle_t *h);
int64_t ea_mp_nvar(const ea_mp_handle_t *h);              /* mod.nvar: nvar + ngen (T > 1), mpacopf_model.jl:97-102 */
ea_handle_t *ea_mp_period(ea_mp_handle_t *h, int32_t t); /* mod.models[t+1] */
/* init_solution!(mod::ModelMpacopf, ...) (mpacopf_init_solution_gpu.jl:16-35); gen_membuf persists. */
int  ea_mp_init_solution(ea_mp_handle_t *h, double rho_pq, double rho_va);
/* mod.solution[t+1].<field> (enum ea_ramp_field), n = ngen, reference generator order; t >= 1. */
int  ea_mp_get_ramp_vector(ea_mp_handle_t *h, int32_t t, int field, double *host, int64_t n);
int  ea_mp_set_ramp_vector(ea_mp_handle_t *h, int32_t t, int field, const double *host, int64_t n);
/* gen_membuf rows 7 (multiplier of the ramp equality) and 8 (its penalty xi) of period t >= 1. */
int  ea_mp_get_gen_membuf(ea_mp_handle_t *h, int32_t t, int row, double *host, int64_t n);
/* The operators of the generic loop for mod::ModelMpacopf, file by file:
 * mpacopf_admm_prepoststep_gpu.jl:1-38 (outer / inner prestep), mpacopf_admm_update_x_gpu.jl:1-46
 * (generators of period 1 closed form, periods >= 2 AL + TRON with n = 3, then every period's
 * branches), mpacopf_admm_update_xbar_gpu.jl + mpacopf_bus_kernel_gpu.jl (ramp-aware bus update),
 * mpacopf_admm_update_{z,l,lz,residual}_gpu.jl, mpacopf_admm_prepoststep_gpu.jl:40-78 (poststep:
 * objval = sum over periods, err_ramp = largest ramp violation). */
int  ea_mp_outer_prestep(ea_mp_handle_t *h, double *norm_z_prev);
int  ea_mp_inner_prestep(ea_mp_handle_t *h);
int  ea_mp_update_x(ea_mp_handle_t *h, int64_t inner, int32_t max_auglag, double mu_max, double scale);
int  ea_mp_update_xbar(ea_mp_handle_t *h);
int  ea_mp_update_z(ea_mp_handle_t *h, double beta);
int  ea_mp_update_l(ea_mp_handle_t *h, double beta);
int  ea_mp_update_lz(ea_mp_handle_t *h, double beta, double max_multiplier);
int  ea_mp_update_residual(ea_mp_handle_t *h, double out[4]);
int  ea_mp_poststep(ea_mp_handle_t *h, double *objval, double *err_ramp);
/* Fused forms (all periods advance together, termination test on the device):
 * the inner `while` of one outer iteration, and admm_two_level end to end. */
int  ea_mp_run_inner(ea_mp_handle_t *h, int64_t outer, double beta, int64_t inner_iterlim, int32_t max_auglag,
                     double mu_max, double scale, int32_t chunk, int64_t *inner_done, double out[4]);
int  ea_mp_admm_two_level(ea_mp_handle_t *h, const ea_params_t *par, ea_info_t *info, double *err_ramp);
/* Launch accounting of the fused loop: out = { device seconds in ea_mp_run_inner, kernels launched,
 * inner iterations executed, 0 }. */
int  ea_mp_get_kernel_times(ea_mp_handle_t *h, double out[4]);

/* ---- one-level ADMM on the SQP sub-problem: `ModelQpsub` (src/models/qpsub/, src/algorithms/admm_one_level.jl;
 * SURVEY.md section 8f row 4) -----------------------------------------------------------------------------------
 * The QP of one SQP iteration of ACOPF, decomposed like the ACOPF itself: generators (closed form with shifted
 * bounds / costs), buses (the same consensus update with shifted loads) and branches - per branch a box-constrained
 * QP in (t_ij, t_ji, w_i, w_j, theta_i, theta_j) built from the SQP Hessian block and the linearised constraints
 * 1h / 1i (eliminated) and 1j / 1k (line limits, augmented Lagrangian), solved by TRON.
 * ea_qpsub_data_t carries the fields the SQP driver fills in on `ModelQpsub` (qpsub_model.jl:63-92), row-major:
 * Hs nline x 6 x 6 (variables w_ijR, w_ijI, w_i, w_j, theta_i, theta_j), LH_1h / LH_1i nline x 4, LH_1j / LH_1k
 * nline x 2, RH_* nline, ls / us nline x 6, line_res nline x 4 (NULL = zeros), generator arrays ngen, Pd / Qd nbus. */
typedef struct ea_qpsub_data {
    const double *Hs;
    const double *LH_1h, *RH_1h, *LH_1i, *RH_1i, *LH_1j, *RH_1j, *LH_1k, *RH_1k;
    const double *ls, *us;
    const double *line_res;
    const double *pgmax, *pgmin, *qgmax, *qgmin, *c1, *c2;
    const double *Pd, *Qd;
} ea_qpsub_data_t;

/* per-line arrays of ModelQpsub readable with ea_qp_get_line_array */
enum ea_qp_array {
    EA_QP_SQP_LINE = 0,   /* 6 x nline: w_ijR, w_ijI, w_i, w_j, theta_i, theta_j of the last branch solve (mod.sqp_line) */
    EA_QP_MEMBUF = 1,     /* 5 x nline: lambda_1h, lambda_1i, lambda_1j, lambda_1k, penalty (mod.qpsub_membuf)          */
    EA_QP_LAMBDA = 2      /* 4 x nline: multipliers of 14h, 14i, 14j, 14k handed back to the SQP (mod.lambda)            */
};

typedef struct ea_qp_handle ea_qp_handle_t;
const char *ea_qp_last_error(const ea_qp_handle_t *h);
/* ModelQpsub{T,TD,TI,TM}(env) (qpsub_model.jl:133-310) + the field copies of solve_qpsub (solve_qpsub.jl:83-104).
 * Hs must be symmetric (the SQP driver builds it so; the device code reads the lower triangle). obj_scale is inert
 * on this path, as in the reference (solve_qpsub stores it after the constructor's scaling step has run). */
int  ea_qp_create(const ea_grid_t *grid, const ea_qpsub_data_t *data, int device, ea_qp_handle_t **out);
void ea_qp_destroy(ea_qp_handle_t *h);
int64_t ea_qp_nvar(const ea_qp_handle_t *h);
/* init_solution! (qpsub_init_solution_gpu.jl:1-98) */
int  ea_qp_init_solution(ea_qp_handle_t *h, double rho_pq, double rho_va);
/* Solution fields (enum ea_field; EA_V_PREV is mod.v_prev), reference layout, n = nvar. */
int  ea_qp_get_vector(ea_qp_handle_t *h, int field, double *host, int64_t n);
int  ea_qp_set_vector(ea_qp_handle_t *h, int field, const double *host, int64_t n);
/* which = enum ea_qp_array; rows x nline, column-major like the reference's matrices (n = rows * nline). */
int  ea_qp_get_line_array(ea_qp_handle_t *h, int which, double *host, int64_t n);
int  ea_qp_set_line_array(ea_qp_handle_t *h, int which, const double *host, int64_t n);
/* admm_update_x: generator_kernel_two_level on the qpsub bounds / costs (qpsub_generator_kernel_gpu.jl), then
 * auglag_linelimit_qpsub (qpsub_auglag_Ab_linelimit_kernel_red_gpu.jl + qpsub_tron_linelimit_kernel.jl).
 * inner = info.inner (the penalty restarts at 10 when it is 1). */
int  ea_qp_update_x(ea_qp_handle_t *h, int64_t inner, int32_t max_auglag, double mu_max, double scale);
/* admm_update_xbar: v_prev <- v_curr, then the bus kernel with qpsub_Pd / qpsub_Qd (qpsub_admm_update_xbar_gpu.jl). */
int  ea_qp_update_xbar(ea_qp_handle_t *h);
/* admm_update_l_single: l += rho (u - v) (qpsub_admm_update_l_single_gpu.jl). */
int  ea_qp_update_l_single(ea_qp_handle_t *h);
/* admm_update_residual (qpsub_admm_update_residual_gpu.jl): out = { primres, dualres, mismatch, objval, auglag }. */
int  ea_qp_update_residual(ea_qp_handle_t *h, double out[5]);
/* admm_poststep (qpsub_admm_prepoststep_gpu.jl): objval / auglag and what is handed back to the SQP driver:
 * dw_sol, dtheta_sol (nbus each: averages over the incident branch ends) and dual_infeas (ngen + 6 nline). Any
 * pointer may be NULL. The other outputs of the reference are views of the state: dpg_sol / dqg_sol / dline_fl are
 * entries of u_curr, dline_var is sqp_line, lambda is EA_QP_LAMBDA. */
int  ea_qp_poststep(ea_qp_handle_t *h, double *objval, double *auglag, double *dw_sol, double *dtheta_sol,
                    double *dual_infeas);
/* admm_one_level end to end (admm_one_level.jl:1-81): every iteration and the termination test run on the device
 * (3 launches per iteration, replayed from a CUDA graph of `chunk` iterations; the host polls once per chunk).
 * par->outer_iterlim, outer_eps, max_auglag, mu_max, scale are read; inner_iterlim is 1 and beta 0 by definition. */
int  ea_qp_admm_one_level(ea_qp_handle_t *h, const ea_params_t *par, ea_info_t *info);
/* "chunk" (iterations per graph replay, default 32), "use_graph" (0/1, default 1), "count_work" (0/1, default 1) */
int  ea_qp_set_option(ea_qp_handle_t *h, const char *name, double value);
/* out = { branch sub-problems solved, AL iterations (TRON solves), objective evaluations, largest AL count of one
 * call } since creation */
int  ea_qp_get_counters(ea_qp_handle_t *h, int64_t out[4]);
/* out = { device seconds inside ea_qp_admm_one_level, iterations executed there, kernels launched, graph replays } */
int  ea_qp_get_kernel_times(ea_qp_handle_t *h, double out[4]);

/* ---- batch of independent load scenarios of one grid (BASELINE config 5) ---------------------------------------------
 * The reference has no batched solve: it runs solve_acopf per scenario, swapping loads into a model the way its rolling
 * horizon driver does (src/models/acopf/acopf_admm_rolling_gpu.jl:42-43). Here S scenarios of one grid are solved
 * TOGETHER on one device: every scenario is an ordinary model (ea_batch_scenario returns its handle - borrowed, do not
 * destroy - so `mod.solution`, membuf etc. are the single-model accessors) whose iterates are those of a stand-alone
 * solve, bit for bit; the scenarios share the launches (one branch kernel over all (scenario, branch) pairs of the
 * scenarios still iterating, one bus kernel with a termination test per scenario). Pd, Qd: S x nbus, MW / MVAr. */
typedef struct ea_batch_handle ea_batch_handle_t;
const char *ea_batch_last_error(const ea_batch_handle_t *h);   /* h may be NULL: last ea_batch_create error */
int  ea_batch_create(const ea_grid_t *grid, int device, int32_t S, const double *Pd, const double *Qd,
                     ea_batch_handle_t **out);
void ea_batch_destroy(ea_batch_handle_t *h);
int32_t ea_batch_size(const ea_batch_handle_t *h);
ea_handle_t *ea_batch_scenario(ea_batch_handle_t *h, int32_t s);
int  ea_batch_init_solution(ea_batch_handle_t *h, double rho_pq, double rho_va);       /* init_solution! of every scenario */
/* admm_two_level (src/algorithms/admm_two_level.jl:1-88) of every scenario; infos: S entries (mod.info per scenario;
 * time_overall = the batch's wall time). */
int  ea_batch_admm_two_level(ea_batch_handle_t *h, const ea_params_t *par, ea_info_t *infos);
int  ea_batch_set_option(ea_batch_handle_t *h, const char *name, double value);        /* "chunk", "use_graph" */
int  ea_batch_get_times(ea_batch_handle_t *h, double out[2]);      /* device seconds inside ea_batch_admm_two_level, launches */

int  ea_get_counters(ea_handle_t *h, ea_counters_t *out);
int  ea_reset_counters(ea_handle_t *h);
/* Options: "count_work" (0/1, atomics for ea_counters_t in the branch kernel, default 1),
 * "chunk" (inner iterations enqueued per host poll in ea_run_inner*, default 16),
 * "kernel_timing" (0/1/2: bracket every kernel of the fused loop with CUDA events; 2: every ITERATION - the summed
 * duration is reported as the x-update's, the bus kernel's as 0),
 * "use_graph" (0/1, default 1: ea_run_inner* replays the chunk of iterations from a CUDA graph),
 * "l2_flush_mb" (measurement: with kernel_timing, a write of this many MB precedes every iteration, outside the
 * event brackets, so that every timed kernel starts from a cold L2),
 * "l2_flush_clean" (0/1, measurement: the write is followed by a read of a second buffer of the same size, which
 * leaves the L2 holding clean lines instead of dirty ones waiting for their write-back). */
int  ea_set_option(ea_handle_t *h, const char *name, double value);
/* Launch accounting since the last ea_reset_counters: out = { device seconds spent in
 * ea_run_inner* (events on the library's stream), x-update launches, their summed
 * duration [s] (kernel_timing only), bus-kernel launches, their summed duration [s],
 * other launches, 0, 0 }. Feeds info.time_* and bench.py (print_statistics.jl:12-19). */
int  ea_get_kernel_times(ea_handle_t *h, double out[8]);
/* Diagnostics: f, gradient (6) and Hessian (6x6 row-major) of the branch objective for
 * n points on the device; param = 31 doubles per point (one membuf column), Y = 8. */
int  ea_diag_branch_eval(int device, int64_t n, const double *x, const double *param,
                         const double *Y, double scale, double *f, double *g, double *H);

/* Diagnostics: measured FP64 FMA throughput of the device in TFLOP/s (8 independent
 * DFMA chains per thread, best of 5); the roofline denominator of the branch kernel. */
/* Solve n branch sub-problems outside the ADMM loop with the branch driver of k_xupdate (the state machine of
 * branch.cuh), one problem per lane or (per_warp) one per warp - a lone lane, the regime of the kernel's tail.
 * prob: n x 56 doubles (layout of tests/golden/hard_branches.npz), sol: n x 13, work: n x 6, cycles: n SM cycles.
 * Tests (device = host build of the same code on the hardest branches of a real solve) and latency probes. */
int  ea_diag_branch_solve(int device, int per_warp, int64_t n, const double *prob, int32_t max_auglag,
                          double mu_max, double scale, double *sol, int32_t *work, int64_t *cycles, double *kernel_ms);
int  ea_diag_fp64_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* EXAADMM_B200_H */
