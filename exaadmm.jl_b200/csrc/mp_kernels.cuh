// mp_kernels.cuh — multi-period ACOPF (src/models/mpacopf/, `ModelMpacopf`): T single-period
// models on one device, coupled through the generator ramp constraints.
//
// Per inner iteration the periods reuse the single-period kernels (k_xupdate for the branches,
// k_bus<FUSED> with the ramp-aware generator terms); this file adds
//   k_mp_gen      the x-update of the generators of periods 2..T (AL + TRON, n = 3, genramp.cuh),
//   k_mp_finish   z / lambda of the ramp consensus, the per-period norms (period vectors + ramp
//                 vectors), their max over periods and the termination test for ALL periods,
//   k_mp_ramp_op  the step-wise element operators on the ramp vectors (operator-API parity).
// Ramp vectors are stored T x ngen (row 0 unused) in generator SLOT order (generators sorted by
// bus, the order of the state vectors' generator part).
#pragma once
#include "genramp.cuh"
#include "kernels.cuh"

namespace ea {

struct MpDev {
    int T, ngen;
    const Dev *devs;               // device array [T]: the periods
    Ctrl *const *pctrl;            // device array [T]: period t's loop control block (== devs[t].ctrl)
    Ctrl *ctrl;                    // model-level control block: res = max over periods
    double *r_u, *r_s, *r_l, *r_rho, *r_lz, *r_z[2], *r_rp, *r_rd, *r_axby;   // SolutionRamping (mpacopf_model.jl:1-16)
    double *g_mu, *g_xi;           // gen_membuf rows 7-8: multiplier and penalty of the ramp equality
    const double *ramp_rate, *c0;  // ngen, slot order
    double *sums;                  // T x 4: sums of squares over the period vectors (written by k_bus<FUSED>)
    double *res_t;                 // T x 4: per-period norms including the ramp vectors (models[i].info.*)
    double *partials;              // (T-1) x 4 x MP_FIN_MAXB
    unsigned *ticket;
};

constexpr int GBLOCK = 128;
constexpr int MP_FIN_BLOCK = 256;
constexpr int MP_FIN_MAXB = 64;

// ---------------------------------------------------------------------------
// x-update of the generators of period t = blockIdx.y (mpacopf_admm_update_x_gpu.jl:1-31): period 1
// closed form, periods >= 2 auglag_generator_kernel.
//   major_arg > 0: step-wise call; == 0: fused loop, read from the model-level control block.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GBLOCK)
k_mp_gen(MpDev m, branch::PowTable T, long long major_arg, int zsel_arg, int max_auglag, double xi_max) {
    long long major = major_arg;
    int zsel = zsel_arg;
    if (major_arg == 0) {
        if (m.ctrl->done) return;
        major = m.ctrl->inner + 1;
        zsel = m.ctrl->zsel;
    }
    const int t = blockIdx.y;
    const int k = blockIdx.x * GBLOCK + threadIdx.x;
    if (k >= m.ngen) return;
    const Dev &d = m.devs[t];
    if (t == 0) {                          // period 1: the plain closed-form update (mpacopf_admm_update_x_gpu.jl:8)
        generator_update(d, d.zbuf[zsel], k);
        return;
    }
    const double *vprev = m.devs[t - 1].v;
    const size_t o = (size_t)t * m.ngen + k;
    const double2 v = *reinterpret_cast<const double2 *>(d.v + 2 * k);
    const double2 z = *reinterpret_cast<const double2 *>(d.zbuf[zsel] + 2 * k);
    const double2 l = *reinterpret_cast<const double2 *>(d.l + 2 * k);
    const double2 rho = *reinterpret_cast<const double2 *>(d.rho + 2 * k);
    double2 u;
    u.y = fmax(d.qgmin[k], fmin(d.qgmax[k], (-(l.y + rho.y * (-v.y + z.y))) / rho.y));
    const double pmin = d.pgmin[k], pmax = d.pgmax[k], ramp = m.ramp_rate[k];
    const double xl[3] = { pmin, pmin, -ramp }, xu[3] = { pmax, pmax, ramp };
    double x[3];
    x[0] = fmin(xu[0], fmax(xl[0], d.u[2 * k]));
    x[1] = fmin(xu[1], fmax(xl[1], m.r_u[o]));
    x[2] = fmin(xu[2], fmax(xl[2], m.r_s[o]));
    const genramp::Problem P = { l.x, m.r_l[o], rho.x, m.r_rho[o], v.x - z.x, vprev[2 * k] - m.r_z[zsel][o],
                                 d.c2[k], d.c1[k], m.c0[k], d.baseMVA, 1.0 /* the reference passes scale = 1 here */ };
    double mu = m.g_mu[o];
    double xi = (major <= 1) ? 10.0 : m.g_xi[o];
    int evals = 0, cg = 0, it = 0;
    genramp::solve(P, mu, xi, x, xl, xu, max_auglag, xi_max, T, evals, cg, it);
    u.x = x[0];
    *reinterpret_cast<double2 *>(d.u + 2 * k) = u;
    m.r_u[o] = x[1];
    m.r_s[o] = x[2];
    m.g_mu[o] = mu;
    m.g_xi[o] = xi;
}

// ---------------------------------------------------------------------------
// Branch x-update of ALL periods in one persistent launch: the work queue runs over the T x nline
// (period, branch) pairs, so the device sees T times the work of a single-period launch and ONE tail
// (the serial chain of the slowest branch) instead of T. Same lane state machine as k_xupdate
// (kernels.cuh); only the bookkeeping of which period a lane's branch belongs to is added.
//
// BATCH = true: the T models are independent load scenarios of one grid (BASELINE config 5; rolling-horizon style load
// swaps, acopf_admm_rolling_gpu.jl:42-43) instead of coupled periods. Every scenario has its own loop control block
// (inner counter, z selector, done flag: they converge at different iterations), `active` lists the scenarios still
// iterating; the queue then runs over n_active x nline pairs, and the closed-form generator update of those scenarios
// is done here too (k_mp_gen belongs to the coupled model).
// ---------------------------------------------------------------------------
struct BatchView { const int *active; const int *n_active; int ngen; };

template <bool BATCH>
__global__ void __launch_bounds__(XBLOCK, EA_XMINB)
k_xupdate_mp(MpDev m, BatchView bv, int nline, branch::PowTable T, long long major_arg, int zsel_arg, int max_auglag,
             double mu_max, double scale) {
    extern __shared__ double tile[];                       // TILE_ROWS x XBLOCK
    long long major = major_arg;
    int zsel = zsel_arg;
    int n_models = m.T;
    if (BATCH) {
        n_models = *bv.n_active;
        if (n_models == 0) return;
        for (int k = blockIdx.x * XBLOCK + threadIdx.x; k < n_models * bv.ngen; k += gridDim.x * XBLOCK) {
            const Dev &d = m.devs[bv.active[k / bv.ngen]];
            if (!d.ctrl->done) generator_update(d, d.zbuf[d.ctrl->zsel], k % bv.ngen);
        }
    } else if (major_arg == 0) {
        if (m.ctrl->done) return;
        major = m.ctrl->inner + 1;
        zsel = m.ctrl->zsel;
    }
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int total = n_models * nline;
    double *col = tile + threadIdx.x;
    const branch::Objective<branch::TileView<XBLOCK>> eval{ { col }, scale };
    branch::Lane L;
    L.cold = tile + branch::TILE_COLD * XBLOCK + threadIdx.x; L.cs = XBLOCK;
    // (no spreading of small grids over all warps here, unlike k_xupdate: the generator kernel runs beside this one
    //  on the side stream and needs the SMs this launch leaves free - measured: -5 % at 1354 buses x 6 periods)
    L.phase = branch::NEED;
    L.step_pending = false;
    int G = -1;                                             // (period, branch) pair of this lane: G = t * nline + I
    // work of this CTA in shared memory (a finishing lane adds its branch), not in registers live across the whole loop
    __shared__ unsigned s_work[8];
    if (threadIdx.x < 8) s_work[threadIdx.x] = 0u;
    __syncthreads();
    Counters *counters = m.devs[0].counters;
    const int count_work = m.devs[0].count_work;

#pragma unroll 1
    for (;;) {
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            const bool need = (L.phase == branch::NEED);
            const unsigned mk = __ballot_sync(full, need);
            if (mk) {
                int base = 0;
                if (lane == __ffs(mk) - 1) base = atomicAdd(&m.ctrl->next_line, __popc(mk));
                base = __shfl_sync(full, base, __ffs(mk) - 1);
                if (need) {
                    G = base + __popc(mk & ((1u << lane) - 1u));
                    if (G < total) {
                        const int a = G / nline;
                        const int t = BATCH ? bv.active[a] : a;
                        const Dev &d = m.devs[t];
                        bool skip = false;
                        if (BATCH) {
                            // a scenario that finished earlier in this chunk of launches stays in the list until the
                            // host has seen it: its branches are skipped (the lane takes another pair next time round)
                            skip = d.ctrl->done != 0;
                            major = d.ctrl->inner + 1; zsel = d.ctrl->zsel;
                        }
                        if (!skip) { load_branch(d, d.zbuf[zsel], G - a * nline, major, col, L); branch::begin(L, T); }
                    } else L.phase = branch::DONE;
                }
            }
            double xl[6], xu[6];
            load_bounds(col, xl, xu);
            if (branch::eval_pass(L, eval, pass, xl, xu, max_auglag, mu_max, T)) {
                const int a = G / nline;
                store_branch(m.devs[BATCH ? bv.active[a] : a], G - a * nline, L);
                if (count_work) {
                    atomicAdd(&s_work[0], 1u); atomicAdd(&s_work[1], (unsigned)L.it_al); atomicAdd(&s_work[2], (unsigned)L.evals);
                    atomicAdd(&s_work[3], (unsigned)L.cg);
                    if (L.shifts) atomicAdd(&s_work[4], (unsigned)L.shifts);
                    if (L.rejected) atomicAdd(&s_work[5], (unsigned)L.rejected);
                    if (L.hit_max) atomicAdd(&s_work[6], (unsigned)L.hit_max);
                    atomicMax(&s_work[7], (unsigned)L.evals);
                }
                L.phase = branch::NEED;
            }
        }
        if (__all_sync(full, L.phase == branch::DONE || L.phase == branch::NEED)) {
            if (__all_sync(full, L.phase == branch::DONE)) break;
        }
        double xl[6], xu[6];
        load_bounds(col, xl, xu);
        branch::compute(L, xl, xu);
    }
    if (count_work) {
        __syncthreads();
        if (threadIdx.x < 7) { if (s_work[threadIdx.x]) atomicAdd(&counters->v[threadIdx.x], (unsigned long long)s_work[threadIdx.x]); }
        else if (threadIdx.x == 7) atomicMax(&counters->v[7], (unsigned long long)s_work[7]);
    }
}

// ---------------------------------------------------------------------------
// End of a fused iteration. grid = (blocks, max(T-1, 1)); row y handles the ramp vectors of period
// t = y + 1: z and lambda of the ramp consensus (mpacopf_admm_update_{z,l}_gpu.jl) and their residual
// sums. The last block adds everything up in a fixed order, forms the per-period norms
// sqrt(||period||^2 + ||ramp||^2) and their max (mpacopf_admm_update_residual_gpu.jl:30-60), runs the
// termination test of admm_two_level.jl:60-62 and advances every period's control block.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(MP_FIN_BLOCK) k_mp_finish(MpDev m) {
    __shared__ double red[4 * (MP_FIN_BLOCK / 32)];
    __shared__ bool is_last;
    Ctrl *c = m.ctrl;
    if (c->done) return;
    const int zsel = c->zsel;
    const double beta = c->beta;
    const int nbx = gridDim.x;
    double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
    if (m.T > 1) {
        const int t = blockIdx.y + 1;
        const double *vprev = m.devs[t - 1].v;
        const double *zold = m.r_z[zsel];
        double *znew = m.r_z[zsel ^ 1];
        for (int k = blockIdx.x * MP_FIN_BLOCK + threadIdx.x; k < m.ngen; k += nbx * MP_FIN_BLOCK) {
            const size_t o = (size_t)t * m.ngen + k;
            const double u = m.r_u[o], vp = vprev[2 * k], lz = m.r_lz[o], zo = zold[o];
            const double zn = z_update(lz, m.r_l[o], m.r_rho[o], u, vp, beta);
            znew[o] = zn;
            m.r_l[o] = l_update(lz, beta, zn);
            const double rp = u - vp + zn, rd = zn - zo, ab = rp - zn;
            acc[0] += rp * rp; acc[1] += rd * rd; acc[2] += zn * zn; acc[3] += ab * ab;
        }
        block_sum4<MP_FIN_BLOCK>(acc, red);
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) m.partials[((size_t)blockIdx.y * 4 + k) * nbx + blockIdx.x] = acc[k];
        __threadfence();
        const unsigned tk = atomicAdd(m.ticket, 1u);
        is_last = (tk == (unsigned)(nbx * gridDim.y - 1));
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // per-period norms, one thread per (period, norm): sqrt(||period vectors||^2 + ||ramp vectors||^2)
    for (int i = threadIdx.x; i < 4 * m.T; i += MP_FIN_BLOCK) {
        const int t = i >> 2, k = i & 3;
        double n = sqrt(__ldcg(&m.sums[i]));
        if (t > 0) {
            double s = 0.0;
            for (int b = 0; b < nbx; ++b) s += __ldcg(&m.partials[((size_t)(t - 1) * 4 + k) * nbx + b]);
            const double rn = sqrt(s);
            n = sqrt(n * n + rn * rn);
        }
        m.res_t[i] = n;
    }
    __syncthreads();
    __shared__ int s_done;
    __shared__ long long s_inner;
    if (threadIdx.x < 4) {                                   // max over periods (mpacopf_admm_update_residual_gpu.jl:54-59)
        double top = 0.0;
        for (int t = 0; t < m.T; ++t) top = fmax(top, m.res_t[4 * t + threadIdx.x]);
        c->res[threadIdx.x] = top;
        if (threadIdx.x == 0) {
            const long long inner = c->inner + 1;
            s_inner = inner;
            s_done = (top <= c->eps_pri || inner >= c->inner_limit) ? 1 : 0;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < m.T; t += MP_FIN_BLOCK) {
        Ctrl *p = m.pctrl[t];
        p->inner = s_inner; p->zsel = zsel ^ 1; p->next_line = 0; p->done = s_done;
    }
    if (threadIdx.x == 0) {
        c->inner = s_inner;
        c->zsel = zsel ^ 1;
        c->done = s_done;
        c->next_line = 0;
        *m.ticket = 0u;
    }
}

__global__ void k_mp_ctrl_begin(MpDev m, double beta, double eps_pri, long long inner0, long long inner_limit, int zsel) {
    const int t = threadIdx.x;
    if (t > m.T) return;
    Ctrl *c = (t == m.T) ? m.ctrl : m.pctrl[t];
    c->beta = beta; c->eps_pri = eps_pri; c->inner = inner0; c->inner_limit = inner_limit;
    c->done = 0; c->zsel = zsel; c->ticket = 0u; c->next_line = 0;
    if (t == m.T) *m.ticket = 0u;
}

// ---------------------------------------------------------------------------
// Step-wise operators on the ramp vectors of period t = blockIdx.y + 1 (one block per period).
// ---------------------------------------------------------------------------
enum MpOp : int {
    MP_INIT = 0,        // mpacopf_init_solution_gpu.jl:1-35 (a = rho_pq)
    MP_Z = 1,           // mpacopf_admm_update_z_gpu.jl:1-16   (a = beta)
    MP_L = 2,           // mpacopf_admm_update_l_gpu.jl        (a = beta)
    MP_LZ = 3,          // mpacopf_admm_update_lz_gpu.jl       (a = beta, b = MAX_MULTIPLIER)
    MP_RESIDUAL = 4,    // mpacopf_admm_update_residual_gpu.jl:36-44 -> out[4t..4t+3] = sums of squares
    MP_ZNORM = 5,       // mpacopf_admm_prepoststep_gpu.jl:10-12  -> out[4t+2] = ||z_curr||^2
    MP_ZPREV = 6        // mpacopf_admm_prepoststep_gpu.jl:28-32: z_prev = z_curr
};

__global__ void __launch_bounds__(MP_FIN_BLOCK) k_mp_ramp_op(MpDev m, int zsel, int op, double a, double b, double *out) {
    __shared__ double red[4 * (MP_FIN_BLOCK / 32)];
    const int t = blockIdx.y + 1;
    const double *vprev = m.devs[t - 1].v;
    double *zc = m.r_z[zsel], *zp = m.r_z[zsel ^ 1];
    double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (int k = threadIdx.x; k < m.ngen; k += MP_FIN_BLOCK) {
        const size_t o = (size_t)t * m.ngen + k;
        switch (op) {
        case MP_INIT: {
            const double up = vprev[2 * k];
            m.r_u[o] = up;
            m.r_s[o] = m.devs[t].u[2 * k] - up;
            m.r_rho[o] = a;
            m.r_l[o] = 0.0; m.r_lz[o] = 0.0; zc[o] = 0.0; zp[o] = 0.0; m.r_rp[o] = 0.0; m.r_rd[o] = 0.0; m.r_axby[o] = 0.0;
            break;
        }
        case MP_Z: zc[o] = z_update(m.r_lz[o], m.r_l[o], m.r_rho[o], m.r_u[o], vprev[2 * k], a); break;
        case MP_L: m.r_l[o] = l_update(m.r_lz[o], a, zc[o]); break;
        case MP_LZ: m.r_lz[o] = fmax(-b, fmin(b, m.r_lz[o] + (a * zc[o]))); break;
        case MP_RESIDUAL: {
            const double rp = m.r_u[o] - vprev[2 * k] + zc[o], rd = zc[o] - zp[o], ab = rp - zc[o];
            m.r_rp[o] = rp; m.r_rd[o] = rd; m.r_axby[o] = ab;
            acc[0] += rp * rp; acc[1] += rd * rd; acc[2] += zc[o] * zc[o]; acc[3] += ab * ab;
            break;
        }
        case MP_ZNORM: acc[2] += zc[o] * zc[o]; break;
        case MP_ZPREV: zp[o] = zc[o]; break;
        }
    }
    if (op == MP_RESIDUAL || op == MP_ZNORM) {
        block_sum4<MP_FIN_BLOCK>(acc, red);
        if (threadIdx.x < 4) out[4 * t + threadIdx.x] = acc[threadIdx.x];
    }
}

// ramp vector <-> reference generator order: dst[gen_of_slot[k]] = src[k] (get) or the reverse (set)
__global__ void k_mp_permute(int n, const int *gen_of_slot, const double *src, double *dst, int to_reference) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (to_reference) dst[gen_of_slot[k]] = src[k];
    else dst[k] = src[gen_of_slot[k]];
}

}  // namespace ea

namespace ea {
// scenario batch: the control blocks of all scenarios into one array (one D2H copy per chunk of iterations)
__global__ void k_gather_ctrl(int n, Ctrl *const *pctrl, Ctrl *out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) out[s] = *pctrl[s];
}
}  // namespace ea
