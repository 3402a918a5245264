// qp_kernels.cuh — kernels of the one-level ADMM on the SQP sub-problem (`ModelQpsub`, src/models/qpsub/).
//
// The model reuses the single-period handle for its Solution vectors (bus-sorted HBM layout of kernels.cuh), the
// generator update (closed form, on the shifted bounds / costs) and the bus kernel (on the shifted loads); what is
// new is the branch kernel (qpsub.cuh) and the tail of the iteration. One ADMM iteration is THREE launches:
//   k_qp_xupdate   generators + branches (AL + TRON on the reduced QP) + v_prev <- v          (admm_update_x)
//   k_bus<false>   consensus update of xbar                                                    (admm_update_xbar)
//   k_qp_tail      l += rho (u - v); rp, rd, Ax+By; norms, objective, augmented Lagrangian;
//                  termination test and outer-iteration counter on the device                  (admm_update_l_single,
//                                                                                               admm_update_residual)
// against the reference's 3 + 1 + 1 + 9 launches, 3 CUBLAS norms and 4 device-to-host reductions per iteration.
#pragma once
#include "kernels.cuh"
#include "qpsub.cuh"

namespace ea {

struct QpCtrl {
    double res[8];            // primres, dualres, mismatch, objval, auglag, -, -, -
    double outer_tol, dual_tol;
    long long outer, outer_limit;
    int done;                 // 1: converged or outer == outer_limit; later launches are no-ops
    int solved;
    unsigned ticket;
    int next_line;            // work queue head of k_qp_xupdate (reset for the next x-update by k_qp_tail / k_qp_ctrl_begin)
};

struct QpDev {                // per-line arrays are SoA: row k of line I at [k * nline + I]
    const double *Hs;         // 21 x nline, packed lower triangle of the 6 x 6 blocks
    const double *lin;        // 16 x nline: LH_1h[4], RH_1h, LH_1i[4], RH_1i, LH_1j[2], RH_1j, LH_1k[2], RH_1k
    const double *lsus;       // 12 x nline: ls[6], us[6]
    const double *res;        // 4 x nline: line_res
    double *sqp_line;         // 6 x nline
    double *membuf;           // 5 x nline (rows 1-2 unused: 1h / 1i are eliminated)
    double *lambda;           // 4 x nline
    double *v_prev;           // nint
    QpCtrl *ctrl;
    double *partials;         // 8 x max_blocks
    unsigned long long *counters;   // branch calls, AL iterations, evaluations, max AL iterations of one call
};

constexpr int QBLOCK = 64;

// init_solution! (qpsub_init_solution_gpu.jl:1-98): generator midpoints, sqp_line = (ls + us) / 2, flows = supY sqp_line,
// rho_pq on generators / rho_va on ALL 8 branch entries. The host zeroes every vector and lambda first.
__global__ void k_qp_init_solution(Dev d, QpDev q, double rho_pq, double rho_va) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < d.gpad) d.rho[t] = (t < 2 * d.ngen) ? rho_pq : 0.0;
    if (t < d.ngen) {
        d.v[2 * t] = 0.5 * (d.pgmin[t] + d.pgmax[t]);
        d.v[2 * t + 1] = 0.5 * (d.qgmin[t] + d.qgmax[t]);
    }
    if (t < d.nline) {
        const int nl = d.nline;
        double Y[8], S[4][4], sq[6];
#pragma unroll
        for (int k = 0; k < 8; ++k) Y[k] = d.Y[k * nl + t];
        qpsub::sup_rows(Y, S);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            sq[k] = (q.lsus[k * nl + t] + q.lsus[(6 + k) * nl + t]) / 2;
            q.sqp_line[k * nl + t] = sq[k];
        }
        double fl[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) fl[r] = S[r][0] * sq[0] + S[r][1] * sq[1] + S[r][2] * sq[2] + S[r][3] * sq[3];
        d4 f, o, r;
        f.p = fl[0]; f.q = fl[1]; f.w = sq[2]; f.t = sq[4];
        o.p = fl[2]; o.q = fl[3]; o.w = sq[3]; o.t = sq[5];
        r.p = r.q = r.w = r.t = rho_va;
        double *vh = d.v + d.gpad, *rh = d.rho + d.gpad;
        st4(vh, d.slot_from[t], f);  st4(vh, d.slot_to[t], o);
        st4(rh, d.slot_from[t], r);  st4(rh, d.slot_to[t], r);
    }
}

// admm_update_x: generator_kernel_two_level on the qpsub bounds / costs + auglag_linelimit_qpsub.
//   major_arg > 0: info.inner supplied by the host (step-wise API)
//   major_arg == 0: fused loop (info.inner is 1 in every iteration of admm_one_level: mu restarts at 10), exits at once
//                   when the loop has finished, and also saves v_prev <- v for the dual residual.
// Persistent grid, one LANE per branch at a time, as in k_xupdate: branches come from a work queue (ctrl->next_line), a
// lane that finishes its branch takes the next one, and the AL / TRON loops are the state machine of qpsub.cuh - per
// round every live lane takes one TRON step, so the lanes of a warp run the same instructions although their branches
// need 1 ... 30 AL iterations. (Round 1: one lane ran one branch to completion inside nested loops, 4.4 of 32 lanes
// active per instruction.) The per-branch constants (47 doubles) and the bounds sit in a shared-memory column per lane.
constexpr int QTILE_ROWS = qpsub::S_ROWS + 8;                // + xl[2..5], xu[2..5]

__global__ void __launch_bounds__(QBLOCK)
k_qp_xupdate(Dev d, QpDev q, branch::PowTable T, long long major_arg, int zsel, int max_auglag, double mu_max, double scale,
             int do_gens, int do_lines) {
    __shared__ double tile[QTILE_ROWS * QBLOCK];
    const bool fused = major_arg == 0;
    if (fused && q.ctrl->done) return;
    const long long major = fused ? 1 : major_arg;
    const double *z = d.zbuf[zsel];
    const int tid = blockIdx.x * QBLOCK + threadIdx.x, nthr = gridDim.x * QBLOCK;
    if (do_gens)
        for (int k = tid; k < d.ngen; k += nthr) {
            if (fused) *reinterpret_cast<double2 *>(q.v_prev + 2 * k) = *reinterpret_cast<const double2 *>(d.v + 2 * k);
            generator_update(d, z, k);
        }
    if (!do_lines) return;
    const int nl = d.nline;
    const unsigned full = 0xffffffffu;
    unsigned work[3] = { 0, 0, 0 };
    int mx = 0;
    qpsub::TileStore<QBLOCK> st{ tile + threadIdx.x };
    double *bnd = tile + qpsub::S_ROWS * QBLOCK + threadIdx.x;
    // with fewer branches than resident lanes the branches are spread over all warps, a few lanes each (2869-like grid,
    // 4582 branches: 4 lanes per warp on 1184 warps instead of 32 lanes on 144): a round costs a warp less the fewer
    // different stages its lanes are in
    const int nwarps = gridDim.x * (QBLOCK / 32), lane = threadIdx.x & 31;
    const int lanes_on = min(32, max(1, (nl + nwarps - 1) / nwarps));
    qpsub::Lane L;
    L.phase = (lane < lanes_on) ? qpsub::NEED : qpsub::DONE;
    int I = -1;
#pragma unroll 1
    for (;;) {
        // refill: lanes without a branch take the next ones from the queue
        const bool need = (L.phase == qpsub::NEED);
        const unsigned m = __ballot_sync(full, need);
        if (m) {
            int base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(&q.ctrl->next_line, __popc(m));
            base = __shfl_sync(full, base, __ffs(m) - 1);
            if (need) {
                I = base + __popc(m & ((1u << lane) - 1u));
                if (I >= nl) L.phase = qpsub::DONE;
                else {
                    qpsub::Inputs in;
                    const int sf = d.slot_from[I], sto = d.slot_to[I];
                    {
                        const double *vh = d.v + d.gpad, *zh = z + d.gpad, *lh = d.l + d.gpad, *rh = d.rho + d.gpad;
                        const d4 lf = ld4(lh, sf), lt = ld4(lh, sto), rf = ld4(rh, sf), rt = ld4(rh, sto);
                        const d4 vf = ld4(vh, sf), vt = ld4(vh, sto), zf = ld4(zh, sf), zt = ld4(zh, sto);
                        if (fused) { st4(q.v_prev + d.gpad, sf, vf); st4(q.v_prev + d.gpad, sto, vt); }
                        const double lam[8] = { lf.p, lf.q, lt.p, lt.q, lf.w, lt.w, lf.t, lt.t };
                        const double rho[8] = { rf.p, rf.q, rt.p, rt.q, rf.w, rt.w, rf.t, rt.t };
                        const double xt[8] = { vf.p - zf.p, vf.q - zf.q, vt.p - zt.p, vt.q - zt.q, vf.w - zf.w, vt.w - zt.w, vf.t - zf.t, vt.t - zt.t };
#pragma unroll
                        for (int k = 0; k < 8; ++k) { in.lam[k] = lam[k]; in.rho[k] = rho[k]; in.xt[k] = xt[k]; in.Y[k] = d.Y[k * nl + I]; }
                    }
#pragma unroll
                    for (int k = 0; k < 21; ++k) in.H[k] = q.Hs[k * nl + I];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        in.res[k] = q.res[k * nl + I];
                        in.LH_1h[k] = q.lin[k * nl + I];
                        in.LH_1i[k] = q.lin[(5 + k) * nl + I];
                    }
                    in.RH_1h = q.lin[4 * nl + I]; in.RH_1i = q.lin[9 * nl + I];
                    in.LH_1j[0] = q.lin[10 * nl + I]; in.LH_1j[1] = q.lin[11 * nl + I]; in.RH_1j = q.lin[12 * nl + I];
                    in.LH_1k[0] = q.lin[13 * nl + I]; in.LH_1k[1] = q.lin[14 * nl + I]; in.RH_1k = q.lin[15 * nl + I];
                    double x0[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        x0[k] = q.sqp_line[(2 + k) * nl + I];
                        bnd[k * QBLOCK] = q.lsus[(2 + k) * nl + I];
                        bnd[(4 + k) * QBLOCK] = q.lsus[(8 + k) * nl + I];
                    }
                    const double mu = (major == 1) ? 10.0 : q.membuf[4 * nl + I];
                    qpsub::begin(L, in, st, x0, q.membuf[2 * nl + I], q.membuf[3 * nl + I], mu, T);
                }
            }
        }
        if (__all_sync(full, L.phase == qpsub::DONE)) break;
        qpsub::start_pass(L, st, scale);
        double xl[6], xu[6];
        xl[0] = 0.0; xl[1] = 0.0; xu[0] = 200000.0; xu[1] = 200000.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { xl[2 + k] = bnd[k * QBLOCK]; xu[2 + k] = bnd[(4 + k) * QBLOCK]; }
        if (qpsub::step_pass(L, st, xl, xu, max_auglag, mu_max, T)) {
            double Y[8], res[4];
#pragma unroll
            for (int k = 0; k < 8; ++k) Y[k] = d.Y[k * nl + I];
#pragma unroll
            for (int k = 0; k < 4; ++k) res[k] = q.res[k * nl + I];
            const double h2[2] = { q.lin[I], q.lin[nl + I] }, i2[2] = { q.lin[5 * nl + I], q.lin[6 * nl + I] };
            qpsub::Result R;
            qpsub::finish(L, st, Y, res, h2, i2, R);
            const int sf = d.slot_from[I], sto = d.slot_to[I];
            d4 of, ot;
            of.p = R.u[0]; of.q = R.u[1]; of.w = R.u[4]; of.t = R.u[6];
            ot.p = R.u[2]; ot.q = R.u[3]; ot.w = R.u[5]; ot.t = R.u[7];
            st4(d.u + d.gpad, sf, of);
            st4(d.u + d.gpad, sto, ot);
#pragma unroll
            for (int k = 0; k < 6; ++k) q.sqp_line[k * nl + I] = R.sqp[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) q.lambda[k * nl + I] = R.lambda[k];
            q.membuf[2 * nl + I] = L.lam_j; q.membuf[3 * nl + I] = L.lam_k; q.membuf[4 * nl + I] = L.mu;
            work[0] += 1; work[1] += (unsigned)R.it; work[2] += (unsigned)R.evals;
            mx = max(mx, R.it);
            L.phase = qpsub::NEED;
        }
    }
    if (d.count_work) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) work[k] += __shfl_down_sync(full, work[k], o);
            mx = max(mx, __shfl_down_sync(full, mx, o));
        }
        if (lane == 0 && work[0]) {
#pragma unroll
            for (int k = 0; k < 3; ++k) atomicAdd(&q.counters[k], (unsigned long long)work[k]);
            atomicMax(&q.counters[3], (unsigned long long)mx);
        }
    }
}

// admm_update_xbar inside the fused loop: the bus kernel of kernels.cuh (xbar only), skipped once the loop has
// finished so that the launches left in a chunk do not touch the final state.
__global__ void __launch_bounds__(BBLOCK, 5)
k_qp_bus(Dev d, const QpCtrl *c, int zsel) {
    __shared__ double red[4 * (BBLOCK / 32)];
    if (c->done) return;
    bus_body<false>(d, zsel, 0.0, red);
}

// admm_update_l_single (qpsub_admm_update_l_single_gpu.jl): l += rho (u - v)
__global__ void k_qp_l_single(int n, double *l, const double *rho, const double *u, const double *v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) l[i] = l[i] + rho[i] * (u[i] - v[i]);
}

constexpr int QTBLOCK = 256;

template <int K>
__device__ __forceinline__ void block_sumK(double (&acc)[K], double *smem /* K * QTBLOCK / 32 */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) smem[k * (QTBLOCK / 32) + wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < QTBLOCK / 32; ++w) v += smem[threadIdx.x * (QTBLOCK / 32) + w];
        smem[threadIdx.x * (QTBLOCK / 32)] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = smem[k * (QTBLOCK / 32)];
    __syncthreads();
}

// Tail of the iteration. MODE 0: admm_update_residual alone (step-wise API); MODE 1: fused loop = l update + residual +
// termination test; MODE 2: objective / augmented Lagrangian from the stored rp (admm_poststep).
// Sums (deterministic: per-block partials added in a fixed order by the last block):
//   0 |rp|^2   1 |rd|^2   2 sum l rp   3 sum rho rp^2   4 sum lz z   5 generator cost   6 sum 0.5 x'Hs x   7 sum z^2
template <int MODE>
__global__ void __launch_bounds__(QTBLOCK)
k_qp_tail(Dev d, QpDev q, int zsel, double beta, double *out) {
    __shared__ double red[8 * (QTBLOCK / 32)];
    __shared__ bool is_last;
    if (MODE == 1 && q.ctrl->done) return;
    const int tid = blockIdx.x * QTBLOCK + threadIdx.x, nthr = gridDim.x * QTBLOCK;
    const double *z = d.zbuf[zsel];
    double acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int i = tid; i < d.nint; i += nthr) {
        const double rho = d.rho[i];
        double l = d.l[i], rp;
        if (MODE == 2) rp = d.rp[i];
        else {
            const double u = d.u[i], v = d.v[i];
            rp = u - v;
            if (MODE == 1) { l = l + rho * rp; d.l[i] = l; }
            const double rd = rho * (v - q.v_prev[i]);
            d.rp[i] = rp; d.rd[i] = rd; d.axby[i] = rp;
            acc[0] += rp * rp; acc[1] += rd * rd;
        }
        const double zi = z[i];
        acc[2] += l * rp; acc[3] += rho * (rp * rp); acc[4] += d.lz[i] * zi; acc[7] += zi * zi;
    }
    for (int k = tid; k < d.ngen; k += nthr) {
        const double pg = d.baseMVA * d.u[2 * k];
        acc[5] += d.c2[k] * (pg * pg) + d.c1[k] * pg;
    }
    const int nl = d.nline;
    for (int I = tid; I < nl; I += nthr) {
        double x[6], xHx = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) x[k] = q.sqp_line[k * nl + I];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j) s += q.Hs[tron::tri(i, j) * nl + I] * x[j];
            xHx += x[i] * s;
        }
        acc[6] += 0.5 * xHx;
    }
    block_sumK<8>(acc, red);
    const int nb = gridDim.x;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) q.partials[k * nb + blockIdx.x] = acc[k];
        __threadfence();
        is_last = (atomicAdd(&q.ctrl->ticket, 1u) == (unsigned)(nb - 1));
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double a[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int b = threadIdx.x; b < nb; b += QTBLOCK) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += __ldcg(&q.partials[k * nb + b]);
    }
    block_sumK<8>(a, red);
    if (threadIdx.x == 0) {
        q.ctrl->ticket = 0u;
        if (MODE == 1) q.ctrl->next_line = 0;
        const double objval = a[5] + a[6];
        // qpsub_admm_update_residual_gpu.jl:30-51; par.beta = 0 on the one-level path
        const double auglag = objval + a[4] + 0.5 * beta * a[7] + a[2] + 0.5 * a[3];
        if (MODE == 2) { out[0] = objval; out[1] = auglag; return; }
        const double primres = sqrt(a[0]), dualres = sqrt(a[1]);
        double *r = (MODE == 1) ? q.ctrl->res : out;
        r[0] = primres; r[1] = dualres; r[2] = primres; r[3] = objval; r[4] = auglag;
        if (MODE == 1) {                                            // admm_one_level.jl:36-67
            const long long outer = q.ctrl->outer + 1;
            q.ctrl->outer = outer;
            const bool solved = primres <= q.ctrl->outer_tol && dualres <= q.ctrl->dual_tol;
            if (solved) q.ctrl->solved = 1;
            if (solved || outer >= q.ctrl->outer_limit) q.ctrl->done = 1;
        }
    }
}

__global__ void k_qp_ctrl_begin(QpCtrl *c, double outer_tol, double dual_tol, long long outer0, long long limit) {
    c->outer_tol = outer_tol; c->dual_tol = dual_tol; c->outer = outer0; c->outer_limit = limit;
    c->done = (outer0 >= limit) ? 1 : 0; c->solved = 0; c->ticket = 0u; c->next_line = 0;
    for (int k = 0; k < 8; ++k) c->res[k] = 0.0;
}

}  // namespace ea
