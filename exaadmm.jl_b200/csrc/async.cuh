// async.cuh — the inner loop of the two-level ADMM as ONE persistent dataflow kernel.
//
// The synchronous fused loop (kernels.cuh) runs, per inner iteration, the branch kernel and then the bus kernel. The
// branch kernel lasts as long as its slowest branch - a branch whose line limit becomes active runs ~25 TRON solves
// in a row, ~20x the mean - so most of the machine idles for most of every iteration.
// Nothing in the algorithm asks for that barrier: the x-update of branch b at iteration r needs only the bus updates
// of its two end buses at r-1, and the update of bus i at r needs only the x-updates of its incident branches at r.
// The only global step is the termination test (norms over all entries). This kernel therefore runs the iteration
// as a task graph:
//   X(b, r)  branch b, iteration r: ready when both end buses finished r-1
//   B(i, r)  bus i (its generators included), iteration r: ready when all incident branches finished r
//   R(r)     norms + termination test of iteration r: ready when all buses finished r
// Branches far from a straggler run ahead of it, up to AD-1 iterations beyond the last evaluated termination test;
// state lives in a ring of AD iteration slots, so when R(K) says stop the state of iteration K (and z of K-1) is
// still intact and everything computed for r > K is simply discarded. The arithmetic of every task is that of the
// synchronous kernels (same device functions), so the iterates are the same; only the schedule differs.
//
// Roles by CTA: block 0 = reducer (R tasks, fixed summation order -> deterministic norms), blocks 1..nb-1 = bus
// workers (one bus per lane, in lockstep), the rest = branch workers (the state machine of branch.cuh, one branch
// per lane, refilled from the X queue). Tasks travel through two multi-producer / multi-consumer ring queues
// (ticket per consumer, entry = (iteration << 32) | (id + 1), 0 = empty); readiness is tracked by monotonic arrival
// counters (bus: arrivals / degree, branch: arrivals / 2). All CTAs are resident (grid <= occupancy), every wait is a
// poll with a deadline: on timeout the kernel raises `abort` and returns instead of hanging.
#pragma once
#include "kernels.cuh"

namespace ea {

constexpr int AD = 6;                       // ring depth: look-ahead of up to AD - 1 iterations
constexpr int APOLL_CHECK = 64;             // idle polls between two looks at the clock
constexpr int BCH = 4;                      // bus task: branch ends loaded per chunk

struct AsyncCtrl {
    unsigned xq_head, xq_tail, bq_head, bq_tail;
    int tested;                             // last iteration whose termination test came out negative
    int halt;                               // 0 running, 1 finished (stop_iter is the last iteration), 2 gave up (timeout)
    int stop_iter, pad0;
    int limit;                              // last iteration allowed (relative to the start of the run)
    int pad;
    double eps_pri, beta;
    double res[4];
    unsigned iter_cnt[AD];                  // buses finished, per ring slot
    int complete[AD];                       // iteration number whose buses are all done, per ring slot
    unsigned long long deadline;            // globaltimer value after which waiting gives up
    unsigned long long n_x, n_b;            // tasks executed (diagnostics)
    // cycle split of the workers (count_work > 1), summed over warps: branch workers 0 refill, 1 evaluation, 2 completion,
    // 3 compute, 4 idle, 5 rounds; bus workers 8 poll, 9 task, 10 arrivals, 11 idle, 12 batches, 13 lanes run
    unsigned long long diag[16];
};

struct AsyncDev {
    double *u[AD], *v[AD], *z[AD], *l[AD];  // ring of state vectors (nint each)
    double *als[AD];                        // 3 x nline: lambda_s1, lambda_s2, mu
    double *bsum[AD];                       // 4 x nbus: per-bus residual sums
    unsigned long long *xq, *bq;
    unsigned xmask, bmask;
    unsigned *bus_cnt, *br_cnt;
    AsyncCtrl *c;
    long long major0;                       // info.inner of iteration r is major0 + r
    int nb;                                 // CTAs 0..nb-1 are the reducer + bus workers
    const int *iso_bus;                     // buses without branches (they follow themselves), n_iso of them
    int n_iso;
};

__device__ __forceinline__ int ld_acq(const int *p) {
    int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long ld_acq64(const unsigned long long *p) {
    unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_rel(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_rel64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
__device__ __forceinline__ d4 ld4cg(const double *base, int slot) {          // data produced by other SMs: bypass L1
    const double2 a = __ldcg(reinterpret_cast<const double2 *>(base + 4 * (size_t)slot));
    const double2 b = __ldcg(reinterpret_cast<const double2 *>(base + 4 * (size_t)slot + 2));
    d4 r; r.p = a.x; r.q = a.y; r.w = b.x; r.t = b.y; return r;
}
__device__ __forceinline__ double2 ld2cg(const double *p) { return __ldcg(reinterpret_cast<const double2 *>(p)); }

__device__ __forceinline__ int ld_rlx(const int *p) {
    int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long ld_rlx64(const unsigned long long *p) {
    unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_rlx64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// atomicAdd with release semantics: MEMBAR.ALL.GPU + ATOMG. Unlike __threadfence() (MEMBAR.SC.GPU + CCTL.IVALL) it does
// not invalidate the SM's L1, where this kernel's register spills live; the consumers read task data with L1-bypassing
// loads issued after (control-dependent on) the relaxed load that saw the counter / queue entry, so no acquire fence
// (another CCTL.IVALL) is needed on their side either.
__device__ __forceinline__ unsigned atom_add_rel(unsigned *p, unsigned v) {
    unsigned old; asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory"); return old;
}
__device__ __forceinline__ unsigned long long task_entry(int id, int r) {
    return ((unsigned long long)(unsigned)r << 32) | (unsigned)(id + 1);
}
// Push one task per participating lane with ONE atomic on the queue tail for the whole warp (the tails are the
// hottest words of the kernel). Called by all 32 lanes; the data the task depends on must already be fenced.
__device__ __forceinline__ void warp_push(unsigned long long *q, unsigned mask, unsigned *tail, bool has,
                                          unsigned long long entry, int lane) {
    const unsigned m = __ballot_sync(0xffffffffu, has);
    if (!m) return;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(tail, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (has) st_rlx64(&q[(base + __popc(m & ((1u << lane) - 1u))) & mask], entry);
}

// out of time: raise the abort flag (never overrides a regular finish)
__device__ __forceinline__ void async_give_up(AsyncCtrl *c) { atomicCAS(&c->halt, 0, 2); }


// ---- X(b, r): stage branch I from ring slot r-1 (load_branch of kernels.cuh with L1-bypassing loads) ----------------
__device__ __forceinline__ void async_load_branch(const Dev &d, const AsyncDev &a, int I, int r, double *col, branch::Lane &L) {
    const int nl = d.nline, in = (r - 1) % AD;
    const int sf = d.slot_from[I], st = d.slot_to[I];
    const double *uh = a.u[in] + d.gpad, *vh = a.v[in] + d.gpad, *zh = a.z[in] + d.gpad, *lh = a.l[in] + d.gpad, *rh = d.rho + d.gpad;
    const d4 lf = ld4cg(lh, sf), lt = ld4cg(lh, st);
    const d4 rf = ld4(rh, sf), rt = ld4(rh, st);                              // rho is constant during a run
    const d4 vf = ld4cg(vh, sf), vt = ld4cg(vh, st), zf = ld4cg(zh, sf), zt = ld4cg(zh, st);
    const d4 uf = ld4cg(uh, sf), ut = ld4cg(uh, st);
    const double lam[8] = { lf.p, lf.q, lt.p, lt.q, lf.w, lt.w, lf.t, lt.t };
    const double rho[8] = { rf.p, rf.q, rt.p, rt.q, rf.w, rt.w, rf.t, rt.t };
    const double xt[8] = { vf.p - zf.p, vf.q - zf.q, vt.p - zt.p, vt.q - zt.q, vf.w - zf.w, vt.w - zt.w, vf.t - zf.t, vt.t - zt.t };
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        col[k * XBLOCK] = lam[k];
        col[(8 + k) * XBLOCK] = rho[k];
        col[(16 + k) * XBLOCK] = xt[k];
        col[(24 + k) * XBLOCK] = d.Y[k * nl + I];
    }
    double b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { b[k] = d.xlu[k * nl + I]; col[(32 + k) * XBLOCK] = b[k]; }
    const double ra = d.rateA[I];
    col[40 * XBLOCK] = ra;
    L.x[0] = fmin(b[1], fmax(b[0], sqrt(uf.w)));
    L.x[1] = fmin(b[3], fmax(b[2], sqrt(ut.w)));
    L.x[2] = fmin(b[5], fmax(b[4], uf.t));
    L.x[3] = fmin(b[7], fmax(b[6], ut.t));
    L.x[4] = fmin(0.0, fmax(-ra, -(uf.p * uf.p + uf.q * uf.q)));
    L.x[5] = fmin(0.0, fmax(-ra, -(ut.p * ut.p + ut.q * ut.q)));
    const double *als = a.als[in];
    L.ls[0] = __ldcg(als + I);
    L.ls[1] = __ldcg(als + nl + I);
    L.mu = (a.major0 + r == 1) ? 10.0 : __ldcg(als + 2 * nl + I);          // auglag_gpu.jl:75-80
}

// ---- B(i, r): generators of the bus (closed form) + consensus update + z / lambda + residual sums of its entries -----
// Same sequence of operations as bus_scalar<true> (kernels.cuh); inputs from ring slot r-1 (z, lambda, xbar of the
// generators) and r (u of the incident branch ends), outputs to slot r.
__device__ __forceinline__ void async_bus_task(const Dev &d, const AsyncDev &a, int b, int r, double beta) {
    const int in = (r - 1) % AD, out = r % AD;
    const int hs = d.hstart[b], he = d.hstart[b + 1], gs = d.gstart[b], ge = d.gstart[b + 1];
    const double *zi = a.z[in], *li = a.l[in];
    double *uo = a.u[out], *vo = a.v[out], *zo = a.z[out], *lo = a.l[out];
    const double B = d.baseMVA;
    double common_wi = 0.0, common_ti = 0.0, inv_p = 0.0, inv_q = 0.0, rs_w = 0.0, rs_t = 0.0;
    double rhs1 = 0.0, rhs2 = 0.0, inv_pg = 0.0, inv_qg = 0.0;
    double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (int k = gs; k < ge; ++k) {
        // generator_update (kernels.cuh) on slot r-1
        const double2 x = ld2cg(a.v[in] + 2 * k), zz = ld2cg(zi + 2 * k), l = ld2cg(li + 2 * k);
        const double2 rho = *reinterpret_cast<const double2 *>(d.rho + 2 * k);
        double2 u;
        u.x = fmax(d.pgmin_curr[k], fmin(d.pgmax_curr[k],
                   (-(d.c1[k] * B + l.x + rho.x * (-x.x + zz.x))) / (2 * d.c2[k] * (B * B) + rho.x)));
        u.y = fmax(d.qgmin[k], fmin(d.qgmax[k], (-(l.y + rho.y * (-x.y + zz.y))) / rho.y));
        *reinterpret_cast<double2 *>(uo + 2 * k) = u;
        // bus_gen_gather
        const double iy = tron::ddiv(1.0, rho.y);
        rhs2 += (u.y + zz.y) + (l.y * iy);
        inv_qg += iy;
        const double ix = tron::ddiv(1.0, rho.x);
        rhs1 += (u.x + zz.x) + (l.x * ix);
        inv_pg += ix;
    }
    rhs1 -= d.pd_pu[b];
    rhs2 -= d.qd_pu[b];
    const double *uh = uo + d.gpad, *zh = zi + d.gpad, *lh = li + d.gpad, *rh = d.rho + d.gpad;
    // the ends in chunks of BCH: all loads of a chunk are issued before the first use (one L2 round trip per chunk
    // instead of one per end); the accumulation order is that of the plain loop
    for (int s0 = hs; s0 < he; s0 += BCH) {
        d4 uu[BCH], zz[BCH], ll[BCH], rr[BCH];
#pragma unroll
        for (int j = 0; j < BCH; ++j) {
            const int s = min(s0 + j, he - 1);
            uu[j] = ld4cg(uh, s); zz[j] = ld4cg(zh, s); ll[j] = ld4cg(lh, s); rr[j] = ld4(rh, s);
        }
#pragma unroll
        for (int j = 0; j < BCH; ++j) {
            if (s0 + j < he) {
                const d4 &u = uu[j], &z = zz[j], &l = ll[j], &r4 = rr[j];
                const double irp = tron::ddiv(1.0, r4.p), irq = tron::ddiv(1.0, r4.q);
                common_wi += l.w + r4.w * (u.w + z.w);
                common_ti += l.t + r4.t * (u.t + z.t);
                inv_p += irp;
                inv_q += irq;
                rs_w += r4.w;
                rs_t += r4.t;
                rhs1 -= (u.p + z.p) + (l.p * irp);
                rhs2 -= (u.q + z.q) + (l.q * irq);
            }
        }
    }
    const BusSolve bs = bus_solve(d, b, common_wi, common_ti, inv_p, inv_q, rs_w, rs_t, rhs1, rhs2, inv_pg, inv_qg);
    for (int k = gs; k < ge; ++k) {                        // bus_gen_scatter<true>
        const double2 u = ld2cg(uo + 2 * k);                                  // written by this lane above
        const double2 z = ld2cg(zi + 2 * k), l = ld2cg(li + 2 * k);
        const double2 rr = *reinterpret_cast<const double2 *>(d.rho + 2 * k);
        const double2 lz = *reinterpret_cast<const double2 *>(d.lz + 2 * k);
        double2 v, zn, ln;
        v.x = (u.x + z.x) + (l.x - bs.mu1) * tron::ddiv(1.0, rr.x);
        v.y = (u.y + z.y) + (l.y - bs.mu2) * tron::ddiv(1.0, rr.y);
        zn.x = z_update(lz.x, l.x, rr.x, u.x, v.x, beta);
        zn.y = z_update(lz.y, l.y, rr.y, u.y, v.y, beta);
        ln.x = l_update(lz.x, beta, zn.x);
        ln.y = l_update(lz.y, beta, zn.y);
        *reinterpret_cast<double2 *>(vo + 2 * k) = v;
        *reinterpret_cast<double2 *>(zo + 2 * k) = zn;
        *reinterpret_cast<double2 *>(lo + 2 * k) = ln;
        const double rpx = u.x - v.x + zn.x, rpy = u.y - v.y + zn.y;
        const double rdx = zn.x - z.x, rdy = zn.y - z.y;
        const double abx = rpx - zn.x, aby = rpy - zn.y;
        acc[0] += rpx * rpx + rpy * rpy;
        acc[1] += rdx * rdx + rdy * rdy;
        acc[2] += zn.x * zn.x + zn.y * zn.y;
        acc[3] += abx * abx + aby * aby;
    }
    for (int s0 = hs; s0 < he; s0 += BCH) {                // bus_end_scatter<true>, chunked like the gather
        d4 uu[BCH], zz[BCH], ll[BCH], rr[BCH], lzz[BCH];
#pragma unroll
        for (int j = 0; j < BCH; ++j) {
            const int s = min(s0 + j, he - 1);
            uu[j] = ld4cg(uh, s); zz[j] = ld4cg(zh, s); ll[j] = ld4cg(lh, s); rr[j] = ld4(rh, s); lzz[j] = ld4(d.lz + d.gpad, s);
        }
#pragma unroll
        for (int j = 0; j < BCH; ++j) {
            if (s0 + j < he) {
                const int s = s0 + j;
                const d4 &u = uu[j], &z = zz[j], &l = ll[j], &r4 = rr[j], &lz = lzz[j];
                const double irp = tron::ddiv(1.0, r4.p), irq = tron::ddiv(1.0, r4.q);
                d4 v, zn, ln;
                v.p = (u.p + z.p) + (l.p + bs.mu1) * irp;
                v.q = (u.q + z.q) + (l.q + bs.mu2) * irq;
                v.w = bs.wi;
                v.t = bs.ti;
                zn.p = z_update(lz.p, l.p, r4.p, u.p, v.p, beta);
                zn.q = z_update(lz.q, l.q, r4.q, u.q, v.q, beta);
                zn.w = z_update(lz.w, l.w, r4.w, u.w, v.w, beta);
                zn.t = z_update(lz.t, l.t, r4.t, u.t, v.t, beta);
                ln.p = l_update(lz.p, beta, zn.p);
                ln.q = l_update(lz.q, beta, zn.q);
                ln.w = l_update(lz.w, beta, zn.w);
                ln.t = l_update(lz.t, beta, zn.t);
                st4(vo + d.gpad, s, v);
                st4(zo + d.gpad, s, zn);
                st4(lo + d.gpad, s, ln);
                const double ua[4] = { u.p, u.q, u.w, u.t }, va[4] = { v.p, v.q, v.w, v.t };
                const double za[4] = { zn.p, zn.q, zn.w, zn.t }, zold[4] = { z.p, z.q, z.w, z.t };
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double rp = ua[k] - va[k] + za[k];
                    const double rd = za[k] - zold[k];
                    const double ab = rp - za[k];
                    acc[0] += rp * rp; acc[1] += rd * rd; acc[2] += za[k] * za[k]; acc[3] += ab * ab;
                }
            }
        }
    }
    double *bs4 = a.bsum[out] + 4 * (size_t)b;
    *reinterpret_cast<double2 *>(bs4) = make_double2(acc[0], acc[1]);
    *reinterpret_cast<double2 *>(bs4 + 2) = make_double2(acc[2], acc[3]);
}


__global__ void __launch_bounds__(XBLOCK, EA_XMINB)
k_async(Dev d, AsyncDev a, branch::PowTable T, int max_auglag, double mu_max, double scale) {
    extern __shared__ double tile[];
    AsyncCtrl *c = a.c;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;

    // ------------------------------------------------------------------------------------------------ reducer
    if (blockIdx.x == 0) {
        __shared__ int go;
        double *red = tile;
        const double eps_pri = c->eps_pri;
        const int limit = c->limit;
        for (int r = 1;; ++r) {
            __syncthreads();
            if (threadIdx.x == 0) {
                int polls = 0, ok = 1;
                while (ld_rlx(&c->complete[r % AD]) != r) {
                    if (ld_rlx(&c->halt)) { ok = 0; break; }
                    if (++polls % APOLL_CHECK == 0 && gtime() > c->deadline) { async_give_up(c); ok = 0; break; }
                    __nanosleep(100);
                }
                go = ok;
            }
            __syncthreads();
            if (!go) return;
            const double *bs = a.bsum[r % AD];
            double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
            for (int i = threadIdx.x; i < d.nbus; i += XBLOCK) {
                const double2 p = ld2cg(bs + 4 * (size_t)i), q = ld2cg(bs + 4 * (size_t)i + 2);
                acc[0] += p.x; acc[1] += p.y; acc[2] += q.x; acc[3] += q.y;
            }
            block_sum4<XBLOCK>(acc, red);
            __syncthreads();
            if (threadIdx.x == 0) {
                const double primres = sqrt(acc[0]);
                const bool stop = primres <= eps_pri || r >= limit;   // admm_two_level.jl:34,60-62
                if (stop) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) c->res[k] = sqrt(acc[k]);
                    c->stop_iter = r;
                    atomicCAS(&c->halt, 0, 1);
                } else st_rel(&c->tested, r);
                go = stop ? 0 : 1;
            }
            __syncthreads();
            if (!go) return;
        }
    }

    // ------------------------------------------------------------------------------------------------ bus workers
    if (blockIdx.x < a.nb) {
        const double beta = c->beta;
        const int limit = c->limit;
        int st = 0, I = -1, r = 0;
        unsigned ticket = 0;
        int polls = 0;
        unsigned long long done_tasks = 0;
        const bool prof = d.count_work > 1;
        long long tq = 0, tt = 0, ta = 0, ti = 0, nbatch = 0, nrun = 0, t0 = 0, t1 = 0;
        for (;;) {
            if (prof) t0 = clock64();
            const int hflag = ld_rlx(&c->halt);
            const int tested = ld_rlx(&c->tested);
            const unsigned m = __ballot_sync(full, st == 0);           // tickets, one atomic per warp
            if (m) {
                unsigned base = 0;
                if (lane == __ffs(m) - 1) base = atomicAdd(&c->bq_head, (unsigned)__popc(m));
                base = __shfl_sync(full, base, __ffs(m) - 1);
                if (st == 0) { ticket = base + __popc(m & lt); st = 1; }
            }
            if (st == 1) {
                const unsigned long long e = ld_rlx64(&a.bq[ticket & a.bmask]);
                if (e) { a.bq[ticket & a.bmask] = 0ull; I = (int)(unsigned)(e & 0xffffffffull) - 1; r = (int)(e >> 32); st = 2; }
            }
            if (__any_sync(full, hflag != 0)) break;                   // (lanes may have read the flag at different times)
            const bool run = st == 2 && r <= tested + AD - 1;          // only buses without branches can reach the gate
            if (!__any_sync(full, run)) {
                if (++polls % APOLL_CHECK == 0 && lane == 0 && gtime() > c->deadline) async_give_up(c);
                __nanosleep(100);
                if (prof) ti += clock64() - t0;
                continue;
            }
            if (prof) { t1 = clock64(); tq += t1 - t0; nbatch++; nrun += __popc(__ballot_sync(full, run)); }
            if (run) async_bus_task(d, a, I, r, beta);
            if (prof) { __syncwarp(); t0 = clock64(); tt += t0 - t1; }
            // arrivals at the incident branches: atomics of a chunk are issued together, the branches that became
            // ready are pushed with one tail atomic per warp and chunk
            const int hs = run ? d.hstart[I] : 0, he = run ? d.hstart[I + 1] : 0;
            const int maxdeg = __reduce_max_sync(full, he - hs);
            for (int j0 = 0; j0 < maxdeg; j0 += BCH) {
                int bid[BCH]; unsigned old[BCH];
#pragma unroll
                for (int j = 0; j < BCH; ++j) {
                    const bool on = hs + j0 + j < he;
                    bid[j] = on ? d.slot_line[hs + j0 + j] : -1;
                }
#pragma unroll
                for (int j = 0; j < BCH; ++j) old[j] = bid[j] >= 0 ? atom_add_rel(&a.br_cnt[bid[j]], 1u) : 0u;   // outputs first
#pragma unroll
                for (int j = 0; j < BCH; ++j)
                    warp_push(a.xq, a.xmask, &c->xq_tail, bid[j] >= 0 && ((old[j] + 1u) & 1u) == 0u && r + 1 <= limit,
                              task_entry(bid[j], r + 1), lane);
            }
            warp_push(a.bq, a.bmask, &c->bq_tail, run && hs == he && r + 1 <= limit, task_entry(I, r + 1), lane);   // no branches: the bus follows itself
            // iteration bookkeeping, one atomic per group of lanes on the same iteration
            const unsigned act = __ballot_sync(full, run);
            if (run) {
                const unsigned peers = __match_any_sync(act, r);
                if (lane == __ffs(peers) - 1) {
                    const unsigned n = atom_add_rel(&c->iter_cnt[r % AD], (unsigned)__popc(peers)) + (unsigned)__popc(peers);
                    if (n == (unsigned)d.nbus) {
                        c->iter_cnt[r % AD] = 0u;
                        st_rel(&c->complete[r % AD], r);
                    }
                }
                done_tasks++;
                st = 0;
            }
            if (prof) { __syncwarp(); ta += clock64() - t0; }
        }
        if (d.count_work && done_tasks) atomicAdd(&c->n_b, done_tasks);
        if (prof && lane == 0) {
            atomicAdd(&c->diag[8], (unsigned long long)tq); atomicAdd(&c->diag[9], (unsigned long long)tt);
            atomicAdd(&c->diag[10], (unsigned long long)ta); atomicAdd(&c->diag[11], (unsigned long long)ti);
            atomicAdd(&c->diag[12], (unsigned long long)nbatch); atomicAdd(&c->diag[13], (unsigned long long)nrun);
        }
        return;
    }

    // ------------------------------------------------------------------------------------------------ branch workers
    // Every global round trip of the task plumbing is issued early and consumed late, so that the TRON arithmetic of
    // the round hides it: flags and the poll of the lane's NEXT task at the top of the round, tickets one round ahead.
    // A lane takes the ticket of its next task when it starts the current one; if that task arrives while the lane is
    // still busy for more than AHOLD rounds (a long augmented-Lagrangian chain) the lane hands it back to the queue.
    double *col = tile + threadIdx.x;
    const branch::Objective<branch::TileView<XBLOCK>> eval{ { col }, scale };
    const int limit = c->limit;
    constexpr int AHOLD = 2;
    branch::Lane L;
    L.phase = branch::NEED;
    L.step_pending = false;
    int I = -1, r = 0;                       // current task
    int st2 = 0, I2 = -1, r2 = 0, age2 = 0;  // next task: 0 none, 1 ticket taken, 2 arrived, 3 handed back (no prefetch until idle)
    unsigned ticket2 = 0;
    int polls = 0;
    unsigned long long done_tasks = 0;
    const bool prof = d.count_work > 1;
    long long tr = 0, te = 0, tc = 0, tp = 0, ti = 0, nround = 0, t0 = 0, t1 = 0;
#pragma unroll 1
    for (;;) {
        if (prof) t0 = clock64();
        const int hflag = ld_rlx(&c->halt);
        const int tested = ld_rlx(&c->tested);
        unsigned long long e = 0ull;
        if (st2 == 1) e = ld_rlx64(&a.xq[ticket2 & a.xmask]);
        double xl[6], xu[6];
        load_bounds(col, xl, xu);
        // ---- pass 0: lanes holding a trial point
        const bool fin = branch::eval_pass(L, eval, 0, xl, xu, max_auglag, mu_max, T);
        if (prof) { __syncwarp(); t1 = clock64(); te += t1 - t0; }
        if (st2 == 1 && e) {                                            // the next task has arrived
            a.xq[ticket2 & a.xmask] = 0ull;
            I2 = (int)(unsigned)(e & 0xffffffffull) - 1; r2 = (int)(e >> 32); age2 = 0;
            st2 = (r2 > limit) ? 0 : 2;                                 // beyond the last iteration allowed: void
        }
        // ---- completion: u and the AL state into ring slot r, then tell the two end buses
        const unsigned finm = __ballot_sync(full, fin);
        if (finm) {
            int bus0 = -1, bus1 = -1;
            bool fire0 = false, fire1 = false;
            if (fin) {
                const int nl = d.nline, out = r % AD;
                d4 of, ot;
                of.p = L.Fc[0]; of.q = L.Fc[1]; of.w = L.x[0] * L.x[0]; of.t = L.x[2];
                ot.p = L.Fc[2]; ot.q = L.Fc[3]; ot.w = L.x[1] * L.x[1]; ot.t = L.x[3];
                st4(a.u[out] + d.gpad, d.slot_from[I], of);
                st4(a.u[out] + d.gpad, d.slot_to[I], ot);
                double *als = a.als[out];
                als[I] = L.ls[0]; als[nl + I] = L.ls[1]; als[2 * nl + I] = L.mu;
                bus0 = d.br_from[I]; bus1 = d.br_to[I];
                const unsigned deg0 = (unsigned)(d.hstart[bus0 + 1] - d.hstart[bus0]);
                const unsigned deg1 = (unsigned)(d.hstart[bus1 + 1] - d.hstart[bus1]);
                const unsigned old0 = atom_add_rel(&a.bus_cnt[bus0], 1u);      // u and the AL state first
                const unsigned old1 = atom_add_rel(&a.bus_cnt[bus1], 1u);
                fire0 = (old0 + 1u) % deg0 == 0u;
                fire1 = (old1 + 1u) % deg1 == 0u;
                done_tasks++;
                L.phase = branch::NEED;
            }
            warp_push(a.bq, a.bmask, &c->bq_tail, fire0, task_entry(bus0, r), lane);
            warp_push(a.bq, a.bmask, &c->bq_tail, fire1, task_entry(bus1, r), lane);
        }
        if (prof) { __syncwarp(); t0 = clock64(); tc += t0 - t1; }
        // ---- next task for the lanes that are free
        const bool need = L.phase == branch::NEED;
        if (need && st2 == 2 && r2 <= tested + AD - 1) {               // ring slot r2 is free to write
            I = I2; r = r2;
            async_load_branch(d, a, I, r, col, L);
            branch::begin(L, T);
            st2 = 0;
        }
        {   // tickets, one atomic per warp: lanes that have just started a task, and idle lanes without one
            const bool want = st2 == 0 || (st2 == 3 && need);
            const unsigned m = __ballot_sync(full, want);
            if (m) {
                unsigned base = 0;
                if (lane == __ffs(m) - 1) base = atomicAdd(&c->xq_head, (unsigned)__popc(m));
                base = __shfl_sync(full, base, __ffs(m) - 1);
                if (want) { ticket2 = base + __popc(m & lt); st2 = 1; }
            }
        }
        {   // a task that has waited too long behind a busy lane goes back to the queue
            const bool back = st2 == 2 && L.phase != branch::NEED && ++age2 > AHOLD;
            warp_push(a.xq, a.xmask, &c->xq_tail, back, task_entry(I2, r2), lane);
            if (back) st2 = 3;
        }
        if (prof) { __syncwarp(); t1 = clock64(); tr += t1 - t0; }
        // ---- pass 1: lanes starting a TRON solve
        load_bounds(col, xl, xu);
        branch::eval_pass(L, eval, 1, xl, xu, max_auglag, mu_max, T);
        if (prof) { __syncwarp(); t0 = clock64(); te += t0 - t1; }
        if (__any_sync(full, hflag != 0)) break;                       // (lanes may have read the flag at different times)
        if (__all_sync(full, L.phase == branch::NEED)) {               // nothing in flight in this warp
            if (++polls % APOLL_CHECK == 0 && lane == 0 && gtime() > c->deadline) async_give_up(c);
            __nanosleep(100);
            if (prof) ti += clock64() - t0;
            continue;
        }
        branch::compute(L, xl, xu);
        if (prof) { __syncwarp(); tp += clock64() - t0; nround++; }
    }
    if (d.count_work && done_tasks) atomicAdd(&c->n_x, done_tasks);
    if (prof && lane == 0) {
        atomicAdd(&c->diag[0], (unsigned long long)tr); atomicAdd(&c->diag[1], (unsigned long long)te);
        atomicAdd(&c->diag[2], (unsigned long long)tc); atomicAdd(&c->diag[3], (unsigned long long)tp);
        atomicAdd(&c->diag[4], (unsigned long long)ti); atomicAdd(&c->diag[5], (unsigned long long)nround);
    }
}

// Set-up of a run: control block, counters, queues, the X tasks of iteration 1 (and the B tasks of buses without
// branches). State of "iteration 0" has been copied into ring slot 0 by the host.
__global__ void k_async_begin(Dev d, AsyncDev a, double beta, double eps_pri, int limit, unsigned long long budget_ns) {
    AsyncCtrl *c = a.c;
    const int t = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
    if (t == 0) {
        c->xq_head = 0u; c->bq_head = 0u; c->xq_tail = (unsigned)d.nline;
        c->tested = 0; c->halt = 0; c->stop_iter = 0; c->limit = limit;
        c->eps_pri = eps_pri; c->beta = beta;
        for (int k = 0; k < AD; ++k) { c->iter_cnt[k] = 0u; c->complete[k] = 0; }
        for (int k = 0; k < 4; ++k) c->res[k] = 0.0;
        c->deadline = gtime() + budget_ns;
        c->n_x = 0ull; c->n_b = 0ull;
        for (int k = 0; k < 16; ++k) c->diag[k] = 0ull;
        c->bq_tail = (unsigned)a.n_iso;
    }
    for (int i = t; i < a.n_iso; i += n) a.bq[i] = (1ull << 32) | (unsigned)(a.iso_bus[i] + 1);
    for (int i = t; i < d.nline; i += n) { a.xq[i] = (1ull << 32) | (unsigned)(i + 1); a.br_cnt[i] = 0u; }
    for (int i = t; i < d.nbus; i += n) a.bus_cnt[i] = 0u;
}

}  // namespace ea
