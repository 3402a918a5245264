// ubench_coop.cu — the north-star's "one warp per branch with shuffle reductions", measured where it has the most to
// offer: the fused f / gradient / Hessian evaluation of one branch (the widest part of a TRON iteration: four flows
// that can be computed side by side). Three variants evaluate the SAME chain of 2048 dependent evaluations (the point
// of the next evaluation depends on the result of the previous one, as in a TRON solve) on one warp:
//   lane      branch::eval_fgh on one lane (what k_xupdate does);
//   coop4     4 lanes, one flow each: P, Q, F, G of its flow, then the sums over the flows by two butterfly stages of
//             64-bit shuffles (14 partial sums), the assembly replicated in the 4 lanes;
//   coop4x8   the same with 8 branches per warp (lanes 4b .. 4b+3): the packing the north-star text suggests.
// Prints SM cycles per evaluation. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I exaadmm.jl_b200/csrc.
#include <cstdio>
#include <cuda_runtime.h>
#include "branch.cuh"

#define N_EV 2048

__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }

__device__ __forceinline__ void make_data(branch::Data &D) {
    const double Y[8] = { 3.2, -29.1, -3.1, 28.3, 3.0, -27.5, -3.1, 28.3 };
    for (int k = 0; k < 8; ++k) { D.lam[k] = 0.3 * (k + 1); D.rho[k] = (k < 4) ? 3e4 : 3e5; D.xt[k] = (k < 4) ? 0.1 * k : 1.0; D.Y[k] = Y[k]; }
    D.xt[6] = 0.01; D.xt[7] = -0.01;
}

__global__ void k_lane(double *out, long long *cyc) {
    branch::Data D; make_data(D);
    double x[6] = { 1.01, 0.99, 0.01, -0.01, -0.05, -0.05 }, ls[2] = { 1.0, 2.0 }, g[6], F[4], f;
    branch::Sym6 A;
    const long long t0 = clk();
#pragma unroll 1
    for (int i = 0; i < N_EV; ++i) {
        branch::eval_fgh(branch::StructView{ &D }, ls, 100.0, 1e-5, x, f, g, A, F);
        x[2] += 1e-9 * (g[2] + A.a22);           // the next point depends on this evaluation
        x[0] -= 1e-12 * f;
    }
    const long long t1 = clk();
    if (threadIdx.x == 0) { out[0] = x[0] + x[2]; cyc[0] = t1 - t0; }
}

// one flow per lane (lane & 3 = flow), sums by butterflies within each group of 4 lanes
__device__ __forceinline__ double bfly4(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ void eval_coop4(const branch::Data &D, const double (&ls)[2], double mu, double scale,
                                           const double (&x)[6], double &f, double (&g)[6], branch::Sym6 &A) {
    const int k = threadIdx.x & 3, j = k >> 1;
    const double vi = x[0], vj = x[1];
    double s, c;
    sincos(x[2] - x[3], &s, &c);
    const double vv = vi * vj, vi2 = vi * vi, vj2 = vj * vj;
    const double a = (k == 0) ? D.Y[0] : ((k == 1) ? -D.Y[1] : 0.0);
    const double b = (k == 2) ? D.Y[4] : ((k == 3) ? -D.Y[5] : 0.0);
    const double ga = (k == 0) ? D.Y[2] : ((k == 1) ? -D.Y[3] : ((k == 2) ? D.Y[6] : -D.Y[7]));
    const double de = (k == 0) ? D.Y[3] : ((k == 1) ? D.Y[2] : ((k == 2) ? -D.Y[7] : -D.Y[6]));
    const double P = ga * c + de * s, Q = de * c - ga * s;
    const double F = a * vi2 + b * vj2 + vv * P;
    const double F2 = F * F;
    const double cj = F2 + __shfl_xor_sync(0xffffffffu, F2, 1) + x[4 + j];      // constraint of my side
    const double m = ls[j] + mu * cj;
    const double lam = D.lam[k], rho = D.rho[k], dev = F - D.xt[k];
    double fv = lam * F + 0.5 * (rho * (dev * dev)) + ((k & 1) ? 0.0 : (ls[j] * cj + 0.5 * (mu * (cj * cj))));
    const double G0 = 2.0 * a * vi + vj * P, G1 = 2.0 * b * vj + vi * P, G2 = vv * Q;
    const double r = lam + rho * dev, w = r + 2.0 * m * F, kap = rho + 2.0 * m;
    const double tF = 2.0 * F;
    double d0 = tF * G0, d1 = tF * G1, d2 = tF * G2;
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d2 += __shfl_xor_sync(0xffffffffu, d2, 1);
    const double k0 = kap * G0, k1 = kap * G1, k2 = kap * G2;
    // my side's mu d d' term is added by the even lane of the side
    const double e = (k & 1) ? 0.0 : mu;
    double As = bfly4(w * a), Bs = bfly4(w * b), Ps = bfly4(w * P), Qs = bfly4(w * Q);
    double H00 = bfly4(k0 * G0 + e * d0 * d0), H01 = bfly4(k0 * G1 + e * d0 * d1), H02 = bfly4(k0 * G2 + e * d0 * d2);
    double H11 = bfly4(k1 * G1 + e * d1 * d1), H12 = bfly4(k1 * G2 + e * d1 * d2), H22 = bfly4(k2 * G2 + e * d2 * d2);
    fv = bfly4(fv);
    const double m0 = __shfl_sync(0xffffffffu, m, (threadIdx.x & ~3)), m1 = __shfl_sync(0xffffffffu, m, (threadIdx.x & ~3) + 2);
    const double da0 = __shfl_sync(0xffffffffu, d0, (threadIdx.x & ~3)), da1 = __shfl_sync(0xffffffffu, d1, (threadIdx.x & ~3));
    const double da2 = __shfl_sync(0xffffffffu, d2, (threadIdx.x & ~3));
    const double db0 = __shfl_sync(0xffffffffu, d0, (threadIdx.x & ~3) + 2), db1 = __shfl_sync(0xffffffffu, d1, (threadIdx.x & ~3) + 2);
    const double db2 = __shfl_sync(0xffffffffu, d2, (threadIdx.x & ~3) + 2);
    H00 += 2.0 * As; H11 += 2.0 * Bs; H01 += Ps; H02 += vj * Qs; H12 += vi * Qs; H22 -= vv * Ps;
    const double rho4 = D.rho[4], rho5 = D.rho[5], rho6 = D.rho[6], rho7 = D.rho[7];
    const double dwi = vi2 - D.xt[4], dwj = vj2 - D.xt[5], dti = x[2] - D.xt[6], dtj = x[3] - D.xt[7];
    fv += D.lam[4] * vi2 + 0.5 * (rho4 * (dwi * dwi)) + D.lam[5] * vj2 + 0.5 * (rho5 * (dwj * dwj))
        + D.lam[6] * x[2] + 0.5 * (rho6 * (dti * dti)) + D.lam[7] * x[3] + 0.5 * (rho7 * (dtj * dtj));
    f = scale * fv;
    const double ri = D.lam[4] + rho4 * dwi, rj = D.lam[5] + rho5 * dwj;
    const double gy2 = vv * Qs;
    H00 += 2.0 * ri + 4.0 * rho4 * vi2; H11 += 2.0 * rj + 4.0 * rho5 * vj2;
    g[0] = scale * (2.0 * As * vi + vj * Ps + 2.0 * vi * ri); g[1] = scale * (2.0 * Bs * vj + vi * Ps + 2.0 * vj * rj);
    g[2] = scale * (gy2 + D.lam[6] + rho6 * dti); g[3] = scale * (-gy2 + D.lam[7] + rho7 * dtj);
    g[4] = scale * m0; g[5] = scale * m1;
    const double smu = scale * mu;
    A.a00 = scale * H00; A.a10 = scale * H01; A.a11 = scale * H11;
    A.a20 = scale * H02; A.a21 = scale * H12; A.a22 = scale * (H22 + rho6);
    A.a32 = -(scale * H22); A.a33 = scale * (H22 + rho7);
    A.a40 = smu * da0; A.a41 = smu * da1; A.a42 = smu * da2; A.a44 = smu;
    A.a50 = smu * db0; A.a51 = smu * db1; A.a52 = smu * db2;
}

__global__ void k_coop(double *out, long long *cyc, int lanes_on) {
    branch::Data D; make_data(D);
    double x[6] = { 1.01, 0.99, 0.01, -0.01, -0.05, -0.05 }, ls[2] = { 1.0, 2.0 }, g[6], f = 0.0;
    branch::Sym6 A;
    x[0] += 1e-3 * (threadIdx.x >> 2);              // different branches in different lane groups
    const long long t0 = clk();
    if ((int)threadIdx.x < lanes_on) {               // whole groups of 4 lanes
#pragma unroll 1
        for (int i = 0; i < N_EV; ++i) {
            eval_coop4(D, ls, 100.0, 1e-5, x, f, g, A);
            x[2] += 1e-9 * (g[2] + A.a22);
            x[0] -= 1e-12 * f;
        }
    }
    const long long t1 = clk();
    if (threadIdx.x == 0) { out[1] = x[0] + x[2]; cyc[0] = t1 - t0; }
}

int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 8);
    for (int rep = 0; rep < 2; ++rep) {
        k_lane<<<1, 32>>>(out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("eval on one lane (k_xupdate)             %8.0f cycles per evaluation\n", (double)h / N_EV);
        k_coop<<<1, 32>>>(out, cyc, 4); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("eval on 4 lanes, shuffle reductions      %8.0f cycles per evaluation\n", (double)h / N_EV);
        k_coop<<<1, 32>>>(out, cyc, 32); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("8 branches x 4 lanes in one warp         %8.0f cycles per evaluation round (8 evaluations)\n", (double)h / N_EV);
    }
    double v[2]; cudaMemcpy(v, out, 16, cudaMemcpyDeviceToHost);
    printf("checksums (same chain, must agree to rounding): %.15g %.15g; status %s\n", v[0], v[1], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
