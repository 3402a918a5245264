// ubench.cu — latency / issue-rate probes for the instructions the branch solver is made of (B200, sm_100a).
// One warp, one CTA; every probe is a dependent chain (latency) or a set of independent chains (issue rate) timed
// with clock64. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench ubench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define N_IT 2048

__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }

template <int ILP> __global__ void k_dfma(double *out, long long *cyc, int active) {
    if ((int)threadIdx.x >= active) return;
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    const double b = 1.0000001, c = 1e-9;
    const long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    }
    const long long t1 = clk();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_dadd(double *out, long long *cyc) {
    double a = 1.0 + threadIdx.x;
    const double c = 1e-9;
    const long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a = a + c;
    }
    const long long t1 = clk();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_shfl64(double *out, long long *cyc) {
    double a = 1.0 + threadIdx.x;
    const long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a = __shfl_xor_sync(0xffffffffu, a, 1);
    }
    const long long t1 = clk();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl32(double *out, long long *cyc) {
    int a = threadIdx.x;
    const long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a = __shfl_xor_sync(0xffffffffu, a, 1) + 1;
    }
    const long long t1 = clk();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// shuffle + dependent DFMA (the pattern of a cross-lane reduction step)
__global__ void k_shfl_fma(double *out, long long *cyc) {
    double a = 1.0 + threadIdx.x;
    const long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a = fma(__shfl_xor_sync(0xffffffffu, a, 1), 0.999, a);
    }
    const long long t1 = clk();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_lds(double *out, long long *cyc) {
    __shared__ int next[256];
    __shared__ double val[256];
    for (int i = threadIdx.x; i < 256; i += 32) { next[i] = (i * 37 + 11) & 255; val[i] = i; }
    __syncwarp();
    int p = threadIdx.x;
    const long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) p = next[p];
    }
    const long long t1 = clk();
    // LDS.64 feeding a DFMA chain
    double a = 1.0;
    int q = threadIdx.x;
    const long long t2 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { a = fma(a, 0.5, val[q]); q = (q + 1) & 255; }
    }
    const long long t3 = clk();
    out[threadIdx.x] = p + a;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; }
}

__global__ void k_rcp(double *out, long long *cyc) {
    double a = 1.5 + 1e-3 * threadIdx.x;
    const long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r + 0.25; }
    }
    const long long t1 = clk();
    double b = 1.5 + 1e-3 * threadIdx.x;
    const long long t2 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b)); b = r + 0.25; }
    }
    const long long t3 = clk();
    out[threadIdx.x] = a + b;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; }
}

__global__ void k_ieee(double *out, long long *cyc) {
    double a = 1.5 + 1e-3 * threadIdx.x;
    long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) a = 1.0 / a + 0.25;
    }
    long long t1 = clk();
    cyc[0] = t1 - t0;
    double b = 1.5 + 1e-3 * threadIdx.x;
    t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) b = sqrt(b) + 0.25;
    }
    t1 = clk();
    cyc[1] = t1 - t0;
    double c = 0.3 + 1e-3 * threadIdx.x;
    t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { double s, co; sincos(c, &s, &co); c = s * 0.5 + co * 0.25; }
    }
    t1 = clk();
    cyc[2] = t1 - t0;
    out[threadIdx.x] = a + b + c;
}

// local memory (register spill) round trip: store + dependent load through L1
__global__ void k_local(double *out, long long *cyc, int n) {
    double buf[64];
    for (int i = 0; i < 64; ++i) buf[i] = i + threadIdx.x;
    int p = threadIdx.x & 63;
    double a = 0.0;
    const long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { a = fma(a, 0.5, buf[p]); p = (p + n) & 63; }
    }
    const long long t1 = clk();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// branch + loop overhead: a data-dependent while loop with a tiny body
__global__ void k_branch(double *out, long long *cyc, int n) {
    double a = 1.0 + threadIdx.x;
    int mode = 0;
    const long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < N_IT * 8; ++it) {
        if (mode == 0) { a = a + 1e-9; if (a > 1e300) mode = n; }
        else if (mode == 1) { a = a * 0.5; mode = 2; }
        else { a = a - 1.0; mode = 0; }
    }
    const long long t1 = clk();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double *out; long long *cyc, h[4];
    cudaMalloc(&out, 1024 * sizeof(double)); cudaMalloc(&cyc, 4 * sizeof(long long));
    const double per = 1.0 / (N_IT * 8.0);
    auto get = [&]() { cudaDeviceSynchronize(); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); };
    for (int rep = 0; rep < 2; ++rep) {
        k_dfma<1><<<1, 32>>>(out, cyc, 32); get(); if (rep) printf("DFMA dependent (32 lanes)      %6.2f cyc\n", h[0] * per);
        k_dfma<1><<<1, 32>>>(out, cyc, 1);  get(); if (rep) printf("DFMA dependent (1 lane)        %6.2f cyc\n", h[0] * per);
        k_dfma<2><<<1, 32>>>(out, cyc, 1);  get(); if (rep) printf("DFMA ILP2 per instr (1 lane)   %6.2f cyc\n", h[0] * per / 2);
        k_dfma<4><<<1, 32>>>(out, cyc, 1);  get(); if (rep) printf("DFMA ILP4 per instr (1 lane)   %6.2f cyc\n", h[0] * per / 4);
        k_dfma<8><<<1, 32>>>(out, cyc, 1);  get(); if (rep) printf("DFMA ILP8 per instr (1 lane)   %6.2f cyc\n", h[0] * per / 8);
        k_dfma<8><<<1, 32>>>(out, cyc, 32); get(); if (rep) printf("DFMA ILP8 per instr (32 lanes) %6.2f cyc\n", h[0] * per / 8);
        k_dfma<8><<<1, 32>>>(out, cyc, 16); get(); if (rep) printf("DFMA ILP8 per instr (16 lanes) %6.2f cyc\n", h[0] * per / 8);
        k_dadd<<<1, 32>>>(out, cyc);        get(); if (rep) printf("DADD dependent                 %6.2f cyc\n", h[0] * per);
        k_shfl32<<<1, 32>>>(out, cyc);      get(); if (rep) printf("SHFL.32 + IADD dependent       %6.2f cyc\n", h[0] * per);
        k_shfl64<<<1, 32>>>(out, cyc);      get(); if (rep) printf("SHFL.64 dependent              %6.2f cyc\n", h[0] * per);
        k_shfl_fma<<<1, 32>>>(out, cyc);    get(); if (rep) printf("SHFL.64 + DFMA dependent       %6.2f cyc\n", h[0] * per);
        k_lds<<<1, 32>>>(out, cyc);         get(); if (rep) printf("LDS.32 pointer chase           %6.2f cyc   LDS.64 -> DFMA chain (addr independent) %6.2f cyc\n", h[0] * per, h[1] * per);
        k_rcp<<<1, 32>>>(out, cyc);         get(); if (rep) printf("rcp.approx.f64 + DADD          %6.2f cyc   rsqrt.approx.f64 + DADD %6.2f cyc\n", h[0] * per, h[1] * per);
        k_ieee<<<1, 1>>>(out, cyc);         get(); if (rep) printf("IEEE 1/x + DADD %6.2f cyc   sqrt + DADD %6.2f cyc   sincos + 2 ops %6.2f cyc\n", h[0] * per, h[1] * per, h[2] * per);
        k_local<<<1, 32>>>(out, cyc, 1);    get(); if (rep) printf("local LD -> DFMA chain         %6.2f cyc\n", h[0] * per);
        k_branch<<<1, 32>>>(out, cyc, 0);   get(); if (rep) printf("branchy loop iteration         %6.2f cyc\n", h[0] * per);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
