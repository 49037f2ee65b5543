// Dependent-chain latency microbenchmarks for the ops on the Viterbi critical path (B200, sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void k(double* out, long long* cyc, double seed, int srcl)
{
    __shared__ double sm[64];
    double v = seed + threadIdx.x * 1e-9, w = seed * 0.5;
    long long t0, t1;
    // DADD chain
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) v = __dadd_rn(v, w);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DSETP + select chain
    double a = v, b = seed;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) { bool r = b > a; double na = r ? b : a; b = r ? a : b; a = na; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    v += a + b;
    // 64-bit shuffle chain
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) v = __shfl_sync(0xffffffffu, v, srcl);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 32-bit shuffle chain
    int iv = (int)v;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) iv = __shfl_sync(0xffffffffu, iv, (iv + srcl) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // STS + syncwarp + LDS roundtrip
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { sm[threadIdx.x] = v; __syncwarp(); v = sm[(threadIdx.x + srcl) & 31]; __syncwarp(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // integer select/shift chain (traceback-like): st = (w[st] >> c) & 7
    unsigned w0 = 0x12345670u + iv, w1 = 0x76543210u, w2 = 0x01234567u, w3 = 0x11223344u, w4 = 0x44332211u;
    unsigned st = iv & 3;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) {
        unsigned word = w0;
        word = st == 1 ? w1 : word; word = st == 2 ? w2 : word; word = st == 3 ? w3 : word; word = st == 4 ? w4 : word;
        st = (word >> ((i & 7) * 4)) & 3;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // DMNMX (fmax) chain
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) v = fmax(v, w + i);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // DADD -> DADD -> DSETP -> FSEL (one candidate update)
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) { double c = __dadd_rn(__dadd_rn(w, v), seed); v = c > v ? c : v; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = t1 - t0;
    out[threadIdx.x] = v + st + iv;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8);
    for (int rep = 0; rep < 2; rep++) k<<<1, 32>>>(out, cyc, 1.000001, 3);
    long long h[8];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const char* names[8] = {"DADD", "DSETP+2xSEL(swap)", "SHFL64", "SHFL32(dep idx)", "STS+LDS roundtrip", "sel-chain traceback step", "fmax(DMNMX?)+DADD", "DADD,DADD,DSETP,FSEL"};
    for (int i = 0; i < 8; i++) printf("%-28s %.1f clk/iter\n", names[i], (double)h[i] / N);
    return 0;
}
