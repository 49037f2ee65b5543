// FP64 pipe throughput per SM on B200: independent DADD / DFMA / DSETP+FSEL chains, W warps in one CTA.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 2048;
template <int OP>
__global__ void k(double* out, long long* cyc, double seed)
{
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x * 1e-6;
    const double w = seed * 0.37;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < N; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = __dadd_rn(a[i], w);
            if (OP == 1) a[i] = __fma_rn(a[i], w, seed);
            if (OP == 2) a[i] = a[i] > a[(i + 1) & 7] ? a[i] : w + i;     // DSETP + FSEL x2 (+ one DADD hoisted)
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[threadIdx.x] = s;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    const char* names[3] = {"DADD", "DFMA", "DSETP+FSEL"};
    for (int op = 0; op < 3; op++)
        for (int w : {1, 2, 4, 8, 16, 32}) {
            for (int rep = 0; rep < 2; rep++) {
                if (op == 0) k<0><<<1, 32 * w>>>(out, cyc, 1.000001);
                if (op == 1) k<1><<<1, 32 * w>>>(out, cyc, 1.000001);
                if (op == 2) k<2><<<1, 32 * w>>>(out, cyc, 1.000001);
            }
            long long h;
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-12s warps=%2d  %.2f clk per warp-instr per warp, %.3f warp-instr/clk/SM\n", names[op], w, (double)h / (N * 8), (double)w * N * 8 / h);
        }
    return 0;
}
