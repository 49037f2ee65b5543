// SHFL / LDS dependent-chain latency and issue cost on B200.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 4096;
__global__ void k(int* out, long long* cyc, int rot)
{
    __shared__ int sm[64];
    const int lane = threadIdx.x & 31;
    const int src = (lane + rot) & 31;
    int x = lane * 3 + rot;
    long long t0, t1;
    t0 = clock64();
#pragma unroll 32
    for (int i = 0; i < N; i++) { asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(x) : "r"(x), "r"(src)); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 10 independent shuffles then a dependent combine (like one Viterbi exchange)
    int y[10];
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int q = 0; q < 10; q++) asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(y[q]) : "r"(x + q), "r"(src));
        x = y[0];
#pragma unroll
        for (int q = 1; q < 10; q++) x ^= y[q];
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // 2 independent shuffles then combine
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int q = 0; q < 2; q++) asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(y[q]) : "r"(x + q), "r"(src));
        x = y[0] ^ y[1];
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // STS.128 -> syncwarp -> 5x LDS.128 -> combine
    __shared__ __align__(16) int4 s4[64];
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        s4[lane] = make_int4(x, x + 1, x + 2, x + 3);
        __syncwarp();
        int acc = 0;
#pragma unroll
        for (int q = 0; q < 5; q++) { const int4 v = s4[(src + q) & 31]; acc ^= v.x ^ v.w; }
        x = acc;
        __syncwarp();
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // STS.32 -> LDS.32 only
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        sm[lane] = x;
        __syncwarp();
        x = sm[src] + 1;
        __syncwarp();
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    out[threadIdx.x] = x;
}
int main()
{
    int* out; long long* cyc;
    cudaMalloc(&out, 32 * 4); cudaMalloc(&cyc, 8 * 8);
    for (int rep = 0; rep < 2; rep++) k<<<1, 32>>>(out, cyc, 3);
    long long h[8];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const char* names[5] = {"SHFL.IDX dependent chain", "10 SHFL + xor combine", "2 SHFL + xor", "STS.128 + 5 LDS.128 + combine", "STS.32 + LDS.32 + add"};
    for (int i = 0; i < 5; i++) printf("%-32s %.1f clk/iter\n", names[i], (double)h[i] / N);
    return 0;
}
