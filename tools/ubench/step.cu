// Cycles per Viterbi step for different ways of exchanging V between the S lanes of a chain (B200, sm_100a).
// One warp (or W warps) runs N dependent steps; em / lt come from shared memory like in the real kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int S = 5;
constexpr int G = 32 / S;
constexpr int N = 4096;

template <int MODE>
__device__ __forceinline__ unsigned step(double em, const double* lt, int src0, int j, double& V, uint4* slot0, double* vslot0, unsigned& anc)
{
    const double ninf = -HUGE_VAL;
    double c[S];
    int id[S];
    unsigned an[S];
    if (MODE == 0) {                       // 64-bit shuffles
#pragma unroll
        for (int k = 0; k < S; k++) {
            const double vk = __shfl_sync(0xffffffffu, V, src0 + k);
            c[k] = __dadd_rn(__dadd_rn(em, vk), lt[k]);
            id[k] = k;
        }
    } else if (MODE == 1) {                // STS.128 + LDS.128 (V + anc)
        slot0[j] = make_uint4((unsigned)__double2loint(V), (unsigned)__double2hiint(V), anc, 0u);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < S; k++) {
            const uint4 x = slot0[k];
            an[k] = x.z;
            c[k] = __dadd_rn(__dadd_rn(em, __hiloint2double((int)x.y, (int)x.x)), lt[k]);
            id[k] = k;
        }
    } else if (MODE == 2) {                // STS.64 + LDS.128/64 (V only)
        vslot0[j] = V;
        __syncwarp();
        const double2 a = reinterpret_cast<const double2*>(vslot0)[0];
        const double2 b = reinterpret_cast<const double2*>(vslot0)[1];
        const double e = vslot0[4];
        const double v[5] = {a.x, a.y, b.x, b.y, e};
#pragma unroll
        for (int k = 0; k < S; k++) {
            c[k] = __dadd_rn(__dadd_rn(em, v[k]), lt[k]);
            id[k] = k;
        }
    } else if (MODE == 7) {                // smem V only + warp barrier after the loads
        vslot0[j] = V;
        __syncwarp();
        int4 a, b; int2 e;
        const unsigned addr = (unsigned)__cvta_generic_to_shared(vslot0);
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(addr));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(addr));
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+32];" : "=r"(e.x), "=r"(e.y) : "r"(addr));
        __syncwarp();
        const double v[5] = {__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), __hiloint2double(b.y, b.x), __hiloint2double(b.w, b.z), __hiloint2double(e.y, e.x)};
#pragma unroll
        for (int k = 0; k < S; k++) {
            c[k] = __dadd_rn(__dadd_rn(em, v[k]), lt[k]);
            id[k] = k;
        }
    } else if (MODE == 8) {                // smem V only, loads tied together by a runtime-zero token
        vslot0[j] = V;
        __syncwarp();
        int4 a, b; int2 e;
        const unsigned addr = (unsigned)__cvta_generic_to_shared(vslot0);
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(addr));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(addr));
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+32];" : "=r"(e.x), "=r"(e.y) : "r"(addr));
        const int tok = (a.x | b.x | e.x) & (int)anc;     // anc is 0 at run time in this mode
        a.x ^= tok; a.z ^= tok; b.x ^= tok; b.z ^= tok; e.x ^= tok;
        const double v[5] = {__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), __hiloint2double(b.y, b.x), __hiloint2double(b.w, b.z), __hiloint2double(e.y, e.x)};
#pragma unroll
        for (int k = 0; k < S; k++) {
            c[k] = __dadd_rn(__dadd_rn(em, v[k]), lt[k]);
            id[k] = k;
        }
    } else if (MODE >= 10 && MODE <= 14) {   // decomposition of mode 8
        vslot0[j] = V;
        __syncwarp();
        int4 a, b; int2 e;
        const unsigned addr = (unsigned)__cvta_generic_to_shared(vslot0);
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(addr));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(addr));
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+32];" : "=r"(e.x), "=r"(e.y) : "r"(addr));
        const double v[5] = {__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), __hiloint2double(b.y, b.x), __hiloint2double(b.w, b.z), __hiloint2double(e.y, e.x)};
        if (MODE == 10) {                  // exchange only: V = xor of all
            V = __hiloint2double((a.y ^ a.w ^ b.y ^ b.w ^ e.y) & 0x3fffffff, a.x ^ a.z ^ b.x ^ b.z ^ e.x);
            return 0;
        }
#pragma unroll
        for (int k = 0; k < S; k++) {
            c[k] = __dadd_rn(__dadd_rn(em, v[k]), lt[k]);
            id[k] = k;
        }
        if (MODE == 11) {                  // no tournament: one compare only
            V = c[4] > c[0] ? c[1] : c[2];
            return c[3] > 0 ? 1 : 0;
        }
        // MODE 12: tournament on values only, no index tracking
        const double m01 = c[1] > c[0] ? c[1] : c[0], m23 = c[3] > c[2] ? c[3] : c[2];
        const double m = m23 > m01 ? m23 : m01;
        V = c[4] > m ? c[4] : m;
        if (MODE == 12) return 0;
        if (MODE == 13) {                  // argmax afterwards: first k whose candidate equals the maximum (FP64 compares)
            unsigned arg = 4u;
            arg = c[3] == V ? 3u : arg;
            arg = c[2] == V ? 2u : arg;
            arg = c[1] == V ? 1u : arg;
            arg = c[0] == V ? 0u : arg;
            return ((unsigned)__double2hiint(V) == 0xfff00000u && __double2loint(V) == 0) ? 7u : arg;
        }
        {                                  // MODE 14: same with integer compares on the bit patterns
            const long long vb = __double_as_longlong(V);
            unsigned arg = 4u;
            arg = __double_as_longlong(c[3]) == vb ? 3u : arg;
            arg = __double_as_longlong(c[2]) == vb ? 2u : arg;
            arg = __double_as_longlong(c[1]) == vb ? 1u : arg;
            arg = __double_as_longlong(c[0]) == vb ? 0u : arg;
            return vb == (long long)0xfff0000000000000ull ? 7u : arg;
        }
    } else if (MODE == 9) {                // shuffles tied together by a runtime-zero token
        int lo[S], hi[S];
        const int vlo = __double2loint(V), vhi = __double2hiint(V);
#pragma unroll
        for (int k = 0; k < S; k++) {
            asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(lo[k]) : "r"(vlo), "r"(src0 + k));
            asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(hi[k]) : "r"(vhi), "r"(src0 + k));
        }
        const int tok = (lo[0] | lo[1] | lo[2] | lo[3] | lo[4] | hi[0] | hi[1] | hi[2] | hi[3] | hi[4]) & (int)anc;
#pragma unroll
        for (int k = 0; k < S; k++) {
            c[k] = __dadd_rn(__dadd_rn(em, __hiloint2double(hi[k], lo[k] ^ tok)), lt[k]);
            id[k] = k;
        }
    } else if (MODE == 5) {                // compute only: every candidate uses the lane's own V
#pragma unroll
        for (int k = 0; k < S; k++) {
            c[k] = __dadd_rn(__dadd_rn(em, V), lt[k]);
            id[k] = k;
        }
    } else if (MODE == 6) {                // exchange only
        int lo[S], hi[S];
        const int vlo = __double2loint(V), vhi = __double2hiint(V);
#pragma unroll
        for (int k = 0; k < S; k++) {
            asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(lo[k]) : "r"(vlo), "r"(src0 + k));
            asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(hi[k]) : "r"(vhi), "r"(src0 + k));
        }
        int xl = lo[0], xh = hi[0];
#pragma unroll
        for (int k = 1; k < S; k++) { xl ^= lo[k]; xh ^= hi[k]; }
        V = __hiloint2double(xh & 0x3fffffff, xl);
        return 0;
    } else if (MODE == 3) {                // shuffles, all issued first through volatile asm
        int lo[S], hi[S];
        const int vlo = __double2loint(V), vhi = __double2hiint(V);
#pragma unroll
        for (int k = 0; k < S; k++) {
            asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(lo[k]) : "r"(vlo), "r"(src0 + k));
            asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(hi[k]) : "r"(vhi), "r"(src0 + k));
        }
#pragma unroll
        for (int k = 0; k < S; k++) {
            c[k] = __dadd_rn(__dadd_rn(em, __hiloint2double(hi[k], lo[k])), lt[k]);
            id[k] = k;
        }
    }
#pragma unroll
    for (int n = S; n > 1; n = (n + 1) / 2) {
#pragma unroll
        for (int p = 0; p + 1 < n; p += 2) {
            const bool right = c[p + 1] > c[p];
            c[p / 2] = right ? c[p + 1] : c[p];
            id[p / 2] = right ? id[p + 1] : id[p];
        }
        if (n & 1) { c[n / 2] = c[n - 1]; id[n / 2] = id[n - 1]; }
    }
    V = c[0];
    unsigned arg = c[0] > ninf ? (unsigned)id[0] : 7u;
    if (MODE == 1) {
        unsigned w = an[0];
#pragma unroll
        for (int k = 1; k < S; k++) w = arg == (unsigned)k ? an[k] : w;
        const unsigned mine = arg == 7u ? (an[0] >> 4) & 0xFu : w & 0xFu;
        anc = mine | (j == 0 ? (an[0] & 0xFu) << 4 : 0u);
    }
    return arg;
}

// software-pipelined: the argmax of step q-1 is derived while the exchange loads of step q are in flight
template <int ARGMODE>
__device__ __forceinline__ unsigned step_pipe(double em, const double* lt, int j, double& V, double* vslot0, double* cp /* [S] candidates of the previous step */, int src0 = 0)
{
    int4 a, b; int2 e;
    if (ARGMODE < 2) {
    vslot0[j] = V;
    __syncwarp();
    const unsigned addr = (unsigned)__cvta_generic_to_shared(vslot0);
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(addr));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(addr));
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+32];" : "=r"(e.x), "=r"(e.y) : "r"(addr));
    } else {
        const int vlo = __double2loint(V), vhi = __double2hiint(V);
        asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(a.x) : "r"(vlo), "r"(src0 + 0));
        asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(a.y) : "r"(vhi), "r"(src0 + 0));
        asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(a.z) : "r"(vlo), "r"(src0 + 1));
        asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(a.w) : "r"(vhi), "r"(src0 + 1));
        asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(b.x) : "r"(vlo), "r"(src0 + 2));
        asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(b.y) : "r"(vhi), "r"(src0 + 2));
        asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(b.z) : "r"(vlo), "r"(src0 + 3));
        asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(b.w) : "r"(vhi), "r"(src0 + 3));
        asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(e.x) : "r"(vlo), "r"(src0 + 4));
        asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(e.y) : "r"(vhi), "r"(src0 + 4));
    }
    // previous step's argmax: first candidate equal to the maximum (V still holds it)
    unsigned arg = 4u;
    if (ARGMODE != 1) {
        arg = cp[3] == V ? 3u : arg; arg = cp[2] == V ? 2u : arg; arg = cp[1] == V ? 1u : arg; arg = cp[0] == V ? 0u : arg;
    } else {
        const long long vb = __double_as_longlong(V);
        arg = __double_as_longlong(cp[3]) == vb ? 3u : arg; arg = __double_as_longlong(cp[2]) == vb ? 2u : arg;
        arg = __double_as_longlong(cp[1]) == vb ? 1u : arg; arg = __double_as_longlong(cp[0]) == vb ? 0u : arg;
    }
    arg = ((unsigned)__double2hiint(V) == 0xfff00000u && __double2loint(V) == 0) ? 7u : arg;
    const double v[5] = {__hiloint2double(a.y, a.x), __hiloint2double(a.w, a.z), __hiloint2double(b.y, b.x), __hiloint2double(b.w, b.z), __hiloint2double(e.y, e.x)};
    double c[S];
#pragma unroll
    for (int k = 0; k < S; k++) c[k] = __dadd_rn(__dadd_rn(em, v[k]), lt[k]);
    const double m01 = c[1] > c[0] ? c[1] : c[0], m23 = c[3] > c[2] ? c[3] : c[2];
    if (ARGMODE == 3) {                    // three-way final: the three compares are independent
        const bool p = m23 > m01, q4 = c[4] > m01, r4 = c[4] > m23;
        const double t = p ? m23 : m01;
        V = (q4 && r4) ? c[4] : t;
    } else {
        const double m = m23 > m01 ? m23 : m01;
        V = c[4] > m ? c[4] : m;
    }
#pragma unroll
    for (int k = 0; k < S; k++) cp[k] = c[k];
    return arg;
}

// thread per chain: every lane keeps all S values of V
__device__ __forceinline__ unsigned step_tpc(const double* em, const double* lt, double* V)
{
    double nv[S];
    unsigned args = 0;
#pragma unroll
    for (int j = 0; j < S; j++) {
        double c[S];
        int id[S];
#pragma unroll
        for (int k = 0; k < S; k++) { c[k] = __dadd_rn(__dadd_rn(em[j], V[k]), lt[j * S + k]); id[k] = k; }
#pragma unroll
        for (int n = S; n > 1; n = (n + 1) / 2) {
#pragma unroll
            for (int p = 0; p + 1 < n; p += 2) {
                const bool right = c[p + 1] > c[p];
                c[p / 2] = right ? c[p + 1] : c[p];
                id[p / 2] = right ? id[p + 1] : id[p];
            }
            if (n & 1) { c[n / 2] = c[n - 1]; id[n / 2] = id[n - 1]; }
        }
        nv[j] = c[0];
        args |= (unsigned)id[0] << (4 * j);
    }
#pragma unroll
    for (int j = 0; j < S; j++) V[j] = nv[j];
    return args;
}

template <int MODE>
__global__ void k(double* out, long long* cyc, const double* emg, const double* ltg)
{
    __shared__ double lt[16 * 26];
    __shared__ double em[16 * 8];
    __shared__ __align__(16) uint4 xch[8][2][G * 5];
    __shared__ __align__(16) double vx[8][2][G * 6];
    for (int i = threadIdx.x; i < 16 * 26; i += blockDim.x) lt[i] = ltg[i];
    for (int i = threadIdx.x; i < 16 * 8; i += blockDim.x) em[i] = emg[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int g = lane / S;
    const int j = lane - g * S;
    if (g >= G) g = G - 1;
    const int src0 = g * S;
    double V = j == 0 ? 0.0 : -1.0;
    unsigned anc = (MODE == 8 || MODE == 9) ? (unsigned)(ltg[0] > 1e30) : j, acc = 0;
    long long t0 = clock64();
    if (MODE >= 15 && MODE <= 18) {
        double cp[S];
#pragma unroll
        for (int s = 0; s < S; s++) cp[s] = V;
        for (int it = 0; it < N / 16; it++) {
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const unsigned a = step_pipe<MODE - 15>(em[q * 8 + j], lt + q * 26 + j * S, j, V, &vx[warp][q & 1][g * 6], cp, src0);
                acc = (acc << 1) ^ a;
            }
        }
    } else if (MODE != 4) {
        for (int it = 0; it < N / 16; it++) {
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const unsigned a = step<MODE>(em[q * 8 + j], lt + q * 26 + j * S, src0, j, V, &xch[warp][q & 1][g * 5], &vx[warp][q & 1][g * 6], anc);
                acc = (acc << 1) ^ a;
            }
        }
    } else {
        double Vv[S];
#pragma unroll
        for (int s = 0; s < S; s++) Vv[s] = s == 0 ? 0.0 : -1.0 - lane * 1e-3;
        for (int it = 0; it < N / 16; it++) {
#pragma unroll
            for (int q = 0; q < 16; q++) acc = (acc << 1) ^ step_tpc(em + q * 8, lt + q * 26, Vv);
        }
        V = Vv[0] + Vv[1] + Vv[2] + Vv[3] + Vv[4];
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = V + acc + anc;
}

template <int MODE>
void run(const char* name, int warps, double* out, long long* cyc, double* em, double* lt)
{
    for (int rep = 0; rep < 2; rep++) k<MODE><<<1, 32 * warps>>>(out, cyc, em, lt);
    long long h[8];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-44s warps=%d  %.1f clk/step  (%s)\n", name, warps, (double)h[0] / N, cudaGetErrorString(e));
}

int main()
{
    double *out, *em, *lt;
    long long* cyc;
    cudaMalloc(&out, 8 * 256 * 8);
    cudaMalloc(&cyc, 64 * 8);
    cudaMalloc(&em, 16 * 8 * 8);
    cudaMalloc(&lt, 16 * 26 * 8);
    double hem[128], hlt[16 * 26];
    for (int i = 0; i < 128; i++) hem[i] = -1.0 - (i % 7) * 0.37;
    for (int i = 0; i < 16 * 26; i++) hlt[i] = -0.5 - (i % 11) * 0.9;
    cudaMemcpy(em, hem, sizeof hem, cudaMemcpyHostToDevice);
    cudaMemcpy(lt, hlt, sizeof hlt, cudaMemcpyHostToDevice);
    for (int w : {1, 2, 4, 8}) {
        run<0>("shfl64 exchange", w, out, cyc, em, lt);
        run<3>("shfl32 x10 (volatile asm)", w, out, cyc, em, lt);
        run<1>("smem uint4 exchange (V + anc)", w, out, cyc, em, lt);
        run<2>("smem double exchange (V only)", w, out, cyc, em, lt);
        run<7>("smem V only + syncwarp after loads", w, out, cyc, em, lt);
        run<8>("smem V only + zero token", w, out, cyc, em, lt);
        run<9>("shfl x10 + zero token", w, out, cyc, em, lt);
        run<10>("smem exchange only", w, out, cyc, em, lt);
        run<11>("smem exchange + DADDs + 1 compare", w, out, cyc, em, lt);
        run<12>("smem exchange + DADDs + value tournament", w, out, cyc, em, lt);
        run<13>("smem + value tournament + argmax by DSETP.EQ", w, out, cyc, em, lt);
        run<14>("smem + value tournament + argmax by int compare", w, out, cyc, em, lt);
        run<15>("pipelined argmax (DSETP.EQ)", w, out, cyc, em, lt);
        run<16>("pipelined argmax (int compare)", w, out, cyc, em, lt);
        run<17>("shfl + value tournament + pipelined argmax", w, out, cyc, em, lt);
        run<18>("shfl + 3-way final + pipelined argmax", w, out, cyc, em, lt);
        run<5>("compute only (own V)", w, out, cyc, em, lt);
        run<6>("exchange only (10 shfl + xor)", w, out, cyc, em, lt);
    }
    return 0;
}
