// Micro-benchmarks of the instructions the FJSP kernels' dependent chains are made of (B200, sm_100a):
// latency of dependent DADD / DMUL / DFMA / IMAD / LDS / REDUX / __ddiv_rn / I2F.F64 chains, and the DP
// throughput with many warps.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int OP> __global__ void lat(double *out, long long *cyc, double a, double b, int nwarps_active)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 1) % 1024);
    __syncthreads();
    if ((threadIdx.x >> 5) >= nwarps_active) return;
    double x = a + threadIdx.x;
    unsigned u = threadIdx.x * 2654435761u;
    int idx = threadIdx.x & 1023;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = __dadd_rn(x, b);
        if (OP == 1) x = __dmul_rn(x, b);
        if (OP == 2) x = __fma_rn(x, b, a);
        if (OP == 3) u = u * 2654435761u + 12345u;
        if (OP == 4) { idx = (int)sm[idx] ; }
        if (OP == 5) u = __reduce_min_sync(0xffffffffu, u) + threadIdx.x;
        if (OP == 6) x = __ddiv_rn(x, b);
        if (OP == 7) { x = (double)(int)u; u = (unsigned)(long long)x + 3u; }
        if (OP == 8) u = __shfl_xor_sync(0xffffffffu, u, 1) + 1u;
        if (OP == 9) u = __ballot_sync(0xffffffffu, u & 1) + threadIdx.x;
        if (OP == 10) x = sqrt(x) + b;
        if (OP == 11) u = __popc(u) + threadIdx.x * 3u;
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + u + idx;
}
template <int OP> void run(const char *name, int warps)
{
    double *out; long long *cyc, h[32];
    cudaMalloc(&out, 8 * 1024 * 8); cudaMalloc(&cyc, 8 * 32 * 8);
    lat<OP><<<1, 1024>>>(out, cyc, 1.000001, 1.0000001, warps);
    cudaDeviceSynchronize();
    lat<OP><<<1, 1024>>>(out, cyc, 1.000001, 1.0000001, warps);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-10s warps %2d: %.1f cycles per op (dependent chain of %d)\n", name, warps, (double)mx / N, N);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {1, 4, 8, 16, 32}) {
        run<0>("DADD", w); run<1>("DMUL", w); run<2>("DFMA", w);
    }
    for (int w : {1, 32}) {
        run<3>("IMAD", w); run<4>("LDS", w); run<5>("REDUX", w); run<6>("DDIV", w); run<7>("I2F/F2I", w); run<8>("SHFL", w); run<9>("BALLOT", w);
        run<10>("DSQRT", w); run<11>("POPC", w);
    }
    return 0;
}
