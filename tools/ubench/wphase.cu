// Micro-benchmark of the LP fast path's "w = B^-1 A_q + ratio test" phase for one driver warp, two rows per
// lane, as in fj_lpf_iterate: (A) B^-1 through generic 64-bit pointers (what the kernel does when the compiler
// cannot prove the address space), (B) through 32-bit shared-window addresses (ld.shared).  Reports cycles per
// phase execution.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wphase wphase.cu
#include <cstdio>
#include <cuda_runtime.h>
#define R 85
#define RS 85
__device__ __forceinline__ double dot4(double v0, double v1, double v2, double v3, double a, double r)
{
    double acc = __dadd_rn(v0, __dmul_rn(v1, a));
    acc = __dadd_rn(acc, __dmul_rn(v2, r));
    return __dadd_rn(acc, __dmul_rn(v3, -r));
}
__device__ __forceinline__ double ldsd(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
template <int MODE> __global__ void k(double *out, long long *cyc, double *gBT, int reps, int qsel)
{
    extern __shared__ double sm[];
    double *BT = MODE == 2 ? gBT : sm;                      // (R + 1) * RS
    double *xb = sm + (R + 1) * RS, *w = xb + 128;
    int *bvar = (int *)(w + 128);
    uint2 *cidx = (uint2 *)(bvar + 128);
    double2 *coef = (double2 *)(cidx + 64);
    for (int i = threadIdx.x; i < (R + 1) * RS; i += blockDim.x) BT[i] = (i % 7 == 0) ? 0.37 * i : 0.0;
    for (int i = threadIdx.x; i < 128; i += blockDim.x) { xb[i] = (i % 3) ? 0.0 : 1.0; bvar[i] = 300 + i; w[i] = 0; }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) { cidx[i] = make_uint2(((i % 10) << 3) | ((10 + i) << 19), ((40 + i % 20) << 3) | ((60 + i % 20) << 19)); coef[i] = make_double2(-0.01 * (i + 1), 0.02 * (i + 1)); }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, first = lane, step = 64;
    double rk = 0; int ri = 0x7fffffff, rrow = 0;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        const int qin = (qsel + rep) & 63;
        const uint2 ix = cidx[qin];
        const double2 cf = coef[qin];
        if (MODE != 1) {
            const double *c0 = BT + (size_t)((ix.x & 0xfff8u) >> 3) * RS, *c1 = BT + (size_t)((ix.x >> 16 & 0xfff8u) >> 3) * RS;
            const double *c2 = BT + (size_t)((ix.y & 0xfff8u) >> 3) * RS, *c3 = BT + (size_t)((ix.y >> 16 & 0xfff8u) >> 3) * RS;
            for (int i = first; i < R; i += 2 * step) {
                const int i2 = i + step < R ? i + step : i;
                const double w1 = dot4(c0[i], c1[i], c2[i], c3[i], cf.x, cf.y), w2 = dot4(c0[i2], c1[i2], c2[i2], c3[i2], cf.x, cf.y);
                const double x1 = xb[i], x2 = xb[i2];
                const int b1 = bvar[i], b2 = bvar[i2];
                w[i] = w1; if (i2 != i) w[i2] = w2;
                const bool t1 = w1 > 1e-9, t2 = i2 != i && w2 > 1e-9;
                if (__any_sync(__activemask(), t1 || t2)) {
                    const double r1 = __ddiv_rn(x1 > 0 ? x1 : 1.0, t1 ? w1 : 1.0), r2 = __ddiv_rn(x2 > 0 ? x2 : 1.0, t2 ? w2 : 1.0);
                    if (t1 && (ri == 0x7fffffff || r1 < rk || (r1 == rk && b1 < ri))) { rk = r1; ri = b1; rrow = i; }
                    if (t2 && (ri == 0x7fffffff || r2 < rk || (r2 == rk && b2 < ri))) { rk = r2; ri = b2; rrow = i2; }
                }
            }
        } else {
            const unsigned bt = (unsigned)__cvta_generic_to_shared(BT), xa = (unsigned)__cvta_generic_to_shared(xb), wa = (unsigned)__cvta_generic_to_shared(w);
            const unsigned o0 = bt + ((ix.x & 0xfff8u) >> 3) * RS * 8, o1 = bt + ((ix.x >> 16 & 0xfff8u) >> 3) * RS * 8;
            const unsigned o2 = bt + ((ix.y & 0xfff8u) >> 3) * RS * 8, o3 = bt + ((ix.y >> 16 & 0xfff8u) >> 3) * RS * 8;
            for (int i = first; i < R; i += 2 * step) {
                const int i2 = i + step < R ? i + step : i;
                const double w1 = dot4(ldsd(o0 + i * 8), ldsd(o1 + i * 8), ldsd(o2 + i * 8), ldsd(o3 + i * 8), cf.x, cf.y);
                const double w2 = dot4(ldsd(o0 + i2 * 8), ldsd(o1 + i2 * 8), ldsd(o2 + i2 * 8), ldsd(o3 + i2 * 8), cf.x, cf.y);
                const double x1 = ldsd(xa + i * 8), x2 = ldsd(xa + i2 * 8);
                const int b1 = bvar[i], b2 = bvar[i2];
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(wa + i * 8), "d"(w1));
                if (i2 != i) asm volatile("st.shared.f64 [%0], %1;" ::"r"(wa + i2 * 8), "d"(w2));
                const bool t1 = w1 > 1e-9, t2 = i2 != i && w2 > 1e-9;
                if (__any_sync(__activemask(), t1 || t2)) {
                    const double r1 = __ddiv_rn(x1 > 0 ? x1 : 1.0, t1 ? w1 : 1.0), r2 = __ddiv_rn(x2 > 0 ? x2 : 1.0, t2 ? w2 : 1.0);
                    if (t1 && (ri == 0x7fffffff || r1 < rk || (r1 == rk && b1 < ri))) { rk = r1; ri = b1; rrow = i; }
                    if (t2 && (ri == 0x7fffffff || r2 < rk || (r2 == rk && b2 < ri))) { rk = r2; ri = b2; rrow = i2; }
                }
            }
        }
        __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) cyc[0] = (t1 - t0) / reps;
    out[lane] = rk + ri + rrow;
}
int main()
{
    double *out, *gBT; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64); cudaMalloc(&gBT, (R + 1) * RS * 8);
    const int smem = ((R + 1) * RS + 256) * 8 + 128 * 4 + 64 * 8 + 64 * 16;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 0; mode < 3; ++mode)
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) k<0><<<1, 256, smem>>>(out, cyc, gBT, 200, 3);
            if (mode == 1) k<1><<<1, 256, smem>>>(out, cyc, gBT, 200, 3);
            if (mode == 2) k<2><<<1, 256, smem>>>(out, cyc, gBT, 200, 3);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("mode %d (%s): %lld cycles per w+ratio phase (2 rows per lane, R = %d)\n", mode, mode == 0 ? "B^-1 in shared memory, C++ pointers" : mode == 1 ? "ld.shared / st.shared, 32-bit addresses" : "B^-1 in global memory (L1/L2)", h, R);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
