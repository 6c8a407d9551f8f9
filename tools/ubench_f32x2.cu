// ubench_f32x2.cu -- measures FFMA vs FFMA2 (packed fp32x2) issue throughput per SM on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_f32x2 tools/ubench_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int MODE>
__global__ void k(float *out, int iters, float a, float b) {
    float x[16]; u64 y[8];
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < 8; i++) y[i] = ((u64)__float_as_uint(x[2*i]) << 32) | __float_as_uint(x[2*i+1]);
    u64 ab = ((u64)__float_as_uint(a) << 32) | __float_as_uint(a), bb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) y[i] = fma2(y[i], ab, bb);
        }
    }
    float s = 0;
    for (int i = 0; i < 16; i++) s += x[i];
    for (int i = 0; i < 8; i++) s += __uint_as_float((unsigned)(y[i] >> 32)) + __uint_as_float((unsigned)y[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int mode = 0; mode < 2; mode++) for (int warps = 4; warps <= 32; warps *= 2) {
        float ms = 0;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, 1.0001f, 0.5f); else k<1><<<148, warps * 32>>>(d, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        double inst = (double)iters * (mode == 0 ? 16 : 8) * warps;              // warp-instructions per SM
        double lanes_fma = (double)iters * 16 * warps * 32;                      // scalar FMAs per SM
        printf("mode %s warps/SM %2d: %.3f ms  %.2f warp-inst/ns/SM  %.1f fp32-FMA/clk/SM (at 1.9 GHz)\n",
               mode ? "FFMA2" : "FFMA ", warps, ms, inst / (ms * 1e6), lanes_fma / (ms * 1e-3 * 1.9e9));
    }
    return 0;
}
