// stockham.cuh -- block-cooperative radix-4 (+ one radix-2) Stockham autosort FFT in shared memory, forward
// direction (e^{-2 pi i k n / nc}), any power of two nc.  `a` holds the input, `b` is scratch of the same size; the
// function returns whichever of the two holds the result.  All threads of the block must call it.
#pragma once
#include "common.cuh"

__device__ __forceinline__ float2 af_twiddle(int k, int m) {   // exp(-2 pi i k / m)
    float s, c;
    sincospif(-2.0f * (float)k / (float)m, &s, &c);
    return make_float2(c, s);
}

__device__ __forceinline__ float2 *af_stockham(float2 *a, float2 *b, int nc, int log2nc) {
    // Stockham autosort passes: P = product of radices already applied
    int P = 1, rem = log2nc;
    while (rem >= 2) {
        const int t = nc >> 2;
        for (int i = threadIdx.x; i < t; i += blockDim.x) {
            const int k = i & (P - 1);
            float2 u0 = a[i], u1 = a[i + t], u2 = a[i + 2 * t], u3 = a[i + 3 * t];
            if (k) {
                float2 w1 = af_twiddle(k, 4 * P);
                float2 w2 = af_cmul(w1, w1), w3 = af_cmul(w2, w1);
                u1 = af_cmul(u1, w1); u2 = af_cmul(u2, w2); u3 = af_cmul(u3, w3);
            }
            float2 s02 = make_float2(u0.x + u2.x, u0.y + u2.y), d02 = make_float2(u0.x - u2.x, u0.y - u2.y);
            float2 s13 = make_float2(u1.x + u3.x, u1.y + u3.y);
            float2 d13 = make_float2(u1.y - u3.y, -(u1.x - u3.x));          // (u1-u3) * (-i)
            const int j = ((i - k) << 2) + k;
            b[j] = make_float2(s02.x + s13.x, s02.y + s13.y);
            b[j + P] = make_float2(d02.x + d13.x, d02.y + d13.y);
            b[j + 2 * P] = make_float2(s02.x - s13.x, s02.y - s13.y);
            b[j + 3 * P] = make_float2(d02.x - d13.x, d02.y - d13.y);
        }
        __syncthreads();
        float2 *tmp = a; a = b; b = tmp;
        P <<= 2; rem -= 2;
    }
    if (rem == 1) {
        const int t = nc >> 1;
        for (int i = threadIdx.x; i < t; i += blockDim.x) {
            const int k = i & (P - 1);
            float2 u0 = a[i], u1 = a[i + t];
            if (k) u1 = af_cmul(u1, af_twiddle(k, 2 * P));
            const int j = ((i - k) << 1) + k;
            b[j] = make_float2(u0.x + u1.x, u0.y + u1.y);
            b[j + P] = make_float2(u0.x - u1.x, u0.y - u1.y);
        }
        __syncthreads();
        float2 *tmp = a; a = b; b = tmp;
    }

    return a;
}
