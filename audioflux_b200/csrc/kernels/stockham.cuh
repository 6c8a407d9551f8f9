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

// `tw` (may be null): device table tw[j] = exp(-2 pi i j / nc), j < nc, built once per (device, nc) on the host in double
// precision (af_twiddle_table).  With it every butterfly twiddle is one cached load instead of a sincospif evaluation.
__device__ __forceinline__ float2 af_tw(const float2 *tw, int k, int shift, int m) {
    return tw ? __ldg(tw + ((size_t)k << shift)) : af_twiddle(k, m);
}

__device__ __forceinline__ float2 *af_stockham(float2 *a, float2 *b, int nc, int log2nc, const float2 *tw = nullptr) {
    // Stockham autosort passes: P = product of radices already applied
    int P = 1, rem = log2nc;
    while (rem >= 2) {
        const int t = nc >> 2;
        for (int i = threadIdx.x; i < t; i += blockDim.x) {
            const int k = i & (P - 1);
            float2 u0 = a[i], u1 = a[i + t], u2 = a[i + 2 * t], u3 = a[i + 3 * t];
            if (k) {
                float2 w1 = af_tw(tw, k, rem - 2, 4 * P);              // nc / (4P) = 2^(rem-2)
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
            if (k) u1 = af_cmul(u1, af_tw(tw, k, 0, 2 * P));           // last pass: 2P = nc
            const int j = ((i - k) << 1) + k;
            b[j] = make_float2(u0.x + u1.x, u0.y + u1.y);
            b[j + P] = make_float2(u0.x - u1.x, u0.y - u1.y);
        }
        __syncthreads();
        float2 *tmp = a; a = b; b = tmp;
    }

    return a;
}

// host side: cached device tables per (device, log2 n): [0, n) exp(-2 pi i j / n) and, behind it, [0, n] exp(-2 pi i j / (2n))
// (the real-FFT post-pass twiddles of a 2n-point real transform packed into n complex points)
const float2 *af_twiddle_table(int log2n);
