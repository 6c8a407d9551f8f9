// istft.cu -- inverse STFT: per-frame complex inverse FFT, synthesis window, overlap-add, window-sum normalisation.
//
// Replaces stftObj_istft (src/stft_algorithm.c:304-409): for every frame  y_t = Re(IFFT_n(X_t))  (fftObj_ifft:
// conj -> forward FFT -> conj, divided by n, src/dsp/fft_algorithm.c:559-623), then
//     data[j] <- ( data[j] + sum_t y_t[j - t hop] w^e[j - t hop] ) / max-guarded( sum_t w^(e+1)[j - t hop] )
// with e = 1 ('weight', methodType 0) or e = 0 ('overlap-add'); a window sum below 1e-6 is replaced by 1.
// Two kernels: frames (one CTA per frame, shared-memory Stockham FFT) and a gather over the <= n/hop frames that
// cover an output sample, summed in ascending frame order like the reference's loop (bit-stable, no atomics).
#include <math.h>
#include "common.cuh"
#include "stockham.cuh"

namespace {

// spec planes [rows][width] with width = n (full spectrum) or n/2+1 (half: the rest is the Hermitian mirror)
__global__ void k_istft_frames(const float *__restrict__ re, const float *__restrict__ im, int width, int n, int log2n,
                               const float *__restrict__ window, int weightMode, float *__restrict__ frames,
                               const float2 *__restrict__ tw) {
    extern __shared__ float2 smem[];
    float2 *a = smem, *b = smem + n;
    const long long row = blockIdx.x;
    const float *r = re + row * width, *q = im + row * width;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        float xr, xi;
        if (k < width) { xr = r[k]; xi = q[k]; }
        else { xr = r[n - k]; xi = -q[n - k]; }
        a[k] = make_float2(xr, -xi);                 // conj in; the conj out only flips the unused imaginary part
    }
    __syncthreads();
    a = af_stockham(a, b, n, log2n, tw);
    const float inv = 1.0f / (float)n;
    float *f = frames + row * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float v = a[j].x * inv;
        if (weightMode && window) v *= window[j];
        f[j] = v;
    }
}

// same transform for frames whose ping-pong buffers do not fit shared memory (fftLength 16384: 2 x 128 KB): in-place
// radix-2 decimation-in-frequency passes over ONE buffer, result in bit-reversed order, undone while the frame is
// written out.  Only this size takes the path; the Stockham kernel above is 2-3x faster where it fits.
__global__ void k_istft_frames_inplace(const float *__restrict__ re, const float *__restrict__ im, int width, int n, int log2n,
                                       const float *__restrict__ window, int weightMode, float *__restrict__ frames,
                                       const float2 *__restrict__ tw) {
    extern __shared__ float2 smem[];
    float2 *a = smem;
    const long long row = blockIdx.x;
    const float *r = re + row * width, *q = im + row * width;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        float xr, xi;
        if (k < width) { xr = r[k]; xi = q[k]; }
        else { xr = r[n - k]; xi = -q[n - k]; }
        a[k] = make_float2(xr, -xi);
    }
    __syncthreads();
    for (int half = n >> 1, shift = 0; half >= 1; half >>= 1, shift++) {
        for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
            const int k = i & (half - 1), base = ((i - k) << 1) + k;
            const float2 u = a[base], v = a[base + half];
            a[base] = make_float2(u.x + v.x, u.y + v.y);
            float2 d = make_float2(u.x - v.x, u.y - v.y);
            if (k) d = af_cmul(d, af_tw(tw, k, shift, 2 * half));               // exp(-2 pi i k / (2 half)) = tw[k << shift]
            a[base + half] = d;
        }
        __syncthreads();
    }
    const float inv = 1.0f / (float)n;
    float *f = frames + row * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float v = a[__brev((unsigned)j) >> (32 - log2n)].x * inv;
        if (weightMode && window) v *= window[j];
        f[j] = v;
    }
}

__global__ void k_istft_ola(const float *__restrict__ frames, int n, int hop, int timeLength, int dataLength,
                            const float *__restrict__ window, int weightMode, float *__restrict__ data, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long clip = i / dataLength;
    const int j = (int)(i - clip * dataLength);
    int t0 = j - n + 1;
    t0 = t0 <= 0 ? 0 : (t0 + hop - 1) / hop;
    int t1 = j / hop;
    if (t1 > timeLength - 1) t1 = timeLength - 1;
    float acc = data[i], norm = 0.0f;
    const float *f = frames + clip * (long long)timeLength * n;
    for (int t = t0; t <= t1; t++) {
        const int k = j - t * hop;
        acc = acc + f[(long long)t * n + k];
        const float w = window ? window[k] : 1.0f;
        norm = norm + (weightMode ? w * w : w);      // w^(e+1)
    }
    if (norm < 1e-6f) norm = 1.0f;
    data[i] = acc / norm;
}

}  // namespace

extern "C" int af_launch_istft(const float *re, const float *im, int width, int fftLength, int slideLength, int timeLength,
                               int batch, const float *window, int methodType, float *frames, float *data, void *stream) {
    if (timeLength <= 0 || batch <= 0) return AF_OK;
    int log2n = 0;
    while ((1 << log2n) < fftLength) log2n++;
    size_t smem = sizeof(float2) * 2 * (size_t)fftLength;
    const bool inplace = smem > 200 * 1024;                     /* 16384 points: one buffer, in-place passes */
    if (inplace) smem /= 2;
    if (smem > 200 * 1024) return af_fail(AF_ERR_UNSUPPORTED, "istft: fftLength %d > 16384 does not fit the shared-memory FFT", fftLength);
    if (smem > 48 * 1024) {
        cudaError_t e = inplace ? cudaFuncSetAttribute(k_istft_frames_inplace, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                : cudaFuncSetAttribute(k_istft_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return af_cuda_check(e, "cudaFuncSetAttribute(k_istft_frames)");
    }
    const int weightMode = methodType == 0;
    int threads = fftLength / 4; if (threads < 32) threads = 32; if (threads > 1024) threads = 1024;
    cudaStream_t st = (cudaStream_t)stream;
    if (inplace)
        k_istft_frames_inplace<<<(unsigned)((long long)batch * timeLength), threads, smem, st>>>(re, im, width, fftLength, log2n, window,
                                                                                                weightMode, frames, af_twiddle_table(log2n));
    else
        k_istft_frames<<<(unsigned)((long long)batch * timeLength), threads, smem, st>>>(re, im, width, fftLength, log2n, window,
                                                                                        weightMode, frames, af_twiddle_table(log2n));
    AF_LAUNCH_CHECK("k_istft_frames");
    const int dataLength = (timeLength - 1) * slideLength + fftLength;
    const long long total = (long long)batch * dataLength;
    k_istft_ola<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(frames, fftLength, slideLength, timeLength, dataLength, window,
                                                                weightMode, data, total);
    AF_LAUNCH_CHECK("k_istft_ola");
    return AF_OK;
}
