// cqt.cu -- constant-Q transform kernels: fixed /2 decimator and per-octave kernel correlation.
//
// Replaces the reference's recursive-octave loop `_cqtObj_cqt` (src/cqt_algorithm.c:845-1061):
//   per octave: resampleObj_resample (/2, src/dsp/resample_algorithm.c:430-521) ; stftObj_stft with a
//   rect window and centre zero padding (src/stft_algorithm.c:601-694) ; compaction ; `__mcdot1` of the
//   half spectrum with the thresholded spectral kernels ; scaling (:972-989, 1020-1041).
//
// B200-first restatement: for a real frame x, sum_k X[k] K_b[k] == sum_n x[n] kappa_b[n] with
// kappa_b = DFT of the (thresholded, half-spectrum) kernel row, so one octave is a strided complex
// correlation  out[t][b] = sum_n xpad[t*hop + n] * kappa_b[n]  -- dense FP32 FMA work on data that
// stays in shared memory, with no per-frame FFT, no T x 512 spectra in HBM and no compaction pass.
// The signal tile is staged polyphase-major (xs[r][u] = x[hop*u + r]) so that the 32 lanes of a warp
// (32 consecutive frames) read consecutive words for every (r, a) step whatever the hop is.
#include <math.h>
#include "common.cuh"

namespace {

constexpr int kDecTile = 256;     // outputs per CTA

__global__ void __launch_bounds__(kDecTile) k_decimate2(const float *__restrict__ in, int inLength, long long inStride,
                                                        const float *__restrict__ left32, const float *__restrict__ right31,
                                                        float *__restrict__ out, long long outStride) {
    __shared__ float sx[2 * kDecTile + 64];
    __shared__ float sl[32], sr[32];
    const int outLength = inLength / 2;
    const int o0 = blockIdx.x * kDecTile;
    const float *x = in + (long long)blockIdx.y * inStride;
    if (threadIdx.x < 32) { sl[threadIdx.x] = left32[threadIdx.x]; sr[threadIdx.x] = threadIdx.x < 31 ? right31[threadIdx.x] : 0.0f; }
    const int m0 = 2 * o0 - 31;                               // first input sample needed by this tile
    for (int i = threadIdx.x; i < 2 * kDecTile + 64; i += kDecTile) {
        const int m = m0 + i;
        sx[i] = (m >= 0 && m < inLength) ? x[m] : 0.0f;
    }
    __syncthreads();
    const int o = o0 + threadIdx.x;
    if (o >= outLength) return;
    const int c = 2 * threadIdx.x + 31;                       // position of x[2*o] in sx
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; j++) acc = fmaf(sl[j], sx[c - j], acc);
#pragma unroll
    for (int j = 0; j < 31; j++) acc = fmaf(sr[j], sx[c + 1 + j], acc);
    out[(long long)blockIdx.y * outStride + o] = acc / sqrtf(0.5f);
}

struct OctParams {
    const float *sig; long long sigStride; int sigLength, validLength;
    int N, hop, T, bpo;
    const float2 *kappa;          // [bpo][N] (re, im)
    const float *scale;           // [bpo]
    float *outRe, *outIm; long long outStride; int num, colOff;
    int TT;                       // frames per CTA (even)
    int rowLen;                   // polyphase row pitch (floats)
    int rowsA;                    // ceil(N / hop)
    int nChunk;                   // kernel taps resident in shared memory at a time
};

constexpr int kBinsPerPass = 12;  // bins whose kernels sit in shared memory together
constexpr int kFT = 2;            // frames per thread
constexpr int kJG = 4;            // lane-groups over bins; 3 bins per thread

__global__ void k_cqt_octave(OctParams p) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float2 *sk = reinterpret_cast<float2 *>(smemRaw);                       // [kBinsPerPass][nChunk]
    float *xs = reinterpret_cast<float *>(smemRaw + sizeof(float2) * (size_t)kBinsPerPass * p.nChunk);
    const int clip = blockIdx.y;
    const int t0 = blockIdx.x * p.TT;
    const float *sig = p.sig + (long long)clip * p.sigStride;
    const int h = p.hop, N = p.N;

    // stage the tile's span of the zero-padded signal, polyphase-major
    const int span = (p.TT - 1) * h + N;
    const long long m0 = (long long)t0 * h - N / 2;          // signal index of padded position t0*h
    for (int i = threadIdx.x; i < span; i += blockDim.x) {
        const long long m = m0 + i;
        const float v = (m >= 0 && m < p.validLength) ? sig[m] : 0.0f;
        xs[(i % h) * p.rowLen + i / h] = v;
    }

    const int tl = threadIdx.x % (p.TT / kFT);               // frame lane inside the tile
    const int jg = threadIdx.x / (p.TT / kFT);               // bin group 0..3
    const int half = p.TT / kFT;

    for (int j0 = 0; j0 < p.bpo; j0 += kBinsPerPass) {
        const int nb = min(kBinsPerPass, p.bpo - j0);
        float ar[kFT][3], ai[kFT][3];
#pragma unroll
        for (int f = 0; f < kFT; f++)
#pragma unroll
            for (int u = 0; u < 3; u++) { ar[f][u] = 0.0f; ai[f][u] = 0.0f; }
        // the kernels of this pass are streamed through shared memory in chunks of p.nChunk taps
        for (int n0 = 0; n0 < N; n0 += p.nChunk) {
            const int n1 = min(N, n0 + p.nChunk), cw = n1 - n0;
            __syncthreads();                                   // xs staged / previous chunk consumed
            for (int i = threadIdx.x; i < kBinsPerPass * cw; i += blockDim.x) {
                const int j = i / cw, n = i - j * cw;
                sk[(size_t)j * p.nChunk + n] = j < nb ? p.kappa[(size_t)(j0 + j) * N + n0 + n] : make_float2(0.f, 0.f);
            }
            __syncthreads();
            const float2 *k0 = sk + (size_t)(jg * 3 + 0) * p.nChunk - n0, *k1 = sk + (size_t)(jg * 3 + 1) * p.nChunk - n0,
                         *k2 = sk + (size_t)(jg * 3 + 2) * p.nChunk - n0;
            for (int r = 0; r < h; r++) {
                const float *row = xs + r * p.rowLen + tl;
                int a = n0 > r ? (n0 - r + h - 1) / h : 0;
                for (int n = a * h + r; n < n1; a++, n += h) {
                    const float x0 = row[a], x1 = row[a + half];
                    const float2 c0 = k0[n], c1 = k1[n], c2 = k2[n];
                    ar[0][0] = fmaf(x0, c0.x, ar[0][0]); ai[0][0] = fmaf(x0, c0.y, ai[0][0]);
                    ar[0][1] = fmaf(x0, c1.x, ar[0][1]); ai[0][1] = fmaf(x0, c1.y, ai[0][1]);
                    ar[0][2] = fmaf(x0, c2.x, ar[0][2]); ai[0][2] = fmaf(x0, c2.y, ai[0][2]);
                    ar[1][0] = fmaf(x1, c0.x, ar[1][0]); ai[1][0] = fmaf(x1, c0.y, ai[1][0]);
                    ar[1][1] = fmaf(x1, c1.x, ar[1][1]); ai[1][1] = fmaf(x1, c1.y, ai[1][1]);
                    ar[1][2] = fmaf(x1, c2.x, ar[1][2]); ai[1][2] = fmaf(x1, c2.y, ai[1][2]);
                }
            }
        }
#pragma unroll
        for (int f = 0; f < kFT; f++) {
            const int t = t0 + tl + f * half;
            if (t >= p.T) continue;
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int j = j0 + jg * 3 + u;
                if (j >= p.bpo || jg * 3 + u >= nb) continue;
                const float s = p.scale[j];
                const long long o = (long long)clip * p.outStride + (long long)t * p.num + p.colOff + j;
                p.outRe[o] = ar[f][u] * s;
                p.outIm[o] = ai[f][u] * s;
            }
        }
    }
}

}  // namespace

extern "C" int af_launch_decimate2(const float *in, int inLength, int inStride, int batch, const float *left32,
                                   const float *right31, float *out, int outStride, void *stream) {
    const int outLength = inLength / 2;
    if (outLength <= 0 || batch <= 0) return AF_OK;
    if (batch > 65535) return af_fail(AF_ERR_ARG, "decimate2: batch %d > 65535 per launch", batch);
    dim3 grid((unsigned)((outLength + kDecTile - 1) / kDecTile), (unsigned)batch);
    k_decimate2<<<grid, kDecTile, 0, (cudaStream_t)stream>>>(in, inLength, inStride, left32, right31, out, outStride);
    AF_LAUNCH_CHECK("k_decimate2");
    return AF_OK;
}

extern "C" int af_launch_cqt_octave(const float *sig, int sigLength, int sigStride, int batch, int validLength,
                                    int fftLength, int hop, int timeLength, int bpo, const float *kappa2,
                                    const float *scale, int num, int colOff,
                                    float *outRe, float *outIm, void *stream) {
    if (batch <= 0 || timeLength <= 0) return AF_OK;
    if (hop < 1) return af_fail(AF_ERR_ARG, "cqt octave: hop < 1");
    if (batch > 65535) return af_fail(AF_ERR_ARG, "cqt octave: batch %d > 65535 per launch", batch);
    OctParams p;
    p.sig = sig; p.sigStride = sigStride; p.sigLength = sigLength; p.validLength = validLength;
    p.N = fftLength; p.hop = hop; p.T = timeLength; p.bpo = bpo;
    p.kappa = reinterpret_cast<const float2 *>(kappa2); p.scale = scale;
    p.outRe = outRe; p.outIm = outIm; p.outStride = (long long)timeLength * num; p.num = num; p.colOff = colOff;
    p.rowsA = (fftLength + hop - 1) / hop;
    p.nChunk = fftLength < 512 ? fftLength : 512;
    const size_t kBytes = sizeof(float2) * (size_t)kBinsPerPass * p.nChunk;
    static const int ttChoices[] = {256, 192, 128, 64, 32, 16, 8};
    int TT = 0;
    size_t smem = 0;
    for (int c = 0; c < 7; c++) {
        TT = ttChoices[c];
        int rowLen = TT + p.rowsA + 1;
        // pitch chosen so consecutive samples (r fastest) land in different banks while staging
        if (hop >= 32) rowLen |= 1; else { int want = 32 / hop; rowLen = ((rowLen + 31) / 32) * 32 + want; }
        p.rowLen = rowLen;
        smem = kBytes + sizeof(float) * (size_t)hop * rowLen;
        if (smem <= (TT > 64 ? 100 : 200) * 1024) break;
    }
    if (smem > 220 * 1024) return af_fail(AF_ERR_UNSUPPORTED, "cqt octave: fftLength %d with hop %d exceeds shared memory", fftLength, hop);
    p.TT = TT;
    cudaError_t e = cudaFuncSetAttribute(k_cqt_octave, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return af_cuda_check(e, "cudaFuncSetAttribute(k_cqt_octave)");
    dim3 grid((unsigned)((timeLength + TT - 1) / TT), (unsigned)batch);
    k_cqt_octave<<<grid, (TT / kFT) * kJG, smem, (cudaStream_t)stream>>>(p);
    AF_LAUNCH_CHECK("k_cqt_octave");
    return AF_OK;
}
