// mfcc_fused.cu -- fused framed-STFT(2048) -> |X|^2 / |X| -> banded filter bank -> log10 / cbrt ->
// ortho DCT-II -> first ccNum coefficients, one persistent kernel, samples read from HBM once.
//
// Replaces, for fftLength = 2048, the whole chain
//   stftObj_stft  (src/stft_algorithm.c:696-715, 790-801)  -> __mccut (src/reassign_algorithm.c:600-604)
//   -> __mcsquare / sqrtf (src/bft_algorithm.c:489-497) -> __mdot1 (:515-518)
//   -> log10f clamp / powf(1/3) (src/feature/xxcc_algorithm.c:124-140) -> fftObj_dct (:142-149) -> cut (:151-155)
// of the reference, which makes 6 passes over T x 2048 floats per clip plus a dense 1025 x 128 dot.
//
// Structure (one CTA per SM, persistent, static tile schedule => bit-identical results for any
// batch size / GPU count):
//   * warp W (of kFrameWarps) owns frame f0+W of the current tile; a tile is `framesPerTile`
//     consecutive frames of one clip whose sample span [(f0*hop), (f0+F-1)*hop + 2048) is brought
//     into shared memory ONCE by a 1-D TMA bulk copy (cp.async.bulk + mbarrier complete_tx),
//     double buffered through a full/empty mbarrier ring fed by a dedicated producer warp;
//   * the 2048 real samples are packed as 1024 complex points and transformed as 32 x 32:
//     lane n1 holds z[n1 + 32*n2] in registers (one complex value per 64-bit register pair, all
//     butterflies in packed FADD2/FMUL2/FFMA2), a generated straight-line 32-point DFT runs over n2,
//     twiddles W_1024^(n1*k) come half from a conflict-free transposed shared table and half from one
//     extra multiply by W_64^n1, a 33-padded shared transpose regroups the data (warp-private,
//     __syncwarp only), a second 32-point DFT runs over n1;  lane l then holds Z[l + 32*kb];
//   * real-FFT post-pass pairs bin k with 1024-k through one warp shuffle per component (both
//     powers |E +- W*O|^2 come from one evaluation);
//   * the banded bank is applied lane-per-filter from a zero-padded transposed weight table whose
//     per-filter start bins are shifted down (host planner) until the 32 lanes of a group read 32
//     different banks -> conflict-free;
//   * DCT-II: 4 lane-groups split the 128 inputs, 8 lanes x CT coefficients each, 2 xor-shuffles reduce.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "fft32_gen.cuh"

namespace {

constexpr int kN = 2048;            // fftLength
constexpr int kNC = 1024;           // packed complex points
constexpr int kFrameWarps = 12;     // consumer warps = max frames per tile
constexpr int kThreads = (kFrameWarps + 1) * 32;
constexpr int kStages = 2;
constexpr int kMaxNum = 128;        // filters (padded)
constexpr int kScratchFloats = 33 * 32 * 2;     // per warp: transpose buffer, later P / log-mel
constexpr int kPsPad = 1152;        // Ps[0..1024], zeros up to kPsPad, log-mel at [kPsPad, kPsPad+128)

struct Plan {                       // host-side descriptor of the device tables
    float *dWindowHalf;             // 2048, window * 0.5
    float2 *dTw1;                   // [32 ka][32 n1]  W_1024^(n1*ka)
    float2 *dTw2;                   // [32]            W_2048^lane (post-pass base twiddle)
    float *dMelW;                   // transposed zero-padded weights, group after group: [len_g][32]
    int *dMelStart;                 // 128
    float *dDct;                    // 4 quarter blocks, each 32 rows x ctStride (+8 pad between blocks)
    int melGroupLen[4];
    int melGroups;
    int melWFloats;
    int num, ccNum, ct, dataType;
};

struct Params {
    const float *data;
    float *out;
    const float *windowHalf;
    const float2 *tw1, *tw2;
    const float *melW;
    const int *melStart;
    const float *dct;
    long long dataStride;
    int batch, timeLength, hop;
    int framesPerTile, tilesPerClip;
    long long totalTiles;
    int spanFloats;                 // floats per stage buffer
    int melGroups, melWFloats;
    int melGroupLen[4];
    int ccNum, rectify, dataType;
};

// shared-memory carve-up (bytes), all 16-byte aligned
struct Smem {
    int spanOff, scratchOff, windowOff, tw1Off, tw2Off, melWOff, melStartOff, dctOff, barOff, total;
};

__host__ __device__ inline Smem carve(int spanFloats, int melWFloats, int ct) {
    Smem s; int o = 0;
    s.spanOff = o;     o += kStages * spanFloats * 4;
    s.scratchOff = o;  o += kFrameWarps * kScratchFloats * 4;
    s.windowOff = o;   o += kN * 4;
    s.tw1Off = o;      o += 1024 * 8;
    s.tw2Off = o;      o += 32 * 8;
    s.melWOff = o;     o += ((melWFloats * 4 + 15) / 16) * 16;
    s.melStartOff = o; o += kMaxNum * 4;
    s.dctOff = o;      o += 4 * (32 * ct * 8 + 8) * 4;
    s.barOff = o;      o += 2 * kStages * 8;
    s.total = o;
    return s;
}

template <int CT>
__global__ void __launch_bounds__(kThreads, 1) k_mfcc_fused(Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Smem L = carve(p.spanFloats, p.melWFloats, CT);
    float *span = reinterpret_cast<float *>(smem + L.spanOff);
    float *scratchAll = reinterpret_cast<float *>(smem + L.scratchOff);
    float2 *sWin2 = reinterpret_cast<float2 *>(smem + L.windowOff);
    float2 *sTw1 = reinterpret_cast<float2 *>(smem + L.tw1Off);
    float2 *sTw2 = reinterpret_cast<float2 *>(smem + L.tw2Off);
    float *sMelW = reinterpret_cast<float *>(smem + L.melWOff);
    int *sMelStart = reinterpret_cast<int *>(smem + L.melStartOff);
    float *sDct = reinterpret_cast<float *>(smem + L.dctOff);
    uint64_t *fullBar = reinterpret_cast<uint64_t *>(smem + L.barOff);
    uint64_t *emptyBar = fullBar + kStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time: tables -> shared, barriers ----
    for (int i = threadIdx.x; i < kN; i += kThreads) reinterpret_cast<float *>(sWin2)[i] = p.windowHalf[i];
    for (int i = threadIdx.x; i < 1024; i += kThreads) sTw1[i] = p.tw1[i];
    for (int i = threadIdx.x; i < 32; i += kThreads) sTw2[i] = p.tw2[i];
    for (int i = threadIdx.x; i < p.melWFloats; i += kThreads) sMelW[i] = p.melW[i];
    for (int i = threadIdx.x; i < kMaxNum; i += kThreads) sMelStart[i] = p.melStart[i];
    for (int i = threadIdx.x; i < 4 * (32 * CT * 8 + 8); i += kThreads) sDct[i] = p.dct[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) { af_mbar_init(&fullBar[s], 1); af_mbar_init(&emptyBar[s], kFrameWarps); }
        af_fence_barrier_init();
    }
    __syncthreads();

    const int F = p.framesPerTile;

    if (warp == kFrameWarps) {
        // ================= producer: one lane streams tile spans with TMA =================
        if (lane == 0) {
            int it = 0;
            for (long long tile = blockIdx.x; tile < p.totalTiles; tile += gridDim.x, ++it) {
                const int stage = it % kStages;
                const uint32_t round = (uint32_t)(it / kStages);
                af_mbar_wait(&emptyBar[stage], (round & 1u) ^ 1u);
                const long long clip = tile / p.tilesPerClip;
                const int f0 = (int)(tile % p.tilesPerClip) * F;
                const int nf = min(F, p.timeLength - f0);
                const uint32_t bytes = (uint32_t)(((nf - 1) * p.hop + kN) * 4);
                af_mbar_arrive_expect_tx(&fullBar[stage], bytes);
                af_tma_load_1d(span + (size_t)stage * p.spanFloats,
                               p.data + clip * p.dataStride + (long long)f0 * p.hop, bytes, &fullBar[stage]);
            }
        }
        return;
    }

    // ================= consumers: warp `warp` computes frame f0 + warp of every tile =================
    float *scratch = scratchAll + (size_t)warp * kScratchFloats;
    c64 *scr2 = reinterpret_cast<c64 *>(scratch);
    const int q = lane >> 3, c8 = lane & 7;
    const float *dctQ = sDct + q * (32 * CT * 8 + 8);
    const int partner = (32 - lane) & 31;
    const c64 w16 = c_from(sTw1[16 * 32 + lane]);            // W_1024^(16*lane) = W_64^lane
    const c64 wBase = c_from(sTw2[lane]);                    // W_2048^lane
    const c64 *sWinC = reinterpret_cast<const c64 *>(sWin2);
    const c64 *sTw1C = reinterpret_cast<const c64 *>(sTw1);

    int it = 0;
    for (long long tile = blockIdx.x; tile < p.totalTiles; tile += gridDim.x, ++it) {
        const int stage = it % kStages;
        const uint32_t round = (uint32_t)(it / kStages);
        const long long clip = tile / p.tilesPerClip;
        const int f0 = (int)(tile % p.tilesPerClip) * F;
        const int nf = min(F, p.timeLength - f0);
        const bool active = warp < nf;

        af_mbar_wait(&fullBar[stage], round & 1u);

        c64 z[32];
        if (active) {
            // ---- A: load 2048 samples (1024 packed pairs), apply 0.5*window ----
            const c64 *sp = reinterpret_cast<const c64 *>(span + (size_t)stage * p.spanFloats + warp * p.hop);
#pragma unroll
            for (int j = 0; j < 32; j++) z[j] = v_mul(sp[lane + 32 * j], sWinC[lane + 32 * j]);
        }
        __syncwarp();
        if (lane == 0) af_mbar_arrive(&emptyBar[stage]);     // span slot may be refilled
        if (!active) continue;

        // ---- B: 1024-point FFT as 32 x 32 ----
        af_fft32(z);                                          // over n2; Y[n1=lane][ka] at AF_BR5(ka)
#pragma unroll
        for (int ka = 0; ka < 32; ka++) {                     // times W_1024^(lane*ka) = W^(lane*(ka&15)) * W^(16*lane*(ka>>4))
            c64 y = z[AF_BR5(ka)];
            if (ka >= 16) y = c_mul(y, w16);
            if (ka & 15) y = c_mul(y, sTw1C[(ka & 15) * 32 + lane]);
            scr2[ka * 33 + lane] = y;
        }
        __syncwarp();
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) z[n1] = scr2[lane * 33 + n1];
        __syncwarp();
        af_fft32(z);                                          // over n1; Z[lane + 32*kb] at AF_BR5(kb)

        // ---- C: real-FFT post-pass + power / magnitude -> Ps[0..1024] ----
        // (window pre-scaled by 1/2, so E' = Z[k] + conj Z[N-k] and O' = -i (Z[k] - conj Z[N-k]) need no halving;
        //  X[k] = E' + W O', conj X[N-k] = E' - W O' with W = W_2048^k = W_2048^lane * W_64^kb)
#pragma unroll
        for (int kb = 0; kb < 16; kb++) {
            const c64 zk = z[AF_BR5(kb)];
            float pr, pi;
            c_unpack(z[AF_BR5(31 - kb)], pr, pi);
            pr = __shfl_sync(0xffffffffu, pr, partner);
            pi = __shfl_sync(0xffffffffu, pi, partner);
            c64 zp = c_pack(pr, pi);
            if (lane == 0) zp = z[AF_BR5((32 - kb) & 31)];
            const c64 zc = c_conj(zp);
            const c64 e = c_add(zk, zc);
            c64 o = c_mul_mi(c_sub(zk, zc));
            o = af_mul_w64(o, kb);                             // compile-time constant twiddle
            const c64 wo = c_mul(o, wBase);
            float pk = c_norm2(c_add(e, wo)), pn = c_norm2(c_sub(e, wo));
            if (p.dataType == SpectralData_Mag) { pk = sqrtf(pk); pn = sqrtf(pn); }
            const int k = lane + 32 * kb;
            scratch[k] = pk;
            scratch[kNC - k] = pn;
        }
        if (lane == 0) {                                       // k = 512 pairs with itself
            float pk = 4.0f * c_norm2(z[AF_BR5(16)]);
            if (p.dataType == SpectralData_Mag) pk = sqrtf(pk);
            scratch[512] = pk;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {                          // zero pad behind Ps for padded band reads
            const int idx = kNC + 1 + lane + 32 * i;
            if (idx < kPsPad) scratch[idx] = 0.0f;
        }
        __syncwarp();

        // ---- D: banded filter bank (lane = filter within group, bank-conflict-free starts) + rectify ----
        {
            const float *wg = sMelW + lane;
            for (int g = 0; g < p.melGroups; g++) {
                const int len = p.melGroupLen[g];              // multiple of 4
                const float *ps = scratch + sMelStart[g * 32 + lane];
                float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
                for (int i = 0; i < len; i += 4) {
                    acc0 = fmaf(ps[i], wg[i * 32], acc0);
                    acc1 = fmaf(ps[i + 1], wg[(i + 1) * 32], acc1);
                    acc2 = fmaf(ps[i + 2], wg[(i + 2) * 32], acc2);
                    acc3 = fmaf(ps[i + 3], wg[(i + 3) * 32], acc3);
                }
                float v = (acc0 + acc1) + (acc2 + acc3);
                if (p.rectify == CepstralRectify_CubicRoot) v = powf(v, 1.0f / 3.0f);
                else v = log10f(v < 1e-8f ? 1e-8f : v);
                scratch[kPsPad + g * 32 + lane] = v;
                wg += len * 32;
            }
            for (int g = p.melGroups; g < 4; g++) scratch[kPsPad + g * 32 + lane] = 0.0f;
        }
        __syncwarp();

        // ---- E: DCT-II: lane (q, c8) sums inputs of quarter q into coefficients c8 + 8t ----
        {
            float acc[CT];
#pragma unroll
            for (int t = 0; t < CT; t++) acc[t] = 0.0f;
            const float4 *l4 = reinterpret_cast<const float4 *>(scratch + kPsPad + q * 32);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float4 lv = l4[i];
                const float lm[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float *d = dctQ + (i * 4 + u) * (CT * 8) + c8;
#pragma unroll
                    for (int t = 0; t < CT; t++) acc[t] = fmaf(lm[u], d[t * 8], acc[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < CT; t++) {
                acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 8);
                acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 16);
            }
            if (q == 0) {
                float *o = p.out + ((long long)clip * p.timeLength + f0 + warp) * p.ccNum;
#pragma unroll
                for (int t = 0; t < CT; t++) {
                    const int c = c8 + 8 * t;
                    if (c < p.ccNum) o[c] = acc[t];
                }
            }
        }
        __syncwarp();
    }
}

void free_plan(Plan *pl) {
    if (!pl) return;
    af_dev_free(pl->dWindowHalf); af_dev_free(pl->dTw1); af_dev_free(pl->dTw2);
    af_dev_free(pl->dMelW); af_dev_free(pl->dMelStart); af_dev_free(pl->dDct);
    free(pl);
}

}  // namespace

// Mel plan: filters are processed in groups of 32 (lane = filter).  Each filter's first tap is moved
// down by delta in [0, 31] (extra taps get zero weight) until the 32 start bins of a group fall in 32
// different shared-memory banks, longest filters first; group length = max(len + delta), rounded up to 4.
static int plan_mel(const AfBands *bands, int num, int *startShifted /* kMaxNum */, int *groupLen /* 4 */) {
    int total = 0;
    for (int m = 0; m < kMaxNum; m++) startShifted[m] = 0;
    for (int g = 0; g < 4; g++) groupLen[g] = 0;
    for (int g = 0; g * 32 < num; g++) {
        int order[32], cnt = 0;
        for (int m = g * 32; m < num && m < g * 32 + 32; m++) order[cnt++] = m;
        for (int i = 1; i < cnt; i++)                         // insertion sort, longest first
            for (int j = i; j > 0 && bands->len[order[j]] > bands->len[order[j - 1]]; j--) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
        unsigned used = 0;
        int len = 0;
        for (int i = 0; i < cnt; i++) {
            const int m = order[i], s0 = bands->start[m];
            int delta = 0;
            while (delta < 32 && s0 - delta >= 0 && (used >> ((s0 - delta) & 31) & 1u)) delta++;
            if (delta >= 32 || s0 - delta < 0) delta = 0;     // no free bank reachable: accept a conflict
            used |= 1u << ((s0 - delta) & 31);
            startShifted[m] = s0 - delta;
            if (bands->len[m] + delta > len) len = bands->len[m] + delta;
        }
        len = (len + 3) & ~3;
        groupLen[g] = len;
        total += len * 32;
    }
    return total;
}

extern "C" int af_mfcc_fused_supported(int fftLength, int num, int ccNum, const AfBands *bands) {
    if (fftLength != kN || num < 1 || num > kMaxNum || ccNum < 1 || ccNum > 64 || !bands) return 0;
    int starts[kMaxNum], groupLen[4];
    const int floats = plan_mel(bands, num, starts, groupLen);
    for (int g = 0; g < 4; g++) if (groupLen[g] > kPsPad - (kNC + 1)) return 0;   // padded reads must stay inside the zero pad
    return floats * 4 <= 24 * 1024;                              // weight table budget in shared memory
}

extern "C" void af_mfcc_plan_free(void *plan) { free_plan(static_cast<Plan *>(plan)); }

extern "C" int af_mfcc_plan_build(void **planOut, int fftLength, int num, int ccNum, const float *window,
                                  const float *bank, const AfBands *bands, const float *dct /* ccNum x num */,
                                  int dataType) {
    *planOut = NULL;
    if (!af_mfcc_fused_supported(fftLength, num, ccNum, bands)) return af_fail(AF_ERR_UNSUPPORTED, "fused MFCC plan: unsupported configuration");
    Plan *pl = static_cast<Plan *>(calloc(1, sizeof(Plan)));
    if (!pl) return AF_ERR_NOMEM;
    pl->num = num; pl->ccNum = ccNum; pl->dataType = dataType;
    pl->ct = ccNum <= 16 ? 2 : ccNum <= 24 ? 3 : ccNum <= 40 ? 5 : 8;
    int rc = AF_OK;

    float *wh = static_cast<float *>(malloc(sizeof(float) * kN));
    for (int i = 0; i < kN; i++) wh[i] = 0.5f * window[i];
    rc = af_dev_upload(reinterpret_cast<void **>(&pl->dWindowHalf), wh, sizeof(float) * kN);
    free(wh);

    float2 *tw = static_cast<float2 *>(malloc(sizeof(float2) * 1024));
    for (int ka = 0; ka < 32 && rc == AF_OK; ka++)
        for (int n1 = 0; n1 < 32; n1++) {
            double a = -2.0 * M_PI * (double)(ka * n1) / 1024.0;
            tw[ka * 32 + n1] = make_float2((float)cos(a), (float)sin(a));
        }
    if (rc == AF_OK) rc = af_dev_upload(reinterpret_cast<void **>(&pl->dTw1), tw, sizeof(float2) * 1024);
    for (int k = 0; k < 32; k++) {
        double a = -2.0 * M_PI * (double)k / 2048.0;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    if (rc == AF_OK) rc = af_dev_upload(reinterpret_cast<void **>(&pl->dTw2), tw, sizeof(float2) * 32);
    free(tw);

    // transposed, zero-padded band weights: group g -> [len_g][32], starts shifted for bank-conflict-free reads
    const int width = kNC + 1;
    pl->melGroups = (num + 31) / 32;
    int starts[kMaxNum];
    const int total = plan_mel(bands, num, starts, pl->melGroupLen);
    pl->melWFloats = total;
    float *mw = static_cast<float *>(calloc((size_t)(total > 0 ? total : 1), sizeof(float)));
    int off = 0;
    for (int g = 0; g < pl->melGroups; g++) {
        for (int l = 0; l < 32; l++) {
            const int m = g * 32 + l;
            if (m >= num) continue;
            const int delta = bands->start[m] - starts[m];
            for (int i = 0; i < bands->len[m]; i++)
                mw[off + (i + delta) * 32 + l] = bank[(size_t)m * width + bands->start[m] + i];
        }
        off += pl->melGroupLen[g] * 32;
    }
    if (rc == AF_OK) rc = af_dev_upload(reinterpret_cast<void **>(&pl->dMelW), mw, sizeof(float) * (size_t)(total > 0 ? total : 1));
    if (rc == AF_OK) rc = af_dev_upload(reinterpret_cast<void **>(&pl->dMelStart), starts, sizeof(int) * kMaxNum);
    free(mw);

    // DCT table: quarter q block = rows m = 32q..32q+31, each row CT*8 floats (coefficient c at [c]),
    // blocks separated by 8 floats so the four lane-groups hit disjoint banks
    const int ct = pl->ct, blk = 32 * ct * 8 + 8;
    float *dt = static_cast<float *>(calloc((size_t)4 * blk, sizeof(float)));
    for (int m = 0; m < num; m++)
        for (int c = 0; c < ccNum; c++)
            dt[(m / 32) * blk + (m % 32) * (ct * 8) + c] = dct[(size_t)c * num + m];
    if (rc == AF_OK) rc = af_dev_upload(reinterpret_cast<void **>(&pl->dDct), dt, sizeof(float) * (size_t)4 * blk);
    free(dt);

    if (rc != AF_OK) { free_plan(pl); return rc; }
    *planOut = pl;
    return AF_OK;
}

extern "C" int af_launch_mfcc_fused(void *plan, const float *data, int dataLength, int batch, int timeLength,
                                    int slideLength, int rectifyType, float *out, void *stream) {
    Plan *pl = static_cast<Plan *>(plan);
    if (!pl) return af_fail(AF_ERR_ARG, "fused MFCC: no plan");
    if (batch <= 0 || timeLength <= 0) return AF_OK;
    if (slideLength % 4 || dataLength % 4 || (reinterpret_cast<uintptr_t>(data) & 15))
        return af_fail(AF_ERR_UNSUPPORTED, "fused MFCC needs 16-byte aligned clips and slideLength %% 4 == 0 (TMA bulk copy)");

    Params p;
    memset(&p, 0, sizeof(p));
    p.data = data; p.out = out; p.windowHalf = pl->dWindowHalf; p.tw1 = pl->dTw1; p.tw2 = pl->dTw2;
    p.melW = pl->dMelW; p.melStart = pl->dMelStart; p.dct = pl->dDct;
    p.dataStride = dataLength; p.batch = batch; p.timeLength = timeLength; p.hop = slideLength;
    p.melGroups = pl->melGroups; p.melWFloats = pl->melWFloats;
    for (int g = 0; g < 4; g++) p.melGroupLen[g] = pl->melGroupLen[g];
    p.ccNum = pl->ccNum; p.rectify = rectifyType; p.dataType = pl->dataType;

    // frames per tile: as many as fit the shared-memory budget (<= kFrameWarps)
    const int budget = 227 * 1024;
    int F = kFrameWarps;
    for (; F >= 1; F--) {
        int spanFloats = (F - 1) * slideLength + kN;
        if (carve(spanFloats, pl->melWFloats, pl->ct).total <= budget) break;
    }
    if (F < 1) return af_fail(AF_ERR_UNSUPPORTED, "fused MFCC: slideLength %d too large for shared memory", slideLength);
    if (F > timeLength) F = timeLength;
    p.framesPerTile = F;
    p.spanFloats = (F - 1) * slideLength + kN;
    p.tilesPerClip = (timeLength + F - 1) / F;
    p.totalTiles = (long long)p.tilesPerClip * batch;
    const int smemBytes = carve(p.spanFloats, pl->melWFloats, pl->ct).total;

    int sms = af_sm_count();
    if (sms <= 0) sms = 148;
    long long grid = p.totalTiles < sms ? p.totalTiles : sms;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
#define AF_MFCC_LAUNCH(CT_)                                                                              \
    e = cudaFuncSetAttribute(k_mfcc_fused<CT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);  \
    if (e != cudaSuccess) return af_cuda_check(e, "cudaFuncSetAttribute(k_mfcc_fused)");                 \
    k_mfcc_fused<CT_><<<(unsigned)grid, kThreads, smemBytes, st>>>(p)
    switch (pl->ct) {
    case 2: AF_MFCC_LAUNCH(2); break;
    case 3: AF_MFCC_LAUNCH(3); break;
    case 5: AF_MFCC_LAUNCH(5); break;
    default: AF_MFCC_LAUNCH(8); break;
    }
#undef AF_MFCC_LAUNCH
    AF_LAUNCH_CHECK("k_mfcc_fused");
    return AF_OK;
}
