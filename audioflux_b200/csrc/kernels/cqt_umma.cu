// cqt_umma.cu -- one CQT octave on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Same restatement as cqt.cu (reference: src/cqt_algorithm.c:951-1048, the per-octave STFT + sparse spectral dot,
// rewritten as the strided correlation out[t][b] = sum_n xpad[t*hop + n] kappa_b[n]):
//     D[128 frames x 32] = A[128 x 512] . B[512 x 32],   A[t][n] = xpad[(t0 + t) hop + n]  (a Hankel matrix),
//     B = the 12 time-domain kernels as interleaved (re, im) columns (24 used).
// The Hankel operand is NEVER materialised: the staged signal itself is the UMMA A operand.  In the canonical K-major
// shared-memory layouts the 8 rows of a core-matrix group sit 16 / 32 / 64 / 128 bytes apart (no swizzle / 32B / 64B /
// 128B swizzle); a signal stored linearly has "rows" hop*4 bytes apart, so
//     hop  4 -> SWIZZLE_NONE  (row pitch 16 B; LBO = 16 B, SBO = 128 B),
//     hop  8 -> SWIZZLE_32B   (row pitch 32 B, SBO = 256 B),   hop 16 -> SWIZZLE_64B (64 B, SBO = 512 B),
//     hop 32 -> SWIZZLE_128B  (128 B, SBO = 1024 B),
//     hop 64 / 128 -> 2 / 4 phase planes of 128-byte rows (plane phi holds rows P u + phi), SWIZZLE_128B each,
// and the next K atom of the Hankel matrix is simply the SAME buffer one row further down (start address + row pitch,
// descriptor base_offset = row index mod 8).  The swizzle is an XOR on absolute shared-memory address bits, so the
// signal is stored through the same XOR and every shifted view reads it back consistently.
// fp32 accuracy with TF32 tensor cores: x = hi + lo (hi = top 19 bits, lo = x - hi exact); hi*hi + lo*hi + hi*lo, three
// MMAs per 8 taps, the kernels pre-split on the host.  One elected thread issues 192 MMAs per 128-frame tile;
// the other threads stage the next signal tile / run the epilogue of the co-resident CTA.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace {

constexpr int kUmmaM = 128;          // frames per tile (TMEM lanes)
constexpr int kUmmaN = 32;           // accumulator columns (24 used: 12 bins x (re, im))
constexpr int kUmmaChunkK = 128;     // taps per kernel chunk in shared memory
constexpr int kUmmaBBytes = kUmmaN * kUmmaChunkK * 4;      // one part (hi or lo) of one chunk: 16 KB (a chunk image holds both: 32 KB,
                                                            // per 32-tap K atom 64 rows x 128 B: rows 0-31 = hi, 32-63 = lo)

struct UmmaParams {
    const float *sig; long long sigStride; int validLength;
    int N, hop, T, padLeft;       // padLeft: N/2 (centre padding) or 0 (streaming)
    const unsigned char *bimg;     // [N / 128 chunks][2 parts (hi, lo)][16 KB] pre-swizzled shared-memory images of B
    const float *scale;            // [12]
    float *outRe, *outIm; long long outStride; int num, colOff;
    int mode;                      // 0: hop 4 (no swizzle), 1: hop 8 (32B), 2: hop 16 (64B), 3: hop 32 * planes (128B),
                                   // 4: hop 2 = two hop-4 problems (even / odd frames; the odd one reads a copy shifted by 2 samples)
    int planes, rowsPerPlane, sigBytes;   // mode 3: phase planes and rows (128 B each) per plane; bytes of one signal copy
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t smemAddr, uint32_t lboBytes, uint32_t sboBytes, uint32_t layout, uint32_t baseOff) {
    uint64_t d = 0;
    d |= (uint64_t)((smemAddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lboBytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sboBytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;                      // descriptor version (sm_100)
    d |= (uint64_t)(baseOff & 7u) << 49;
    d |= (uint64_t)(layout & 7u) << 61;
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(af_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// byte offset of sample s (floats from the tile start) inside one signal copy
__device__ __forceinline__ uint32_t sig_offset(const UmmaParams &p, int s) {
    if (p.mode == 0 || p.mode == 4) return (uint32_t)s * 4u;
    uint32_t a;
    if (p.mode == 3) {
        const int row = s >> 5, kk = s & 31;                       // 128-byte rows of 32 samples
        const int phi = row % p.planes, u = row / p.planes;
        a = (uint32_t)(phi * p.rowsPerPlane + u) * 128u + (uint32_t)kk * 4u;
        return a ^ (((a >> 7) & 7u) << 4);                          // 128B swizzle: bits [4,7) ^= bits [7,10)
    }
    a = (uint32_t)s * 4u;
    if (p.mode == 2) return a ^ (((a >> 7) & 3u) << 4);             // 64B swizzle: bits [4,6) ^= bits [7,9)
    return a ^ (((a >> 7) & 1u) << 4);                              // 32B swizzle: bit 4 ^= bit 7
}

// ============================================================================================
// Persistent, warp-specialised version: one CTA per SM loops over tiles; three groups of warps run as a pipeline
// connected by mbarriers, so the staging of tile k+1, the MMAs of tile k and the epilogue of tile k-1 overlap:
//   warps 0-3  epilogue : accFull[a] -> tcgen05.ld (TMEM lane = frame) -> accEmpty[a] -> 16-byte stores
//   warp  4    issuer   : sigFull[s], accEmpty[a] -> 192 tcgen05.mma (x2 for hop 2) -> tcgen05.commit -> sigEmpty[s], accFull[a]
//   warps 5-11 stagers  : sigEmpty[s] -> signal tile (hi / lo, swizzled) -> fence.proxy.async -> sigFull[s]
// The kernels (B, 128 KB hi + lo) stay RESIDENT in shared memory when the signal tiles leave room (hop <= 32: loaded once
// per CTA by TMA); for hop 64 / 128 they are streamed per tile through the two 32 KB slots as in k_cqt_octave_umma.
// Two accumulators in TMEM (four for hop 2) let the tensor core start tile k+1 while tile k is being read out.
// ============================================================================================
#define AF_TMEM_LD32(R, TADDR)                                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                              \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                               \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"               \
                 : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]),       \
                   "=r"(R[8]), "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), \
                   "=r"(R[16]), "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), \
                   "=r"(R[24]), "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31]) \
                 : "r"(TADDR))
constexpr int kPEpiWarps = 4, kPIssueWarps = 2, kPStageWarps = 7;
constexpr int kUmmaAccCols = 64;    // TMEM columns of one accumulator: 0-31 = hi.hi + lo.hi, 32-63 = hi.lo (summed in the epilogue)
constexpr int kPThreads = (kPEpiWarps + kPIssueWarps + kPStageWarps) * 32;

struct UmmaPParams {
    UmmaParams u;
    int tilesPerClip, stages, bResident;
    long long totalTiles;
};

__global__ void __launch_bounds__(kPThreads, 1) k_cqt_octave_umma_p(UmmaPParams pp) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const UmmaParams &p = pp.u;
    const int chunks = p.N / kUmmaChunkK;
    const int bBytes = pp.bResident ? chunks * 2 * kUmmaBBytes : 2 * 2 * kUmmaBBytes;
    unsigned char *sB = smem;
    unsigned char *sSig = smem + bBytes;                              // [stage][hi | lo][sigBytes]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sSig + (size_t)pp.stages * 2 * p.sigBytes);
    uint64_t *sigFull = bars, *sigEmpty = bars + 2, *accFull = bars + 4, *accEmpty = bars + 6, *bFull = bars + 8, *bEmpty = bars + 10;
    uint32_t *tmemSlot = reinterpret_cast<uint32_t *>(bars + 12);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (uniform for the compiler)
    const int par2 = p.mode == 4 ? 2 : 1;
    const int framesPerTile = kUmmaM * par2;
    const int h = p.hop;
    const int copyBytes = p.mode == 4 ? p.sigBytes / 2 : p.sigBytes;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            af_mbar_init(&sigFull[i], kPStageWarps); af_mbar_init(&sigEmpty[i], 1);
            af_mbar_init(&accFull[i], 1); af_mbar_init(&accEmpty[i], kPEpiWarps);
            af_mbar_init(&bFull[i], 1); af_mbar_init(&bEmpty[i], 1);
        }
        af_fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(af_smem_u32(tmemSlot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmemBase = *tmemSlot;

    if (warp >= kPEpiWarps + kPIssueWarps) {
        // ================= stagers =================
        const int sw = warp - (kPEpiWarps + kPIssueWarps), nst = kPStageWarps * 32, tid = sw * 32 + lane;
        int k = 0;
        for (long long tile = blockIdx.x; tile < pp.totalTiles; tile += gridDim.x, ++k) {
            const int s = k % pp.stages;
            af_mbar_wait(&sigEmpty[s], (((uint32_t)(k / pp.stages)) & 1u) ^ 1u);
            const int clip = (int)(tile / pp.tilesPerClip), t0 = (int)(tile % pp.tilesPerClip) * framesPerTile;
            const float *sig = p.sig + (long long)clip * p.sigStride;
            const long long m0 = (long long)t0 * h - p.padLeft;
            const int total = copyBytes / 4;
            const bool vec = ((p.sigStride & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.sig) & 15) == 0) && ((m0 & 3) == 0);
            unsigned char *sHi = sSig + (size_t)s * 2 * p.sigBytes, *sLo = sHi + p.sigBytes;
            for (int cp = 0; cp < par2; cp++) {
                const long long mc = m0 + 2 * cp;
                unsigned char *dHi = sHi + cp * copyBytes, *dLo = sLo + cp * copyBytes;
                constexpr int kLd = 8;
                for (int i0 = tid * 4; i0 < total; i0 += 4 * nst * kLd) {
                    float4 q[kLd];
#pragma unroll
                    for (int b = 0; b < kLd; b++) {
                        const int i = i0 + b * 4 * nst;
                        const long long m = mc + i;
                        q[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        if (i >= total) continue;
                        if (vec && cp == 0 && m >= 0 && m + 3 < p.validLength) q[b] = *reinterpret_cast<const float4 *>(sig + m);
                        else {
                            float *v = reinterpret_cast<float *>(&q[b]);
#pragma unroll
                            for (int u = 0; u < 4; u++) v[u] = (m + u >= 0 && m + u < p.validLength) ? sig[m + u] : 0.0f;
                        }
                    }
#pragma unroll
                    for (int b = 0; b < kLd; b++) {
                        const int i = i0 + b * 4 * nst;
                        if (i >= total) continue;
                        const float *v = reinterpret_cast<const float *>(&q[b]);
                        float4 hi4, lo4;
                        float *hp = reinterpret_cast<float *>(&hi4), *lp = reinterpret_cast<float *>(&lo4);
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            hp[u] = __uint_as_float(__float_as_uint(v[u]) & 0xffffe000u);
                            lp[u] = v[u] - hp[u];
                        }
                        const uint32_t off = sig_offset(p, i);
                        *reinterpret_cast<float4 *>(dHi + off) = hi4;
                        *reinterpret_cast<float4 *>(dLo + off) = lo4;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) af_mbar_arrive(&sigFull[s]);
        }
    } else if (warp >= kPEpiWarps) {
        // ================= MMA issuers (+ kernel loads) =================
        // The whole warp runs the loop on warp-uniform values (descriptors live in uniform registers); one elected lane
        // issues the asynchronous operations.  With the kernels resident and two signal stages there are TWO issuer warps,
        // each with its own stage and TMEM accumulator (tiles k = iw, iw + 2, ...): a single thread cannot feed the tensor
        // pipe with N = 64 MMAs.
        const int iw = warp - kPEpiWarps;
        const int nIssue = (pp.bResident && pp.stages == 2) ? kPIssueWarps : 1;
        const bool leader = elect_one();
        if (iw < nIssue) {
            constexpr uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(kUmmaM >> 4) << 24);
            constexpr uint32_t idesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(kUmmaM >> 4) << 24);
            uint32_t layoutA, sboA, lboA;
            if (p.mode == 0 || p.mode == 4) { layoutA = 0; sboA = 128; lboA = 16; }
            else if (p.mode == 1) { layoutA = 6; sboA = 256; lboA = 0; }
            else if (p.mode == 2) { layoutA = 4; sboA = 512; lboA = 0; }
            else { layoutA = 2; sboA = 1024; lboA = 0; }
            const int lgP = 31 - __clz(p.planes);                     // planes is a power of two (hop / 32)
            const int atoms = p.N / 32;                               // 32-tap K atoms (128 B of a signal row, 8 KB of a B chunk image)
            long long bLoads = 0, bUses = 0;                          // streamed mode: chunk loads issued / consumed so far
            if (pp.bResident) {
                if (iw == 0 && leader) {
                    af_mbar_arrive_expect_tx(&bFull[0], (uint32_t)(chunks * 2 * kUmmaBBytes));
                    for (int c = 0; c < chunks; c++)
                        af_tma_load_1d(sB + (size_t)c * 2 * kUmmaBBytes, p.bimg + (size_t)c * 2 * kUmmaBBytes, 2 * kUmmaBBytes, &bFull[0]);
                }
                af_mbar_wait(&bFull[0], 0);
            } else {
                for (; bLoads < 2; bLoads++)
                    if (leader) {
                        af_mbar_arrive_expect_tx(&bFull[bLoads], 2 * kUmmaBBytes);
                        af_tma_load_1d(sB + bLoads * 2 * kUmmaBBytes, p.bimg + (size_t)(bLoads % chunks) * 2 * kUmmaBBytes, 2 * kUmmaBBytes, &bFull[bLoads]);
                    }
            }
            const long long myTiles = (pp.totalTiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
            const long long totalUses = myTiles * chunks;
            const uint32_t cpStep = (uint32_t)(copyBytes >> 4);
            for (long long k = iw; k < myTiles; k += nIssue) {
                const int s = (int)(k % pp.stages), a = (int)(k & 1);
                af_mbar_wait(&sigFull[s], ((uint32_t)(k / pp.stages)) & 1u);
                af_mbar_wait(&accEmpty[a], (((uint32_t)(k >> 1)) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // descriptors differ only in the 14-bit start-address field (16-byte units): one base per operand, the K
                // steps just add to it
                const uint32_t aHi = af_smem_u32(sSig + (size_t)s * 2 * p.sigBytes);
                const uint64_t dAhi0 = umma_desc(aHi, lboA, sboA, layoutA, 0);
                const uint64_t dAlo0 = umma_desc(aHi + (uint32_t)p.sigBytes, lboA, sboA, layoutA, 0);
                const uint32_t tmemD = tmemBase + (uint32_t)a * (kUmmaAccCols * par2);
                uint32_t acc = 0;
                uint64_t dB0 = 0;
                for (int j = 0; j < atoms; j++) {
                    const int c = j >> 2;
                    int slot = 0;
                    if ((j & 3) == 0) {
                        uint32_t bBase;
                        if (pp.bResident) bBase = af_smem_u32(sB + (size_t)c * 2 * kUmmaBBytes);
                        else {
                            slot = (int)(bUses & 1);
                            af_mbar_wait(&bFull[slot], (uint32_t)(bUses >> 1) & 1u);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            bBase = af_smem_u32(sB + slot * 2 * kUmmaBBytes);
                        }
                        dB0 = umma_desc(bBase, 0, 1024, 2, 0);
                    }
                    // A: Hankel view of the signal copy at tap 32 j (the swizzle is an XOR on absolute address bits: a view
                    // that starts some rows further down needs no descriptor base offset)
                    const uint32_t aOff = p.mode == 3 ? (uint32_t)((j & (p.planes - 1)) * p.rowsPerPlane + (j >> lgP)) * 8u : (uint32_t)j * 8u;
                    const uint64_t dAh = dAhi0 + aOff, dAl = dAlo0 + aOff;
                    const uint64_t dB = dB0 + (uint64_t)((j & 3) * (8192 >> 4));
                    if (leader) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {                  // 8-tap K steps inside the atom: +32 bytes each
                            umma_tf32(tmemD, dAh + 2 * q, dB + 2 * q, idesc64, q == 0 ? acc : 1u);    // [hi.hi | hi.lo] -> columns 0-31 | 32-63
                            umma_tf32(tmemD, dAl + 2 * q, dB + 2 * q, idesc32, 1);                     // lo.hi -> columns 0-31
                            if (par2 == 2) {
                                umma_tf32(tmemD + kUmmaAccCols, dAh + cpStep + 2 * q, dB + 2 * q, idesc64, q == 0 ? acc : 1u);
                                umma_tf32(tmemD + kUmmaAccCols, dAl + cpStep + 2 * q, dB + 2 * q, idesc32, 1);
                            }
                        }
                    }
                    acc = 1;
                    if (!pp.bResident && (j & 3) == 3) {
                        // this chunk's MMAs are issued: its slot is refilled (with the chunk two uses ahead; the stream of
                        // chunks is periodic over the tiles) once they have read it -- but the wait for that comes only after
                        // the NEXT chunk's MMAs are in the pipe, so the tensor core does not drain
                        slot = (int)(bUses & 1);
                        if (leader) umma_commit(&bEmpty[slot]);
                        if (bUses >= 1 && bLoads < totalUses) {
                            const int ps = (int)((bUses - 1) & 1);       // previous use's slot
                            af_mbar_wait(&bEmpty[ps], (uint32_t)((bUses - 1) >> 1) & 1u);
                            if (leader) {
                                af_mbar_arrive_expect_tx(&bFull[ps], 2 * kUmmaBBytes);
                                af_tma_load_1d(sB + ps * 2 * kUmmaBBytes, p.bimg + (size_t)(bLoads % chunks) * 2 * kUmmaBBytes, 2 * kUmmaBBytes, &bFull[ps]);
                            }
                            bLoads++;
                        }
                        bUses++;
                    }
                }
                if (leader) {
                    umma_commit(&sigEmpty[s]);                         // the tile's MMAs have read the signal stage
                    umma_commit(&accFull[a]);                          // ... and the accumulator is complete
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue =================
        int k = 0;
        for (long long tile = blockIdx.x; tile < pp.totalTiles; tile += gridDim.x, ++k) {
            const int a = k & 1;
            const int clip = (int)(tile / pp.tilesPerClip), t0 = (int)(tile % pp.tilesPerClip) * framesPerTile;
            af_mbar_wait(&accFull[a], ((uint32_t)(k >> 1)) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int cp = 0; cp < par2; cp++) {
                uint32_t r[32], q[32];
                const uint32_t taddr = tmemBase + ((uint32_t)(warp * 32) << 16) + (uint32_t)a * (kUmmaAccCols * par2) + (uint32_t)cp * kUmmaAccCols;
                AF_TMEM_LD32(r, taddr);
                AF_TMEM_LD32(q, taddr + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cp == par2 - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) af_mbar_arrive(&accEmpty[a]);       // the tensor core may overwrite this accumulator
                }
                const int t = t0 + (warp * 32 + lane) * par2 + cp;
                if (t >= p.T) continue;
                const long long o = (long long)clip * p.outStride + (long long)t * p.num + p.colOff;
                float re[12], im[12];
#pragma unroll
                for (int j = 0; j < 12; j++) {
                    const float sc = p.scale[j];
                    re[j] = (__uint_as_float(r[2 * j]) + __uint_as_float(q[2 * j])) * sc;
                    im[j] = (__uint_as_float(r[2 * j + 1]) + __uint_as_float(q[2 * j + 1])) * sc;
                }
                if (((o & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.outRe) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.outIm) & 15) == 0)) {
#pragma unroll
                    for (int v = 0; v < 3; v++) {
                        *reinterpret_cast<float4 *>(p.outRe + o + 4 * v) = make_float4(re[4 * v], re[4 * v + 1], re[4 * v + 2], re[4 * v + 3]);
                        *reinterpret_cast<float4 *>(p.outIm + o + 4 * v) = make_float4(im[4 * v], im[4 * v + 1], im[4 * v + 2], im[4 * v + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 12; j++) { p.outRe[o + j] = re[j]; p.outIm[o + j] = im[j]; }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "n"(256) : "memory");
}

}  // namespace

// Pre-swizzled shared-memory images of the B operand: per 128-tap chunk ONE K-major 128B-swizzled [64 rows][128 k] matrix
// whose rows 0-31 are the TF32 hi parts of the 24 (+8 zero) columns and rows 32-63 the lo parts; a K atom (32 taps) is
// 64 rows x 128 B = 8192 B (8 groups of 8 rows).  A_hi . [B_hi | B_lo] is then one N = 64 MMA (A fetched once for both
// products) and A_lo . B_hi an N = 32 MMA on the first 32 rows of the same image.
extern "C" void af_cqt_umma_bimage(const float *kappa2 /* [12][N] (re, im) */, int N, unsigned char *out /* N/128 * 32 KB */) {
    memset(out, 0, (size_t)(N / kUmmaChunkK) * 2 * kUmmaBBytes);
    for (int c = 0; c < N / kUmmaChunkK; c++)
        for (int n = 0; n < 24; n++)
            for (int k = 0; k < kUmmaChunkK; k++) {
                const int b = n >> 1, part = n & 1;
                const float v = kappa2[((size_t)b * N + c * kUmmaChunkK + k) * 2 + part];
                uint32_t u;
                memcpy(&u, &v, 4);
                u &= 0xffffe000u;
                float hi;
                memcpy(&hi, &u, 4);
                const float lo = v - hi;
                const int ka = k >> 5, kk = k & 31;
                for (int half = 0; half < 2; half++) {
                    const int row = n + 32 * half, g = row >> 3, r = row & 7;
                    const size_t off = (size_t)ka * 8192 + (size_t)g * 1024 + (size_t)r * 128 + (size_t)(((kk >> 2) ^ r) * 16) + (size_t)(kk & 3) * 4;
                    memcpy(out + (size_t)c * 2 * kUmmaBBytes + off, half ? &lo : &hi, 4);
                }
            }
}

// tile geometry: mode, phase planes, bytes of the staged signal (all copies of one precision) and the dynamic shared memory
static size_t cqt_umma_geometry(int fftLength, int hop, int *mode, int *planes, int *rowsPerPlane, int *sigBytes) {
    *mode = hop == 2 ? 4 : hop == 4 ? 0 : hop == 8 ? 1 : hop == 16 ? 2 : 3;
    *planes = hop >= 32 ? hop / 32 : 1;
    *rowsPerPlane = 0;
    if (*mode == 3) {
        *rowsPerPlane = ((kUmmaM + fftLength / 32 / *planes + 1 + 7) / 8) * 8;     // rows u = t + j / planes, whole 8-row groups
        *sigBytes = *planes * *rowsPerPlane * 128;
    } else if (*mode == 4) {
        const int span = (kUmmaM - 1) * 4 + fftLength;                                // each parity is a hop-4 problem
        *sigBytes = 2 * (((span * 4 + 1023) / 1024) * 1024);                          // two copies (shift 0 / 2 samples)
    } else {
        const int span = (kUmmaM - 1) * hop + fftLength;
        *sigBytes = ((span * 4 + 1023) / 1024) * 1024;
    }
    return (size_t)2 * 2 * kUmmaBBytes + 2 * (size_t)*sigBytes + 64;
}

extern "C" int af_cqt_umma_supported(int fftLength, int hop, int bpo) {
    if (bpo != 12 || fftLength % kUmmaChunkK != 0 || fftLength < 2 * kUmmaChunkK) return 0;
    if (!(hop == 2 || hop == 4 || hop == 8 || hop == 16 || hop == 32 || hop == 64 || hop == 128)) return 0;
    int mode, planes, rpp, sb;
    return cqt_umma_geometry(fftLength, hop, &mode, &planes, &rpp, &sb) + 256 <= (size_t)227 * 1024;
}

extern "C" int af_launch_cqt_octave_umma(const float *sig, int sigStride, int batch, int validLength, int fftLength, int hop,
                                         int padLeft, int timeLength, const unsigned char *bimg, const float *scale, int num, int colOff,
                                         float *outRe, float *outIm, void *stream) {
    if (batch <= 0 || timeLength <= 0) return AF_OK;
    if (!af_cqt_umma_supported(fftLength, hop, 12)) return af_fail(AF_ERR_UNSUPPORTED, "cqt octave (tcgen05): fftLength %d hop %d", fftLength, hop);
    if (batch > 65535) return af_fail(AF_ERR_ARG, "cqt octave: batch %d > 65535 per launch", batch);
    UmmaParams p;
    p.sig = sig; p.sigStride = sigStride; p.validLength = validLength;
    p.N = fftLength; p.hop = hop; p.T = timeLength; p.padLeft = padLeft;
    p.bimg = bimg; p.scale = scale;
    p.outRe = outRe; p.outIm = outIm; p.outStride = (long long)timeLength * num; p.num = num; p.colOff = colOff;
    (void)cqt_umma_geometry(fftLength, hop, &p.mode, &p.planes, &p.rowsPerPlane, &p.sigBytes);
    const int framesPerTile = kUmmaM * (p.mode == 4 ? 2 : 1);
    {
        // persistent pipelined kernel: kernels resident + two signal stages when they fit, else streamed kernels
        UmmaPParams pp;
        pp.u = p;
        pp.tilesPerClip = (timeLength + framesPerTile - 1) / framesPerTile;
        pp.totalTiles = (long long)pp.tilesPerClip * batch;
        const size_t bAll = (size_t)(fftLength / kUmmaChunkK) * 2 * kUmmaBBytes, bSlots = (size_t)2 * 2 * kUmmaBBytes, lim = (size_t)227 * 1024 - 256;
        const size_t sig2 = 2 * (size_t)p.sigBytes;                  // hi + lo of one stage
        size_t smemP = 0;
        if (bAll + 2 * sig2 <= lim) { pp.bResident = 1; pp.stages = 2; smemP = bAll + 2 * sig2; }
        else if (bSlots + 2 * sig2 <= lim) { pp.bResident = 0; pp.stages = 2; smemP = bSlots + 2 * sig2; }
        else if (bSlots + sig2 <= lim) { pp.bResident = 0; pp.stages = 1; smemP = bSlots + sig2; }
        if (smemP) {
            smemP += 256;
            cudaError_t e2 = cudaFuncSetAttribute(k_cqt_octave_umma_p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemP);
            if (e2 != cudaSuccess) return af_cuda_check(e2, "cudaFuncSetAttribute(k_cqt_octave_umma_p)");
            int sms = af_sm_count();
            if (sms <= 0) sms = 148;
            const long long g = pp.totalTiles < sms ? pp.totalTiles : sms;
            k_cqt_octave_umma_p<<<(unsigned)g, kPThreads, smemP, (cudaStream_t)stream>>>(pp);
            AF_LAUNCH_CHECK("k_cqt_octave_umma_p");
            return AF_OK;
        }
    }
    return af_fail(AF_ERR_UNSUPPORTED, "cqt octave (tcgen05): the tile does not fit shared memory");
}
