/* af_filterbank.c -- auditory scales, band edges and dense filter banks (setup time, host).
 *
 * Behavioural spec (constants the device kernels consume; every integer outcome -- band-edge
 * bins, supports -- must equal the reference's, so the float32 evaluation order of the
 * published scale formulas is kept):
 *   scales      /root/reference/src/filterbank/auditory_filterBank.c:1023-1190
 *   band edges  :594-677      range revision :946-1021     bank styles :210-500
 *   range rules /root/reference/src/bft_algorithm.c:158-230, src/cwt_algorithm.c:137-196
 * Unlike the reference, writes are clipped to the bank's columns (the reference writes out of
 * bounds when a revised edge leaves [0, samplate/2], e.g. Linspace/Log with the default range).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "../af_internal.h"

/* ---- Hz <-> scale value.  `ref` = bin spacing (Linear) or bins per octave (Octave). ---- */
static float fre_to_scale(float fre, int scale, float ref) {
    switch (scale) {
    case SpectralFilterBankScale_Linear: return roundf(fre / ref);
    case SpectralFilterBankScale_Mel: return 2595 * log10f(1 + fre / 700);            /* O'Shaughnessy */
    case SpectralFilterBankScale_Bark: {                                              /* Traunmueller */
        float z = 26.81 * fre / (1960 + fre) - 0.53;
        if (z < 2) z = z + 0.15 * (2 - z);
        else if (z > 20.1) z = z + 0.22 * (z - 20.1);
        return z;
    }
    case SpectralFilterBankScale_Erb: {                                               /* Glasberg-Moore */
        float a = 21.3654;
        return a * log10f(1 + fre * 0.004368);
    }
    case SpectralFilterBankScale_Octave: return roundf(ref * log2(fre / 440));
    case SpectralFilterBankScale_Log: return log2(fre / 440);
    default: return fre;                                                              /* Linspace */
    }
}

static float scale_to_fre(float v, int scale, float ref) {
    switch (scale) {
    case SpectralFilterBankScale_Linear: return v * ref;
    case SpectralFilterBankScale_Mel: return 700 * (powf(10, v / 2595) - 1);
    case SpectralFilterBankScale_Bark: {
        if (v < 2) v = (v - 0.3) / 0.85;
        else if (v > 20.1) v = (v + 4.422) / 1.22;
        return 1960 * (v + 0.53) / (26.28 - v);
    }
    case SpectralFilterBankScale_Erb: {
        float a = 21.3654;
        return (powf(10, v / a) - 1) / 0.004368;
    }
    case SpectralFilterBankScale_Octave: return pow(2, v / ref) * 440;
    case SpectralFilterBankScale_Log: return pow(2, v) * 440;
    default: return v;
    }
}

static void linspace_f32(float start, float stop, int n, float *out) {
    float step = (stop - start) / (n - 1 > 0 ? n - 1 : 1);
    for (int i = 0; i < n; i++) out[i] = start + i * step;
}

static void default_log_range(float *lo, float *hi) {
    *lo = powf(2, -45 / 12.0) * 440;      /* C1 */
    *hi = powf(2, 38 / 12.0) * 440;       /* B7 */
}

int af_revise_range(int num, int fftLength, int samplate, const float *lowFre, const float *highFre,
                    int scale, int bpo, AfRange *out) {
    const int logLike = (scale == SpectralFilterBankScale_Octave || scale == SpectralFilterBankScale_Log);
    float lo = 0, hi = samplate / 2.0;
    if (lowFre && *lowFre >= 0 && *lowFre < samplate / 2.0) lo = *lowFre;
    if (lo == 0 && logLike) default_log_range(&lo, &hi);
    if (highFre && *highFre > 0 && *highFre <= samplate / 2.0) hi = *highFre;
    if (hi < lo) {
        lo = 0; hi = samplate / 2.0;
        if (logLike) default_log_range(&lo, &hi);
    }
    out->lowIndex = out->highIndex = 0;
    if (scale == SpectralFilterBankScale_Linear) {
        float det = samplate / (float)fftLength;
        float l = roundf(lo / det), h = l + num - 1;
        lo = l * det; hi = h * det;
        out->lowIndex = roundf(lo / det);
        out->highIndex = roundf(hi / det);
        if (hi > samplate / 2.0) return -1;
    } else if (scale == SpectralFilterBankScale_Octave) {
        float l = fre_to_scale(lo, scale, bpo), h = l + num - 1;
        lo = scale_to_fre(l, scale, bpo); hi = scale_to_fre(h, scale, bpo);
        if (hi > samplate / 2.0) return -1;
    }
    out->low = lo; out->high = hi;
    return 0;
}

/* widen [low, high] so that the num centres sit inside num+2 edge points (non-edge styles) */
static float widen_range(int num, int lengthForLinear, int samplate, int scale, int bpo,
                         float *low, float *high) {
    float ref = 0, lo = *low, hi = *high;
    if (scale == SpectralFilterBankScale_Octave) {
        ref = (bpo >= 4 && bpo <= 48) ? bpo : 12;
        float l = fre_to_scale(lo, scale, ref) - 1, h = l + num - 1 + 2;
        lo = scale_to_fre(l, scale, ref); hi = scale_to_fre(h, scale, ref);
    } else if (scale == SpectralFilterBankScale_Linear) {
        ref = samplate * 1.0 / lengthForLinear;
        float l = roundf(lo / ref) - 1, h = l + num - 1 + 2;
        lo = l * ref; hi = h * ref;
    } else if (scale == SpectralFilterBankScale_Linspace) {
        float d = (hi - lo) / (num - 1);
        lo = lo - d; hi = hi + d;
    } else if (scale == SpectralFilterBankScale_Log) {
        float l = fre_to_scale(lo, scale, 0), h = fre_to_scale(hi, scale, 0);
        float d = (h - l) / (num - 1);
        lo = scale_to_fre(l - d, scale, 0); hi = scale_to_fre(h + d, scale, 0);
    }
    *low = lo; *high = hi;
    return ref;
}

void af_band_edges(int num, int fftLength, int samplate, float lowFre, float highFre, int scale,
                   int bpo, int slaneyBins, int forCwt, float *freEdge, int *binEdge) {
    (void)forCwt;
    float ref = widen_range(num, fftLength, samplate, scale, bpo, &lowFre, &highFre);
    const int n = num + 2;
    linspace_f32(fre_to_scale(lowFre, scale, ref), fre_to_scale(highFre, scale, ref), n, freEdge);
    for (int i = 0; i < n; i++) freEdge[i] = scale_to_fre(freEdge[i], scale, ref);
    if (!binEdge) return;
    if (!slaneyBins) {
        for (int i = 0; i < n; i++) binEdge[i] = roundf(fftLength * freEdge[i] / samplate);
    } else {
        /* first FFT-grid frequency strictly above the edge; grid = linspace(0, sr - sr/n, n) */
        float step = ((samplate - samplate / (float)fftLength) - 0.0f) / (fftLength - 1 > 0 ? fftLength - 1 : 1);
        for (int i = 0; i < n; i++) {
            int j = 0;
            while (j < fftLength && !(0.0f + j * step > freEdge[i])) j++;
            binEdge[i] = j < fftLength ? j : 0;
        }
    }
}

static void put(float *bank, int width, int row, int col, float v) {
    if (col >= 0 && col < width) bank[(size_t)row * width + col] = v;
}

static void window_half_fill(float *bank, int width, int row, int style, int from, int to, int rising) {
    /* one flank of a window-designed filter (flux: auditory_filterBank.c:249-316) */
    static const int map[] = {0, 0, 0, 0, 0, Window_Hann, Window_Hamm, Window_Blackman, Window_Bohman,
                              Window_Kaiser, Window_Gauss};
    int span = to - from;               /* > 0 */
    int L = 2 * span + 1;
    double *w = (double *)malloc(sizeof(double) * (size_t)L);
    if (!w) return;
    af_window_symmetric(map[style], L, NULL, w);
    if (rising) for (int j = from, k = 0; j <= to; j++, k++) put(bank, width, row, j, (float)w[k]);
    else for (int j = from + 1, k = L / 2 + 1; j <= to; j++, k++) put(bank, width, row, j, (float)w[k]);
    free(w);
}

int af_auditory_filterbank(int num, int fftLength, int samplate, int scale, int style, int norm,
                           float lowFre, float highFre, int bpo, float *bank, float *freBandArr,
                           int *binBandArr) {
    if (num < 1 || fftLength < 2 || !bank) return AF_ERR_ARG;
    if (style == SpectralFilterBankStyle_Gammatone)
        return af_fail(AF_ERR_UNSUPPORTED, "gammatone filter banks are not implemented yet");
    const int width = fftLength / 2 + 1;
    float *fre = (float *)calloc((size_t)num + 2, sizeof(float));
    int *bin = (int *)calloc((size_t)num + 2, sizeof(int));
    if (!fre || !bin) { free(fre); free(bin); return AF_ERR_NOMEM; }
    memset(bank, 0, sizeof(float) * (size_t)num * width);
    af_band_edges(num, fftLength, samplate, lowFre, highFre, scale, bpo,
                  style == SpectralFilterBankStyle_Slaney, 0, fre, bin);

    if (scale == SpectralFilterBankScale_Linear) {
        for (int i = 1; i <= num; i++) { bin[i] -= 1; put(bank, width, i - 1, bin[i], 1.0f); }
    } else if (style == SpectralFilterBankStyle_Slaney) {
        /* triangles measured in Hz on the FFT grid */
        float step = ((samplate - samplate / (float)fftLength) - 0.0f) / (fftLength - 1 > 0 ? fftLength - 1 : 1);
        for (int i = 0; i < num; i++) {
            float up = fre[i + 1] - fre[i], down = fre[i + 2] - fre[i + 1];
            for (int j = bin[i]; j < bin[i + 1]; j++) put(bank, width, i, j, ((0.0f + j * step) - fre[i]) / up);
            for (int j = bin[i + 1]; j < bin[i + 2]; j++) put(bank, width, i, j, (fre[i + 2] - (0.0f + j * step)) / down);
        }
    } else if (style == SpectralFilterBankStyle_ETSI) {
        /* triangles measured in bins */
        for (int i = 1; i <= num; i++) {
            int l = bin[i - 1], c = bin[i], r = bin[i + 1];
            if (c > l) for (int j = l; j <= c; j++) put(bank, width, i - 1, j, 1.0 * (j - l) / (c - l));
            for (int j = c + 1; j <= r; j++) put(bank, width, i - 1, j, 1.0 * (r - j) / (r - c));
        }
    } else if (style == SpectralFilterBankStyle_Point) {
        for (int i = 1; i <= num; i++) put(bank, width, i - 1, bin[i], 1.0f);
    } else if (style == SpectralFilterBankStyle_Rect) {
        for (int i = 1; i <= num; i++)
            for (int j = bin[i - 1]; j <= bin[i + 1]; j++) put(bank, width, i - 1, j, 1.0f);
    } else {
        for (int i = 1; i <= num; i++) {
            int l = bin[i - 1], c = bin[i], r = bin[i + 1];
            if (c > l) window_half_fill(bank, width, i - 1, style, l, c, 1);
            if (r > c) window_half_fill(bank, width, i - 1, style, c, r, 0);
        }
    }

    if (scale != SpectralFilterBankScale_Linear &&
        (norm == SpectralFilterBankNormal_Area || norm == SpectralFilterBankNormal_BandWidth)) {
        for (int i = 0; i < num; i++) {
            float wt = 0;
            float *row = bank + (size_t)i * width;
            if (norm == SpectralFilterBankNormal_Area) for (int j = 0; j < width; j++) wt += row[j];
            else wt = (fre[i + 2] - fre[i]) / 2;
            for (int j = 0; j < width; j++) if (row[j]) row[j] = row[j] / wt;   /* exact zeros stay zero */
        }
    }
    if (freBandArr) memcpy(freBandArr, fre + 1, sizeof(float) * (size_t)num);
    if (binBandArr) memcpy(binBandArr, bin + 1, sizeof(int) * (size_t)num);
    free(fre); free(bin);
    return AF_OK;
}

int afb200_auditoryFilterBank(int num, int fftLength, int samplate, int scaleType, int styleType,
                              int normType, float lowFre, float highFre, int binPerOctave,
                              float *bank, float *freBandArr, int *binBandArr) {
    return af_auditory_filterbank(num, fftLength, samplate, scaleType, styleType, normType, lowFre,
                                  highFre, binPerOctave, bank, freBandArr, binBandArr);
}

int af_bands_build(const float *bank, int num, int width, AfBands *b) {
    b->num = num; b->width = width; b->nnz = 0; b->maxLen = 0;
    b->start = (int *)calloc((size_t)num, sizeof(int));
    b->len = (int *)calloc((size_t)num, sizeof(int));
    if (!b->start || !b->len) return AF_ERR_NOMEM;
    for (int i = 0; i < num; i++) {
        const float *row = bank + (size_t)i * width;
        int first = -1, last = -1;
        for (int j = 0; j < width; j++) if (row[j] != 0.0f) { if (first < 0) first = j; last = j; }
        b->start[i] = first < 0 ? 0 : first;
        b->len[i] = first < 0 ? 0 : last - first + 1;
        b->nnz += b->len[i];
        if (b->len[i] > b->maxLen) b->maxLen = b->len[i];
    }
    return AF_OK;
}

void af_bands_free(AfBands *b) { free(b->start); free(b->len); b->start = b->len = NULL; }

/* ortho-normalised DCT-II rows: D[k][j] = s_k cos(pi (j + 1/2) k / N)
 * (reference: src/dsp/fft_algorithm.c:625-674, src/dsp/dct_algorithm.c:81-110, 170-181) */
void af_dct2_matrix(int num, int ccNum, float *out) {
    for (int k = 0; k < ccNum; k++) {
        double s = sqrt((k == 0 ? 1.0 : 2.0) / num);
        for (int j = 0; j < num; j++) out[(size_t)k * num + j] = (float)(s * cos(M_PI * (j + 0.5) * k / num));
    }
}

void af_fft_twiddles(int n, float *c, float *s) {
    for (int i = 0; i < n / 2; i++) { c[i] = (float)cos(2.0 * M_PI * i / n); s[i] = (float)-sin(2.0 * M_PI * i / n); }
}
