/*
 * oracle/channelizer.c -- L3 restatement: firpfbch2 (2x oversampled) and firpfbch
 * (critically sampled) polyphase filterbank channelizers, in yagi's idiom
 * (per-branch Window + reverse-loaded sub-filter + sequential dotprod + one
 * unnormalised M-point Fft::run per frame).
 *
 * TEST INFRASTRUCTURE ONLY (see yagi_oracle.h).
 *
 * PARITY UNPINNED against the reference for these two objects:
 * /root/reference/src/multichannel/mod.rs is empty; the control flow follows the
 * upstream liquid-dsp algorithm that yagi tracks by test name
 * (LIQUID_COMPAT.md:1765-1798), restated in SURVEY.md Appendix A.1/A.2.  The
 * pieces it is assembled from ARE pinned:
 *   Window push/read ............ src/buffer/window.rs:66-85
 *   sub-filter reverse load ..... src/filter/fir/firpfb.rs:45-52
 *   dot product order ........... src/dotprod/mod.rs:36-39
 *   Fft direction / scaling ..... src/fft/mod.rs:13-26,45-48
 *   tap normalisation idiom ..... src/filter/resampler/resamp.rs:49-51
 *   two-bank toggle idiom ....... src/filter/resampler/resamp2.rs:104-151
 *   constructor validation ...... src/filter/fir/firdecim.rs:38-44
 */
#include "yagi_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ================================================================ firpfbch2 */

struct orc_firpfbch2_s {
    int          type;
    uint32_t     M, M2, m;
    size_t       h_len;        /* 2*M*m taps used */
    size_t       h_sub_len;    /* 2*m             */
    float*       h;            /* prototype, as given (h_len entries kept) */
    float*       h_sub;        /* [M][h_sub_len], reversed per firpfb.rs:45-52 */
    orc_fft*     ifft;         /* Backward, unnormalised */
    ocf32*       X;            /* IFFT input  */
    ocf32*       x;            /* IFFT output */
    orc_window** w0;           /* [M] */
    orc_window** w1;           /* [M] */
    int          flag;
};

int orc_firpfbch2_crcf_create(int type, uint32_t M, uint32_t m, const float* h, size_t h_len_given, orc_firpfbch2** out)
{
    if (type != ORC_ANALYZER && type != ORC_SYNTHESIZER) return ORC_ECONFIG;
    if (M < 2 || (M % 2)) return ORC_ECONFIG;
    if (m < 1) return ORC_ECONFIG;
    if (h == NULL || h_len_given < (size_t)2 * M * m) return ORC_ECONFIG;

    orc_firpfbch2* q = (orc_firpfbch2*)calloc(1, sizeof(*q));
    q->type = type;
    q->M = M;
    q->M2 = M / 2;
    q->m = m;
    q->h_len = (size_t)2 * M * m;
    q->h_sub_len = 2 * m;
    q->h = (float*)malloc(sizeof(float) * q->h_len);
    memcpy(q->h, h, sizeof(float) * q->h_len);
    q->h_sub = (float*)malloc(sizeof(float) * M * q->h_sub_len);
    for (uint32_t i = 0; i < M; i++)
        for (size_t n = 0; n < q->h_sub_len; n++)
            q->h_sub[i * q->h_sub_len + (q->h_sub_len - n - 1)] = q->h[i + n * M];
    q->ifft = orc_fft_create(M, 1);
    q->X = (ocf32*)calloc(M, sizeof(ocf32));
    q->x = (ocf32*)calloc(M, sizeof(ocf32));
    q->w0 = (orc_window**)malloc(sizeof(orc_window*) * M);
    q->w1 = (orc_window**)malloc(sizeof(orc_window*) * M);
    for (uint32_t i = 0; i < M; i++) {
        q->w0[i] = orc_window_create((uint32_t)q->h_sub_len);
        q->w1[i] = orc_window_create((uint32_t)q->h_sub_len);
    }
    q->flag = 0;
    *out = q;
    return ORC_OK;
}

/* Appendix A.1 create_kaiser: n = 2Mm+1, fc = 1/M (analyser) | 0.5/M (synthesiser),
 * h <- h * M / sum(h) (resamp.rs:49-51 idiom), last tap unused. */
int orc_firpfbch2_crcf_create_kaiser(int type, uint32_t M, uint32_t m, float as, orc_firpfbch2** out)
{
    if (type != ORC_ANALYZER && type != ORC_SYNTHESIZER) return ORC_ECONFIG;
    if (M < 2 || (M % 2)) return ORC_ECONFIG;
    if (m < 1) return ORC_ECONFIG;
    const uint32_t n = 2 * M * m + 1;
    float* hf = (float*)malloc(sizeof(float) * n);
    const float fc = (type == ORC_ANALYZER) ? 1.0f / (float)M : 0.5f / (float)M;
    int rc = orc_fir_design_kaiser(n, fc, as, 0.0f, hf);
    if (rc) { free(hf); return rc; }
    float sum = 0.0f;
    for (uint32_t i = 0; i < n; i++) sum += hf[i];
    for (uint32_t i = 0; i < n; i++) hf[i] = hf[i] * (float)M / sum;
    rc = orc_firpfbch2_crcf_create(type, M, m, hf, n, out);
    free(hf);
    return rc;
}

int orc_firpfbch2_crcf_clone(const orc_firpfbch2* s, orc_firpfbch2** out)
{
    orc_firpfbch2* q = NULL;
    int rc = orc_firpfbch2_crcf_create(s->type, s->M, s->m, s->h, s->h_len, &q);
    if (rc) return rc;
    for (uint32_t i = 0; i < s->M; i++) {
        orc_window_destroy(q->w0[i]);
        orc_window_destroy(q->w1[i]);
        q->w0[i] = orc_window_clone(s->w0[i]);
        q->w1[i] = orc_window_clone(s->w1[i]);
    }
    q->flag = s->flag;
    *out = q;
    return ORC_OK;
}

void orc_firpfbch2_crcf_destroy(orc_firpfbch2* q)
{
    if (!q) return;
    for (uint32_t i = 0; i < q->M; i++) {
        orc_window_destroy(q->w0[i]);
        orc_window_destroy(q->w1[i]);
    }
    free(q->w0); free(q->w1);
    free(q->X); free(q->x);
    orc_fft_destroy(q->ifft);
    free(q->h_sub); free(q->h);
    free(q);
}

void orc_firpfbch2_crcf_reset(orc_firpfbch2* q)
{
    for (uint32_t i = 0; i < q->M; i++) {
        orc_window_reset(q->w0[i]);
        orc_window_reset(q->w1[i]);
    }
    q->flag = 0;
}

/* Appendix A.1 execute_analyzer: M/2 in -> M out */
static void firpfbch2_execute_analyzer(orc_firpfbch2* q, const ocf32* x, ocf32* y)
{
    const uint32_t M = q->M, M2 = q->M2;
    /* 1. load buffers in blocks of M/2 starting at the appropriate base index */
    const uint32_t base = q->flag ? M : M2;
    for (uint32_t i = 0; i < M2; i++)
        orc_window_push(q->w0[base - i - 1], x[i]);
    /* 2. branch filters; result lands at the (rotated) window index */
    const uint32_t offset = q->flag ? M2 : 0;
    for (uint32_t i = 0; i < M; i++) {
        const uint32_t b = (offset + i) % M;
        q->X[b] = orc_dotprod_rcc(q->h_sub + (size_t)i * q->h_sub_len, orc_window_read(q->w0[b]), q->h_sub_len);
    }
    /* 3. unnormalised backward transform, scale by 1/M */
    orc_fft_run(q->ifft, q->X, q->x);
    for (uint32_t i = 0; i < M; i++) {
        y[i].re = q->x[i].re / (float)M;
        y[i].im = q->x[i].im / (float)M;
    }
    q->flag = 1 - q->flag;
}

/* Appendix A.1 execute_synthesizer: M in -> M/2 out */
static void firpfbch2_execute_synthesizer(orc_firpfbch2* q, const ocf32* x, ocf32* y)
{
    const uint32_t M = q->M, M2 = q->M2;
    /* 1. u = IFFT(x) * (1/M) * (M/2), applied as two f32 multiplies like upstream */
    memcpy(q->X, x, sizeof(ocf32) * M);
    orc_fft_run(q->ifft, q->X, q->x);
    const float s0 = 1.0f / (float)M;
    const float s1 = (float)M2;
    for (uint32_t i = 0; i < M; i++) {
        q->x[i].re *= s0; q->x[i].im *= s0;
        q->x[i].re *= s1; q->x[i].im *= s1;
    }
    /* 2. push into the bank selected by the flag */
    orc_window** buf = (q->flag == 0) ? q->w1 : q->w0;
    for (uint32_t i = 0; i < M; i++) orc_window_push(buf[i], q->x[i]);
    /* 3. weighted overlap-add over the two banks, swapping roles on alternate frames */
    for (uint32_t i = 0; i < M2; i++) {
        const uint32_t b = (q->flag == 0) ? i : i + M2;
        const ocf32* r0 = orc_window_read(q->w0[b]);
        const ocf32* r1 = orc_window_read(q->w1[b]);
        const ocf32* p0 = q->flag ? r0 : r1;
        const ocf32* p1 = q->flag ? r1 : r0;
        const ocf32 y0 = orc_dotprod_rcc(q->h_sub + (size_t)i * q->h_sub_len, p0, q->h_sub_len);
        const ocf32 y1 = orc_dotprod_rcc(q->h_sub + (size_t)(i + M2) * q->h_sub_len, p1, q->h_sub_len);
        y[i].re = y0.re + y1.re;
        y[i].im = y0.im + y1.im;
    }
    q->flag = 1 - q->flag;
}

int orc_firpfbch2_crcf_execute(orc_firpfbch2* q, const ocf32* x, ocf32* y)
{
    if (q->type == ORC_ANALYZER) firpfbch2_execute_analyzer(q, x, y);
    else firpfbch2_execute_synthesizer(q, x, y);
    return ORC_OK;
}

int orc_firpfbch2_crcf_execute_block(orc_firpfbch2* q, const ocf32* x, size_t n_frames, ocf32* y)
{
    const size_t nin = (q->type == ORC_ANALYZER) ? q->M2 : q->M;
    const size_t nout = (q->type == ORC_ANALYZER) ? q->M : q->M2;
    for (size_t k = 0; k < n_frames; k++)
        orc_firpfbch2_crcf_execute(q, x + k * nin, y + k * nout);
    return ORC_OK;
}

int orc_firpfbch2_crcf_get_type(const orc_firpfbch2* q) { return q->type; }
uint32_t orc_firpfbch2_crcf_get_M(const orc_firpfbch2* q) { return q->M; }
uint32_t orc_firpfbch2_crcf_get_m(const orc_firpfbch2* q) { return q->m; }
const float* orc_firpfbch2_crcf_taps(const orc_firpfbch2* q, size_t* len) { if (len) *len = q->h_len; return q->h; }

/* ================================================================= firpfbch */

struct orc_firpfbch_s {
    int          type;
    uint32_t     M, p;
    size_t       h_len;        /* M*p */
    float*       h;
    float*       h_sub;        /* [M][p] reversed */
    orc_fft*     fft;          /* Forward (analyser) | Backward (synthesiser) */
    ocf32*       X;
    ocf32*       x;
    orc_window** w;            /* [M] */
    uint32_t     filter_index;
};

int orc_firpfbch_crcf_create(int type, uint32_t M, uint32_t p, const float* h, size_t h_len_given, orc_firpfbch** out)
{
    if (type != ORC_ANALYZER && type != ORC_SYNTHESIZER) return ORC_ECONFIG;
    if (M == 0) return ORC_ECONFIG;
    if (p == 0) return ORC_ECONFIG;
    if (h == NULL || h_len_given < (size_t)M * p) return ORC_ECONFIG;

    orc_firpfbch* q = (orc_firpfbch*)calloc(1, sizeof(*q));
    q->type = type;
    q->M = M;
    q->p = p;
    q->h_len = (size_t)M * p;
    q->h = (float*)malloc(sizeof(float) * q->h_len);
    memcpy(q->h, h, sizeof(float) * q->h_len);
    q->h_sub = (float*)malloc(sizeof(float) * q->h_len);
    for (uint32_t i = 0; i < M; i++)
        for (uint32_t n = 0; n < p; n++)
            q->h_sub[(size_t)i * p + (p - n - 1)] = q->h[i + (size_t)n * M];
    q->fft = orc_fft_create(M, type == ORC_SYNTHESIZER);
    q->X = (ocf32*)calloc(M, sizeof(ocf32));
    q->x = (ocf32*)calloc(M, sizeof(ocf32));
    q->w = (orc_window**)malloc(sizeof(orc_window*) * M);
    for (uint32_t i = 0; i < M; i++) q->w[i] = orc_window_create(p);
    q->filter_index = M - 1;
    *out = q;
    return ORC_OK;
}

/* Appendix A.2 create_kaiser: fc = 0.5/M, n = 2Mm+1, p = 2m, no normalisation */
int orc_firpfbch_crcf_create_kaiser(int type, uint32_t M, uint32_t m, float as, orc_firpfbch** out)
{
    if (type != ORC_ANALYZER && type != ORC_SYNTHESIZER) return ORC_ECONFIG;
    if (M == 0) return ORC_ECONFIG;
    if (m == 0) return ORC_ECONFIG;
    const uint32_t n = 2 * M * m + 1;
    float* hf = (float*)malloc(sizeof(float) * n);
    int rc = orc_fir_design_kaiser(n, 0.5f / (float)M, as, 0.0f, hf);
    if (rc) { free(hf); return rc; }
    rc = orc_firpfbch_crcf_create(type, M, 2 * m, hf, n, out);
    free(hf);
    return rc;
}

int orc_firpfbch_crcf_clone(const orc_firpfbch* s, orc_firpfbch** out)
{
    orc_firpfbch* q = NULL;
    int rc = orc_firpfbch_crcf_create(s->type, s->M, s->p, s->h, s->h_len, &q);
    if (rc) return rc;
    for (uint32_t i = 0; i < s->M; i++) {
        orc_window_destroy(q->w[i]);
        q->w[i] = orc_window_clone(s->w[i]);
    }
    q->filter_index = s->filter_index;
    *out = q;
    return ORC_OK;
}

void orc_firpfbch_crcf_destroy(orc_firpfbch* q)
{
    if (!q) return;
    for (uint32_t i = 0; i < q->M; i++) orc_window_destroy(q->w[i]);
    free(q->w);
    free(q->X); free(q->x);
    orc_fft_destroy(q->fft);
    free(q->h_sub); free(q->h);
    free(q);
}

void orc_firpfbch_crcf_reset(orc_firpfbch* q)
{
    for (uint32_t i = 0; i < q->M; i++) orc_window_reset(q->w[i]);
    q->filter_index = q->M - 1;
}

int orc_firpfbch_crcf_execute(orc_firpfbch* q, const ocf32* x, ocf32* y)
{
    const uint32_t M = q->M, p = q->p;
    if (q->type == ORC_ANALYZER) {
        /* push M samples, walking the commutator backwards */
        for (uint32_t i = 0; i < M; i++) {
            orc_window_push(q->w[q->filter_index], x[i]);
            q->filter_index = (q->filter_index + M - 1) % M;
        }
        /* branch filters, output order reversed into the FFT input */
        for (uint32_t i = 0; i < M; i++)
            q->X[M - i - 1] = orc_dotprod_rcc(q->h_sub + (size_t)i * p, orc_window_read(q->w[i]), p);
        orc_fft_run(q->fft, q->X, q->x);                       /* Forward, no scaling */
        memcpy(y, q->x, sizeof(ocf32) * M);
    } else {
        memcpy(q->X, x, sizeof(ocf32) * M);
        orc_fft_run(q->fft, q->X, q->x);                       /* Backward, no scaling */
        for (uint32_t i = 0; i < M; i++) {
            orc_window_push(q->w[i], q->x[i]);
            y[i] = orc_dotprod_rcc(q->h_sub + (size_t)i * p, orc_window_read(q->w[i]), p);
        }
    }
    return ORC_OK;
}

int orc_firpfbch_crcf_execute_block(orc_firpfbch* q, const ocf32* x, size_t n_frames, ocf32* y)
{
    for (size_t k = 0; k < n_frames; k++)
        orc_firpfbch_crcf_execute(q, x + k * q->M, y + k * q->M);
    return ORC_OK;
}

const float* orc_firpfbch_crcf_taps(const orc_firpfbch* q, size_t* len) { if (len) *len = q->h_len; return q->h; }
