/*
 * oracle/filters.c -- L2 restatement of the filter objects whose golden vectors pin
 * the building blocks of the channelizer: FirFilter, FirDecimationFilter, FirPfbFilter.
 * TEST INFRASTRUCTURE ONLY (see yagi_oracle.h).  Citations: /root/reference/.
 */
#include "yagi_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------ FirFilter */
/* src/filter/fir/firfilt.rs:10-15.  State is a VecDeque<T> of exactly h_len
 * elements, newest at the front (push = rotate_right(1); w[0] = x, :220-223).
 * execute = w.dotprod(h) * scale (:241-245); the VecDeque dot product sums the
 * deque's two contiguous slices separately and adds them
 * (src/dotprod/mod.rs:99-109), which we reproduce so the f32 summation order
 * is the reference's: after k pushes the ring head sits at (-k mod h_len), so
 * the front slice holds (k mod h_len) elements (or all of them when that is 0). */
struct orc_firfilt_s {
    float*   h;
    size_t   h_len;
    ocf32*   ring;       /* physical ring of h_len elements */
    size_t   head;       /* physical index of logical element 0 (newest) */
    float    scale;
};

int orc_firfilt_crcf_create(const float* h, size_t h_len, orc_firfilt** out)
{
    if (h_len == 0) return ORC_ECONFIG;                       /* firfilt.rs:65-67 */
    orc_firfilt* q = (orc_firfilt*)malloc(sizeof(*q));
    q->h = (float*)malloc(sizeof(float) * h_len);
    memcpy(q->h, h, sizeof(float) * h_len);
    q->h_len = h_len;
    q->ring = (ocf32*)malloc(sizeof(ocf32) * h_len);
    q->scale = 1.0f;
    orc_firfilt_crcf_reset(q);
    *out = q;
    return ORC_OK;
}

void orc_firfilt_crcf_destroy(orc_firfilt* q)
{
    if (!q) return;
    free(q->h);
    free(q->ring);
    free(q);
}

void orc_firfilt_crcf_reset(orc_firfilt* q)
{
    memset(q->ring, 0, sizeof(ocf32) * q->h_len);
    q->head = 0;
}

void orc_firfilt_crcf_set_scale(orc_firfilt* q, float scale) { q->scale = scale; }

void orc_firfilt_crcf_push(orc_firfilt* q, ocf32 x)
{
    q->head = (q->head + q->h_len - 1) % q->h_len;            /* rotate_right(1) */
    q->ring[q->head] = x;                                     /* w[0] = x        */
}

ocf32 orc_firfilt_crcf_execute(const orc_firfilt* q)
{
    /* as_slices(): front = ring[head..], back = ring[..head] */
    const size_t split = q->h_len - q->head;
    ocf32 l = {0.0f, 0.0f}, r = {0.0f, 0.0f};
    for (size_t i = 0; i < split; i++) {                      /* [Complex] x [f32], dotprod/mod.rs:47-53 */
        l.re += q->ring[q->head + i].re * q->h[i];
        l.im += q->ring[q->head + i].im * q->h[i];
    }
    for (size_t i = split; i < q->h_len; i++) {
        r.re += q->ring[i - split].re * q->h[i];
        r.im += q->ring[i - split].im * q->h[i];
    }
    ocf32 y = { (l.re + r.re) * q->scale, (l.im + r.im) * q->scale };
    return y;
}

int orc_firfilt_crcf_execute_block(orc_firfilt* q, const ocf32* x, size_t n, ocf32* y)
{
    for (size_t i = 0; i < n; i++) {                          /* firfilt.rs:272-275 */
        orc_firfilt_crcf_push(q, x[i]);
        y[i] = orc_firfilt_crcf_execute(q);
    }
    return ORC_OK;
}

/* ------------------------------------------------- FirDecimationFilter */
/* src/filter/fir/firdecim.rs:38-57 : taps stored reversed, Window of h_len;
 * execute (:179-191) pushes M samples and takes the dot product after the FIRST. */
struct orc_firdecim_s {
    float*      h_rev;
    size_t      h_len;
    uint32_t    M;
    orc_window* w;
    float       scale;
};

int orc_firdecim_crcf_create(uint32_t M, const float* h, size_t h_len, orc_firdecim** out)
{
    if (h_len == 0) return ORC_ECONFIG;
    if (M == 0) return ORC_ECONFIG;
    orc_firdecim* q = (orc_firdecim*)malloc(sizeof(*q));
    q->h_rev = (float*)malloc(sizeof(float) * h_len);
    for (size_t i = 0; i < h_len; i++) q->h_rev[i] = h[h_len - 1 - i];
    q->h_len = h_len;
    q->M = M;
    q->w = orc_window_create((uint32_t)h_len);
    q->scale = 1.0f;
    *out = q;
    return ORC_OK;
}

void orc_firdecim_crcf_destroy(orc_firdecim* q)
{
    if (!q) return;
    free(q->h_rev);
    orc_window_destroy(q->w);
    free(q);
}

ocf32 orc_firdecim_crcf_execute(orc_firdecim* q, const ocf32* x)
{
    ocf32 y = {0.0f, 0.0f};
    for (uint32_t i = 0; i < q->M; i++) {
        orc_window_push(q->w, x[i]);
        if (i == 0) {
            y = orc_dotprod_rcc(q->h_rev, orc_window_read(q->w), q->h_len);
            y.re *= q->scale;
            y.im *= q->scale;
        }
    }
    return y;
}

/* ------------------------------------------------------- FirPfbFilter */
/* src/filter/fir/firpfb.rs:34-65 : h_sub[h_sub_len-n-1] = h[i + n*num_filters];
 * one shared window; execute(i) = filters[i].dotprod(window.read()) * scale (:277-286).
 * Real-valued instantiation, as used by the reference's golden test (:310-359). */
struct orc_firpfb_s {
    uint32_t num_filters;
    size_t   h_sub_len;
    float*   h_sub;        /* [num_filters][h_sub_len] */
    float*   win;          /* same Window logic on f32: ring with linear view */
    uint32_t w_n, w_mask, w_read, w_alloc;
};

int orc_firpfb_rrrf_create(uint32_t num_filters, const float* h, size_t h_len, orc_firpfb** out)
{
    if (num_filters == 0) return ORC_ECONFIG;
    if (h_len == 0) return ORC_ECONFIG;
    orc_firpfb* q = (orc_firpfb*)malloc(sizeof(*q));
    q->num_filters = num_filters;
    q->h_sub_len = h_len / num_filters;
    q->h_sub = (float*)malloc(sizeof(float) * num_filters * q->h_sub_len);
    for (uint32_t i = 0; i < num_filters; i++)
        for (size_t n = 0; n < q->h_sub_len; n++)
            q->h_sub[i * q->h_sub_len + (q->h_sub_len - n - 1)] = h[i + n * num_filters];
    uint32_t k = 0, t = (uint32_t)q->h_sub_len;
    while (t) { k++; t >>= 1; }
    q->w_n = 1u << k;
    q->w_mask = q->w_n - 1;
    q->w_alloc = q->w_n + (uint32_t)q->h_sub_len - 1;
    q->w_read = 0;
    q->win = (float*)calloc(q->w_alloc, sizeof(float));
    *out = q;
    return ORC_OK;
}

void orc_firpfb_rrrf_destroy(orc_firpfb* q)
{
    if (!q) return;
    free(q->h_sub);
    free(q->win);
    free(q);
}

void orc_firpfb_rrrf_push(orc_firpfb* q, float x)
{
    q->w_read = (q->w_read + 1) & q->w_mask;
    if (q->w_read == 0) memmove(q->win, q->win + q->w_n, sizeof(float) * (q->h_sub_len - 1));
    q->win[q->w_read + q->h_sub_len - 1] = x;
}

int orc_firpfb_rrrf_execute(orc_firpfb* q, uint32_t i, float* y)
{
    if (i >= q->num_filters) return ORC_ECONFIG;
    *y = orc_dotprod_rrr(q->h_sub + (size_t)i * q->h_sub_len, q->win + q->w_read, q->h_sub_len);
    return ORC_OK;
}
