/*
 * oracle/core.c -- L0/L1 restatement: Kaiser design math, Window, dotprod, Fft.
 * TEST INFRASTRUCTURE ONLY (see yagi_oracle.h).  Citations: /root/reference/.
 */
#include "yagi_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI_F 3.14159265358979323846f

/* ------------------------------------------------------------------ math */

/* src/math/gamma.rs:7-22 -- recursion up to z >= 10, then the Stirling-like form */
float orc_lngammaf(float z)
{
    if (z <= 0.0f) return NAN;                       /* reference panics */
    if (z < 10.0f) return orc_lngammaf(z + 1.0f) - logf(z);
    float g = 0.5f * (logf(2.0f * ORC_PI_F) - logf(z));
    g += z * (logf(z + (1.0f / (12.0f * z - 0.1f / z))) - 1.0f);
    return g;
}

/* src/math/bessel.rs:9-41 -- 64-term log-domain series */
float orc_lnbesselif(float nu, float z)
{
    if (z == 0.0f) return nu == 0.0f ? 0.0f : -INFINITY;
    if (nu == 0.5f) return 0.5f * logf(2.0f / (ORC_PI_F * z)) + logf(sinhf(z));
    if (z < 1e-3f * sqrtf(nu + 1.0f)) return -orc_lngammaf(nu + 1.0f) + nu * logf(0.5f * z);

    float t0 = nu * logf(0.5f * z);
    float y = 0.0f;
    for (int k = 0; k < 64; k++) {
        float t1 = 2.0f * (float)k * logf(0.5f * z);
        float t2 = orc_lngammaf((float)k + 1.0f);
        float t3 = orc_lngammaf(nu + (float)k + 1.0f);
        y += expf(t1 - t2 - t3);
    }
    return t0 + logf(y);
}

/* src/math/bessel.rs:44-67 (besselif with nu = 0) */
float orc_besseli0f(float z)
{
    const float nu = 0.0f;
    if (z == 0.0f) return 1.0f;
    if (z < 1e-3f * sqrtf(nu + 1.0f))
        return powf(0.5f * z, nu) / expf(orc_lngammaf(nu + 1.0f));   /* gammaf(z>=0) = exp(lngammaf), gamma.rs:25-42 */
    return expf(orc_lnbesselif(nu, z));
}

/* src/math/mod.rs:63-69 */
float orc_sincf(float x)
{
    if (fabsf(x) < 0.01f)
        return cosf(ORC_PI_F * x / 2.0f) * cosf(ORC_PI_F * x / 4.0f) * cosf(ORC_PI_F * x / 8.0f);
    return sinf(ORC_PI_F * x) / (ORC_PI_F * x);
}

/* src/math/windows.rs:76-90 */
int orc_kaiser(uint32_t i, uint32_t wlen, float beta, float* out)
{
    if (i >= wlen) return ORC_EVALUE;
    if (beta < 0.0f) return ORC_EVALUE;
    float t = (float)i - (float)(wlen - 1) / 2.0f;
    float r = 2.0f * t / (float)(wlen - 1);
    float a = orc_besseli0f(beta * sqrtf(1.0f - r * r));
    float b = orc_besseli0f(beta);
    *out = a / b;
    return ORC_OK;
}

/* src/filter/fir/design/kaiser.rs:62-72 */
float orc_kaiser_beta_as(float as)
{
    float a = fabsf(as);
    if (a > 50.0f) return 0.1102f * (a - 8.7f);
    if (a > 21.0f) return 0.5842f * powf(a - 21.0f, 0.4f) + 0.07886f * (a - 21.0f);
    return 0.0f;
}

/* src/filter/fir/design/kaiser.rs:16-51 */
int orc_fir_design_kaiser(uint32_t n, float fc, float as, float mu, float* h)
{
    if (mu <= -0.5f || mu > 0.5f) return ORC_ECONFIG;
    if (fc <= 0.0f || fc > 0.5f) return ORC_ECONFIG;
    if (n == 0) return ORC_ECONFIG;
    if (as <= 0.0f) return ORC_ECONFIG;
    float beta = orc_kaiser_beta_as(as);
    for (uint32_t i = 0; i < n; i++) {
        float t = (float)i - ((float)n - 1.0f) / 2.0f + mu;
        float h1 = orc_sincf(2.0f * fc * t);
        float h2;
        int rc = orc_kaiser(i, n, beta, &h2);
        if (rc) return rc;
        h[i] = h1 * h2;
    }
    return ORC_OK;
}

/* ---------------------------------------------------------------- window */

/* src/buffer/window.rs:3-10 */
struct orc_window_s {
    ocf32*   v;
    uint32_t len;          /* n                                   */
    uint32_t n;            /* 2^msb_index(n)                      */
    uint32_t mask;
    uint32_t read_index;
    uint32_t num_allocated;
};

/* src/utility/bits.rs:112-114 : 32 - leading_zeros(x) */
static uint32_t msb_index(uint32_t x)
{
    uint32_t k = 0;
    while (x) { k++; x >>= 1; }
    return k;
}

/* src/buffer/window.rs:13-33 */
orc_window* orc_window_create(uint32_t n)
{
    if (n == 0) return NULL;
    orc_window* w = (orc_window*)malloc(sizeof(*w));
    uint32_t m = msb_index(n);
    w->len = n;
    w->n = 1u << m;
    w->mask = w->n - 1;
    w->num_allocated = w->n + n - 1;
    w->v = (ocf32*)malloc(sizeof(ocf32) * w->num_allocated);
    orc_window_reset(w);
    return w;
}

void orc_window_destroy(orc_window* w)
{
    if (!w) return;
    free(w->v);
    free(w);
}

orc_window* orc_window_clone(const orc_window* w)
{
    orc_window* c = (orc_window*)malloc(sizeof(*c));
    *c = *w;
    c->v = (ocf32*)malloc(sizeof(ocf32) * w->num_allocated);
    memcpy(c->v, w->v, sizeof(ocf32) * w->num_allocated);
    return c;
}

/* src/buffer/window.rs:61-64 */
void orc_window_reset(orc_window* w)
{
    w->read_index = 0;
    memset(w->v, 0, sizeof(ocf32) * w->num_allocated);
}

/* src/buffer/window.rs:77-85 */
void orc_window_push(orc_window* w, ocf32 value)
{
    w->read_index = (w->read_index + 1) & w->mask;
    if (w->read_index == 0)
        memmove(w->v, w->v + w->n, sizeof(ocf32) * (w->len - 1));      /* copy_within(n..n+len-1, 0) */
    w->v[w->read_index + w->len - 1] = value;
}

/* src/buffer/window.rs:66-68 */
const ocf32* orc_window_read(const orc_window* w) { return w->v + w->read_index; }
uint32_t orc_window_len(const orc_window* w) { return w->len; }
uint32_t orc_window_allocated(const orc_window* w) { return w->num_allocated; }

/* --------------------------------------------------------------- dotprod */

/* src/dotprod/mod.rs:33-39 : iter().zip().map(a*b).sum(), one f32 accumulator per
 * component, index order 0 -> n-1.  `volatile`-free but written so the compiler
 * cannot reassociate (no -ffast-math in the oracle build). */
ocf32 orc_dotprod_rcc(const float* h, const ocf32* x, size_t n)
{
    ocf32 acc = {0.0f, 0.0f};
    for (size_t i = 0; i < n; i++) {
        acc.re += h[i] * x[i].re;
        acc.im += h[i] * x[i].im;
    }
    return acc;
}

/* src/dotprod/mod.rs:19-25 */
float orc_dotprod_rrr(const float* h, const float* x, size_t n)
{
    float acc = 0.0f;
    for (size_t i = 0; i < n; i++) acc += h[i] * x[i];
    return acc;
}

/* ------------------------------------------------------------------- fft */
/*
 * src/fft/mod.rs:39-48 wraps rustfft ^6.2.0 (Cargo.toml:21; not vendored, lock
 * file git-ignored).  rustfft's published algorithm for power-of-two lengths is
 * a radix-4 decimation with precomputed f32 twiddles (computed in f64, rounded
 * once); other lengths use mixed-radix / Rader / Bluestein.  Here:
 *   - power of two  : iterative Stockham autosort, radix-4 stages (+ one radix-2
 *                     stage when log2 n is odd), f32 arithmetic, twiddles from
 *                     f64 cos/sin rounded once;
 *   - other lengths : recursive mixed radix on the smallest prime factor with a
 *                     direct DFT at prime lengths.
 * Pinned at the Fft::run boundary by the reference's 33 golden pairs
 * (src/fft/test_data.rs, tolerance 2e-4 as src/fft/mod.rs:125-151).
 */
struct orc_fft_s {
    uint32_t n;
    int      backward;
    int      pow2;
    ocf32*   tw;         /* tw[k] = exp(-+ j 2 pi k / n), k < n */
    ocf32*   scratch;    /* 2n */
};

orc_fft* orc_fft_create(uint32_t n, int backward)
{
    if (n == 0) return NULL;
    orc_fft* p = (orc_fft*)malloc(sizeof(*p));
    p->n = n;
    p->backward = backward ? 1 : 0;
    p->pow2 = (n & (n - 1)) == 0;
    p->tw = (ocf32*)malloc(sizeof(ocf32) * n);
    p->scratch = (ocf32*)malloc(sizeof(ocf32) * 2 * n);
    for (uint32_t k = 0; k < n; k++) {
        double a = 2.0 * M_PI * (double)k / (double)n;
        p->tw[k].re = (float)cos(a);
        p->tw[k].im = (float)(backward ? sin(a) : -sin(a));
    }
    return p;
}

void orc_fft_destroy(orc_fft* p)
{
    if (!p) return;
    free(p->tw);
    free(p->scratch);
    free(p);
}

static inline ocf32 cmul(ocf32 a, ocf32 b)
{
    ocf32 r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
    return r;
}
static inline ocf32 cadd(ocf32 a, ocf32 b) { ocf32 r = { a.re + b.re, a.im + b.im }; return r; }
static inline ocf32 csub(ocf32 a, ocf32 b) { ocf32 r = { a.re - b.re, a.im - b.im }; return r; }

/* Stockham autosort, power-of-two n.  x -> y ping-pong; returns pointer holding the result. */
static ocf32* fft_pow2(const orc_fft* p, ocf32* x, ocf32* y)
{
    const uint32_t n = p->n;
    const ocf32* tw = p->tw;
    const float sgn = p->backward ? 1.0f : -1.0f;      /* multiply by (sgn * j) */
    uint32_t l = n;        /* remaining sub-transform length */
    uint32_t s = 1;        /* stride */
    while (l >= 4) {
        const uint32_t q = l / 4;
        for (uint32_t j = 0; j < q; j++) {
            const ocf32 w1 = tw[(j * s) % n];
            const ocf32 w2 = tw[(2 * j * s) % n];
            const ocf32 w3 = tw[(3 * j * s) % n];
            for (uint32_t k = 0; k < s; k++) {
                const ocf32 a = x[k + s * (j + 0 * q)];
                const ocf32 b = x[k + s * (j + 1 * q)];
                const ocf32 c = x[k + s * (j + 2 * q)];
                const ocf32 d = x[k + s * (j + 3 * q)];
                const ocf32 apc = cadd(a, c), amc = csub(a, c);
                const ocf32 bpd = cadd(b, d), bmd = csub(b, d);
                const ocf32 jbmd = { -sgn * bmd.im, sgn * bmd.re };     /* (sgn*j) * (b - d) */
                y[k + s * (4 * j + 0)] = cadd(apc, bpd);
                y[k + s * (4 * j + 1)] = cmul(w1, cadd(amc, jbmd));
                y[k + s * (4 * j + 2)] = cmul(w2, csub(apc, bpd));
                y[k + s * (4 * j + 3)] = cmul(w3, csub(amc, jbmd));
            }
        }
        l = q;
        s *= 4;
        ocf32* t = x; x = y; y = t;
    }
    if (l == 2) {
        for (uint32_t k = 0; k < s; k++) {
            const ocf32 a = x[k], b = x[k + s];
            y[k] = cadd(a, b);
            y[k + s] = csub(a, b);
        }
        ocf32* t = x; x = y; y = t;
    }
    return x;
}

/* generic mixed radix (decimation in time on the smallest prime factor) */
static void fft_generic(const orc_fft* p, const ocf32* in, ocf32* out, uint32_t n, uint32_t stride, ocf32* scratch)
{
    const uint32_t N = p->n;
    if (n == 1) { out[0] = in[0]; return; }
    uint32_t f = 0;
    for (uint32_t c = 2; c * c <= n; c++) if (n % c == 0) { f = c; break; }
    if (f == 0) {
        /* prime length: direct DFT, f32 accumulation, twiddle index reduced mod n exactly */
        const uint32_t tstep = N / n;
        for (uint32_t k = 0; k < n; k++) {
            ocf32 acc = {0.0f, 0.0f};
            for (uint32_t i = 0; i < n; i++) {
                const ocf32 w = p->tw[(((uint64_t)i * k) % n) * tstep];
                acc = cadd(acc, cmul(in[(size_t)i * stride], w));
            }
            out[k] = acc;
        }
        return;
    }
    const uint32_t m = n / f;
    /* f sub-transforms of length m over the decimated inputs */
    for (uint32_t r = 0; r < f; r++)
        fft_generic(p, in + (size_t)r * stride, scratch + (size_t)r * m, m, stride * f, out);
    /* combine: X[k + m q] = sum_r W_n^{r (k + m q)} Y_r[k] */
    const uint32_t tstep = N / n;
    for (uint32_t k = 0; k < m; k++) {
        for (uint32_t q = 0; q < f; q++) {
            const uint32_t kk = k + m * q;
            ocf32 acc = scratch[k];
            for (uint32_t r = 1; r < f; r++) {
                const ocf32 w = p->tw[(((uint64_t)r * kk) % n) * tstep];
                acc = cadd(acc, cmul(scratch[(size_t)r * m + k], w));
            }
            out[kk] = acc;
        }
    }
}

/* src/fft/mod.rs:45-48 : out <- in, then in-place process */
void orc_fft_run(const orc_fft* p, const ocf32* in, ocf32* out)
{
    const uint32_t n = p->n;
    if (n == 1) { out[0] = in[0]; return; }
    if (p->pow2) {
        ocf32* a = p->scratch;
        ocf32* b = p->scratch + n;
        memcpy(a, in, sizeof(ocf32) * n);
        ocf32* r = fft_pow2(p, a, b);
        memcpy(out, r, sizeof(ocf32) * n);
    } else {
        /* the recursion needs its own scratch per level; allocate n per level lazily */
        ocf32* tmp_in = (ocf32*)malloc(sizeof(ocf32) * n);
        memcpy(tmp_in, in, sizeof(ocf32) * n);
        /* scratch usage: each level uses `n_level` entries of scratch and recurses with
         * `out` as the child's scratch; children overwrite disjoint regions. */
        ocf32* scratch = (ocf32*)malloc(sizeof(ocf32) * n);
        fft_generic(p, tmp_in, out, n, 1, scratch);
        free(scratch);
        free(tmp_in);
    }
}
