/*
 * yagi_oracle.h -- CPU restatement of yagi's polyphase-channelizer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, link or call it, and there only as the
 * checker or the timed CPU baseline.  The shipped library (libyagi_b200.so)
 * never links or calls this code.
 *
 * Parity status
 *   - Building blocks (Window, dotprod, Fft::run, FirFilter, FirDecim,
 *     FirPfbFilter sub-filter layout) are PINNED against the reference's own
 *     golden vectors (tests/test_oracle_golden.py).
 *   - The channelizer objects themselves (firpfbch2 / firpfbch) are PARITY
 *     UNPINNED against the reference: /root/reference/src/multichannel/mod.rs
 *     is a 0-byte file and yagi cannot be compiled here (no Rust toolchain).
 *     They are restated from the upstream liquid-dsp algorithm (SURVEY.md
 *     Appendix A) in yagi's idiom and accepted by the self-checking
 *     properties the upstream autotests assert (tests/test_oracle_properties.py).
 *
 * All citations are relative to /root/reference/.
 */
#ifndef YAGI_ORACLE_H
#define YAGI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } ocf32;          /* num_complex::Complex<f32> layout */

/* status codes mirror src/error.rs:6-14 */
enum { ORC_OK = 0, ORC_EINTERNAL = 1, ORC_ECONFIG = 2, ORC_EVALUE = 3,
       ORC_ERANGE = 4, ORC_EMODE = 5, ORC_ENOCONV = 6 };
enum { ORC_ANALYZER = 0, ORC_SYNTHESIZER = 1 };

/* ---- L0: math / design (src/math, src/filter/fir/design/kaiser.rs) ---- */
float orc_lngammaf(float z);                       /* src/math/gamma.rs:7-22 */
float orc_lnbesselif(float nu, float z);           /* src/math/bessel.rs:9-41 */
float orc_besseli0f(float z);                      /* src/math/bessel.rs:44-67 */
float orc_sincf(float x);                          /* src/math/mod.rs:63-69 */
int   orc_kaiser(uint32_t i, uint32_t wlen, float beta, float* out);   /* src/math/windows.rs:76-90 */
float orc_kaiser_beta_as(float as);                /* src/filter/fir/design/kaiser.rs:62-72 */
int   orc_fir_design_kaiser(uint32_t n, float fc, float as, float mu, float* h); /* kaiser.rs:16-51 */

/* ---- L1: window, dotprod, fft ---- */
typedef struct orc_window_s orc_window;            /* src/buffer/window.rs:3-10 */
orc_window* orc_window_create(uint32_t n);         /* :13-33  */
void  orc_window_destroy(orc_window* w);
orc_window* orc_window_clone(const orc_window* w);
void  orc_window_reset(orc_window* w);             /* :61-64  */
void  orc_window_push(orc_window* w, ocf32 v);     /* :77-85  */
const ocf32* orc_window_read(const orc_window* w); /* :66-68, oldest first */
uint32_t orc_window_len(const orc_window* w);
uint32_t orc_window_allocated(const orc_window* w);

ocf32 orc_dotprod_rcc(const float* h, const ocf32* x, size_t n);  /* src/dotprod/mod.rs:33-39 */
float orc_dotprod_rrr(const float* h, const float* x, size_t n);  /* src/dotprod/mod.rs:19-25 */

/* Fft::run, src/fft/mod.rs:39-48: forward = e^{-j..}, backward = e^{+j..}, both unnormalised */
typedef struct orc_fft_s orc_fft;
orc_fft* orc_fft_create(uint32_t n, int backward);
void  orc_fft_destroy(orc_fft* p);
void  orc_fft_run(const orc_fft* p, const ocf32* in, ocf32* out);

/* ---- L2: filter objects used to pin the building blocks ---- */
typedef struct orc_firfilt_s orc_firfilt;          /* src/filter/fir/firfilt.rs */
int   orc_firfilt_crcf_create(const float* h, size_t h_len, orc_firfilt** out);   /* :63-79 */
void  orc_firfilt_crcf_destroy(orc_firfilt* q);
void  orc_firfilt_crcf_reset(orc_firfilt* q);
void  orc_firfilt_crcf_set_scale(orc_firfilt* q, float scale);
void  orc_firfilt_crcf_push(orc_firfilt* q, ocf32 x);                              /* :220-223 */
ocf32 orc_firfilt_crcf_execute(const orc_firfilt* q);                              /* :241-245 */
int   orc_firfilt_crcf_execute_block(orc_firfilt* q, const ocf32* x, size_t n, ocf32* y); /* :267-278 */

typedef struct orc_firdecim_s orc_firdecim;        /* src/filter/fir/firdecim.rs */
int   orc_firdecim_crcf_create(uint32_t M, const float* h, size_t h_len, orc_firdecim** out); /* :38-57 */
void  orc_firdecim_crcf_destroy(orc_firdecim* q);
ocf32 orc_firdecim_crcf_execute(orc_firdecim* q, const ocf32* x);                  /* :179-191 */

typedef struct orc_firpfb_s orc_firpfb;            /* src/filter/fir/firpfb.rs (real-valued rrrf, as its golden test) */
int   orc_firpfb_rrrf_create(uint32_t num_filters, const float* h, size_t h_len, orc_firpfb** out); /* :34-65 */
void  orc_firpfb_rrrf_destroy(orc_firpfb* q);
void  orc_firpfb_rrrf_push(orc_firpfb* q, float x);                                /* :255-257 */
int   orc_firpfb_rrrf_execute(orc_firpfb* q, uint32_t i, float* y);                /* :277-286 */

/* ---- L3: the channelizers ---- */
typedef struct orc_firpfbch2_s orc_firpfbch2;
int   orc_firpfbch2_crcf_create(int type, uint32_t M, uint32_t m, const float* h, size_t h_len, orc_firpfbch2** out);
int   orc_firpfbch2_crcf_create_kaiser(int type, uint32_t M, uint32_t m, float as, orc_firpfbch2** out);
int   orc_firpfbch2_crcf_clone(const orc_firpfbch2* q, orc_firpfbch2** out);
void  orc_firpfbch2_crcf_destroy(orc_firpfbch2* q);
void  orc_firpfbch2_crcf_reset(orc_firpfbch2* q);
int   orc_firpfbch2_crcf_execute(orc_firpfbch2* q, const ocf32* x, ocf32* y);      /* one frame */
int   orc_firpfbch2_crcf_execute_block(orc_firpfbch2* q, const ocf32* x, size_t n_frames, ocf32* y);
int   orc_firpfbch2_crcf_get_type(const orc_firpfbch2* q);
uint32_t orc_firpfbch2_crcf_get_M(const orc_firpfbch2* q);
uint32_t orc_firpfbch2_crcf_get_m(const orc_firpfbch2* q);
const float* orc_firpfbch2_crcf_taps(const orc_firpfbch2* q, size_t* len);         /* prototype as given */

typedef struct orc_firpfbch_s orc_firpfbch;
int   orc_firpfbch_crcf_create(int type, uint32_t M, uint32_t p, const float* h, size_t h_len, orc_firpfbch** out);
int   orc_firpfbch_crcf_create_kaiser(int type, uint32_t M, uint32_t m, float as, orc_firpfbch** out);
int   orc_firpfbch_crcf_clone(const orc_firpfbch* q, orc_firpfbch** out);
void  orc_firpfbch_crcf_destroy(orc_firpfbch* q);
void  orc_firpfbch_crcf_reset(orc_firpfbch* q);
int   orc_firpfbch_crcf_execute(orc_firpfbch* q, const ocf32* x, ocf32* y);        /* M in -> M out */
int   orc_firpfbch_crcf_execute_block(orc_firpfbch* q, const ocf32* x, size_t n_frames, ocf32* y);
const float* orc_firpfbch_crcf_taps(const orc_firpfbch* q, size_t* len);

/* ---- CPU baseline runner (bench.py cpu_baseline / --impl reference) ----
 * One firpfbch2 analyser object per thread, each over its own slice of
 * n_per_thread input samples (pre-generated by the caller, x[t*n_per_thread..]).
 * Returns wall seconds of the slowest thread for `passes` passes, or <0 on error. */
double orc_bench_firpfbch2_analysis(uint32_t M, uint32_t m, float as,
                                    const ocf32* x, size_t n_per_thread,
                                    uint32_t n_threads, uint32_t passes,
                                    ocf32* y /* n_threads * 2*n_per_thread, may be NULL */);

#ifdef __cplusplus
}
#endif
#endif
