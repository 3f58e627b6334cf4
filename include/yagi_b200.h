/*
 * yagi_b200.h -- C ABI of the B200-native polyphase channelizer path.
 *
 * This is the drop-in boundary for yagi's (empty) `multichannel` module slot
 * (/root/reference/src/lib.rs:27-28) and the filter it is built on.  The reference has
 * no FFI on this path (its only FFI artefact, c_shim/src/lib.rs:1-72, is an unimplemented
 * export skeleton for bsequence), so each entry point below mirrors the *object protocol*
 * every yagi filter struct follows and cites the reference item it replaces.  The Rust
 * binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - every function returns an int32 status mirroring `enum Error`
 *     (src/error.rs:6-14): 0 OK, 1 Internal, 2 Config, 3 Value, 4 Range, 5 Mode,
 *     6 NoConvergence; yg_last_error() returns the message of the calling thread's
 *     last failure (the String the Rust side puts into the Error variant);
 *   - constructors validate eagerly like the reference's (`Err(Error::Config(..))`,
 *     e.g. src/filter/fir/firdecim.rs:38-44, src/filter/fir/firpfb.rs:35-40);
 *   - a handle is bound to the CUDA device that was current at creation, owns its device
 *     buffers and one stream, is NOT thread-safe (reference methods take `&mut self`),
 *     and distinct handles are independent;
 *   - `*_dev` entry points take DEVICE pointers and are asynchronous on the given
 *     cudaStream_t (passed as void*; NULL = the CUDA default stream, as everywhere in CUDA);
 *     the others take HOST pointers, run on the handle's own stream and return when the
 *     result is in `y`; state updates stay ordered when consecutive calls use different streams;
 *   - samples are interleaved (re, im) f32 pairs == num_complex::Complex<f32>.
 *   - there is no CPU fallback: without a CUDA device every constructor fails with
 *     Internal.
 */
#ifndef YAGI_B200_H
#define YAGI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } yg_cf32;

/* src/error.rs:6-14 */
enum { YG_OK = 0, YG_EINTERNAL = 1, YG_ECONFIG = 2, YG_EVALUE = 3, YG_ERANGE = 4, YG_EMODE = 5, YG_ENOCONV = 6 };
/* upstream LIQUID_ANALYZER / LIQUID_SYNTHESIZER; cf. the 2-variant enum idiom of
 * src/filter/resampler/msresamp2.rs:27-30 */
enum { YG_ANALYZER = 0, YG_SYNTHESIZER = 1 };

/* ------------------------------------------------------------------ library */
int32_t     yg_version(void);                       /* 0xMMmmpp */
const char* yg_last_error(void);                    /* message of this thread's last non-OK status */
int32_t     yg_device_count(int32_t* n);
/* kernels launched by this library in this process so far (every launch is counted; benchmarks take the difference
 * around their timed region) */
int32_t     yg_launch_count(uint64_t* n);
/* page-locked host buffers for the host-pointer entry points (plain malloc'd memory works too, slower) */
int32_t     yg_host_alloc(void** p, size_t bytes);
int32_t     yg_host_free(void* p);

/* fir_design_kaiser(n, fc, as_, mu) -> Vec<f32>      src/filter/fir/design/kaiser.rs:16-51
 * (host-side, cold path; same f32 formulae: math/windows.rs:76-90, math/bessel.rs:9-67,
 *  math/gamma.rs:7-22, math/mod.rs:63-69) */
int32_t yg_fir_design_kaiser(uint32_t n, float fc, float as, float mu, float* h);

/* ------------------------------------------------- firpfbch2_crcf (metric path) */
/* 2x oversampled polyphase channelizer; would be `multichannel::FirPfbCh2<Complex32, f32>`.
 * Algorithm: SURVEY.md Appendix A.1 (upstream firpfbch2, tests named at
 * LIQUID_COMPAT.md:1783-1791).  Analyzer: M/2 samples in -> M channel samples out per
 * frame; synthesizer: M in -> M/2 out. */
typedef struct yg_firpfbch2_crcf_s* yg_firpfbch2_crcf;

/* new(type, M, m, h): M >= 2 and even, m >= 1, h_len >= 2*M*m (first 2*M*m taps used).
 * Replaces: constructor idiom of src/filter/fir/firpfb.rs:34-65 (sub-filter layout :45-52). */
int32_t yg_firpfbch2_crcf_create(int32_t type, uint32_t M, uint32_t m, const float* h, size_t h_len,
                                 yg_firpfbch2_crcf* out);
/* new_kaiser(type, M, m, as_): n = 2Mm+1, fc = 1/M | 0.5/M, h <- h*M/sum(h)
 * Replaces: src/filter/fir/firpfb.rs:95-114 idiom + src/filter/resampler/resamp.rs:49-51. */
int32_t yg_firpfbch2_crcf_create_kaiser(int32_t type, uint32_t M, uint32_t m, float as,
                                        yg_firpfbch2_crcf* out);
/* #[derive(Clone)]: deep copy including stream state (the `_copy` tests, e.g. firpfb.rs:361-396) */
int32_t yg_firpfbch2_crcf_clone(yg_firpfbch2_crcf q, yg_firpfbch2_crcf* out);
int32_t yg_firpfbch2_crcf_destroy(yg_firpfbch2_crcf q);                     /* Drop */
int32_t yg_firpfbch2_crcf_reset(yg_firpfbch2_crcf q);                       /* zero history, flag = 0 */
/* execute(&mut self, x, y): exactly one frame, host pointers (cf. firdecim.rs:179-191) */
int32_t yg_firpfbch2_crcf_execute(yg_firpfbch2_crcf q, const yg_cf32* x, yg_cf32* y);
/* execute_block(&mut self, x, n, y): n_frames consecutive frames (cf. firdecim.rs:193-205) */
int32_t yg_firpfbch2_crcf_execute_block(yg_firpfbch2_crcf q, const yg_cf32* x, size_t n_frames, yg_cf32* y);
/* Same on device pointers, asynchronous on `cuda_stream` (NULL = the CUDA default stream).  Any 8-byte aligned
 * pointers work; the fused kernels (last_path 2 / 3) additionally want d_x and d_y 16-byte aligned (every
 * cudaMalloc pointer plus a whole number of frames is) and otherwise hand the call to the generic kernels. */
int32_t yg_firpfbch2_crcf_execute_block_dev(yg_firpfbch2_crcf q, const yg_cf32* d_x, size_t n_frames,
                                            yg_cf32* d_y, void* cuda_stream);
int32_t yg_firpfbch2_crcf_sync(yg_firpfbch2_crcf q);                        /* wait for all work queued on this handle */
int32_t yg_firpfbch2_crcf_get_type(yg_firpfbch2_crcf q, int32_t* type);
int32_t yg_firpfbch2_crcf_get_M(yg_firpfbch2_crcf q, uint32_t* M);
int32_t yg_firpfbch2_crcf_get_m(yg_firpfbch2_crcf q, uint32_t* m);
int32_t yg_firpfbch2_crcf_get_taps(yg_firpfbch2_crcf q, float* h /* 2*M*m */);
int32_t yg_firpfbch2_crcf_get_device(yg_firpfbch2_crcf q, int32_t* dev);   /* the CUDA device the handle is bound to */
/* Stream state as plain data (what Clone copies; also the time-shard hand-off).
 * Both types keep the tail of their INPUT stream, oldest first: analyzer the last (4m-1)*M/2
 * input samples, synthesizer the last 4m-1 input frames (M each; their IFFTs are recomputed, so the
 * state stays independent of the kernel used).  `flag` = frame parity since reset. */
int32_t yg_firpfbch2_crcf_state_len(yg_firpfbch2_crcf q, size_t* n_cf32);
int32_t yg_firpfbch2_crcf_get_state(yg_firpfbch2_crcf q, yg_cf32* hist, int32_t* flag);
int32_t yg_firpfbch2_crcf_set_state(yg_firpfbch2_crcf q, const yg_cf32* hist, int32_t flag);
/* which kernel the last execute_block* used: 0 none, 1 generic, 2 fused (M = 8 .. 256), 3 large-M path (M = 512 .. 4096) */
int32_t yg_firpfbch2_crcf_last_path(yg_firpfbch2_crcf q, int32_t* path);
/* device time of the dominant kernel of the last execute_block_dev call, in ms (CUDA events
 * recorded on the launching stream around that kernel only); syncs the stream */
int32_t yg_firpfbch2_crcf_last_kernel_ms(yg_firpfbch2_crcf q, float* ms);
/* the same for up to `cap` most recent calls (at most 64 are kept), oldest first; *n = how many */
int32_t yg_firpfbch2_crcf_kernel_times(yg_firpfbch2_crcf q, float* ms, size_t cap, size_t* n);
/* The two timing events per call cost about a microsecond of stream time each; enable = 0 stops recording them
 * (last_kernel_ms / kernel_times then report Mode errors / nothing), enable != 0 (the default) turns them back on. */
int32_t yg_firpfbch2_crcf_set_kernel_timing(yg_firpfbch2_crcf q, int32_t enable);

/* ------------------------------------------------------------- firpfbch_crcf */
/* Critically sampled channelizer (LIQUID_COMPAT.md:1765-1780),
 * batched over n_streams independent streams that share the taps (one reference object per
 * stream).  Layout: x[stream][frame][M], y likewise.  M >= 1, p >= 1, h_len >= M*p. */
typedef struct yg_firpfbch_crcf_s* yg_firpfbch_crcf;
int32_t yg_firpfbch_crcf_create(int32_t type, uint32_t M, uint32_t p, const float* h, size_t h_len,
                                uint32_t n_streams, yg_firpfbch_crcf* out);
int32_t yg_firpfbch_crcf_create_kaiser(int32_t type, uint32_t M, uint32_t m, float as,
                                       uint32_t n_streams, yg_firpfbch_crcf* out);   /* p = 2m, fc = 0.5/M */
int32_t yg_firpfbch_crcf_clone(yg_firpfbch_crcf q, yg_firpfbch_crcf* out);
int32_t yg_firpfbch_crcf_destroy(yg_firpfbch_crcf q);
int32_t yg_firpfbch_crcf_reset(yg_firpfbch_crcf q);
int32_t yg_firpfbch_crcf_execute(yg_firpfbch_crcf q, const yg_cf32* x, yg_cf32* y);  /* one frame per stream */
int32_t yg_firpfbch_crcf_execute_block(yg_firpfbch_crcf q, const yg_cf32* x, size_t n_frames, yg_cf32* y);
int32_t yg_firpfbch_crcf_execute_block_dev(yg_firpfbch_crcf q, const yg_cf32* d_x, size_t n_frames,
                                           yg_cf32* d_y, void* cuda_stream);
int32_t yg_firpfbch_crcf_sync(yg_firpfbch_crcf q);
int32_t yg_firpfbch_crcf_get_type(yg_firpfbch_crcf q, int32_t* type);
int32_t yg_firpfbch_crcf_get_M(yg_firpfbch_crcf q, uint32_t* M);
int32_t yg_firpfbch_crcf_get_p(yg_firpfbch_crcf q, uint32_t* p);
int32_t yg_firpfbch_crcf_get_n_streams(yg_firpfbch_crcf q, uint32_t* n);
int32_t yg_firpfbch_crcf_get_taps(yg_firpfbch_crcf q, float* h /* M*p */);
int32_t yg_firpfbch_crcf_get_device(yg_firpfbch_crcf q, int32_t* dev);
/* which kernel the last execute_block* used for the bulk of the streams: 0 none, 1 generic, 2 fused (M = 8, 16, 32, 64) */
int32_t yg_firpfbch_crcf_last_path(yg_firpfbch_crcf q, int32_t* path);

/* ------------------------------------------------------------- output layout */
/* Frame-major analysis output y[frame][M] -> channel-major out[channel][frame] (one contiguous time series per
 * channel), device pointers, out of place, asynchronous on `cuda_stream`.  Not on the hot path: the step after it in
 * an SDR pipeline. */
int32_t yg_channel_major_dev(const yg_cf32* d_frames, size_t n_frames, uint32_t M, yg_cf32* d_out, void* cuda_stream);

/* -------------------------------------------------------------- firfilt_crcf */
/* Direct-form FIR, real taps x complex samples: FirFilter<Complex32, f32>
 * (src/filter/fir/firfilt.rs:63-79 new, :220-223 push, :241-245 execute, :267-278
 * execute_block, set_scale/get_scale), batched over n_streams independent streams sharing
 * the taps.  y[s][n] = scale * sum_k h[k] x[s][n-k], zero state after reset.  Layout
 * x[stream][n]. */
typedef struct yg_firfilt_crcf_s* yg_firfilt_crcf;
int32_t yg_firfilt_crcf_create(const float* h, size_t h_len, uint32_t n_streams, yg_firfilt_crcf* out);
int32_t yg_firfilt_crcf_create_kaiser(uint32_t n, float fc, float as, float mu, uint32_t n_streams,
                                      yg_firfilt_crcf* out);                 /* firfilt.rs new_kaiser */
int32_t yg_firfilt_crcf_clone(yg_firfilt_crcf q, yg_firfilt_crcf* out);
int32_t yg_firfilt_crcf_destroy(yg_firfilt_crcf q);
int32_t yg_firfilt_crcf_reset(yg_firfilt_crcf q);
int32_t yg_firfilt_crcf_set_scale(yg_firfilt_crcf q, float scale);
int32_t yg_firfilt_crcf_get_scale(yg_firfilt_crcf q, float* scale);
int32_t yg_firfilt_crcf_get_len(yg_firfilt_crcf q, size_t* h_len);
int32_t yg_firfilt_crcf_get_device(yg_firfilt_crcf q, int32_t* dev);
/* which kernel the last execute_block* used: 0 none, 1 generic, 2 register-blocked FFMA2 (<= 256 taps),
 * 4 tensor cores (tcgen05 3xTF32 Toeplitz GEMM, <= 161 taps; the FFMA2 kernel finishes a ragged tail) */
int32_t yg_firfilt_crcf_last_path(yg_firfilt_crcf q, int32_t* path);
int32_t yg_firfilt_crcf_execute_block(yg_firfilt_crcf q, const yg_cf32* x, size_t n, yg_cf32* y);
int32_t yg_firfilt_crcf_execute_block_dev(yg_firfilt_crcf q, const yg_cf32* d_x, size_t n, yg_cf32* d_y,
                                          void* cuda_stream);
int32_t yg_firfilt_crcf_sync(yg_firfilt_crcf q);

#ifdef __cplusplus
}
#endif
#endif /* YAGI_B200_H */
