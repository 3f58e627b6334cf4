// common.cuh -- shared host/device helpers for libyagi_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/yagi_b200.h"

namespace yg {

// ---------------------------------------------------------------- errors
std::string& last_error_ref();
int32_t fail(int32_t code, const char* fmt, ...);

#define YG_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::yg::fail(YG_EINTERNAL, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), \
                              __FILE__, __LINE__, #expr);                                      \
    } while (0)

#define YG_TRY(expr)                 \
    do {                             \
        int32_t _rc = (expr);        \
        if (_rc != YG_OK) return _rc;\
    } while (0)

// Binds the calling thread to a handle's device for the duration of a call.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        cur = dev;
    }
    ~DeviceGuard() { if (prev >= 0 && prev != cur) cudaSetDevice(prev); }
    int cur = -1;
};

#define YG_DEVICE_GUARD(dev)                                                                  \
    ::yg::DeviceGuard _yg_guard(dev);                                                         \
    if (!_yg_guard.ok) return ::yg::fail(YG_EINTERNAL, "cannot switch to CUDA device %d", (int)(dev))

int32_t require_device(int* dev_out);
int sm_count(int dev);                  // multiprocessor count (>= 1)

// Every kernel launch of the library goes through YG_LAUNCH_CHECK() (or count_launch() after a
// cooperative launch), so callers can ask how many kernels a region launched (yg_launch_count).
void count_launch();
unsigned long long launch_count();
#define YG_LAUNCH_CHECK()                  \
    do {                                   \
        ::yg::count_launch();              \
        YG_CUDA(cudaGetLastError());       \
    } while (0)

// Host<->device copies and fills that have COMPLETED on the device when they return, whatever
// stream the next kernel runs on.  (The plain cudaMemset / device-to-device cudaMemcpy are
// asynchronous on the legacy default stream, which the handles' non-blocking streams do not
// synchronise with.)
cudaError_t memcpy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
cudaError_t memset_sync(void* dst, int value, size_t bytes);

// ---------------------------------------------------------------- design (host)
int32_t fir_design_kaiser(uint32_t n, float fc, float as, float mu, float* h);
// twiddle table tw[k] = exp(+j 2 pi k / M), computed in f64 and rounded once
void make_twiddles(uint32_t M, std::vector<float2>& tw);

// Radices of the mixed-radix transform of length M, in pass order: 4 before 2, then the odd prime factors ascending.
struct TiledPass { unsigned char radix[24]; int n_pass; };
// false when M has a prime factor > 255 (or more than 24 factors): the tiled kernels do not take such an M
bool plan_radices(uint32_t M, TiledPass& tp);

// ---------------------------------------------------------------- device buffers
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int32_t reserve(size_t want) {
        if (want <= n) return YG_OK;
        if (p) { cudaFree(p); p = nullptr; n = 0; }
        YG_CUDA(cudaMalloc(&p, want * sizeof(T)));
        n = want;
        return YG_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// Keeps a handle's state updates ordered when consecutive calls use different streams
// (e.g. a device-pointer call on the caller's stream followed by a host-pointer call).
// After a call on a stream that outlives the handle's use of it (the default streams, the handle's own stream:
// `own`) nothing is recorded -- an event record between two launches would keep them from overlapping under
// programmatic dependent launch -- and the event is recorded on that stream later, if and when a call arrives on a
// different one.  After a call on a caller-created stream, which may be destroyed at any time, it is recorded at once.
struct StreamOrder {
    cudaEvent_t ev = nullptr;
    cudaStream_t last = nullptr;
    cudaStream_t own = nullptr;
    bool armed = false;
    bool pending = false;          // work on `last` since the last record
    bool permanent(cudaStream_t st) const
    {
        return st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread || st == own;
    }
    int32_t flush()
    {
        if (pending) {
            if (!ev) YG_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            YG_CUDA(cudaEventRecord(ev, last));
            pending = false;
        }
        return YG_OK;
    }
    int32_t enter(cudaStream_t st)
    {
        if (armed && st != last) {
            YG_TRY(flush());
            YG_CUDA(cudaStreamWaitEvent(st, ev, 0));
        }
        return YG_OK;
    }
    int32_t leave(cudaStream_t st)
    {
        last = st;
        armed = true;
        pending = true;
        if (!permanent(st)) YG_TRY(flush());
        return YG_OK;
    }
    int32_t wait_host()
    {
        if (armed) {
            YG_TRY(flush());
            YG_CUDA(cudaEventSynchronize(ev));
        }
        return YG_OK;
    }
    void destroy() { if (ev) cudaEventDestroy(ev); ev = nullptr; armed = false; pending = false; }
};

// Three-stream chunked host<->device pipeline used by the host-pointer entry points:
// H2D of chunk c+1 and D2H of chunk c-1 overlap the kernels of chunk c.
struct HostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    DevBuf<yg_cf32> dx[2], dy[2];
    bool inited = false;
    int32_t init();
    void destroy();
};

// launch(d_x, frames, d_y, stream) must enqueue everything for `frames` frames on `stream`.
template <typename Launch>
int32_t run_host_pipe(HostPipe& hp, cudaStream_t s_comp, const yg_cf32* x, yg_cf32* y, size_t n_frames,
                      size_t in_per_frame, size_t out_per_frame, size_t frames_per_chunk, Launch&& launch)
{
    YG_TRY(hp.init());
    if (n_frames == 0) return YG_OK;
    // On a mid-loop failure the copies already queued still target the caller's x and y: drain the three
    // streams before reporting it, so the caller may free its buffers as soon as the call returns.
    struct Drain {
        HostPipe& hp; cudaStream_t s; bool armed = true;
        ~Drain() { if (armed) { cudaStreamSynchronize(hp.s_in); cudaStreamSynchronize(s); cudaStreamSynchronize(hp.s_out); } }
    } drain{hp, s_comp};
    if (frames_per_chunk == 0) frames_per_chunk = 1;
    if (frames_per_chunk > n_frames) frames_per_chunk = n_frames;
    const int nbuf = (n_frames > frames_per_chunk) ? 2 : 1;
    for (int b = 0; b < nbuf; b++) {
        YG_TRY(hp.dx[b].reserve(frames_per_chunk * in_per_frame));
        YG_TRY(hp.dy[b].reserve(frames_per_chunk * out_per_frame));
    }
    size_t done = 0;
    for (size_t c = 0; done < n_frames; c++) {
        const int b = (int)(c & 1);
        const size_t f = (n_frames - done < frames_per_chunk) ? (n_frames - done) : frames_per_chunk;
        if (c >= 2) YG_CUDA(cudaStreamWaitEvent(hp.s_in, hp.ev_comp[b], 0));      // dx[b] free
        YG_CUDA(cudaMemcpyAsync(hp.dx[b].p, x + done * in_per_frame, f * in_per_frame * sizeof(yg_cf32),
                                cudaMemcpyHostToDevice, hp.s_in));
        YG_CUDA(cudaEventRecord(hp.ev_in[b], hp.s_in));
        YG_CUDA(cudaStreamWaitEvent(s_comp, hp.ev_in[b], 0));
        if (c >= 2) YG_CUDA(cudaStreamWaitEvent(s_comp, hp.ev_out[b], 0));       // dy[b] free
        YG_TRY(launch(hp.dx[b].p, f, hp.dy[b].p, s_comp));
        YG_CUDA(cudaEventRecord(hp.ev_comp[b], s_comp));
        YG_CUDA(cudaStreamWaitEvent(hp.s_out, hp.ev_comp[b], 0));
        YG_CUDA(cudaMemcpyAsync(y + done * out_per_frame, hp.dy[b].p, f * out_per_frame * sizeof(yg_cf32),
                                cudaMemcpyDeviceToHost, hp.s_out));
        YG_CUDA(cudaEventRecord(hp.ev_out[b], hp.s_out));
        done += f;
    }
    YG_CUDA(cudaStreamSynchronize(hp.s_out));
    YG_CUDA(cudaStreamSynchronize(s_comp));
    drain.armed = false;
    return YG_OK;
}

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// Unnormalised M-point DFT of X (shared memory) by one thread block.
//   backward != 0 : e^{+j 2 pi n k / M}   (Fft Direction::Backward, src/fft/mod.rs:13-26)
//   backward == 0 : e^{-j 2 pi n k / M}
// tw[k] = e^{+j 2 pi k / M} (global, read-only).  `Y` is a second M-entry shared scratch.
// Returns the shared buffer holding the result.  All threads of the block must call it;
// on return the result is visible to all threads (ends with __syncthreads()).
__device__ inline float2* block_dft(float2* X, float2* Y, uint32_t M, const float2* __restrict__ tw, int backward)
{
    __syncthreads();
    if ((M & (M - 1)) == 0) {
        // Stockham autosort: radix-4 passes, plus one radix-2 pass when log2(M) is odd
        uint32_t l = M, s = 1;                       // remaining sub-transform length, stride
        float2* x = X;
        float2* y = Y;
        const float sgn = backward ? 1.0f : -1.0f;   // multiply by (sgn * j)
        while (l >= 4) {
            const uint32_t q = l >> 2;
            for (uint32_t idx = threadIdx.x; idx < (M >> 2); idx += blockDim.x) {
                const uint32_t j = idx / s, k = idx - j * s;
                float2 w1 = __ldg(&tw[j * s]);
                float2 w2 = __ldg(&tw[2 * j * s]);
                float2 w3 = __ldg(&tw[3 * j * s]);
                if (!backward) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
                const float2 a = x[k + s * j];
                const float2 b = x[k + s * (j + q)];
                const float2 c = x[k + s * (j + 2 * q)];
                const float2 d = x[k + s * (j + 3 * q)];
                const float2 apc = cadd(a, c), amc = csub(a, c);
                const float2 bpd = cadd(b, d), bmd = csub(b, d);
                const float2 jbmd = make_float2(-sgn * bmd.y, sgn * bmd.x);
                y[k + s * (4 * j + 0)] = cadd(apc, bpd);
                y[k + s * (4 * j + 1)] = cmul(cadd(amc, jbmd), w1);
                y[k + s * (4 * j + 2)] = cmul(csub(apc, bpd), w2);
                y[k + s * (4 * j + 3)] = cmul(csub(amc, jbmd), w3);
            }
            __syncthreads();
            float2* t = x; x = y; y = t;
            l = q;
            s <<= 2;
        }
        if (l == 2) {
            for (uint32_t k = threadIdx.x; k < s; k += blockDim.x) {
                const float2 a = x[k], b = x[k + s];
                y[k] = cadd(a, b);
                y[k + s] = csub(a, b);
            }
            __syncthreads();
            float2* t = x; x = y; y = t;
        }
        return x;
    }
    // Any other M: mixed-radix Stockham autosort, one pass per prime factor (4 before 2, then odd primes ascending).
    // A pass of radix r maps butterfly j = jh Ns + k (Ns = product of the radices done, k < Ns) from x[j + i M/r] to
    // y[(jh r + q) Ns + k]; every thread computes single outputs, r complex MACs each, with the twiddle of the pass and
    // the r-point DFT folded into ONE table exponent i (k + q Ns) M / (Ns r), reduced mod M exactly in integers.
    // Work per frame: M * (sum of the radices) MACs instead of the M^2 of a direct DFT (M = 1000: 21 vs 1000 per bin).
    uint32_t rem = M, Ns = 1;
    float2* x = X;
    float2* y = Y;
    while (rem > 1) {
        uint32_t r = rem;
        if ((rem & 3) == 0) r = 4;
        else if ((rem & 1) == 0) r = 2;
        else
            for (uint32_t p = 3; p * p <= rem; p += 2)
                if (rem % p == 0) { r = p; break; }
        const uint32_t L = M / r, step = M / (Ns * r);
        for (uint32_t o = threadIdx.x; o < M; o += blockDim.x) {
            const uint32_t k = o % Ns, t = o / Ns, q = t % r, jh = t / r;
            const float2* xi = x + jh * Ns + k;
            const uint32_t e = (k + q * Ns) * step;          // < M
            float2 acc = xi[0];
            uint32_t idx = 0;
            for (uint32_t i = 1; i < r; i++) {
                idx += e;
                if (idx >= M) idx -= M;
                float2 w = __ldg(&tw[idx]);
                if (!backward) w.y = -w.y;
                acc = cadd(acc, cmul(xi[i * L], w));
            }
            y[o] = acc;
        }
        __syncthreads();
        float2* t2 = x; x = y; y = t2;
        Ns *= r;
        rem /= r;
    }
    return x;
}

// The same mixed-radix passes for F independent frames side by side, one output bin per thread and pass:
// x / y = this thread's frame slot (M entries each, shared memory), T = e^{+j 2 pi k / M} in shared memory, o = the
// thread's bin.  Backward transform.  Every thread of the block must call it (barriers inside; threads without a
// frame pass valid = false); returns the row that holds the result.  M = 48: 11 complex MACs per bin instead of 48.
__device__ inline float2* slot_dft(float2* x, float2* y, uint32_t M, const float2* T, uint32_t o, bool valid)
{
    // a pass costs its radix in MACs plus ~4 MACs' worth of index arithmetic and a barrier: below that (M <= 12 or so)
    // the direct DFT (= one pass of radix M) is cheaper
    uint32_t cost = 0;
    for (uint32_t rem = M; rem > 1;) {
        uint32_t r = rem;
        if ((rem & 3) == 0) r = 4;
        else if ((rem & 1) == 0) r = 2;
        else
            for (uint32_t p = 3; p * p <= rem; p += 2)
                if (rem % p == 0) { r = p; break; }
        cost += r + 4;
        rem /= r;
    }
    const bool direct = M <= cost;
    uint32_t rem = M, Ns = 1;
    while (rem > 1) {
        uint32_t r = rem;
        if (direct) r = rem;
        else if ((rem & 3) == 0) r = 4;
        else if ((rem & 1) == 0) r = 2;
        else
            for (uint32_t p = 3; p * p <= rem; p += 2)
                if (rem % p == 0) { r = p; break; }
        if (valid) {
            const uint32_t L = M / r, step = M / (Ns * r);
            const uint32_t k = o % Ns, t = o / Ns, q = t % r, jh = t / r;
            const float2* xi = x + jh * Ns + k;
            const uint32_t e = (k + q * Ns) * step;
            float2 acc = xi[0];
            uint32_t idx = 0;
            for (uint32_t i = 1; i < r; i++) {
                idx += e;
                if (idx >= M) idx -= M;
                acc = cadd(acc, cmul(xi[i * L], T[idx]));
            }
            y[o] = acc;
        }
        __syncthreads();
        float2* t2 = x; x = y; y = t2;
        Ns *= r;
        rem /= r;
    }
    return x;
}

// ------------------------------------------------------------------ tiled generic kernels: mixed-radix passes over F frames
// (firpfbch2.cu, firpfbch.cu) -- F frames of M points side by side in shared memory (in, out: ping-pong buffers), the
// radices chosen on the host, radix-2/3/4/5 butterflies in registers, any other prime factor one output per thread.

__device__ __forceinline__ float2 cmulj(float2 a) { return make_float2(-a.y, a.x); }        // j a

template <int R>
__device__ __forceinline__ void dft_small(float2 (&v)[R])                                     // backward, in place
{
    if (R == 2) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    } else if (R == 3) {
        constexpr float s3 = 0.86602540378443865f;
        const float2 t1 = cadd(v[1], v[2]);
        const float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
        const float2 d = csub(v[1], v[2]);
        const float2 t3 = make_float2(-s3 * d.y, s3 * d.x);                                   // j s3 (v1 - v2)
        v[0] = cadd(v[0], t1); v[1] = cadd(t2, t3); v[2] = csub(t2, t3);
    } else if (R == 4) {
        const float2 apc = cadd(v[0], v[2]), amc = csub(v[0], v[2]);
        const float2 bpd = cadd(v[1], v[3]), jbmd = cmulj(csub(v[1], v[3]));
        v[0] = cadd(apc, bpd); v[1] = cadd(amc, jbmd); v[2] = csub(apc, bpd); v[3] = csub(amc, jbmd);
    } else if (R == 5) {
        constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
        constexpr float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
        const float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
        const float2 a1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
        const float2 a2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
        const float2 b1 = cmulj(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
        const float2 b2 = cmulj(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
        v[0] = cadd(v[0], cadd(t1, t2));
        v[1] = cadd(a1, b1); v[4] = csub(a1, b1); v[2] = cadd(a2, b2); v[3] = csub(a2, b2);
    }
}

template <int R>
__device__ __forceinline__ void tiled_pass(const float2* in, float2* out, const float2* T, uint32_t M, uint32_t Ns, uint32_t nf)
{
    const uint32_t L = M / R, step = M / (Ns * R);
    const uint32_t items = nf * L;
    for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
        const uint32_t fl = it / L, j = it - fl * L;
        const uint32_t jh = j / Ns, k = j - jh * Ns;
        const float2* xi = in + fl * M + j;
        float2 v[R];
        v[0] = xi[0];
#pragma unroll
        for (int i = 1; i < R; i++) v[i] = cmul(xi[i * L], T[i * k * step]);
        dft_small<R>(v);
        float2* yo = out + fl * M + jh * R * Ns + k;
#pragma unroll
        for (int q = 0; q < R; q++) yo[q * Ns] = v[q];
    }
}

// any other (prime) radix: one output per thread, r MACs each (block_dft's form)
__device__ __forceinline__ void tiled_pass_any(const float2* in, float2* out, const float2* T, uint32_t M, uint32_t Ns, uint32_t r,
                                               uint32_t nf)
{
    const uint32_t L = M / r, step = M / (Ns * r);
    const uint32_t items = nf * M;
    for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
        const uint32_t fl = it / M, o = it - fl * M;
        const uint32_t t = o / Ns, k = o - t * Ns, jh = t / r, q = t - jh * r;
        const float2* xi = in + fl * M + jh * Ns + k;
        const uint32_t e = (k + q * Ns) * step;
        float2 acc = xi[0];
        uint32_t idx = 0;
        for (uint32_t i = 1; i < r; i++) {
            idx += e;
            if (idx >= M) idx -= M;
            acc = cadd(acc, cmul(xi[i * L], T[idx]));
        }
        out[it] = acc;
    }
}

// all passes over nf frames held in A (result: returned pointer, A or B); barriers inside, every thread must call it
__device__ __forceinline__ float2* tiled_dft(float2* A, float2* B, const float2* T, uint32_t M, const TiledPass& tp, uint32_t nf)
{
    uint32_t Ns = 1;
    for (int p = 0; p < tp.n_pass; p++) {
        const uint32_t r = tp.radix[p];
        switch (r) {
            case 2: tiled_pass<2>(A, B, T, M, Ns, nf); break;
            case 3: tiled_pass<3>(A, B, T, M, Ns, nf); break;
            case 4: tiled_pass<4>(A, B, T, M, Ns, nf); break;
            case 5: tiled_pass<5>(A, B, T, M, Ns, nf); break;
            default: tiled_pass_any(A, B, T, M, Ns, r, nf); break;
        }
        __syncthreads();
        float2* t = A; A = B; B = t;
        Ns *= r;
    }
    return A;
}

#endif  // __CUDACC__

}  // namespace yg
