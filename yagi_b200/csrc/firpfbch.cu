// firpfbch.cu -- firpfbch_crcf (critically sampled channelizer), batched over independent streams.
//
// Closed forms (frame q since reset, s[t<0] = 0):
//   analysis : V_q[b] = sum_n h[b+nM] s[qM + M-1 - b - nM];  X[M-1-b] = V_q[b];  y_q = DFT_forward(X)
//   synthesis: U_q = IDFT_unnorm(X_q);  y[qM + i] = sum_n h[i+nM] U_{q-n}[i]
#include "common.cuh"
#include "firpfbch_fast.cuh"

#include <algorithm>

using namespace yg;

struct yg_firpfbch_crcf_s {
    int32_t type = 0;
    uint32_t M = 0, p = 0, n_streams = 1;
    size_t L = 0;                  // M*p
    int dev = 0;
    int n_sm = 1;                  // multiprocessor count of `dev` (grid sizing)
    cudaStream_t stream = nullptr;
    StreamOrder order;
    std::vector<float> h;
    DevBuf<float> d_h;
    DevBuf<float2> d_tw;
    size_t state_len = 0;          // per stream: entries of INPUT history kept on the device
    DevBuf<yg_cf32> d_hist[2];
    int cur = 0;
    DevBuf<yg_cf32> d_U;           // synthesiser scratch [stream][(p-1)+n][M]
    DevBuf<yg_cf32> d_stage_x, d_stage_y;
    FirpfbchFastPlan fast;         // fused kernels (M = 64)
    FirpfbchFastPlan tiny;         // fused tiny-M kernels (M = 8, 16, 32)
    int32_t last_path = 0;
};

namespace {

__global__ void k_pfbch_analysis(const float* __restrict__ h, const float2* __restrict__ tw,
                                 const float2* __restrict__ hist, long long Hlen,
                                 const float2* __restrict__ x, float2* __restrict__ y,
                                 uint32_t M, uint32_t p, long long n_frames, long long n_streams)
{
    extern __shared__ float2 sm[];
    float2* X = sm;
    float2* Y = sm + M;
    const long long total = n_frames * n_streams;
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const long long s = w / n_frames, q = w - s * n_frames;
        const float2* xs = x + s * n_frames * (long long)M;
        const float2* hs = hist + s * Hlen;
        const long long tq = q * (long long)M + M - 1;
        for (uint32_t b = threadIdx.x; b < M; b += blockDim.x) {
            float2 acc = make_float2(0.f, 0.f);
            for (int n = (int)p - 1; n >= 0; n--) {
                const long long t = tq - b - (long long)n * M;
                const float2 v = (t >= 0) ? __ldg(&xs[t]) : __ldg(&hs[Hlen + t]);
                const float c = __ldg(&h[b + (size_t)n * M]);
                acc.x = fmaf(c, v.x, acc.x);
                acc.y = fmaf(c, v.y, acc.y);
            }
            X[M - 1 - b] = acc;
        }
        const float2* r = block_dft(X, Y, M, tw, 0);
        float2* ys = y + (s * n_frames + q) * (long long)M;
        for (uint32_t c = threadIdx.x; c < M; c += blockDim.x) ys[c] = r[c];
        __syncthreads();
    }
}

// per-stream history update: new_hist[s][i] = stream_s[n_new - Hlen + i]
__global__ void k_pfbch_update_hist(float2* __restrict__ hist_new, const float2* __restrict__ hist_old,
                                    long long Hlen, const float2* __restrict__ x, long long n_new, long long n_streams)
{
    const long long total = Hlen * n_streams;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long s = g / Hlen, i = g - s * Hlen;
        const long long t = n_new - Hlen + i;
        hist_new[g] = (t >= 0) ? x[s * n_new + t] : hist_old[s * Hlen + Hlen + t];
    }
}

// U layout: [stream][(p-1) + n_frames][M].  U frame g of a stream is the IDFT of virtual input frame g - (p-1):
// negative virtual frames are the tail of the stream's input history (hist holds hist_frames frames per stream).
__global__ void k_pfbch_synth_ifft(const float2* __restrict__ tw, const float2* __restrict__ hist, long long hist_frames,
                                   const float2* __restrict__ x, float2* __restrict__ U,
                                   uint32_t M, uint32_t p, long long n_frames, long long n_streams)
{
    extern __shared__ float2 sm[];
    float2* X = sm;
    float2* Y = sm + M;
    const long long per = n_frames + p - 1;
    const long long total = per * n_streams;
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const long long s = w / per, g = w - s * per;
        const long long v = g - (long long)(p - 1);
        const float2* xs = (v >= 0) ? x + (s * n_frames + v) * (long long)M
                                    : hist + (s * hist_frames + hist_frames + v) * (long long)M;
        for (uint32_t c = threadIdx.x; c < M; c += blockDim.x) X[c] = __ldg(&xs[c]);
        const float2* r = block_dft(X, Y, M, tw, 1);
        float2* us = U + (s * per + g) * (long long)M;
        for (uint32_t c = threadIdx.x; c < M; c += blockDim.x) us[c] = r[c];
        __syncthreads();
    }
}

__global__ void k_pfbch_synth_fir(const float* __restrict__ h, const float2* __restrict__ U, float2* __restrict__ y,
                                  uint32_t M, uint32_t p, long long n_frames, long long n_streams)
{
    const long long per = n_frames * (long long)M;
    const long long total = per * n_streams;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long s = g / per, o = g - s * per;
        const long long q = o / M;
        const uint32_t i = (uint32_t)(o - q * M);
        const float2* us = U + (s * (n_frames + p - 1) + (p - 1) + q) * (long long)M + i;
        float2 acc = make_float2(0.f, 0.f);
        for (int n = (int)p - 1; n >= 0; n--) {
            const float c = __ldg(&h[i + (size_t)n * M]);
            const float2 u = __ldg(us - (long long)n * M);
            acc.x = fmaf(c, u.x, acc.x);
            acc.y = fmaf(c, u.y, acc.y);
        }
        y[g] = acc;
    }
}

int32_t check(yg_firpfbch_crcf q)
{
    if (!q) return fail(YG_EVALUE, "null firpfbch handle");
    return YG_OK;
}

int32_t validate(int32_t type, uint32_t M, uint32_t p, uint32_t n_streams)
{
    if (type != YG_ANALYZER && type != YG_SYNTHESIZER) return fail(YG_ECONFIG, "invalid type %d", type);
    if (M == 0) return fail(YG_ECONFIG, "number of channels must be greater than 0");
    if (p == 0) return fail(YG_ECONFIG, "invalid filter size (must be greater than 0)");
    if (n_streams == 0) return fail(YG_ECONFIG, "number of streams must be greater than 0");
    return YG_OK;
}

int32_t set_smem(const void* fn, size_t bytes)
{
    if (bytes > 48 * 1024) {
        if (bytes > 227 * 1024) return fail(YG_ECONFIG, "M too large (needs %zu B shared memory)", bytes);
        YG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    return YG_OK;
}

int32_t execute_dev_impl(yg_firpfbch_crcf q, const yg_cf32* d_x, size_t n_frames, yg_cf32* d_y, cudaStream_t st);

int32_t execute_dev(yg_firpfbch_crcf q, const yg_cf32* d_x, size_t n_frames, yg_cf32* d_y, cudaStream_t st)
{
    YG_TRY(q->order.enter(st));
    YG_TRY(execute_dev_impl(q, d_x, n_frames, d_y, st));
    return q->order.leave(st);
}

int32_t execute_dev_impl(yg_firpfbch_crcf q, const yg_cf32* d_x, size_t n_frames, yg_cf32* d_y, cudaStream_t st)
{
    if (n_frames == 0) return YG_OK;
    const uint32_t M = q->M, p = q->p;
    const long long S = q->n_streams;
    const long long Hlen = (long long)q->state_len;
    const float2* x = reinterpret_cast<const float2*>(d_x);
    float2* y = reinterpret_cast<float2*>(d_y);
    const size_t smem = 2 * (size_t)M * sizeof(float2);
    const int block = (int)std::min<uint32_t>(256, (M + 31) / 32 * 32);
    // tiny-M fused kernels: groups of 32 / M streams, the remaining streams go to the generic kernels below
    const long long spw = (M <= 32 && M > 0) ? 32 / M : 1;
    const bool use_tiny = q->tiny.supported && S >= spw && n_frames >= 16 &&
                          ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (q->type == YG_ANALYZER) {
        const float2* hist = reinterpret_cast<const float2*>(q->d_hist[q->cur].p);
        long long s_fast = 0;
        if (use_tiny) {
            s_fast = (S / spw) * spw;
            YG_TRY(firpfbch_tiny_launch(q->tiny, hist, Hlen, x, y, (long long)n_frames, s_fast, st));
            q->last_path = 2;
        } else if (q->fast.supported && S >= 4 && n_frames >= 16 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
            // groups of four streams go to the fused kernel, the remaining 0..3 streams to the generic one
            s_fast = (S / 4) * 4;
            YG_TRY(firpfbch_fast_launch(q->fast, hist, Hlen, x, y, (long long)n_frames, s_fast, st));
            q->last_path = 2;
        } else {
            q->last_path = 1;
        }
        if (s_fast < S) {
            const long long rest = S - s_fast;
            const long long per = (long long)n_frames * M;
            YG_TRY(set_smem((const void*)k_pfbch_analysis, smem));
            const int g1 = (int)std::min<long long>((long long)n_frames * rest, q->n_sm * 16);
            k_pfbch_analysis<<<g1, block, smem, st>>>(q->d_h.p, q->d_tw.p, hist + s_fast * Hlen, Hlen, x + s_fast * per,
                                                      y + s_fast * per, M, p, (long long)n_frames, rest);
            YG_LAUNCH_CHECK();
        }
    } else {
        const float2* hist = reinterpret_cast<const float2*>(q->d_hist[q->cur].p);
        const long long hist_frames = Hlen / M;
        long long s_fast = 0;
        if (use_tiny) {
            s_fast = (S / spw) * spw;
            YG_TRY(firpfbch_tiny_launch(q->tiny, hist, Hlen, x, y, (long long)n_frames, s_fast, st));
            q->last_path = 2;
        } else if (q->fast.supported && S >= 4 && n_frames >= 16 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
            s_fast = (S / 4) * 4;
            YG_TRY(firpfbch_fast_synth_launch(q->fast, hist, hist_frames, x, y, (long long)n_frames, s_fast, st));
            q->last_path = 2;
        } else {
            q->last_path = 1;
        }
        if (s_fast < S) {
            const long long rest = S - s_fast;
            const long long per = (long long)n_frames * M;
            const long long ftm = (long long)(n_frames + p - 1) * M;
            YG_TRY(q->d_U.reserve((size_t)ftm * rest));
            float2* U = reinterpret_cast<float2*>(q->d_U.p);
            YG_TRY(set_smem((const void*)k_pfbch_synth_ifft, smem));
            const int g1 = (int)std::min<long long>((long long)(n_frames + p - 1) * rest, q->n_sm * 16);
            k_pfbch_synth_ifft<<<g1, block, smem, st>>>(q->d_tw.p, hist + s_fast * Hlen, hist_frames, x + s_fast * per, U, M, p,
                                                        (long long)n_frames, rest);
            YG_LAUNCH_CHECK();
            const long long total = per * rest;
            const int g3 = (int)std::min<long long>((total + 255) / 256, q->n_sm * 32);
            k_pfbch_synth_fir<<<g3, 256, 0, st>>>(q->d_h.p, U, y + s_fast * per, M, p, (long long)n_frames, rest);
            YG_LAUNCH_CHECK();
        }
    }
    // both types keep the tail of their INPUT stream as state
    if (Hlen > 0) {
        const int nxt = q->cur ^ 1;
        const int g2 = (int)std::min<long long>((Hlen * S + 255) / 256, q->n_sm * 8);
        k_pfbch_update_hist<<<g2, 256, 0, st>>>(reinterpret_cast<float2*>(q->d_hist[nxt].p),
                                                reinterpret_cast<const float2*>(q->d_hist[q->cur].p), Hlen, x,
                                                (long long)n_frames * M, S);
        YG_LAUNCH_CHECK();
        q->cur = nxt;
    }
    return YG_OK;
}

int32_t build(int32_t type, uint32_t M, uint32_t p, const float* h, size_t h_len, uint32_t n_streams,
              yg_firpfbch_crcf* out)
{
    if (!out) return fail(YG_EVALUE, "null output pointer");
    *out = nullptr;
    YG_TRY(validate(type, M, p, n_streams));
    const size_t L = (size_t)M * p;
    if (!h) return fail(YG_EVALUE, "null prototype filter");
    if (h_len < L) return fail(YG_ECONFIG, "prototype filter length (%zu) must be at least M*p (%zu)", h_len, L);
    int dev = 0;
    YG_TRY(require_device(&dev));
    auto* q = new yg_firpfbch_crcf_s();
    q->type = type; q->M = M; q->p = p; q->n_streams = n_streams; q->L = L; q->dev = dev;
    q->n_sm = sm_count(dev);
    q->h.assign(h, h + L);
    auto cleanup = [&](int32_t rc) { yg_firpfbch_crcf_destroy(q); return rc; };
#define TRYQ(expr) do { int32_t _rc = (expr); if (_rc != YG_OK) return cleanup(_rc); } while (0)
#define CUDAQ(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return cleanup(fail(YG_EINTERNAL, "CUDA error %s (%s)", cudaGetErrorString(_e), #expr)); } while (0)
    CUDAQ(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    q->order.own = q->stream;
    TRYQ(q->d_h.reserve(L));
    CUDAQ(yg::memcpy_sync(q->d_h.p, q->h.data(), L * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<float2> tw;
    make_twiddles(M, tw);
    TRYQ(q->d_tw.reserve(M));
    CUDAQ(yg::memcpy_sync(q->d_tw.p, tw.data(), M * sizeof(float2), cudaMemcpyHostToDevice));
    // analyser: the last p-1 input frames; synthesiser: at least 16 (one warm-up batch of the fused kernel)
    q->state_len = (type == YG_ANALYZER) ? (size_t)(p - 1) * M : (size_t)std::max<uint32_t>(p - 1, 16) * M;
    TRYQ(firpfbch_fast_plan(q->fast, type, M, p, q->h.data()));
    TRYQ(firpfbch_tiny_plan(q->tiny, type, M, p, q->h.data()));
    for (int b = 0; b < 2; b++) {
        TRYQ(q->d_hist[b].reserve(std::max<size_t>(1, q->state_len * n_streams)));
        CUDAQ(yg::memset_sync(q->d_hist[b].p, 0, std::max<size_t>(1, q->state_len * n_streams) * sizeof(yg_cf32)));
    }
#undef TRYQ
#undef CUDAQ
    *out = q;
    return YG_OK;
}

}  // namespace

extern "C" {

int32_t yg_firpfbch_crcf_create(int32_t type, uint32_t M, uint32_t p, const float* h, size_t h_len, uint32_t n_streams,
                                yg_firpfbch_crcf* out)
{
    return build(type, M, p, h, h_len, n_streams, out);
}

int32_t yg_firpfbch_crcf_create_kaiser(int32_t type, uint32_t M, uint32_t m, float as, uint32_t n_streams,
                                       yg_firpfbch_crcf* out)
{
    if (!out) return fail(YG_EVALUE, "null output pointer");
    *out = nullptr;
    YG_TRY(validate(type, M, m, n_streams));
    const uint32_t n = 2 * M * m + 1;
    std::vector<float> hf(n);
    YG_TRY(fir_design_kaiser(n, 0.5f / (float)M, as, 0.0f, hf.data()));
    return build(type, M, 2 * m, hf.data(), n, n_streams, out);
}

int32_t yg_firpfbch_crcf_clone(yg_firpfbch_crcf q, yg_firpfbch_crcf* out)
{
    YG_TRY(check(q));
    YG_DEVICE_GUARD(q->dev);
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    yg_firpfbch_crcf c = nullptr;
    YG_TRY(build(q->type, q->M, q->p, q->h.data(), q->h.size(), q->n_streams, &c));
    if (q->state_len) {
        cudaError_t e = yg::memcpy_sync(c->d_hist[c->cur].p, q->d_hist[q->cur].p,
                                   q->state_len * q->n_streams * sizeof(yg_cf32), cudaMemcpyDeviceToDevice);
        if (e != cudaSuccess) { yg_firpfbch_crcf_destroy(c); return fail(YG_EINTERNAL, "CUDA error %s", cudaGetErrorString(e)); }
    }
    *out = c;
    return YG_OK;
}

int32_t yg_firpfbch_crcf_destroy(yg_firpfbch_crcf q)
{
    if (!q) return YG_OK;
    YG_DEVICE_GUARD(q->dev);
    if (q->stream) cudaStreamSynchronize(q->stream);
    q->order.wait_host();
    q->order.destroy();
    q->d_h.release(); q->d_tw.release(); q->d_hist[0].release(); q->d_hist[1].release(); q->d_U.release();
    q->d_stage_x.release(); q->d_stage_y.release();
    firpfbch_fast_release(q->fast);
    firpfbch_fast_release(q->tiny);
    if (q->stream) cudaStreamDestroy(q->stream);
    delete q;
    return YG_OK;
}

int32_t yg_firpfbch_crcf_reset(yg_firpfbch_crcf q)
{
    YG_TRY(check(q));
    YG_DEVICE_GUARD(q->dev);
    YG_TRY(q->order.wait_host());
    if (q->state_len)
        YG_CUDA(cudaMemsetAsync(q->d_hist[q->cur].p, 0, q->state_len * q->n_streams * sizeof(yg_cf32), q->stream));
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    return YG_OK;
}

int32_t yg_firpfbch_crcf_execute_block_dev(yg_firpfbch_crcf q, const yg_cf32* d_x, size_t n_frames, yg_cf32* d_y,
                                           void* cuda_stream)
{
    YG_TRY(check(q));
    if (n_frames && (!d_x || !d_y)) return fail(YG_EVALUE, "null buffer");
    YG_DEVICE_GUARD(q->dev);
    return execute_dev(q, d_x, n_frames, d_y, (cudaStream_t)cuda_stream);
}

int32_t yg_firpfbch_crcf_execute_block(yg_firpfbch_crcf q, const yg_cf32* x, size_t n_frames, yg_cf32* y)
{
    YG_TRY(check(q));
    if (n_frames && (!x || !y)) return fail(YG_EVALUE, "null buffer");
    if (n_frames == 0) return YG_OK;
    YG_DEVICE_GUARD(q->dev);
    // stream-major layout makes time-chunking a strided copy; stage the whole block
    const size_t n = n_frames * (size_t)q->M * q->n_streams;
    YG_TRY(q->d_stage_x.reserve(n));
    YG_TRY(q->d_stage_y.reserve(n));
    YG_CUDA(cudaMemcpyAsync(q->d_stage_x.p, x, n * sizeof(yg_cf32), cudaMemcpyHostToDevice, q->stream));
    YG_TRY(execute_dev(q, q->d_stage_x.p, n_frames, q->d_stage_y.p, q->stream));
    YG_CUDA(cudaMemcpyAsync(y, q->d_stage_y.p, n * sizeof(yg_cf32), cudaMemcpyDeviceToHost, q->stream));
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    return YG_OK;
}

int32_t yg_firpfbch_crcf_execute(yg_firpfbch_crcf q, const yg_cf32* x, yg_cf32* y)
{
    return yg_firpfbch_crcf_execute_block(q, x, 1, y);
}

int32_t yg_firpfbch_crcf_sync(yg_firpfbch_crcf q)
{
    YG_TRY(check(q));
    YG_DEVICE_GUARD(q->dev);
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    return YG_OK;
}

int32_t yg_firpfbch_crcf_get_type(yg_firpfbch_crcf q, int32_t* type) { YG_TRY(check(q)); *type = q->type; return YG_OK; }
int32_t yg_firpfbch_crcf_get_M(yg_firpfbch_crcf q, uint32_t* M) { YG_TRY(check(q)); *M = q->M; return YG_OK; }
int32_t yg_firpfbch_crcf_get_p(yg_firpfbch_crcf q, uint32_t* p) { YG_TRY(check(q)); *p = q->p; return YG_OK; }
int32_t yg_firpfbch_crcf_get_n_streams(yg_firpfbch_crcf q, uint32_t* n) { YG_TRY(check(q)); *n = q->n_streams; return YG_OK; }
int32_t yg_firpfbch_crcf_get_device(yg_firpfbch_crcf q, int32_t* dev) { YG_TRY(check(q)); *dev = q->dev; return YG_OK; }
int32_t yg_firpfbch_crcf_last_path(yg_firpfbch_crcf q, int32_t* path) { YG_TRY(check(q)); *path = q->last_path; return YG_OK; }
int32_t yg_firpfbch_crcf_get_taps(yg_firpfbch_crcf q, float* h)
{
    YG_TRY(check(q));
    if (!h) return fail(YG_EVALUE, "null pointer");
    memcpy(h, q->h.data(), q->L * sizeof(float));
    return YG_OK;
}

}  // extern "C"
