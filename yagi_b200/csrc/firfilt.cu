// firfilt.cu -- firfilt_crcf: direct-form FIR, real taps x complex samples, batched over streams.
//   y[s][n] = scale * sum_k h[k] x[s][n-k]        (src/filter/fir/firfilt.rs:241-245, :267-278)
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

using namespace yg;

namespace yg {
// firfilt_fast.cu
bool firfilt_fast_supported(size_t h_len);
int32_t firfilt_fast_launch(const float* h, size_t h_len, float scale, const float2* hist, long long Hlen, const float2* x,
                            float2* y, long long n, long long n_streams, cudaStream_t st, long long t_begin = 0);
// firfilt_tc.cu: tensor-core (tcgen05, 3xTF32 banded Toeplitz) path for <= 65 taps
long long firfilt_tc_prefix(size_t h_len, long long n, long long n_streams, const void* x, const void* y);
int32_t firfilt_tc_plan(const float* h, size_t h_len, float** d_toep);
bool firfilt_tc_taps_ok(size_t h_len);
int32_t firfilt_tc_launch(const float* d_toep, size_t h_len, float scale, const float2* hist, long long Hlen, const float2* x, float2* y,
                          long long n, long long pitch, long long n_streams, int n_sm, cudaStream_t st);
}  // namespace yg

struct yg_firfilt_crcf_s {
    size_t h_len = 0;
    uint32_t n_streams = 1;
    float scale = 1.0f;
    int dev = 0;
    int n_sm = 1;                  // multiprocessor count of `dev` (grid sizing)
    cudaStream_t stream = nullptr;
    StreamOrder order;
    std::vector<float> h;
    DevBuf<float> d_h;
    size_t state_len = 0;          // h_len - 1 samples per stream
    DevBuf<yg_cf32> d_hist[2];
    int cur = 0;
    DevBuf<yg_cf32> d_stage_x, d_stage_y;
    float* d_toep = nullptr;       // tensor-core path: aliased Toeplitz tables (null: path not planned)
    int tc_mode = 0;               // 0 never, 1 whenever the geometry allows (YG_FIRFILT_TC)
    int32_t last_path = 0;         // 0 none, 1 generic, 2 register-blocked FFMA2, 4 tensor cores
};

namespace {

constexpr int kOutPerThread = 8;
constexpr int kFirfiltTcDefault = 1;     // 3.10 ms vs 4.39 ms on BASELINE config #2 (0.85 vs 0.60 of the HBM roofline)

// Each thread produces kOutPerThread consecutive outputs of one stream.  Taps are staged in shared
// memory, zero-padded by kOutPerThread-1 on both sides so the inner loop needs no bounds test.
// Accumulation order per output: oldest sample (highest tap index) first.
__global__ void k_firfilt(const float* __restrict__ h, int h_len, float scale,
                          const float2* __restrict__ hist, long long Hlen,
                          const float2* __restrict__ x, float2* __restrict__ y, long long n, long long n_streams)
{
    extern __shared__ float hp[];      // hp[i + R-1] = h[i]
    constexpr int R = kOutPerThread;
    const int padded = h_len + 2 * (R - 1);
    for (int i = threadIdx.x; i < padded; i += blockDim.x) {
        const int k = i - (R - 1);
        hp[i] = (k >= 0 && k < h_len) ? h[k] : 0.0f;
    }
    __syncthreads();
    const long long tiles = (n + R - 1) / R;
    const long long total = tiles * n_streams;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long s = g / tiles;
        const long long n0 = (g - s * tiles) * R;
        const float2* xs = x + s * n;
        const float2* hs = hist + s * Hlen;
        float2 acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = make_float2(0.f, 0.f);
        // input index j ascending; x[j] feeds output n0+r through tap k = n0 + r - j
        for (long long j = n0 - (h_len - 1); j <= n0 + R - 1; j++) {
            float2 v = make_float2(0.f, 0.f);
            if (j < n) v = (j >= 0) ? __ldg(&xs[j]) : __ldg(&hs[Hlen + j]);
            const int d = (int)(n0 - j) + (R - 1);       // tap index for r = 0, shifted by the padding
#pragma unroll
            for (int r = 0; r < R; r++) {
                const float c = hp[d + r];
                acc[r].x = fmaf(c, v.x, acc[r].x);
                acc[r].y = fmaf(c, v.y, acc[r].y);
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++)
            if (n0 + r < n) y[s * n + n0 + r] = make_float2(acc[r].x * scale, acc[r].y * scale);
    }
}

__global__ void k_firfilt_update_hist(float2* __restrict__ hist_new, const float2* __restrict__ hist_old,
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

int32_t check(yg_firfilt_crcf q)
{
    if (!q) return fail(YG_EVALUE, "null firfilt handle");
    return YG_OK;
}

int32_t execute_dev_impl(yg_firfilt_crcf q, const yg_cf32* d_x, size_t n, yg_cf32* d_y, cudaStream_t st);

int32_t execute_dev(yg_firfilt_crcf q, const yg_cf32* d_x, size_t n, yg_cf32* d_y, cudaStream_t st)
{
    YG_TRY(q->order.enter(st));
    YG_TRY(execute_dev_impl(q, d_x, n, d_y, st));
    return q->order.leave(st);
}

int32_t execute_dev_impl(yg_firfilt_crcf q, const yg_cf32* d_x, size_t n, yg_cf32* d_y, cudaStream_t st)
{
    if (n == 0) return YG_OK;
    const long long S = q->n_streams;
    const long long Hlen = (long long)q->state_len;
    const long long n_main = (q->tc_mode && q->d_toep) ? firfilt_tc_prefix(q->h_len, (long long)n, S, d_x, d_y) : 0;
    if (n_main > 0) {
        // tcgen05 3xTF32 Toeplitz GEMM (<= 65 taps) on the longest prefix that is whole segments; the FFMA2 kernel
        // finishes the remaining < 8192 (< 512 for short streams) samples of every stream, reading its history from x
        YG_TRY(firfilt_tc_launch(q->d_toep, q->h_len, q->scale, reinterpret_cast<const float2*>(q->d_hist[q->cur].p), Hlen,
                                 reinterpret_cast<const float2*>(d_x), reinterpret_cast<float2*>(d_y), n_main, (long long)n, S,
                                 q->n_sm, st));
        YG_TRY(firfilt_fast_launch(q->h.data(), q->h_len, q->scale, reinterpret_cast<const float2*>(q->d_hist[q->cur].p), Hlen,
                                   reinterpret_cast<const float2*>(d_x), reinterpret_cast<float2*>(d_y), (long long)n, S, st, n_main));
        q->last_path = 4;
    } else if (firfilt_fast_supported(q->h_len) && (long long)n * S >= 4096) {
        // register-blocked FFMA2 kernel (taps as kernel parameters); generic kernel for long filters / tiny calls
        YG_TRY(firfilt_fast_launch(q->h.data(), q->h_len, q->scale, reinterpret_cast<const float2*>(q->d_hist[q->cur].p), Hlen,
                                   reinterpret_cast<const float2*>(d_x), reinterpret_cast<float2*>(d_y), (long long)n, S, st));
        q->last_path = 2;
    } else {
        const long long tiles = ((long long)n + kOutPerThread - 1) / kOutPerThread;
        const int grid = (int)std::min<long long>((tiles * S + 127) / 128, q->n_sm * 32);
        const size_t smem = (q->h_len + 2 * (kOutPerThread - 1)) * sizeof(float);
        if (smem > 48 * 1024) return fail(YG_ECONFIG, "filter too long for this kernel (%zu taps)", q->h_len);
        k_firfilt<<<grid, 128, smem, st>>>(q->d_h.p, (int)q->h_len, q->scale,
                                           reinterpret_cast<const float2*>(q->d_hist[q->cur].p), Hlen,
                                           reinterpret_cast<const float2*>(d_x), reinterpret_cast<float2*>(d_y),
                                           (long long)n, S);
        q->last_path = 1;
        YG_LAUNCH_CHECK();
    }
    if (Hlen > 0) {
        const int nxt = q->cur ^ 1;
        const int g2 = (int)std::min<long long>((Hlen * S + 255) / 256, q->n_sm * 8);
        k_firfilt_update_hist<<<g2, 256, 0, st>>>(reinterpret_cast<float2*>(q->d_hist[nxt].p),
                                                  reinterpret_cast<const float2*>(q->d_hist[q->cur].p), Hlen,
                                                  reinterpret_cast<const float2*>(d_x), (long long)n, S);
        YG_LAUNCH_CHECK();
        q->cur = nxt;
    }
    return YG_OK;
}

int32_t build(const float* h, size_t h_len, uint32_t n_streams, yg_firfilt_crcf* out)
{
    if (!out) return fail(YG_EVALUE, "null output pointer");
    *out = nullptr;
    if (h_len == 0) return fail(YG_ECONFIG, "filter length must be greater than zero");   // firfilt.rs:65-67
    if (!h) return fail(YG_EVALUE, "null filter coefficients");
    if (n_streams == 0) return fail(YG_ECONFIG, "number of streams must be greater than 0");
    int dev = 0;
    YG_TRY(require_device(&dev));
    auto* q = new yg_firfilt_crcf_s();
    q->h_len = h_len; q->n_streams = n_streams; q->dev = dev;
    q->n_sm = sm_count(dev);
    q->h.assign(h, h + h_len);
    auto cleanup = [&](int32_t rc) { yg_firfilt_crcf_destroy(q); return rc; };
#define TRYQ(expr) do { int32_t _rc = (expr); if (_rc != YG_OK) return cleanup(_rc); } while (0)
#define CUDAQ(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return cleanup(fail(YG_EINTERNAL, "CUDA error %s (%s)", cudaGetErrorString(_e), #expr)); } while (0)
    CUDAQ(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    q->order.own = q->stream;
    TRYQ(q->d_h.reserve(h_len));
    CUDAQ(yg::memcpy_sync(q->d_h.p, q->h.data(), h_len * sizeof(float), cudaMemcpyHostToDevice));
    q->state_len = h_len - 1;
    for (int b = 0; b < 2; b++) {
        const size_t n = std::max<size_t>(1, q->state_len * n_streams);
        TRYQ(q->d_hist[b].reserve(n));
        CUDAQ(yg::memset_sync(q->d_hist[b].p, 0, n * sizeof(yg_cf32)));
    }
    {   // tensor-core path (firfilt_tc.cu), on by default; YG_FIRFILT_TC=0 turns it off (A/B runs: tools/tc_probe.py)
        const char* e = getenv("YG_FIRFILT_TC");
        q->tc_mode = e ? (e[0] != '0') : kFirfiltTcDefault;
        cudaDeviceProp prop;
        CUDAQ(cudaGetDeviceProperties(&prop, dev));
        if (q->tc_mode && firfilt_tc_taps_ok(h_len) && prop.major == 10) TRYQ(firfilt_tc_plan(q->h.data(), h_len, &q->d_toep));
    }
#undef TRYQ
#undef CUDAQ
    *out = q;
    return YG_OK;
}

}  // namespace

extern "C" {

int32_t yg_firfilt_crcf_create(const float* h, size_t h_len, uint32_t n_streams, yg_firfilt_crcf* out)
{
    return build(h, h_len, n_streams, out);
}

int32_t yg_firfilt_crcf_create_kaiser(uint32_t n, float fc, float as, float mu, uint32_t n_streams, yg_firfilt_crcf* out)
{
    if (!out) return fail(YG_EVALUE, "null output pointer");
    *out = nullptr;
    if (n == 0) return fail(YG_ECONFIG, "filter length must be greater than zero");
    std::vector<float> h(n);
    YG_TRY(fir_design_kaiser(n, fc, as, mu, h.data()));
    return build(h.data(), n, n_streams, out);
}

int32_t yg_firfilt_crcf_clone(yg_firfilt_crcf q, yg_firfilt_crcf* out)
{
    YG_TRY(check(q));
    YG_DEVICE_GUARD(q->dev);
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    yg_firfilt_crcf c = nullptr;
    YG_TRY(build(q->h.data(), q->h_len, q->n_streams, &c));
    c->scale = q->scale;
    if (q->state_len) {
        cudaError_t e = yg::memcpy_sync(c->d_hist[c->cur].p, q->d_hist[q->cur].p,
                                   q->state_len * q->n_streams * sizeof(yg_cf32), cudaMemcpyDeviceToDevice);
        if (e != cudaSuccess) { yg_firfilt_crcf_destroy(c); return fail(YG_EINTERNAL, "CUDA error %s", cudaGetErrorString(e)); }
    }
    *out = c;
    return YG_OK;
}

int32_t yg_firfilt_crcf_destroy(yg_firfilt_crcf q)
{
    if (!q) return YG_OK;
    YG_DEVICE_GUARD(q->dev);
    if (q->stream) cudaStreamSynchronize(q->stream);
    q->order.wait_host();
    q->order.destroy();
    q->d_h.release(); q->d_hist[0].release(); q->d_hist[1].release();
    q->d_stage_x.release(); q->d_stage_y.release();
    if (q->d_toep) cudaFree(q->d_toep);
    if (q->stream) cudaStreamDestroy(q->stream);
    delete q;
    return YG_OK;
}

int32_t yg_firfilt_crcf_reset(yg_firfilt_crcf q)
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

int32_t yg_firfilt_crcf_set_scale(yg_firfilt_crcf q, float scale) { YG_TRY(check(q)); q->scale = scale; return YG_OK; }
int32_t yg_firfilt_crcf_get_scale(yg_firfilt_crcf q, float* scale) { YG_TRY(check(q)); *scale = q->scale; return YG_OK; }
int32_t yg_firfilt_crcf_last_path(yg_firfilt_crcf q, int32_t* path) { YG_TRY(check(q)); *path = q->last_path; return YG_OK; }
int32_t yg_firfilt_crcf_get_device(yg_firfilt_crcf q, int32_t* dev) { YG_TRY(check(q)); *dev = q->dev; return YG_OK; }
int32_t yg_firfilt_crcf_get_len(yg_firfilt_crcf q, size_t* h_len) { YG_TRY(check(q)); *h_len = q->h_len; return YG_OK; }

int32_t yg_firfilt_crcf_execute_block_dev(yg_firfilt_crcf q, const yg_cf32* d_x, size_t n, yg_cf32* d_y, void* cuda_stream)
{
    YG_TRY(check(q));
    if (n && (!d_x || !d_y)) return fail(YG_EVALUE, "null buffer");
    YG_DEVICE_GUARD(q->dev);
    return execute_dev(q, d_x, n, d_y, (cudaStream_t)cuda_stream);
}

int32_t yg_firfilt_crcf_execute_block(yg_firfilt_crcf q, const yg_cf32* x, size_t n, yg_cf32* y)
{
    YG_TRY(check(q));
    if (n && (!x || !y)) return fail(YG_EVALUE, "null buffer");
    if (n == 0) return YG_OK;
    YG_DEVICE_GUARD(q->dev);
    const size_t tot = n * (size_t)q->n_streams;
    YG_TRY(q->d_stage_x.reserve(tot));
    YG_TRY(q->d_stage_y.reserve(tot));
    YG_CUDA(cudaMemcpyAsync(q->d_stage_x.p, x, tot * sizeof(yg_cf32), cudaMemcpyHostToDevice, q->stream));
    YG_TRY(execute_dev(q, q->d_stage_x.p, n, q->d_stage_y.p, q->stream));
    YG_CUDA(cudaMemcpyAsync(y, q->d_stage_y.p, tot * sizeof(yg_cf32), cudaMemcpyDeviceToHost, q->stream));
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    return YG_OK;
}

int32_t yg_firfilt_crcf_sync(yg_firfilt_crcf q)
{
    YG_TRY(check(q));
    YG_DEVICE_GUARD(q->dev);
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    return YG_OK;
}

}  // extern "C"
