// firpfbch.cu -- firpfbch_crcf (critically sampled channelizer), batched over independent streams.
//
// Closed forms (frame q since reset, s[t<0] = 0):
//   analysis : V_q[b] = sum_n h[b+nM] s[qM + M-1 - b - nM];  X[M-1-b] = V_q[b];  y_q = DFT_forward(X)
//   synthesis: U_q = IDFT_unnorm(X_q);  y[qM + i] = sum_n h[i+nM] U_{q-n}[i]
#include "common.cuh"
#include "firpfbch_fast.cuh"

#include <algorithm>
#include <cstdlib>

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
    // tiled generic kernels (any M whose tile fits shared memory): F frames of one stream per CTA pass
    struct Tiled {
        bool supported = false;
        int F = 0, threads = 256, ctas_per_sm = 1;
        size_t smem = 0;
        TiledPass tp = {};
    } tiled;
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

// ------------------------------------------------------------------ tiled generic kernels
// The firpfbch2 scheme (firpfbch2.cu, k_analysis_tiled): a CTA takes F consecutive frames of ONE stream, stages their
// input span ((F + p - 1) M samples of history ++ x), the taps and the twiddles in shared memory, runs the branch dot
// products out of shared memory and the mixed-radix passes over all F frames.  The analyser's FORWARD transform is the
// backward one on conjugated data: DFT_f(x) = conj(DFT_b(conj(x))).
// Shared: T[M] | taps[p M] floats | Xin[(F + p - 1) M] | A[F M] | B[F M]
__global__ void __launch_bounds__(1024) k_pfbch_analysis_tiled(const float* __restrict__ h, const float2* __restrict__ tw,
                                                               const float2* __restrict__ hist, long long Hlen,
                                                               const float2* __restrict__ x, float2* __restrict__ y,
                                                               uint32_t M, uint32_t p, long long n_frames, long long n_streams,
                                                               uint32_t F, TiledPass tp)
{
    extern __shared__ float2 sm[];
    float2* T = sm;
    float* taps = reinterpret_cast<float*>(sm + M);
    float2* Xin = sm + M + (p * M + 1) / 2;
    float2* A = Xin + (F + p - 1) * M;
    float2* B = A + F * M;
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) T[i] = __ldg(&tw[i]);
    for (uint32_t i = threadIdx.x; i < p * M; i += blockDim.x) taps[i] = __ldg(&h[i]);
    const long long tiles_per = (n_frames + F - 1) / F;
    const long long n_tiles = tiles_per * n_streams;
    for (long long g = blockIdx.x; g < n_tiles; g += gridDim.x) {
        const long long s = g / tiles_per, q0 = (g - s * tiles_per) * F;
        const uint32_t nf = (uint32_t)min((long long)F, n_frames - q0);
        const float2* xs = x + s * n_frames * (long long)M;
        const float2* hs = hist + s * Hlen;
        const long long t_start = (q0 - (long long)p + 1) * (long long)M;
        __syncthreads();                                     // the previous tile has been stored (and T, taps are loaded)
        for (uint32_t i = threadIdx.x; i < (nf + p - 1) * M; i += blockDim.x) {
            const long long t = t_start + i;
            float2 v = make_float2(0.f, 0.f);
            if (t >= 0) v = __ldg(&xs[t]);
            else if (Hlen + t >= 0) v = __ldg(&hs[Hlen + t]);
            Xin[i] = v;
        }
        __syncthreads();
        // sample t_q - b - n M sits at Xin[(fl + p) M - 1 - b - n M]
        for (uint32_t it = threadIdx.x; it < nf * M; it += blockDim.x) {
            const uint32_t fl = it / M, b = it - fl * M;
            const float2* xw = Xin + (fl + p) * M - 1 - b;
            const float* hw = taps + b;
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
            for (int n = (int)p - 1; n >= 0; n--)            // oldest sample first (src/dotprod/mod.rs:36-39); packed FFMA2
                acc = __ffma2_rn(xw[-(int)(n * M)], make_float2(hw[n * M], hw[n * M]), acc);
            A[fl * M + (M - 1 - b)] = make_float2(acc.x, -acc.y);
        }
        __syncthreads();
        const float2* r = tiled_dft(A, B, T, M, tp, nf);
        float2* ys = y + (s * n_frames + q0) * (long long)M;
        for (uint32_t it = threadIdx.x; it < nf * M; it += blockDim.x) {
            const float2 v = r[it];
            ys[it] = make_float2(v.x, -v.y);
        }
    }
}

// The synthesiser's stage 1 in the same shape (U frame g of a stream = IDFT of virtual input frame g - (p - 1)).
// Shared: T[M] | A[F M] | B[F M]
__global__ void __launch_bounds__(1024) k_pfbch_synth_ifft_tiled(const float2* __restrict__ tw, const float2* __restrict__ hist,
                                                                 long long hist_frames, const float2* __restrict__ x,
                                                                 float2* __restrict__ U, uint32_t M, uint32_t p, long long n_frames,
                                                                 long long n_streams, uint32_t F, TiledPass tp)
{
    extern __shared__ float2 sm[];
    float2* T = sm;
    float2* A = sm + M;
    float2* B = A + F * M;
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) T[i] = __ldg(&tw[i]);
    const long long per = n_frames + p - 1;
    const long long tiles_per = (per + F - 1) / F;
    const long long n_tiles = tiles_per * n_streams;
    for (long long g = blockIdx.x; g < n_tiles; g += gridDim.x) {
        const long long s = g / tiles_per, g0 = (g - s * tiles_per) * F;
        const uint32_t nf = (uint32_t)min((long long)F, per - g0);
        __syncthreads();
        for (uint32_t it = threadIdx.x; it < nf * M; it += blockDim.x) {
            const uint32_t fl = it / M, c = it - fl * M;
            const long long v = g0 + fl - (long long)(p - 1);
            const float2* src = (v >= 0) ? x + (s * n_frames + v) * (long long)M
                                         : hist + (s * hist_frames + hist_frames + v) * (long long)M;
            A[it] = __ldg(&src[c]);
        }
        __syncthreads();
        const float2* r = tiled_dft(A, B, T, M, tp, nf);
        float2* us = U + (s * per + g0) * (long long)M;
        for (uint32_t it = threadIdx.x; it < nf * M; it += blockDim.x) us[it] = r[it];
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
            if (q->tiled.supported) {
                const long long tiles = (((long long)n_frames + q->tiled.F - 1) / q->tiled.F) * rest;
                const int gt = (int)std::min<long long>(tiles, (long long)q->n_sm * q->tiled.ctas_per_sm);
                k_pfbch_analysis_tiled<<<gt, q->tiled.threads, q->tiled.smem, st>>>(q->d_h.p, q->d_tw.p, hist + s_fast * Hlen, Hlen,
                                                                                    x + s_fast * per, y + s_fast * per, M, p,
                                                                                    (long long)n_frames, rest, (uint32_t)q->tiled.F,
                                                                                    q->tiled.tp);
            } else {
                YG_TRY(set_smem((const void*)k_pfbch_analysis, smem));
                const int g1 = (int)std::min<long long>((long long)n_frames * rest, q->n_sm * 16);
                k_pfbch_analysis<<<g1, block, smem, st>>>(q->d_h.p, q->d_tw.p, hist + s_fast * Hlen, Hlen, x + s_fast * per,
                                                          y + s_fast * per, M, p, (long long)n_frames, rest);
            }
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
            if (q->tiled.supported) {
                const long long tiles = (((long long)(n_frames + p - 1) + q->tiled.F - 1) / q->tiled.F) * rest;
                const int gt = (int)std::min<long long>(tiles, (long long)q->n_sm * q->tiled.ctas_per_sm);
                k_pfbch_synth_ifft_tiled<<<gt, q->tiled.threads, q->tiled.smem, st>>>(q->d_tw.p, hist + s_fast * Hlen, hist_frames,
                                                                                      x + s_fast * per, U, M, p, (long long)n_frames,
                                                                                      rest, (uint32_t)q->tiled.F, q->tiled.tp);
            } else {
                YG_TRY(set_smem((const void*)k_pfbch_synth_ifft, smem));
                const int g1 = (int)std::min<long long>((long long)(n_frames + p - 1) * rest, q->n_sm * 16);
                k_pfbch_synth_ifft<<<g1, block, smem, st>>>(q->d_tw.p, hist + s_fast * Hlen, hist_frames, x + s_fast * per, U, M, p,
                                                            (long long)n_frames, rest);
            }
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

// Tile geometry of the tiled generic kernels (see plan_tiled in firpfbch2.cu)
int32_t plan_tiled(yg_firpfbch_crcf q)
{
    auto& t = q->tiled;
    t.supported = false;
    const char* e = getenv("YG_GENERIC_TILED");           // debugging knob: 0 keeps the one-frame-per-block kernels
    if (e && e[0] == '0') return YG_OK;
    const size_t M = q->M, P = q->p;
    if (M < 2) return YG_OK;
    if (!plan_radices(q->M, t.tp)) return YG_OK;          // a large prime factor: the one-frame-per-block kernel takes it
    auto bytes = [&](size_t F) {
        return q->type == YG_ANALYZER ? 8 * (M + (P * M + 1) / 2 + (F + P - 1) * M + 2 * F * M) : 8 * (M + 2 * F * M);
    };
    const size_t budget = 190 * 1024;
    size_t F = std::min<size_t>(64, std::max<size_t>(1, 8192 / M));
    while (F > 1 && bytes(F) > budget) F--;
    if (bytes(F) > budget) return YG_OK;
    // (as for firpfbch2: the synthesiser's stage 1 at a large power-of-two M stays on the radix-4 one-frame-per-block kernel)
    if (q->type == YG_SYNTHESIZER && M >= 128 && (M & (M - 1)) == 0) return YG_OK;
    // ... and so is an analyser at a power-of-two M whose tile holds fewer than 8 frames (taps and history are reloaded per
    // tile; measured firpfbch M = 1024, p = 8, F = 3: 35 vs 47 Gsps)
    if (q->type == YG_ANALYZER && (M & (M - 1)) == 0 && F < 8) return YG_OK;
    t.F = (int)F;
    t.smem = bytes(F);
    if (q->type == YG_ANALYZER) YG_CUDA(cudaFuncSetAttribute(k_pfbch_analysis_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    else YG_CUDA(cudaFuncSetAttribute(k_pfbch_synth_ifft_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    t.ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (t.smem + 1024)));
    t.threads = std::min(1024, (2048 / t.ctas_per_sm) & ~31);
    t.supported = true;
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
    TRYQ(plan_tiled(q));
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
