// firpfbch2.cu -- firpfbch2_crcf handle, state management, generic kernels (any even M, any m)
// and dispatch to the fused fast path (firpfbch2_fast.cu).
//
// Closed forms computed here (frame k since reset, s[t<0] = 0):
//   analysis : V_k[b] = sum_n h[b+nM] s[t_k - b - nM],  t_k = (k+1) M/2 - 1
//              y_k    = (1/M) IDFT_unnorm( roll(V_k, (k&1) M/2) )
//   synthesis: u_k = 1/2 IDFT_unnorm(X_k)
//              y[k M/2 + i] = sum_{l<4m} h[i + l M/2] u_{k-l}[(i + (k&1) M/2) mod M]
#include "common.cuh"
#include "firpfbch2_fast.cuh"

#include <algorithm>

using namespace yg;

struct yg_firpfbch2_crcf_s {
    int32_t type = 0;
    uint32_t M = 0, M2 = 0, m = 0;
    size_t L = 0;                 // taps used = 2*M*m
    int dev = 0;
    int n_sm = 1;                  // multiprocessor count of `dev` (grid sizing)
    cudaStream_t stream = nullptr;
    StreamOrder order;
    std::vector<float> h;         // prototype (L)
    DevBuf<float> d_h;
    DevBuf<float2> d_tw;
    size_t state_len = 0;         // cf32 entries of history exposed through get/set_state
    size_t hist_len = 0;          // cf32 entries kept on the device (>= state_len; synthesiser keeps 32 frames)
    DevBuf<yg_cf32> d_hist[2];
    int cur = 0;
    int32_t flag = 0;
    DevBuf<yg_cf32> d_U;          // synthesiser: [4m-1 history frames + n frames][M]
    HostPipe pipe;
    int32_t last_path = 0;
    // ring of event pairs around the dominant kernel of each execute_block_dev call
    static constexpr int kRing = 64;
    cudaEvent_t ev0s[kRing] = {}, ev1s[kRing] = {};
    unsigned long long n_timed = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;      // the pair being recorded by the current call
    bool timed = false;
    bool timing_on = true;        // record CUDA events around the dominant kernel (yg_firpfbch2_crcf_set_kernel_timing)
    void next_events() { ev0 = ev0s[n_timed % kRing]; ev1 = ev1s[n_timed % kRing]; n_timed++; }
    Firpfbch2FastPlan fast;       // fused fast path (may be unsupported for this M/m)
    Firpfbch2FastPlan sfast;      // fused synthesis fast path
    Firpfbch2FastPlan large;      // two-stage large-M analysis path (M = 512 .. 4096)
    Firpfbch2FastPlan small;      // fused small-M analysis kernel (M = 64, 128)
    Firpfbch2FastPlan tiny;       // fused tiny-M analysis kernel (M = 8, 16, 32)
    Firpfbch2FastPlan slarge;     // two-stage large-M synthesis path (M = 1024)
    DevBuf<yg_cf32> d_Uc;         // its L2-sized U scratch
    // tiled generic kernels (any even M whose tile fits shared memory): F frames per CTA pass, mixed-radix passes
    struct Tiled {
        bool supported = false;
        int F = 0;                 // frames per tile
        TiledPass tp = {};
        size_t smem = 0;
        int ctas_per_sm = 1;
        int threads = 256;         // 2048 threads per SM whatever the tile's footprint allows resident
        int wF = 0;                // synthesiser stage 2 (overlap-add): frames per tile, 0 = one thread per output from global
        size_t wsmem = 0;
        int wctas = 1, wthreads = 256;
    } tiled;
};

namespace {

// ------------------------------------------------------------------ generic kernels
// One block per frame (grid-stride).  Shared: 2*M float2.
__global__ void k_analysis_generic(const float* __restrict__ h, const float2* __restrict__ tw,
                                   const float2* __restrict__ hist, long long Hlen,
                                   const float2* __restrict__ x, float2* __restrict__ y,
                                   uint32_t M, uint32_t P /*2m*/, long long f_begin, long long f_end, int flag0)
{
    extern __shared__ float2 sm[];
    float2* X = sm;
    float2* Y = sm + M;
    const uint32_t M2 = M >> 1;
    // frames are numbered from the start of this call; flag0 is the parity of frame 0
    for (long long f = f_begin + blockIdx.x; f < f_end; f += gridDim.x) {
        const int par = (flag0 + (int)(f & 1)) & 1;
        const long long tk = (f + 1) * (long long)M2 - 1;        // relative to x[0]
        for (uint32_t b = threadIdx.x; b < M; b += blockDim.x) {
            float2 acc = make_float2(0.f, 0.f);
            // oldest sample first, like window.read() . h_sub (src/dotprod/mod.rs:36-39)
            for (int n = (int)P - 1; n >= 0; n--) {
                const long long t = tk - b - (long long)n * M;
                const float2 s = (t >= 0) ? __ldg(&x[t]) : __ldg(&hist[Hlen + t]);
                const float c = __ldg(&h[b + (size_t)n * M]);
                acc.x = fmaf(c, s.x, acc.x);
                acc.y = fmaf(c, s.y, acc.y);
            }
            uint32_t dst = b + (par ? M2 : 0);
            if (dst >= M) dst -= M;
            X[dst] = acc;
        }
        const float2* r = block_dft(X, Y, M, tw, 1);
        const float Mf = (float)M;
        for (uint32_t c = threadIdx.x; c < M; c += blockDim.x) {
            const float2 v = r[c];
            y[f * (long long)M + c] = make_float2(v.x / Mf, v.y / Mf);
        }
        __syncthreads();
    }
}

// Small M (M <= 64, any even M): one frame per block would leave most of a warp idle, so a 256-thread block takes
// F = 256 / M consecutive frames per pass, thread = (frame slot, branch) for the dot products and (frame slot, bin)
// for the mixed-radix passes of slot_dft out of shared memory (valid for every M; M = 48: 11 complex MACs per bin
// where the direct DFT this kernel used first needs 48).  Shared: 2*F*M samples + M twiddles.
__global__ void __launch_bounds__(256) k_analysis_generic_small(const float* __restrict__ h, const float2* __restrict__ tw,
                                                                const float2* __restrict__ hist, long long Hlen,
                                                                const float2* __restrict__ x, float2* __restrict__ y,
                                                                uint32_t M, uint32_t P /*2m*/, long long f_begin, long long f_end,
                                                                int flag0)
{
    extern __shared__ float2 sm[];
    const uint32_t F = 256 / M, M2 = M >> 1;
    float2* X = sm;
    float2* Y = sm + F * M;
    float2* T = sm + 2 * F * M;
    const uint32_t fs = threadIdx.x / M, b = threadIdx.x - fs * M;
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) T[i] = __ldg(&tw[i]);
    const long long n_groups = (f_end - f_begin + F - 1) / F;
    const float Mf = (float)M;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long f = f_begin + g * F + fs;
        const bool valid = fs < F && f < f_end;
        __syncthreads();                                     // the previous pass has been read (and T is loaded)
        if (valid) {
            const int par = (flag0 + (int)(f & 1)) & 1;
            const long long tk = (f + 1) * (long long)M2 - 1;
            float2 acc = make_float2(0.f, 0.f);
            for (int n = (int)P - 1; n >= 0; n--) {          // oldest sample first (src/dotprod/mod.rs:36-39)
                const long long t = tk - b - (long long)n * M;
                const float2 s = (t >= 0) ? __ldg(&x[t]) : __ldg(&hist[Hlen + t]);
                const float c = __ldg(&h[b + (size_t)n * M]);
                acc.x = fmaf(c, s.x, acc.x);
                acc.y = fmaf(c, s.y, acc.y);
            }
            uint32_t dst = b + (par ? M2 : 0);
            if (dst >= M) dst -= M;
            X[fs * M + dst] = acc;
        }
        __syncthreads();
        const float2* r = slot_dft(X + (fs < F ? fs : 0) * M, Y + (fs < F ? fs : 0) * M, M, T, b, valid);
        if (valid) {
            const float2 acc = r[b];
            y[f * (long long)M + b] = make_float2(acc.x / Mf, acc.y / Mf);
        }
    }
}

// The synthesiser's stage 1 in the same shape: F frames per pass, direct M-point inverse DFT.
__global__ void __launch_bounds__(256) k_synth_ifft_small(const float2* __restrict__ tw, const float2* __restrict__ hist,
                                                          long long hist_frames, const float2* __restrict__ x,
                                                          float2* __restrict__ U, uint32_t M, long long v_begin, long long v_end)
{
    extern __shared__ float2 sm[];
    const uint32_t F = 256 / M;
    float2* X = sm;
    float2* Y = sm + F * M;
    float2* T = sm + 2 * F * M;
    const uint32_t fs = threadIdx.x / M, c = threadIdx.x - fs * M;
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) T[i] = __ldg(&tw[i]);
    const long long n_groups = (v_end - v_begin + F - 1) / F;
    const float s0 = 1.0f / (float)M;
    const float s1 = (float)(M >> 1);
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long v = v_begin + g * F + fs;
        const bool valid = fs < F && v < v_end;
        __syncthreads();
        if (valid) {
            const float2* src = (v >= 0) ? x + v * (long long)M : hist + (hist_frames + v) * (long long)M;
            X[fs * M + c] = __ldg(&src[c]);
        }
        __syncthreads();
        const float2* r = slot_dft(X + (fs < F ? fs : 0) * M, Y + (fs < F ? fs : 0) * M, M, T, c, valid);
        if (valid) {
            float2 acc = r[c];
            acc.x *= s0; acc.y *= s0;                        // two f32 multiplies, as upstream
            acc.x *= s1; acc.y *= s1;
            U[(v - v_begin) * (long long)M + c] = acc;
        }
    }
}

// new_hist[i] = stream[n_new - Hlen + i], stream = concat(old_hist, x[0..n_new))
__global__ void k_update_hist(float2* __restrict__ hist_new, const float2* __restrict__ hist_old, long long Hlen,
                              const float2* __restrict__ x, long long n_new)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < Hlen;
         i += (long long)gridDim.x * blockDim.x) {
        const long long t = n_new - Hlen + i;
        hist_new[i] = (t >= 0) ? x[t] : hist_old[Hlen + t];
    }
}

// Synthesiser stage 1: U[v] = IDFT_unnorm(X_v) * (1/M) * (M/2)   (two f32 multiplies, as upstream)
// for virtual frames v = v_begin .. v_end-1 of the stream (hist ++ x): frame v >= 0 is x[v], frame
// v < 0 is input history.  U[0] corresponds to v_begin.
__global__ void k_synth_ifft(const float2* __restrict__ tw, const float2* __restrict__ hist, long long hist_frames,
                             const float2* __restrict__ x, float2* __restrict__ U, uint32_t M,
                             long long v_begin, long long v_end)
{
    extern __shared__ float2 sm[];
    float2* X = sm;
    float2* Y = sm + M;
    const float s0 = 1.0f / (float)M;
    const float s1 = (float)(M >> 1);
    for (long long v = v_begin + blockIdx.x; v < v_end; v += gridDim.x) {
        const float2* src = (v >= 0) ? x + v * (long long)M : hist + (hist_frames + v) * (long long)M;
        for (uint32_t c = threadIdx.x; c < M; c += blockDim.x) X[c] = __ldg(&src[c]);
        const float2* r = block_dft(X, Y, M, tw, 1);
        for (uint32_t c = threadIdx.x; c < M; c += blockDim.x) {
            float2 vv = r[c];
            vv.x *= s0; vv.y *= s0;
            vv.x *= s1; vv.y *= s1;
            U[(v - v_begin) * (long long)M + c] = vv;
        }
        __syncthreads();
    }
}

// Synthesiser stage 2: weighted overlap-add.  U points at frame 0 of this call; frames -1..-(4m-1)
// (history) precede it in memory.  One thread per output sample.
__global__ void k_synth_wola(const float* __restrict__ h, const float2* __restrict__ U, float2* __restrict__ y,
                             uint32_t M, uint32_t m, long long n_frames, int par0)
{
    // U points at the first frame of this launch; its 4m-1 predecessors precede it in memory.
    // y points at the first output sample of this launch; par0 = parity of that frame.
    const uint32_t M2 = M >> 1;
    const long long total = n_frames * (long long)M2;
    for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total;
         o += (long long)gridDim.x * blockDim.x) {
        const long long f = o / M2;
        const uint32_t i = (uint32_t)(o - f * M2);
        const int par = (par0 + (int)(f & 1)) & 1;
        uint32_t col = i + (par ? M2 : 0);
        // two banks (even / odd l), each summed oldest first, then added (upstream y0 + y1)
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
        for (int n = (int)(2 * m) - 1; n >= 0; n--) {
            const int l0 = 2 * n, l1 = 2 * n + 1;
            const float c0 = __ldg(&h[i + (size_t)l0 * M2]);
            const float c1 = __ldg(&h[i + (size_t)l1 * M2]);
            const float2 u0 = __ldg(&U[(f - l0) * (long long)M + col]);
            const float2 u1 = __ldg(&U[(f - l1) * (long long)M + col]);
            a0.x = fmaf(c0, u0.x, a0.x); a0.y = fmaf(c0, u0.y, a0.y);
            a1.x = fmaf(c1, u1.x, a1.x); a1.y = fmaf(c1, u1.y, a1.y);
        }
        y[o] = make_float2(a0.x + a1.x, a0.y + a1.y);
    }
}

// ------------------------------------------------------------------ tiled generic kernels (G2)
// The kernels above spend ~500 instructions per output bin on run-time index arithmetic.  For every even M whose tile
// fits shared memory the same work is done by a CTA on F consecutive frames at once: the tile's input span, the taps
// ([tap][branch], the prototype's own order) and the twiddles are staged in shared memory, the dot products read
// only shared memory, and the transform runs as mixed-radix Stockham passes over all F frames with the radix-2/3/4/5
// butterflies in registers (other prime factors: one output per thread, r MACs each).  A pass of radix r maps
// butterfly j = jh Ns + k of frame fl from in[fl M + j + i M/r] (times W_M^{i k M / (Ns r)}, no reduction needed:
// i k M / (Ns r) < M) to out[fl M + (jh r + q) Ns + k].
// Shared: T[M] | taps[P M] floats | Xin[(F-1) M/2 + P M] | A[F M] | B[F M]
__global__ void __launch_bounds__(1024) k_analysis_tiled(const float* __restrict__ h, const float2* __restrict__ tw,
                                                        const float2* __restrict__ hist, long long Hlen,
                                                        const float2* __restrict__ x, float2* __restrict__ y,
                                                        uint32_t M, uint32_t P /*2m*/, long long f_begin, long long f_end, int flag0,
                                                        uint32_t F, TiledPass tp)
{
    extern __shared__ float2 sm[];
    const uint32_t M2 = M >> 1;
    float2* T = sm;
    float* taps = reinterpret_cast<float*>(sm + M);
    float2* Xin = sm + M + (P * M + 1) / 2;
    const uint32_t span = (F - 1) * M2 + P * M;
    float2* A = Xin + span;
    float2* B = A + F * M;
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) T[i] = __ldg(&tw[i]);
    for (uint32_t i = threadIdx.x; i < P * M; i += blockDim.x) taps[i] = __ldg(&h[i]);
    const long long n_tiles = (f_end - f_begin + F - 1) / F;
    const float Mf = (float)M;
    for (long long g = blockIdx.x; g < n_tiles; g += gridDim.x) {
        const long long f0 = f_begin + g * F;
        const uint32_t nf = (uint32_t)min((long long)F, f_end - f0);
        // 1. the tile's input span: samples t_start .. t_start + (nf - 1) M/2 + P M - 1 of the stream (history ++ x)
        const long long t_start = (f0 + 1) * (long long)M2 - (long long)P * M;
        const uint32_t n_in = (nf - 1) * M2 + P * M;
        __syncthreads();                                     // the previous tile has been stored (and T, taps are loaded)
        for (uint32_t i = threadIdx.x; i < n_in; i += blockDim.x) {
            const long long t = t_start + i;
            float2 v = make_float2(0.f, 0.f);
            if (t >= 0) v = __ldg(&x[t]);
            else if (Hlen + t >= 0) v = __ldg(&hist[Hlen + t]);
            Xin[i] = v;
        }
        __syncthreads();
        // 2. branch dot products, oldest sample first (src/dotprod/mod.rs:36-39); sample t_k - b - n M sits at
        //    Xin[fl M/2 + P M - 1 - b - n M]
        for (uint32_t it = threadIdx.x; it < nf * M; it += blockDim.x) {
            const uint32_t fl = it / M, b = it - fl * M;
            const float2* xs = Xin + fl * M2 + P * M - 1 - b;
            const float* hs = taps + b;
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
            for (int n = (int)P - 1; n >= 0; n--)            // one packed FFMA2 per tap (complex sample x broadcast real tap)
                acc = __ffma2_rn(xs[-(int)(n * M)], make_float2(hs[n * M], hs[n * M]), acc);
            const int par = (flag0 + (int)((f0 + fl) & 1)) & 1;
            uint32_t dst = b + (par ? M2 : 0);
            if (dst >= M) dst -= M;
            A[fl * M + dst] = acc;
        }
        __syncthreads();
        // 3. transform, 4. store
        const float2* r = tiled_dft(A, B, T, M, tp, nf);
        float2* yo = y + f0 * (long long)M;
        for (uint32_t it = threadIdx.x; it < nf * M; it += blockDim.x) {
            const float2 v = r[it];
            yo[it] = make_float2(v.x / Mf, v.y / Mf);
        }
    }
}

// The synthesiser's stage 1 in the same shape.  Shared: T[M] | A[F M] | B[F M]
__global__ void __launch_bounds__(1024) k_synth_ifft_tiled(const float2* __restrict__ tw, const float2* __restrict__ hist,
                                                          long long hist_frames, const float2* __restrict__ x,
                                                          float2* __restrict__ U, uint32_t M, long long v_begin, long long v_end,
                                                          uint32_t F, TiledPass tp)
{
    extern __shared__ float2 sm[];
    float2* T = sm;
    float2* A = sm + M;
    float2* B = A + F * M;
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) T[i] = __ldg(&tw[i]);
    const long long n_tiles = (v_end - v_begin + F - 1) / F;
    const float s0 = 1.0f / (float)M;
    const float s1 = (float)(M >> 1);
    for (long long g = blockIdx.x; g < n_tiles; g += gridDim.x) {
        const long long v0 = v_begin + g * F;
        const uint32_t nf = (uint32_t)min((long long)F, v_end - v0);
        __syncthreads();
        for (uint32_t it = threadIdx.x; it < nf * M; it += blockDim.x) {
            const uint32_t fl = it / M, c = it - fl * M;
            const long long v = v0 + fl;
            const float2* src = (v >= 0) ? x + v * (long long)M : hist + (hist_frames + v) * (long long)M;
            A[it] = __ldg(&src[c]);
        }
        __syncthreads();
        const float2* r = tiled_dft(A, B, T, M, tp, nf);
        float2* uo = U + (v0 - v_begin) * (long long)M;
        for (uint32_t it = threadIdx.x; it < nf * M; it += blockDim.x) {
            float2 vv = r[it];
            vv.x *= s0; vv.y *= s0;                          // two f32 multiplies, as upstream
            vv.x *= s1; vv.y *= s1;
            uo[it] = vv;
        }
    }
}

// The synthesiser's stage 2 in the same shape: a CTA stages the F + 4m - 1 frames of U its F output frames reach and the
// taps in shared memory.  Shared: taps[4m M/2] floats | Us[(F + 4m - 1) M]
__global__ void __launch_bounds__(1024) k_synth_wola_tiled(const float* __restrict__ h, const float2* __restrict__ U,
                                                           float2* __restrict__ y, uint32_t M, uint32_t m, long long n_frames,
                                                           int par0, uint32_t F)
{
    extern __shared__ float2 sm[];
    const uint32_t M2 = M >> 1, nh = 4 * m - 1, L = 4 * m * M2;
    float* taps = reinterpret_cast<float*>(sm);
    float2* Us = sm + (L + 1) / 2;
    for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) taps[i] = __ldg(&h[i]);
    const long long n_tiles = (n_frames + F - 1) / F;
    for (long long g = blockIdx.x; g < n_tiles; g += gridDim.x) {
        const long long f0 = g * F;
        const uint32_t nf = (uint32_t)min((long long)F, n_frames - f0);
        __syncthreads();                                     // the previous tile has been read (and the taps are loaded)
        const float2* src = U + (f0 - (long long)nh) * (long long)M;             // the launch's U has its 4m-1 predecessors in front
        for (uint32_t i = threadIdx.x; i < (nf + nh) * M; i += blockDim.x) Us[i] = __ldg(&src[i]);
        __syncthreads();
        for (uint32_t it = threadIdx.x; it < nf * M2; it += blockDim.x) {
            const uint32_t fl = it / M2, i = it - fl * M2;
            const int par = (par0 + (int)((f0 + fl) & 1)) & 1;
            const float2* us = Us + (fl + nh) * M + i + (par ? M2 : 0);          // frame f of column col
            const float* hs = taps + i;
            // two banks (even / odd lag), each summed oldest first, then added (upstream y0 + y1)
            float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll 4
            for (int n = (int)(2 * m) - 1; n >= 0; n--) {
                const float c0 = hs[(2 * n) * M2], c1 = hs[(2 * n + 1) * M2];
                a0 = __ffma2_rn(us[-(int)((2 * n) * M)], make_float2(c0, c0), a0);
                a1 = __ffma2_rn(us[-(int)((2 * n + 1) * M)], make_float2(c1, c1), a1);
            }
            y[f0 * (long long)M2 + it] = make_float2(a0.x + a1.x, a0.y + a1.y);
        }
    }
}

int32_t check(yg_firpfbch2_crcf q)
{
    if (!q) return fail(YG_EVALUE, "null firpfbch2 handle");
    return YG_OK;
}

int32_t validate(int32_t type, uint32_t M, uint32_t m)
{
    if (type != YG_ANALYZER && type != YG_SYNTHESIZER) return fail(YG_ECONFIG, "invalid type %d", type);
    if (M < 2) return fail(YG_ECONFIG, "number of channels must be at least 2");
    if (M % 2) return fail(YG_ECONFIG, "number of channels must be even");
    if (m < 1) return fail(YG_ECONFIG, "filter semi-length must be at least 1");
    return YG_OK;
}

size_t smem_dft(uint32_t M) { return 2 * (size_t)M * sizeof(float2); }

int32_t set_smem(const void* fn, size_t bytes)
{
    if (bytes > 48 * 1024) {
        if (bytes > 227 * 1024) return fail(YG_ECONFIG, "M too large for the generic kernel (needs %zu B shared memory)", bytes);
        YG_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    return YG_OK;
}

int32_t launch_generic_analysis(yg_firpfbch2_crcf q, const float2* hist, const float2* x, float2* y,
                                size_t f_begin, size_t f_end, cudaStream_t st)
{
    if (f_end <= f_begin) return YG_OK;
    if (q->tiled.supported) {                // F frames per CTA pass out of shared memory
        const TiledPass tp = q->tiled.tp;
        const long long tiles = ((long long)(f_end - f_begin) + q->tiled.F - 1) / q->tiled.F;
        const int grid_t = (int)std::min<long long>(tiles, (long long)q->n_sm * q->tiled.ctas_per_sm);
        k_analysis_tiled<<<grid_t, q->tiled.threads, q->tiled.smem, st>>>(q->d_h.p, q->d_tw.p, hist, (long long)q->hist_len, x, y, q->M, 2 * q->m,
                                                             (long long)f_begin, (long long)f_end, q->flag, (uint32_t)q->tiled.F, tp);
        YG_LAUNCH_CHECK();
        return YG_OK;
    }
    if (q->M <= 64) {                        // several frames per block
        const uint32_t F = 256 / q->M;
        const size_t smem_s = (2 * (size_t)F * q->M + q->M) * sizeof(float2);
        const long long groups = ((long long)(f_end - f_begin) + F - 1) / F;
        const int grid_s = (int)std::min<long long>(groups, q->n_sm * 8);
        k_analysis_generic_small<<<grid_s, 256, smem_s, st>>>(q->d_h.p, q->d_tw.p, hist, (long long)q->hist_len, x, y, q->M,
                                                              2 * q->m, (long long)f_begin, (long long)f_end, q->flag);
        YG_LAUNCH_CHECK();
        return YG_OK;
    }
    const size_t smem = smem_dft(q->M);
    YG_TRY(set_smem((const void*)k_analysis_generic, smem));
    const int block = (int)std::min<uint32_t>(256, (q->M + 31) / 32 * 32);
    const int grid = (int)std::min<size_t>(f_end - f_begin, q->n_sm * 16);
    k_analysis_generic<<<grid, block, smem, st>>>(q->d_h.p, q->d_tw.p, hist, (long long)q->hist_len, x, y, q->M,
                                                  2 * q->m, (long long)f_begin, (long long)f_end, q->flag);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

// *hist_done is set when the kernel that took the call also wrote the next state into d_hist[cur ^ 1]
int32_t launch_analysis(yg_firpfbch2_crcf q, const yg_cf32* d_x, size_t n_frames, yg_cf32* d_y, cudaStream_t st, bool* hist_done)
{
    const float2* hist = reinterpret_cast<const float2*>(q->d_hist[q->cur].p);
    const float2* x = reinterpret_cast<const float2*>(d_x);
    float2* y = reinterpret_cast<float2*>(d_y);

    // The fused kernel handles an even-parity start and an even number of frames; what is
    // left over (a leading odd-parity frame, a trailing single frame) and calls too small to
    // fill the machine go to the generic kernel.
    if (q->timing_on) q->next_events();
    // the fused kernels read the input with TMA bulk copies, which need 16-byte aligned sources
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    const bool use_fused = aligned && q->fast.supported && n_frames >= q->fast.min_frames;
    const bool use_large = q->large.supported && n_frames >= q->large.min_frames;
    const bool use_small = aligned && q->small.supported && n_frames >= q->small.min_frames;
    const bool use_tiny = aligned && q->tiny.supported && n_frames >= q->tiny.min_frames && (reinterpret_cast<uintptr_t>(y) & 15) == 0;
    if (use_fused || use_large || use_small || use_tiny) {
        const size_t lead = (q->flag & 1) ? 1 : 0;
        const size_t body = (n_frames - lead) & ~(size_t)1;
        if (q->timing_on) YG_CUDA(cudaEventRecord(q->ev0, st));
        if (use_fused) {
            YG_TRY(firpfbch2_fast_launch(q->fast, hist, (long long)q->hist_len, x, y, lead, body, st,
                                         reinterpret_cast<float2*>(q->d_hist[q->cur ^ 1].p), (long long)n_frames * q->M2));
            *hist_done = true;
        }
        else if (use_small) YG_TRY(firpfbch2_small_launch(q->small, hist, (long long)q->hist_len, x, y, lead, body, st));
        else if (use_tiny) YG_TRY(firpfbch2_tiny_launch(q->tiny, hist, (long long)q->hist_len, x, y, lead, body, st));
        else YG_TRY(firpfbch2_large_launch(q->large, hist, (long long)q->hist_len, x, y, lead, body, st,
                                           reinterpret_cast<float2*>(q->d_hist[q->cur ^ 1].p), (long long)n_frames * q->M2, hist_done));
        if (q->timing_on) YG_CUDA(cudaEventRecord(q->ev1, st));
        q->timed = q->timing_on;
        q->last_path = (use_fused || use_small || use_tiny) ? 2 : 3;
        YG_TRY(launch_generic_analysis(q, hist, x, y, 0, lead, st));
        YG_TRY(launch_generic_analysis(q, hist, x, y, lead + body, n_frames, st));
        return YG_OK;
    }
    if (q->timing_on) YG_CUDA(cudaEventRecord(q->ev0, st));
    YG_TRY(launch_generic_analysis(q, hist, x, y, 0, n_frames, st));
    if (q->timing_on) YG_CUDA(cudaEventRecord(q->ev1, st));
    q->timed = q->timing_on;
    q->last_path = 1;
    return YG_OK;
}

int32_t launch_generic_synthesis(yg_firpfbch2_crcf q, const float2* hist, const float2* x, float2* y,
                                 size_t f_begin, size_t f_end, cudaStream_t st)
{
    if (f_end <= f_begin) return YG_OK;
    const uint32_t M = q->M;
    const long long nh = 4 * (long long)q->m - 1;
    // bounded scratch: at most ~64 MiB of U per pass (the 4m-1 frames before each pass are transformed again)
    const long long chunk = std::max<long long>(2 * nh, ((long long)64 << 20) / ((long long)M * 8));
    YG_TRY(q->d_U.reserve((size_t)(nh + std::min<long long>(chunk, (long long)(f_end - f_begin))) * M));
    float2* U = reinterpret_cast<float2*>(q->d_U.p);
    const size_t smem = smem_dft(M);
    YG_TRY(set_smem((const void*)k_synth_ifft, smem));
    const int block = (int)std::min<uint32_t>(256, (M + 31) / 32 * 32);
    for (long long f0 = (long long)f_begin; f0 < (long long)f_end; f0 += chunk) {
        const long long nf = std::min<long long>(chunk, (long long)f_end - f0);
        if (q->tiled.supported) {
            const TiledPass tp = q->tiled.tp;
            const long long tiles = (nh + nf + q->tiled.F - 1) / q->tiled.F;
            const int grid_t = (int)std::min<long long>(tiles, (long long)q->n_sm * q->tiled.ctas_per_sm);
            k_synth_ifft_tiled<<<grid_t, q->tiled.threads, q->tiled.smem, st>>>(q->d_tw.p, hist, (long long)(q->hist_len / M), x, U, M, f0 - nh, f0 + nf,
                                                                   (uint32_t)q->tiled.F, tp);
        } else if (M <= 64) {            // several frames per block
            const uint32_t F = 256 / M;
            const size_t smem_s = (2 * (size_t)F * M + M) * sizeof(float2);
            const int grid_s = (int)std::min<long long>((nh + nf + F - 1) / F, q->n_sm * 8);
            k_synth_ifft_small<<<grid_s, 256, smem_s, st>>>(q->d_tw.p, hist, (long long)(q->hist_len / M), x, U, M, f0 - nh, f0 + nf);
        } else {
            const int grid = (int)std::min<long long>(nh + nf, q->n_sm * 16);
            k_synth_ifft<<<grid, block, smem, st>>>(q->d_tw.p, hist, (long long)(q->hist_len / M), x, U, M, f0 - nh, f0 + nf);
        }
        YG_LAUNCH_CHECK();
        const long long total = nf * q->M2;
        const int grid2 = (int)std::min<long long>((total + 255) / 256, q->n_sm * 32);
        if (q->tiled.wF > 0) {
            const long long tiles = (nf + q->tiled.wF - 1) / q->tiled.wF;
            const int grid_w = (int)std::min<long long>(tiles, (long long)q->n_sm * q->tiled.wctas);
            k_synth_wola_tiled<<<grid_w, q->tiled.wthreads, q->tiled.wsmem, st>>>(q->d_h.p, U + nh * M, y + f0 * q->M2, M, q->m, nf,
                                                                                  (q->flag + (int)(f0 & 1)) & 1, (uint32_t)q->tiled.wF);
        } else
        k_synth_wola<<<grid2, 256, 0, st>>>(q->d_h.p, U + nh * M, y + f0 * q->M2, M, q->m, nf, (q->flag + (int)(f0 & 1)) & 1);
        YG_LAUNCH_CHECK();
    }
    return YG_OK;
}

int32_t launch_synthesis(yg_firpfbch2_crcf q, const yg_cf32* d_x, size_t n_frames, yg_cf32* d_y, cudaStream_t st, bool* hist_done)
{
    const float2* hist = reinterpret_cast<const float2*>(q->d_hist[q->cur].p);
    const float2* x = reinterpret_cast<const float2*>(d_x);
    float2* y = reinterpret_cast<float2*>(d_y);
    if (q->timing_on) q->next_events();
    // The fused kernel takes an even-parity start and whole rounds of 32 frames; the rest goes to
    // the generic kernels.  All of them read the same (history ++ x) stream and write disjoint
    // output ranges, so their order does not matter.
    const size_t lead = (q->flag & 1) ? 1 : 0;
    const size_t body = (n_frames > lead) ? ((n_frames - lead) / 32) * 32 : 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    const bool use_fused = aligned && q->sfast.supported && body >= q->sfast.min_frames;
    const bool use_large = q->slarge.supported && body >= q->slarge.min_frames;
    if (use_fused || use_large) {
        if (use_large && firpfbch2_large_synth_needs_scratch(q->slarge, hist, x))
            YG_TRY(q->d_Uc.reserve((size_t)firpfbch2_large_synth_scratch_frames(q->M) * q->M));
        if (q->timing_on) YG_CUDA(cudaEventRecord(q->ev0, st));
        if (use_fused && q->M == 256) YG_TRY(firpfbch2_synth_fast_launch(q->sfast, hist, x, y, lead, body, st));
        else if (use_fused && q->M >= 64) YG_TRY(firpfbch2_small_synth_launch(q->sfast, hist, x, y, lead, body, st));
        else if (use_fused) YG_TRY(firpfbch2_tiny_synth_launch(q->sfast, hist, x, y, lead, body, st));
        else YG_TRY(firpfbch2_large_synth_launch(q->slarge, hist, x, y, reinterpret_cast<float2*>(q->d_Uc.p), lead, body, st,
                                                 reinterpret_cast<float2*>(q->d_hist[q->cur ^ 1].p), (long long)n_frames * q->M, hist_done));
        if (q->timing_on) YG_CUDA(cudaEventRecord(q->ev1, st));
        q->timed = q->timing_on;
        q->last_path = use_fused ? 2 : 3;
        YG_TRY(launch_generic_synthesis(q, hist, x, y, 0, lead, st));
        YG_TRY(launch_generic_synthesis(q, hist, x, y, lead + body, n_frames, st));
        return YG_OK;
    }
    if (q->timing_on) YG_CUDA(cudaEventRecord(q->ev0, st));
    YG_TRY(launch_generic_synthesis(q, hist, x, y, 0, n_frames, st));
    if (q->timing_on) YG_CUDA(cudaEventRecord(q->ev1, st));
    q->timed = q->timing_on;
    q->last_path = 1;
    return YG_OK;
}

int32_t execute_dev_impl(yg_firpfbch2_crcf q, const yg_cf32* d_x, size_t n_frames, yg_cf32* d_y, cudaStream_t st);

int32_t execute_dev(yg_firpfbch2_crcf q, const yg_cf32* d_x, size_t n_frames, yg_cf32* d_y, cudaStream_t st)
{
    YG_TRY(q->order.enter(st));
    YG_TRY(execute_dev_impl(q, d_x, n_frames, d_y, st));
    return q->order.leave(st);
}

int32_t execute_dev_impl(yg_firpfbch2_crcf q, const yg_cf32* d_x, size_t n_frames, yg_cf32* d_y, cudaStream_t st)
{
    if (n_frames == 0) return YG_OK;
    bool hist_done = false;
    if (q->type == YG_ANALYZER) YG_TRY(launch_analysis(q, d_x, n_frames, d_y, st, &hist_done));
    else YG_TRY(launch_synthesis(q, d_x, n_frames, d_y, st, &hist_done));
    // both types keep the tail of their INPUT stream as state
    const long long n_new = (long long)n_frames * (q->type == YG_ANALYZER ? q->M2 : q->M);
    const long long Hlen = (long long)q->hist_len;
    const int nxt = q->cur ^ 1;
    if (!hist_done) {
        const int grid = (int)std::min<long long>((Hlen + 255) / 256, 1024);
        k_update_hist<<<grid, 256, 0, st>>>(reinterpret_cast<float2*>(q->d_hist[nxt].p),
                                            reinterpret_cast<const float2*>(q->d_hist[q->cur].p), Hlen,
                                            reinterpret_cast<const float2*>(d_x), n_new);
        YG_LAUNCH_CHECK();
    }
    q->cur = nxt;
    q->flag = (q->flag + (int)(n_frames & 1)) & 1;
    return YG_OK;
}

// Tile geometry of the tiled generic kernels for this object (on its device): the prime factors of M (4 before 2, then
// odd primes ascending), the largest tile of F frames that fits ~190 KB of shared memory (F M <= 8192, F <= 64).
int32_t plan_tiled(yg_firpfbch2_crcf q)
{
    auto& t = q->tiled;
    t.supported = false;
    const char* e = getenv("YG_GENERIC_TILED");           // debugging knob: 0 keeps the one-frame-per-block kernels
    if (e && e[0] == '0') return YG_OK;
    const size_t M = q->M, P = 2 * (size_t)q->m;
    // stage 2 of the synthesiser: F + 4m - 1 frames of U and the taps in shared memory -- pays up to M ~ 100 (M = 24:
    // 27.5 -> 31.5 Gsps); beyond that the tile holds few frames and the one-thread-per-output kernel is as good or better
    if (q->type == YG_SYNTHESIZER && M <= 128) {
        const size_t nh = 2 * P - 1;
        auto wbytes = [&](size_t F) { return 8 * ((P * M + 1) / 2 + (F + nh) * M); };
        size_t F = std::min<size_t>(64, std::max<size_t>(2, 16384 / M));
        while (F > 1 && wbytes(F) > 190 * 1024) F--;
        if (wbytes(F) <= 190 * 1024 && F >= 2) {
            t.wF = (int)F;
            t.wsmem = wbytes(F);
            YG_CUDA(cudaFuncSetAttribute(k_synth_wola_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 190 * 1024));
            t.wctas = (int)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (t.wsmem + 1024)));
            t.wthreads = std::min(1024, (2048 / t.wctas) & ~31);
        }
    }
    if (!plan_radices(q->M, t.tp)) return YG_OK;          // a large prime factor: the one-frame-per-block kernel takes it
    auto bytes = [&](size_t F) {
        return q->type == YG_ANALYZER ? 8 * (M + (P * M + 1) / 2 + (F - 1) * (M / 2) + P * M + 2 * F * M) : 8 * (M + 2 * F * M);
    };
    // (smaller tiles for more resident CTAs were measured: the analyser loses -- every tile reloads P M samples of
    // history and the taps -- and the synthesiser does not gain)
    const size_t budget = 190 * 1024;
    size_t F = std::min<size_t>(64, std::max<size_t>(1, 8192 / M));
    while (F > 1 && bytes(F) > budget) F--;
    if (bytes(F) > budget) return YG_OK;
    // the synthesiser's stage 1 at a large power-of-two M is served better by the one-frame-per-block radix-4 kernel
    // (16 independent blocks per SM; measured M = 256: 19.4 vs 16.8 Gsps, M = 1024: 23.6 vs 16.4)
    if (q->type == YG_SYNTHESIZER && M >= 128 && (M & (M - 1)) == 0) return YG_OK;
    // ... and so is an analyser at a power-of-two M whose tile holds fewer than 8 frames (taps and history are reloaded per
    // tile; measured firpfbch M = 1024, p = 8, F = 3: 35 vs 47 Gsps)
    if (q->type == YG_ANALYZER && (M & (M - 1)) == 0 && F < 8) return YG_OK;
    t.F = (int)F;
    t.smem = bytes(F);
    if (q->type == YG_ANALYZER) YG_CUDA(cudaFuncSetAttribute(k_analysis_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    else YG_CUDA(cudaFuncSetAttribute(k_synth_ifft_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    t.ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (t.smem + 1024)));
    t.threads = std::min(1024, (2048 / t.ctas_per_sm) & ~31);
    t.supported = true;
    return YG_OK;
}

int32_t build(int32_t type, uint32_t M, uint32_t m, const float* h, size_t h_len, yg_firpfbch2_crcf* out)
{
    if (!out) return fail(YG_EVALUE, "null output pointer");
    *out = nullptr;
    YG_TRY(validate(type, M, m));
    const size_t L = 2 * (size_t)M * m;
    if (!h) return fail(YG_EVALUE, "null prototype filter");
    if (h_len < L) return fail(YG_ECONFIG, "prototype filter length (%zu) must be at least 2*M*m (%zu)", h_len, L);
    int dev = 0;
    YG_TRY(require_device(&dev));

    auto* q = new yg_firpfbch2_crcf_s();
    q->type = type; q->M = M; q->M2 = M / 2; q->m = m; q->L = L; q->dev = dev;
    q->n_sm = sm_count(dev);
    q->h.assign(h, h + L);
    auto cleanup = [&](int32_t rc) { yg_firpfbch2_crcf_destroy(q); return rc; };
#define TRYQ(expr) do { int32_t _rc = (expr); if (_rc != YG_OK) return cleanup(_rc); } while (0)
#define CUDAQ(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return cleanup(fail(YG_EINTERNAL, "CUDA error %s (%s)", cudaGetErrorString(_e), #expr)); } while (0)
    CUDAQ(cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking));
    q->order.own = q->stream;
    for (int i = 0; i < yg_firpfbch2_crcf_s::kRing; i++) {
        CUDAQ(cudaEventCreate(&q->ev0s[i]));
        CUDAQ(cudaEventCreate(&q->ev1s[i]));
    }
    TRYQ(q->d_h.reserve(L));
    CUDAQ(yg::memcpy_sync(q->d_h.p, q->h.data(), L * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<float2> tw;
    make_twiddles(M, tw);
    TRYQ(q->d_tw.reserve(M));
    CUDAQ(yg::memcpy_sync(q->d_tw.p, tw.data(), M * sizeof(float2), cudaMemcpyHostToDevice));
    const size_t nh = 4 * (size_t)m - 1;
    q->state_len = (type == YG_ANALYZER) ? nh * q->M2 : nh * M;
    // the synthesiser keeps at least 32 input frames so the fused kernel can warm up a whole round
    q->hist_len = (type == YG_ANALYZER) ? q->state_len : std::max<size_t>(nh, 32) * M;
    for (int b = 0; b < 2; b++) {
        TRYQ(q->d_hist[b].reserve(q->hist_len));
        CUDAQ(yg::memset_sync(q->d_hist[b].p, 0, q->hist_len * sizeof(yg_cf32)));
    }
    TRYQ(plan_tiled(q));
    if (type == YG_ANALYZER) TRYQ(firpfbch2_fast_plan(q->fast, M, m, q->h.data()));
    if (type == YG_ANALYZER) TRYQ(firpfbch2_large_plan(q->large, M, m, q->h.data()));
    if (type == YG_ANALYZER) TRYQ(firpfbch2_small_plan(q->small, M, m, q->h.data()));
    if (type == YG_ANALYZER) TRYQ(firpfbch2_tiny_plan(q->tiny, M, m, q->h.data()));
    else {
        TRYQ(firpfbch2_synth_fast_plan(q->sfast, M, m, q->h.data()));
        if (!q->sfast.supported) TRYQ(firpfbch2_small_synth_plan(q->sfast, M, m, q->h.data()));
        if (!q->sfast.supported) TRYQ(firpfbch2_tiny_synth_plan(q->sfast, M, m, q->h.data()));
        TRYQ(firpfbch2_large_synth_plan(q->slarge, M, m, q->h.data()));
    }
#undef TRYQ
#undef CUDAQ
    *out = q;
    return YG_OK;
}

}  // namespace

// ------------------------------------------------------------------ C ABI
extern "C" {

int32_t yg_firpfbch2_crcf_create(int32_t type, uint32_t M, uint32_t m, const float* h, size_t h_len,
                                 yg_firpfbch2_crcf* out)
{
    return build(type, M, m, h, h_len, out);
}

int32_t yg_firpfbch2_crcf_create_kaiser(int32_t type, uint32_t M, uint32_t m, float as, yg_firpfbch2_crcf* out)
{
    if (!out) return fail(YG_EVALUE, "null output pointer");
    *out = nullptr;
    YG_TRY(validate(type, M, m));
    const uint32_t n = 2 * M * m + 1;
    std::vector<float> hf(n);
    const float fc = (type == YG_ANALYZER) ? 1.0f / (float)M : 0.5f / (float)M;
    YG_TRY(fir_design_kaiser(n, fc, as, 0.0f, hf.data()));
    float sum = 0.0f;
    for (uint32_t i = 0; i < n; i++) sum += hf[i];
    for (uint32_t i = 0; i < n; i++) hf[i] = hf[i] * (float)M / sum;      // resamp.rs:49-51 idiom
    return build(type, M, m, hf.data(), n, out);
}

int32_t yg_firpfbch2_crcf_clone(yg_firpfbch2_crcf q, yg_firpfbch2_crcf* out)
{
    YG_TRY(check(q));
    YG_DEVICE_GUARD(q->dev);
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    yg_firpfbch2_crcf c = nullptr;
    YG_TRY(build(q->type, q->M, q->m, q->h.data(), q->h.size(), &c));
    cudaError_t e = yg::memcpy_sync(c->d_hist[c->cur].p, q->d_hist[q->cur].p, q->hist_len * sizeof(yg_cf32),
                               cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) { yg_firpfbch2_crcf_destroy(c); return fail(YG_EINTERNAL, "CUDA error %s", cudaGetErrorString(e)); }
    c->flag = q->flag;
    *out = c;
    return YG_OK;
}

int32_t yg_firpfbch2_crcf_destroy(yg_firpfbch2_crcf q)
{
    if (!q) return YG_OK;
    YG_DEVICE_GUARD(q->dev);
    if (q->stream) cudaStreamSynchronize(q->stream);
    q->order.wait_host();
    q->order.destroy();
    firpfbch2_fast_release(q->fast);
    firpfbch2_fast_release(q->sfast);
    firpfbch2_fast_release(q->large);
    firpfbch2_fast_release(q->small);
    firpfbch2_fast_release(q->tiny);
    firpfbch2_fast_release(q->slarge);
    q->d_Uc.release();
    q->pipe.destroy();
    q->d_h.release(); q->d_tw.release(); q->d_hist[0].release(); q->d_hist[1].release(); q->d_U.release();
    for (int i = 0; i < yg_firpfbch2_crcf_s::kRing; i++) {
        if (q->ev0s[i]) cudaEventDestroy(q->ev0s[i]);
        if (q->ev1s[i]) cudaEventDestroy(q->ev1s[i]);
    }
    if (q->stream) cudaStreamDestroy(q->stream);
    delete q;
    return YG_OK;
}

int32_t yg_firpfbch2_crcf_reset(yg_firpfbch2_crcf q)
{
    YG_TRY(check(q));
    YG_DEVICE_GUARD(q->dev);
    YG_TRY(q->order.wait_host());
    YG_CUDA(cudaMemsetAsync(q->d_hist[q->cur].p, 0, q->hist_len * sizeof(yg_cf32), q->stream));
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    q->flag = 0;
    return YG_OK;
}

int32_t yg_firpfbch2_crcf_execute_block_dev(yg_firpfbch2_crcf q, const yg_cf32* d_x, size_t n_frames, yg_cf32* d_y,
                                            void* cuda_stream)
{
    YG_TRY(check(q));
    if (n_frames && (!d_x || !d_y)) return fail(YG_EVALUE, "null buffer");
    YG_DEVICE_GUARD(q->dev);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    return execute_dev(q, d_x, n_frames, d_y, st);
}

int32_t yg_firpfbch2_crcf_execute_block(yg_firpfbch2_crcf q, const yg_cf32* x, size_t n_frames, yg_cf32* y)
{
    YG_TRY(check(q));
    if (n_frames && (!x || !y)) return fail(YG_EVALUE, "null buffer");
    YG_DEVICE_GUARD(q->dev);
    const size_t nin = (q->type == YG_ANALYZER) ? q->M2 : q->M;
    const size_t nout = (q->type == YG_ANALYZER) ? q->M : q->M2;
    // ~32 MiB of input per chunk
    size_t fpc = ((size_t)32 << 20) / (nin * sizeof(yg_cf32));
    fpc = std::max<size_t>(32, fpc & ~(size_t)31);      // whole 32-frame rounds keep chunk starts on the fused paths
    return run_host_pipe(q->pipe, q->stream, x, y, n_frames, nin, nout, fpc,
                         [&](const yg_cf32* dx, size_t f, yg_cf32* dy, cudaStream_t st) {
                             return execute_dev(q, dx, f, dy, st);
                         });
}

int32_t yg_firpfbch2_crcf_execute(yg_firpfbch2_crcf q, const yg_cf32* x, yg_cf32* y)
{
    return yg_firpfbch2_crcf_execute_block(q, x, 1, y);
}

int32_t yg_firpfbch2_crcf_sync(yg_firpfbch2_crcf q)
{
    YG_TRY(check(q));
    YG_DEVICE_GUARD(q->dev);
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    return YG_OK;
}

int32_t yg_firpfbch2_crcf_get_type(yg_firpfbch2_crcf q, int32_t* type) { YG_TRY(check(q)); *type = q->type; return YG_OK; }
int32_t yg_firpfbch2_crcf_get_M(yg_firpfbch2_crcf q, uint32_t* M) { YG_TRY(check(q)); *M = q->M; return YG_OK; }
int32_t yg_firpfbch2_crcf_get_m(yg_firpfbch2_crcf q, uint32_t* m) { YG_TRY(check(q)); *m = q->m; return YG_OK; }

int32_t yg_firpfbch2_crcf_get_taps(yg_firpfbch2_crcf q, float* h)
{
    YG_TRY(check(q));
    if (!h) return fail(YG_EVALUE, "null pointer");
    memcpy(h, q->h.data(), q->L * sizeof(float));
    return YG_OK;
}

int32_t yg_firpfbch2_crcf_state_len(yg_firpfbch2_crcf q, size_t* n)
{
    YG_TRY(check(q));
    *n = q->state_len;
    return YG_OK;
}

int32_t yg_firpfbch2_crcf_get_state(yg_firpfbch2_crcf q, yg_cf32* hist, int32_t* flag)
{
    YG_TRY(check(q));
    YG_DEVICE_GUARD(q->dev);
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    if (hist) YG_CUDA(yg::memcpy_sync(hist, q->d_hist[q->cur].p + (q->hist_len - q->state_len), q->state_len * sizeof(yg_cf32), cudaMemcpyDeviceToHost));
    if (flag) *flag = q->flag;
    return YG_OK;
}

int32_t yg_firpfbch2_crcf_set_state(yg_firpfbch2_crcf q, const yg_cf32* hist, int32_t flag)
{
    YG_TRY(check(q));
    if (flag != 0 && flag != 1) return fail(YG_EVALUE, "flag must be 0 or 1");
    YG_DEVICE_GUARD(q->dev);
    YG_CUDA(cudaStreamSynchronize(q->stream));
    YG_TRY(q->order.wait_host());
    if (hist) {
        YG_CUDA(yg::memset_sync(q->d_hist[q->cur].p, 0, q->hist_len * sizeof(yg_cf32)));
        YG_CUDA(yg::memcpy_sync(q->d_hist[q->cur].p + (q->hist_len - q->state_len), hist, q->state_len * sizeof(yg_cf32), cudaMemcpyHostToDevice));
    }
    q->flag = flag;
    return YG_OK;
}

int32_t yg_firpfbch2_crcf_last_path(yg_firpfbch2_crcf q, int32_t* path) { YG_TRY(check(q)); *path = q->last_path; return YG_OK; }

int32_t yg_firpfbch2_crcf_get_device(yg_firpfbch2_crcf q, int32_t* dev) { YG_TRY(check(q)); *dev = q->dev; return YG_OK; }

int32_t yg_firpfbch2_crcf_set_kernel_timing(yg_firpfbch2_crcf q, int32_t enable)
{
    YG_TRY(check(q));
    q->timing_on = enable != 0;
    if (!q->timing_on) { q->timed = false; q->n_timed = 0; }
    return YG_OK;
}

int32_t yg_firpfbch2_crcf_last_kernel_ms(yg_firpfbch2_crcf q, float* ms)
{
    YG_TRY(check(q));
    if (!q->timed) return fail(YG_EMODE, "no timed launch yet");
    YG_DEVICE_GUARD(q->dev);
    YG_CUDA(cudaEventSynchronize(q->ev1));
    YG_CUDA(cudaEventElapsedTime(ms, q->ev0, q->ev1));
    return YG_OK;
}

int32_t yg_firpfbch2_crcf_kernel_times(yg_firpfbch2_crcf q, float* ms, size_t cap, size_t* n)
{
    YG_TRY(check(q));
    if (!ms || !n) return fail(YG_EVALUE, "null pointer");
    YG_DEVICE_GUARD(q->dev);
    const unsigned long long have = std::min<unsigned long long>(q->n_timed, yg_firpfbch2_crcf_s::kRing);
    const size_t take = (size_t)std::min<unsigned long long>(have, cap);
    for (size_t i = 0; i < take; i++) {
        const unsigned long long idx = (q->n_timed - take + i) % yg_firpfbch2_crcf_s::kRing;   // oldest first
        YG_CUDA(cudaEventSynchronize(q->ev1s[idx]));
        YG_CUDA(cudaEventElapsedTime(&ms[i], q->ev0s[idx], q->ev1s[idx]));
    }
    *n = take;
    return YG_OK;
}

}  // extern "C"
